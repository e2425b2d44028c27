// the training-side reuse of the distance/exp engine.
//
//   k_kernel_mat  -- the force-field kernel matrix of the closed-form trainer: one CTA per ordered pair of
//                    training points (i, j) writes the 3N x 3N Hessian-of-Matern block K[blk_i][blk_j] and,
//                    with energy constraints, its energy row/column entries.
//                    Reference: mbgdml/_gdml/train.py:92-292 (_assemble_kernel_mat_wkr), 1105-1337 (driver).
//   k_mat52       -- Matern-5/2 covariances of query structures against the T*P permuted training rows.
//                    Reference: mbgdml/analysis/models.py:31-139 (gdml_mat52_wrk, gdml_mat52).
//
// Kernel-matrix math (train.py:183-218), pi_p(k) = tril_perms_lin[k P + p] mod D, J = full Jacobian (D x 3N,
// row k <-> atoms (a_k > b_k): +g_k at b_k, -g_k at a_k, desc.py:397-444):
//     diff_p = X_i - X_j[pi_p]        n_p = sqrt(5)|diff_p|       b_p = exp(-n_p/sig) 5/(3 sig^4)
//     c_p = (sig^2 + sig n_p) b_p     u_p = J_i^T diff_p          v_p = (J_j[pi_p])^T diff_p
//     K_ij = sum_p 5 b_p u_p v_p^T  -  J_i^T W,      W[k][c] = sum_p c_p J_j[pi_p(k)][c]
// (the reference forms M = sum_p 5 b_p diff_p v_p^T - W and then J_i^T M; pulling J_i^T through the rank-1
// term makes it P outer products of 3N-vectors instead of D x 3N ones).  Energy constraints (train.py:224-289):
//     K[E+i][blk_j] = K[blk_j][E+i] = -sum_p c_p v_p          K[E+j][E+i] = -sum_p (1 + (n/sig)(1 + n/(3 sig))) e^{-n/sig}
// Permutations are processed in chunks of kKmPC so that shared memory does not grow with P.  When the permutation
// set is closed under inversion (always, for symmetry groups; checked on the host) K_ji = K_ij^T and the energy
// entries of the swapped pair follow from u_p, so only the pairs i >= j are computed (the reference's exploit_sym).
#pragma once
#include "common.cuh"

namespace mbg {

constexpr int kKmPC = 16;        // permutations per chunk
constexpr int kKmThreads = 256;  // threads per CTA

struct KernelMatArgs {
  int T, P;
  double sig;
  int use_E;
  int symmetric;     // 1: the permutation set is closed under inversion -> K_ji = K_ij^T; one CTA per pair i >= j
  const double* X;   // [T][D]      R_desc
  const double* G;   // [T][D][3]   compressed Jacobian (desc.py:448-480)
  const int* pi;     // [P][D]      pi_p(k)
  const int* ipi;    // [P][D]      pi_p^-1(k)
  double* K;         // [rows][rows], rows = T 3N (+ T)
  long long ld;      // rows
  // tensor-pipe kernel only: for every descriptor k the permutations ordered by their image pi_p(k):
  const int* srt_p;  // [P][D]  srt_p[t D + k] = the t-th permutation of row k
  const int* srt_k;  // [P][D]  srt_k[t D + k] = its image pi_p(k) (ascending in t)
  // column-block subset (mbg_kernel_mat_cols): pairs (i, j_list[jj]) for all i, block written at column block jj of a
  // [T 3N][n_j 3N] matrix; NULL for the full matrix.  Never with use_E or symmetric.
  const int* j_list;
  int n_j;
};

template <int NA>
struct KmShape {
  static constexpr int N = NA, D = NA * (NA - 1) / 2, N3 = 3 * NA;
  static constexpr int NOUT = (N3 * N3 + kKmThreads - 1) / kKmThreads;  // outputs per thread
  // doubles: Xi, Xj, gi, gj, diff, (fb, cc, kee, pad), u, fu, v, W, out block; ints: pi/ipi chunk
  static constexpr size_t kDoubles = 2 * D + 6 * D + (size_t)kKmPC * D + 4 * kKmPC + 3 * (size_t)kKmPC * N3 +
                                     (size_t)D * N3 + (size_t)N3 * (N3 + 1);
  static constexpr size_t kBytes = kDoubles * 8 + (2 * (size_t)kKmPC * D + 2 * D) * sizeof(int);  // + pair table
};

// pair index k of atoms (a, x), a != x, in np.tril_indices(n, -1) order, and the sign of J[k] at atom a
__device__ __forceinline__ int km_pair(int a, int x) { return a > x ? a * (a - 1) / 2 + x : x * (x - 1) / 2 + a; }

template <int NA>
__global__ void __launch_bounds__(kKmThreads) k_kernel_mat(const __grid_constant__ KernelMatArgs a) {
  using S = KmShape<NA>;
  constexpr int N = S::N, D = S::D, N3 = S::N3, NOUT = S::NOUT, PC = kKmPC, NT = kKmThreads;
  extern __shared__ __align__(16) unsigned char km_smem[];
  int i, j;
  if (a.symmetric) {  // blockIdx.x = i (i + 1) / 2 + j, j <= i
    const long long b = blockIdx.x;
    i = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= b) ++i;
    while ((long long)i * (i + 1) / 2 > b) --i;
    j = (int)(b - (long long)i * (i + 1) / 2);
  } else {
    i = blockIdx.x % a.T, j = blockIdx.x / a.T;
  }
  const int jc = j;  // column block of the output
  if (a.j_list) j = a.j_list[j];
  const int P = a.P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Xi = reinterpret_cast<double*>(km_smem);
  double* Xj = Xi + D;
  double* gi = Xj + D;        // [D][3]
  double* gj = gi + 3 * D;    // [D][3]
  double* diff = gj + 3 * D;  // [PC][D]
  double* fb = diff + PC * D; // [PC] 5 b_p
  double* cc = fb + PC;       // [PC] c_p
  double* kee = cc + PC;      // [PC]
  double* u = kee + 2 * PC;   // [PC][3N]  J_i^T diff_p
  double* fu = u + PC * N3;   // [PC][3N]  5 b_p u_p
  double* v = fu + PC * N3;   // [PC][3N]  (J_j[pi_p])^T diff_p
  double* W = v + PC * N3;    // [D][3N]
  double* ob = W + D * N3;    // [3N][3N + 1] output block staging
  int* pis = reinterpret_cast<int*>(ob + N3 * (N3 + 1));  // [PC][D]
  int* ipis = pis + PC * D;                               // [PC][D]
  int* pa = ipis + PC * D;                                // [D] larger atom of pair k
  int* pb = pa + D;                                       // [D] smaller atom

  for (int k = tid; k < D; k += NT) {
    Xi[k] = a.X[(size_t)i * D + k], Xj[k] = a.X[(size_t)j * D + k];
    int hi = 1;
    while ((hi + 1) * hi / 2 <= k) ++hi;  // k = hi (hi - 1) / 2 + lo, lo < hi  (np.tril_indices order)
    pa[k] = hi, pb[k] = k - hi * (hi - 1) / 2;
  }
  for (int k = tid; k < 3 * D; k += NT) gi[k] = a.G[(size_t)i * 3 * D + k], gj[k] = a.G[(size_t)j * 3 * D + k];
  for (int k = tid; k < D * N3; k += NT) W[k] = 0.0;

  const double sig = a.sig;
  const double sqrt5 = sqrt(5.0);
  double acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = 0.0;
  double erow = 0.0, ecol = 0.0, eacc = 0.0;  // threads < 3N: -sum c_p v_p[tid], +sum c_p u_p[tid]; thread 0: -sum kee_p

  for (int p0 = 0; p0 < P; p0 += PC) {
    const int pc = min(PC, P - p0);
    __syncthreads();  // previous chunk fully consumed (and the loads above visible)
    for (int idx = tid; idx < pc * D; idx += NT) pis[idx] = a.pi[(size_t)p0 * D + idx], ipis[idx] = a.ipi[(size_t)p0 * D + idx];
    __syncthreads();
    // A1: diff_p[k] = X_i[k] - X_j[pi_p(k)]
    for (int idx = tid; idx < pc * D; idx += NT) diff[idx] = Xi[idx % D] - Xj[pis[idx]];
    __syncthreads();
    // A2: norms and radial factors, one warp per permutation
    for (int pl = warp; pl < pc; pl += NT / 32) {
      double s = 0.0;
      for (int k = lane; k < D; k += 32) s = fma(diff[pl * D + k], diff[pl * D + k], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        const double n = sqrt5 * sqrt(s);
        const double ex = exp(-n / sig);
        const double b = ex / (3.0 * sig * sig * sig * sig) * 5.0;
        fb[pl] = 5.0 * b;
        cc[pl] = (sig * sig + sig * n) * b;
        kee[pl] = (1.0 + (n / sig) * (1.0 + n / (3.0 * sig))) * ex;
      }
    }
    __syncthreads();
    // A3: u_p = J_i^T diff_p, v_p = (J_j[pi_p])^T diff_p  -- one thread per (p, atom, axis)
    for (int idx = tid; idx < pc * N3; idx += NT) {
      const int pl = idx / N3, r = idx - pl * N3, at = r / 3, ax = r - at * 3;
      const int* ipl = ipis + pl * D;
      const double* dp = diff + pl * D;
      double su = 0.0, sv = 0.0;
#pragma unroll
      for (int x = 0; x < N; ++x) {
        if (x == at) continue;
        const int k = km_pair(at, x);
        const double sgn = at < x ? 1.0 : -1.0;  // +g at the smaller atom of the pair, -g at the larger
        su = fma(sgn * gi[k * 3 + ax], dp[k], su);
        sv = fma(sgn * gj[k * 3 + ax], dp[ipl[k]], sv);  // row k of J_j sits at position pi_p^-1(k) of J_j[pi_p]
      }
      u[idx] = su, fu[idx] = fb[pl] * su, v[idx] = sv;
    }
    // B1: W[k][c] += c_p J_j[pi_p(k)][c]  -- thread (k, axis) owns its row/axis: no conflicts
    for (int idx = tid; idx < D * 3; idx += NT) {
      const int k = idx / 3, ax = idx - k * 3;
      for (int pl = 0; pl < pc; ++pl) {
        const int k2 = pis[pl * D + k];
        const double val = cc[pl] * gj[k2 * 3 + ax];
        W[k * N3 + pb[k2] * 3 + ax] += val;
        W[k * N3 + pa[k2] * 3 + ax] -= val;
      }
    }
    __syncthreads();
    // B2: rank-1 part, outputs (r, c) strided over the CTA
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      const int e = tid + o * NT;
      if (e < N3 * N3) {
        const int r = e / N3, c = e - r * N3;
        double s = acc[o];
        for (int pl = 0; pl < pc; ++pl) s = fma(fu[pl * N3 + r], v[pl * N3 + c], s);
        acc[o] = s;
      }
    }
    if (a.use_E) {
      if (tid < N3)
        for (int pl = 0; pl < pc; ++pl) erow = fma(-cc[pl], v[pl * N3 + tid], erow), ecol = fma(cc[pl], u[pl * N3 + tid], ecol);
      if (tid == 0)
        for (int pl = 0; pl < pc; ++pl) eacc -= kee[pl];
    }
  }

  // C: K_ij[r][c] = acc - sum_{k containing atom(r)} J_i[k][r] W[k][c], staged so that the block and (symmetric mode)
  // its transpose leave in coalesced rows
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const int e = tid + o * NT;
    if (e < N3 * N3) {
      const int r = e / N3, c = e - r * N3, at = r / 3, ax = r - at * 3;
      double s = acc[o];
#pragma unroll
      for (int x = 0; x < N; ++x) {
        if (x == at) continue;
        const int k = km_pair(at, x);
        const double sgn = at < x ? 1.0 : -1.0;
        s = fma(-sgn * gi[k * 3 + ax], W[k * N3 + c], s);
      }
      ob[r * (N3 + 1) + c] = s;
    }
  }
  __syncthreads();
  double* Kij = a.K + ((size_t)i * N3) * a.ld + (size_t)jc * N3;
  double* Kji = a.K + ((size_t)j * N3) * a.ld + (size_t)i * N3;
  for (int e = tid; e < N3 * N3; e += NT) {
    const int r = e / N3, c = e - r * N3;
    Kij[(size_t)r * a.ld + c] = ob[r * (N3 + 1) + c];
    if (a.symmetric && i != j) Kji[(size_t)r * a.ld + c] = ob[c * (N3 + 1) + r];
  }
  if (a.use_E) {
    const size_t E0 = (size_t)a.T * N3;
    if (tid < N3) {
      a.K[(E0 + i) * a.ld + (size_t)j * N3 + tid] = erow;  // train.py:224-239
      a.K[((size_t)j * N3 + tid) * a.ld + E0 + i] = erow;  // train.py:262-287 (worker j' = T + i, block j)
      if (a.symmetric && i != j) {  // the swapped pair: (J_i[pi_p])^T (X_j - X_i[pi_p]) = -u_q with pi_q = pi_p^-1
        a.K[(E0 + j) * a.ld + (size_t)i * N3 + tid] = ecol;
        a.K[((size_t)i * N3 + tid) * a.ld + E0 + j] = ecol;
      }
    }
    if (tid == 0) {
      a.K[(E0 + j) * a.ld + E0 + i] = eacc;  // train.py:287-289
      if (a.symmetric && i != j) a.K[(E0 + i) * a.ld + E0 + j] = eacc;
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Warp-per-pair form of the kernel-matrix assembly (round 2).  The CTA-per-pair kernel above spends its time in
// ~20 CTA-wide barriers per pair around short, partly narrow phases (ncu: barrier + short-scoreboard stalls,
// shared-memory wavefronts at 94 % of peak, FP64 pipe 17 %).  Here every WARP owns a pair (i, j) of training points
// from start to finish -- only __syncwarp between phases -- with its own slice of shared memory, 8-12 warps per SM
// overlap each other's latencies, and the rank-1 part K += 5 b_p u_p v_p^T runs on a register tile (lane (rg, cg)
// owns RT x CT outputs: RT + CT shared-memory loads per RT CT FMAs).  Same arithmetic per pair as above.
// ---------------------------------------------------------------------------------------------------
constexpr int kKmwPC = 8;         // permutations per chunk: 4 lanes per permutation in the norm phase
constexpr int kKmwMaxWarps = 12;  // warps per CTA (fewer when the per-warp slices do not fit)

template <int NA>
struct KmwShape {
  static constexpr int N = NA, D = NA * (NA - 1) / 2, N3 = 3 * NA;
  static constexpr int RT = (N3 + 3) / 4, CT = (N3 + 7) / 8;  // register tile per lane: 4 row groups x 8 column groups
  static constexpr int NE = (N3 + 31) / 32;                   // energy-constraint entries per lane
  // doubles per warp: Xi, Xj, gi, gj, diff[PC][D], u / fu / v [PC][N3], W[D][N3], (fb, cc, kee, pad)[PC]
  static constexpr int kWarpDoubles = 2 * D + 6 * D + kKmwPC * D + 3 * kKmwPC * N3 + D * N3 + 4 * kKmwPC;
  __host__ __device__ static constexpr bool tables_in_smem(int P) { return D <= 255 && 2 * P * D <= 16 * 1024; }
  __host__ __device__ static constexpr size_t bytes(int nw, int P) {
    return (size_t)nw * kWarpDoubles * 8 + (tables_in_smem(P) ? (size_t)(2 * P * D + 7) / 8 * 8 : 0) + 2 * D * sizeof(int) + 16;
  }
};

template <int NA>
__global__ void __launch_bounds__(kKmwMaxWarps * 32) k_kernel_mat_w(const __grid_constant__ KernelMatArgs a, const long long n_pairs) {
  using S = KmwShape<NA>;
  constexpr int N = S::N, D = S::D, N3 = S::N3, RT = S::RT, CT = S::CT, NE = S::NE, PC = kKmwPC;
  extern __shared__ __align__(16) unsigned char kmw_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int P = a.P;
  const bool tabs = S::tables_in_smem(P);
  double* wbase = reinterpret_cast<double*>(kmw_smem) + (size_t)warp * S::kWarpDoubles;
  unsigned char* tab_s = reinterpret_cast<unsigned char*>(reinterpret_cast<double*>(kmw_smem) + (size_t)nw * S::kWarpDoubles);
  int* pa = reinterpret_cast<int*>(tab_s + (tabs ? (size_t)(2 * P * D + 7) / 8 * 8 : 0));  // [D] larger atom of pair k
  int* pb = pa + D;                                                                          // [D] smaller atom
  double* Xi = wbase;
  double* Xj = Xi + D;
  double* gi = Xj + D;         // [D][3]
  double* gj = gi + 3 * D;     // [D][3]
  double* diff = gj + 3 * D;   // [PC][D]
  double* u = diff + PC * D;   // [PC][3N]  J_i^T diff_p
  double* fu = u + PC * N3;    // [PC][3N]  5 b_p u_p
  double* v = fu + PC * N3;    // [PC][3N]  (J_j[pi_p])^T diff_p
  double* W = v + PC * N3;     // [D][3N]
  double* fb = W + D * N3;     // [PC] 5 b_p
  double* cc = fb + PC;        // [PC] c_p
  double* kee = cc + PC;       // [PC]

  for (int k = tid; k < D; k += blockDim.x) {
    int hi = 1;
    while ((hi + 1) * hi / 2 <= k) ++hi;  // k = hi (hi - 1) / 2 + lo, lo < hi  (np.tril_indices order)
    pa[k] = hi, pb[k] = k - hi * (hi - 1) / 2;
  }
  if (tabs)
    for (int k = tid; k < P * D; k += blockDim.x) tab_s[k] = (unsigned char)a.pi[k], tab_s[P * D + k] = (unsigned char)a.ipi[k];
  __syncthreads();
  auto pi_at = [&](const int p, const int k) -> int { return tabs ? (int)tab_s[p * D + k] : a.pi[(size_t)p * D + k]; };
  auto ipi_at = [&](const int p, const int k) -> int { return tabs ? (int)tab_s[P * D + p * D + k] : a.ipi[(size_t)p * D + k]; };

  const double sig = a.sig, sqrt5 = sqrt(5.0);
  const int rg = lane >> 3, cg = lane & 7;  // this lane's tile: rows rg RT + [0, RT), columns cg CT + [0, CT)
  const size_t E0 = (size_t)a.T * N3;

  for (long long b = (long long)blockIdx.x * nw + warp; b < n_pairs; b += (long long)gridDim.x * nw) {
    int i, j;
    if (a.symmetric) {  // b = i (i + 1) / 2 + j, j <= i
      i = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
      while ((long long)(i + 1) * (i + 2) / 2 <= b) ++i;
      while ((long long)i * (i + 1) / 2 > b) --i;
      j = (int)(b - (long long)i * (i + 1) / 2);
    } else {
      i = (int)(b % a.T), j = (int)(b / a.T);
    }
    const int jc = j;  // column block of the output
    if (a.j_list) j = a.j_list[j];
    __syncwarp();  // the previous pair's reads of this warp's slice are done
    for (int k = lane; k < D; k += 32) Xi[k] = a.X[(size_t)i * D + k], Xj[k] = a.X[(size_t)j * D + k];
    for (int k = lane; k < 3 * D; k += 32) gi[k] = a.G[(size_t)i * 3 * D + k], gj[k] = a.G[(size_t)j * 3 * D + k];
    for (int k = lane; k < D * N3; k += 32) W[k] = 0.0;
    double acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[r][c] = 0.0;
    double erow[NE], ecol[NE], eacc = 0.0;  // -sum c_p v_p[e], +sum c_p u_p[e] for e = lane + 32 q; lane 0: -sum kee_p
#pragma unroll
    for (int q = 0; q < NE; ++q) erow[q] = 0.0, ecol[q] = 0.0;

    for (int p0 = 0; p0 < P; p0 += PC) {
      const int pc = min(PC, P - p0);
      __syncwarp();
      // A1: diff_p[k] = X_i[k] - X_j[pi_p(k)]
      for (int idx = lane; idx < pc * D; idx += 32) {
        const int pl = idx / D, k = idx - pl * D;
        diff[idx] = Xi[k] - Xj[pi_at(p0 + pl, k)];
      }
      __syncwarp();
      // A2: norms and radial factors, four lanes per permutation
      {
        const int pl = lane >> 2, part = lane & 3;
        double s = 0.0;
        if (pl < pc)
          for (int k = part; k < D; k += 4) s = fma(diff[pl * D + k], diff[pl * D + k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0 && pl < pc) {
          const double n = sqrt5 * sqrt(s);
          const double ex = exp(-n / sig);
          const double bq = ex / (3.0 * sig * sig * sig * sig) * 5.0;
          fb[pl] = 5.0 * bq;
          cc[pl] = (sig * sig + sig * n) * bq;
          kee[pl] = (1.0 + (n / sig) * (1.0 + n / (3.0 * sig))) * ex;
        }
      }
      __syncwarp();
      // A3: u_p = J_i^T diff_p, v_p = (J_j[pi_p])^T diff_p  -- one lane per (p, atom, axis)
      for (int idx = lane; idx < pc * N3; idx += 32) {
        const int pl = idx / N3, r = idx - pl * N3, at = r / 3, ax = r - at * 3;
        const double* dp = diff + pl * D;
        double su = 0.0, sv = 0.0;
#pragma unroll
        for (int x = 0; x < N; ++x) {
          if (x == at) continue;
          const int k = km_pair(at, x);
          const double sgn = at < x ? 1.0 : -1.0;  // +g at the smaller atom of the pair, -g at the larger
          su = fma(sgn * gi[k * 3 + ax], dp[k], su);
          sv = fma(sgn * gj[k * 3 + ax], dp[ipi_at(p0 + pl, k)], sv);  // row k of J_j sits at position pi_p^-1(k) of J_j[pi_p]
        }
        u[idx] = su, fu[idx] = fb[pl] * su, v[idx] = sv;
      }
      // B1: W[k][c] += c_p J_j[pi_p(k)][c]  -- lane (k, axis) owns its row/axis: no conflicts
      for (int idx = lane; idx < D * 3; idx += 32) {
        const int k = idx / 3, ax = idx - k * 3;
        for (int pl = 0; pl < pc; ++pl) {
          const int k2 = pi_at(p0 + pl, k);
          const double val = cc[pl] * gj[k2 * 3 + ax];
          W[k * N3 + pb[k2] * 3 + ax] += val;
          W[k * N3 + pa[k2] * 3 + ax] -= val;
        }
      }
      __syncwarp();
      // B2: rank-1 part on the register tile
      for (int pl = 0; pl < pc; ++pl) {
        double fr[RT], vc[CT];
#pragma unroll
        for (int r = 0; r < RT; ++r) fr[r] = (rg * RT + r < N3) ? fu[pl * N3 + rg * RT + r] : 0.0;
#pragma unroll
        for (int c = 0; c < CT; ++c) vc[c] = (cg * CT + c < N3) ? v[pl * N3 + cg * CT + c] : 0.0;
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
          for (int c = 0; c < CT; ++c) acc[r][c] = fma(fr[r], vc[c], acc[r][c]);
      }
      if (a.use_E) {
#pragma unroll
        for (int q = 0; q < NE; ++q) {
          const int e = lane + 32 * q;
          if (e < N3)
            for (int pl = 0; pl < pc; ++pl) erow[q] = fma(-cc[pl], v[pl * N3 + e], erow[q]), ecol[q] = fma(cc[pl], u[pl * N3 + e], ecol[q]);
        }
        if (lane == 0)
          for (int pl = 0; pl < pc; ++pl) eacc -= kee[pl];
      }
    }
    __syncwarp();
    // C: K_ij[r][c] = acc - sum_{k containing atom(r)} J_i[k][r] W[k][c]; the block and (symmetric mode) its transpose
    double* Kij = a.K + ((size_t)i * N3) * a.ld + (size_t)jc * N3;
    double* Kji = a.K + ((size_t)j * N3) * a.ld + (size_t)i * N3;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int row = rg * RT + r;
      if (row >= N3) continue;
      const int at = row / 3, ax = row - at * 3;
      double s[CT];
#pragma unroll
      for (int c = 0; c < CT; ++c) s[c] = acc[r][c];
#pragma unroll
      for (int x = 0; x < N; ++x) {
        if (x == at) continue;
        const int k = km_pair(at, x);
        const double gq = (at < x ? -1.0 : 1.0) * gi[k * 3 + ax];
#pragma unroll
        for (int c = 0; c < CT; ++c)
          if (cg * CT + c < N3) s[c] = fma(gq, W[k * N3 + cg * CT + c], s[c]);
      }
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const int col = cg * CT + c;
        if (col < N3) {
          Kij[(size_t)row * a.ld + col] = s[c];
          if (a.symmetric && i != j) Kji[(size_t)col * a.ld + row] = s[c];
        }
      }
    }
    if (a.use_E) {
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        const int e = lane + 32 * q;
        if (e < N3) {
          a.K[(E0 + i) * a.ld + (size_t)j * N3 + e] = erow[q];  // train.py:224-239
          a.K[((size_t)j * N3 + e) * a.ld + E0 + i] = erow[q];  // train.py:262-287 (worker j' = T + i, block j)
          if (a.symmetric && i != j) {  // the swapped pair: (J_i[pi_p])^T (X_j - X_i[pi_p]) = -u_q with pi_q = pi_p^-1
            a.K[(E0 + j) * a.ld + (size_t)i * N3 + e] = ecol[q];
            a.K[((size_t)i * N3 + e) * a.ld + E0 + j] = ecol[q];
          }
        }
      }
      if (lane == 0) {
        a.K[(E0 + j) * a.ld + E0 + i] = eacc;  // train.py:287-289
        if (a.symmetric && i != j) a.K[(E0 + i) * a.ld + E0 + j] = eacc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Tensor-pipe form of the kernel-matrix assembly (round 2): a warp per pair, every D-wide or P-wide sum a chain of
// DMMA m8n8k4 (the FP64 tensor pipe), no scatter.  With the DENSE Jacobians J_i, J_j (D x 3N, six non-zeros per
// row, never stored: an operand fragment is a compare + select on the compressed G) and per 8 permutations (one
// m-tile):
//     U = Dif J_i            Dif[p][k] = X_i[k] - X_j[pi_p(k)]                    [8 x D] [D x 3N]
//     V = Dif~ J_j           Dif~[p][k] = X_i[pi_p^-1(k)] - X_j[k]                (= (J_j[pi_p])^T diff_p, re-indexed)
//     K += (5 b U)^T V                                                           [3N x 8] [8 x 3N]
// and once per pair, with C[k][k'] = sum_{p: pi_p(k) = k'} c_p (D x D, built by row-owning lanes):
//     W = C J_j              [D x D] [D x 3N]        K -= J_i^T W                [3N x D] [D x 3N]
// Two passes over the permutations so that the accumulators of W and of K are never live together: pass 1 norms ->
// b_p, c_p -> C -> W (to shared memory, over C), pass 2 U, V, rank-8 updates of K, then K -= J_i^T W.
// Shared-memory strides are = 4 (mod 16) doubles: the (k = 4 kk + l, column 8 n + g) operand pattern of m8n8k4
// then tiles the 32 banks exactly.  Same arithmetic as k_kernel_mat up to summation order.
// ---------------------------------------------------------------------------------------------------
constexpr int kKmdMaxWarps = 12;
__host__ __device__ constexpr int kmd_pad4(int n) { return (n - 4 + 15) / 16 * 16 + 4; }  // smallest m >= n, m % 16 == 4
__host__ __device__ constexpr int kmd_max(int a, int b) { return a > b ? a : b; }

template <int NA>
struct KmdShape {
  static constexpr int N = NA, D = NA * (NA - 1) / 2, N3 = 3 * NA;
  static constexpr int KS = (D + 3) / 4, NTc = (N3 + 7) / 8, MTk = (D + 7) / 8;
  static constexpr int CS = kmd_pad4(4 * KS);   // row stride of C (columns k' < 4 KS)
  static constexpr int WS = kmd_pad4(8 * NTc);  // row stride of W, U, V (columns < 8 NTc)
  static constexpr int OS = N3 | 1;             // output staging: odd stride, rows and columns both conflict-free
  static constexpr int kRegion = (kmd_max(kmd_max(4 * KS * CS, 4 * KS * WS), N3 * OS) + 1) & ~1;  // rows k < 4 KS only; even: 16-byte alignment of what follows
  __host__ __device__ static constexpr int p_pad(int P) { return (P + 7) / 8 * 8; }
  // bit kk NTc + n: the 4 x 8 fragment (descriptors 4 kk .. 4 kk + 3, columns 8 n .. 8 n + 7) of a dense Jacobian has a
  // non-zero (row k <-> atoms hi > lo touches columns 3 hi .. 3 hi + 2 and 3 lo .. 3 lo + 2); trimers: 24 of 36
  __host__ __device__ static constexpr unsigned long long nz_mask() {
    unsigned long long m = 0;
    for (int kk = 0; kk < KS; ++kk)
      for (int n = 0; n < NTc; ++n) {
        bool hit = false;
        for (int k = 4 * kk; k < 4 * kk + 4 && k < D; ++k) {
          int hi = 1;
          while ((hi + 1) * hi / 2 <= k) ++hi;
          const int lo = k - hi * (hi - 1) / 2;
          for (int c = 8 * n; c < 8 * n + 8 && c < N3; ++c) hit = hit || c / 3 == hi || c / 3 == lo;
        }
        if (hit) m |= 1ull << (kk * NTc + n);
      }
    return m;
  }
  static constexpr unsigned long long kNz = nz_mask();
  static_assert(KS * NTc <= 64, "fragment mask");
  __host__ __device__ static constexpr bool nz(int kk, int n) { return (kNz >> (kk * NTc + n)) & 1ull; }
  // doubles per warp: Xi, Xj [4 KS]; gi, gj [D][3]; region (C, then W, then the output block); U, V [8][WS]; 5 b_p, c_p [P]
  __host__ __device__ static constexpr int warp_doubles(int P) { return 8 * KS + 6 * D + kRegion + 16 * WS + 2 * p_pad(P); }
  // the four byte tables (pi, pi^-1, srt_p, srt_k) live in shared memory: P <= 255 and 4 P D bytes
  __host__ __device__ static constexpr bool supported(int P) { return P <= 255 && 4 * P * D <= 24 * 1024; }
  __host__ __device__ static constexpr size_t bytes(int nw, int P) {
    return (size_t)nw * warp_doubles(P) * 8 + (size_t)(4 * P * D + 7) / 8 * 8 + 16;
  }
};

__device__ __forceinline__ void kmd_mma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NA>
__global__ void __launch_bounds__(kKmdMaxWarps * 32, 1) k_kernel_mat_d(const __grid_constant__ KernelMatArgs a, const long long n_pairs) {
  using S = KmdShape<NA>;
  constexpr int D = S::D, N3 = S::N3, KS = S::KS, NTc = S::NTc, MTk = S::MTk, CS = S::CS, WS = S::WS, OS = S::OS;
  static_assert(N3 <= 32, "one lane per energy-constraint entry");
  extern __shared__ __align__(16) unsigned char kmd_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int g = lane >> 2, l = lane & 3;
  const int P = a.P, Pp = S::p_pad(P);
  double* wbase = reinterpret_cast<double*>(kmd_smem) + (size_t)warp * S::warp_doubles(P);
  unsigned char* tab_s = reinterpret_cast<unsigned char*>(reinterpret_cast<double*>(kmd_smem) + (size_t)nw * S::warp_doubles(P));
  double* Xi = wbase;           // [4 KS], zero beyond D
  double* Xj = Xi + 4 * KS;     // [4 KS]
  double* gi = Xj + 4 * KS;     // [D][3]
  double* gj = gi + 3 * D;      // [D][3]
  double* reg = gj + 3 * D;     // C [4 KS][CS] -> W [4 KS][WS] -> output block [3N][OS]
  double* Uc = reg + S::kRegion;  // [8][WS]  u_p of the current 8 permutations
  double* Vc = Uc + 8 * WS;       // [8][WS]
  double* fbs = Vc + 8 * WS;      // [Pp] |diff_p|^2, then 5 b_p (0 beyond P)
  double* ccs = fbs + Pp;         // [Pp] c_p

  for (int k = tid; k < P * D; k += blockDim.x) {
    tab_s[k] = (unsigned char)a.pi[k], tab_s[P * D + k] = (unsigned char)a.ipi[k];
    tab_s[2 * P * D + k] = (unsigned char)a.srt_p[k], tab_s[3 * P * D + k] = (unsigned char)a.srt_k[k];
  }
  __syncthreads();
  const unsigned char* pi_s = tab_s;
  const unsigned char* ipi_s = tab_s + P * D;
  const unsigned char* sp_s = tab_s + 2 * P * D;
  const unsigned char* sk_s = tab_s + 3 * P * D;

  // Per-lane constants of the operand pattern J[k = 4 kk + l][column 8 n + g]: per fragment f = kk NTc + n one bit
  // "non-zero" (the column's atom is one of the two atoms of pair k) and one bit "negative" (it is the larger one:
  // -G there, +G at the smaller), and the offset 3 l + axis of the column in a row of G.  The sign is applied to
  // the high word with integer logic: a DADD per fragment on the pipe the DMMAs need costs more than it looks.
  unsigned long long nzw = 0ull, sgw = 0ull;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const int k = 4 * kk + l;
    int hi = 1;
    while ((hi + 1) * hi / 2 <= k) ++hi;
    const int lo = k - hi * (hi - 1) / 2;
#pragma unroll
    for (int n = 0; n < NTc; ++n) {
      const int c = 8 * n + g, at = c / 3;
      if (k < D && c < N3 && S::nz(kk, n)) {
        if (at == lo || at == hi) nzw |= 1ull << (kk * NTc + n);
        if (at == hi) sgw |= 1ull << (kk * NTc + n);
      }
    }
  }
  int goff[NTc];
#pragma unroll
  for (int n = 0; n < NTc; ++n) goff[n] = 3 * l + (8 * n + g) % 3;
  auto jfrag = [&](const double* G3, const int kk, const int n) -> double {  // kk, n: unrolled constants
    const int f = kk * NTc + n;
    double v = 0.0;
    if ((nzw >> f) & 1ull) v = G3[12 * kk + goff[n]];
    const int hw = __double2hiint(v) ^ (int)(((unsigned)(sgw >> f) & 1u) << 31);
    return __hiloint2double(hw, __double2loint(v));
  };

  const double sig = a.sig, sqrt5 = sqrt(5.0);
  const double inv_sig = 1.0 / sig, inv_3sig = 1.0 / (3.0 * sig), b_scale = 5.0 / (3.0 * sig * sig * sig * sig);
  const size_t E0 = (size_t)a.T * N3;

  // pairs strided over the warps: the warps of a CTA work on neighbouring blocks of K at any time (measured: a
  // contiguous range per warp, which would keep X_i, G_i in shared memory, is 3 % slower: 1776 separate write streams)
  for (long long b = (long long)blockIdx.x * nw + warp; b < n_pairs; b += (long long)gridDim.x * nw) {
    int i, j;
    if (a.symmetric) {  // b = i (i + 1) / 2 + j, j <= i
      i = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
      while ((long long)(i + 1) * (i + 2) / 2 <= b) ++i;
      while ((long long)i * (i + 1) / 2 > b) --i;
      j = (int)(b - (long long)i * (i + 1) / 2);
    } else {
      i = (int)(b % a.T), j = (int)(b / a.T);
    }
    const int jc = j;  // column block of the output
    if (a.j_list) j = a.j_list[j];
    __syncwarp();  // the previous pair's reads of this warp's slice are done
    for (int k = lane; k < 4 * KS; k += 32) Xi[k] = k < D ? a.X[(size_t)i * D + k] : 0.0, Xj[k] = k < D ? a.X[(size_t)j * D + k] : 0.0;
    for (int k = lane; k < 3 * D; k += 32) gi[k] = a.G[(size_t)i * 3 * D + k], gj[k] = a.G[(size_t)j * 3 * D + k];
    for (int k = lane; k < 4 * KS * CS; k += 32) reg[k] = 0.0;
    __syncwarp();

    // ---- pass 1: |diff_p|^2 (four lanes per permutation), then 5 b_p, c_p and the energy-energy term per lane -----
    for (int p0 = 0; p0 < Pp; p0 += 8) {
      const unsigned char* pr = pi_s + min(p0 + g, P - 1) * D;
      double s2 = 0.0;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int k = 4 * kk + l;
        const double d = k < D ? Xi[k] - Xj[pr[k]] : 0.0;
        s2 = fma(d, d, s2);
      }
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
      if (l == 0) fbs[p0 + g] = s2;
    }
    __syncwarp();
    double eacc = 0.0;
    for (int p = lane; p < Pp; p += 32) {
      const double n = sqrt5 * sqrt(fbs[p]);
      const double ex = exp(-n * inv_sig);
      const double bq = ex * b_scale;  // 5 / (3 sig^4)
      const bool ok = p < P;
      fbs[p] = ok ? 5.0 * bq : 0.0;
      ccs[p] = ok ? (sig * sig + sig * n) * bq : 0.0;
      if (ok) eacc -= (1.0 + (n * inv_sig) * (1.0 + n * inv_3sig)) * ex;
    }
    __syncwarp();
    // C[k][k'] = sum_{p: pi_p(k) = k'} c_p: the lane owning row k walks its permutations in the order of their images
    // (host-sorted), sums each run in a register and stores it once -- no read-modify-write on shared memory
    if (lane < D) {
      const int k = lane;
      double* row = reg + k * CS;
      double acc = 0.0;
      int prev = sk_s[k];
      for (int t = 0; t < P; ++t) {
        const int k2 = sk_s[t * D + k];
        const double c = ccs[sp_s[t * D + k]];
        if (k2 != prev) row[prev] = acc, acc = 0.0, prev = k2;
        acc += c;
      }
      row[prev] = acc;
    }
    if (D > 32) {  // rows 32 .. D - 1: LPR lanes per row, each a slice of the sorted list; runs cut by a slice meet in an atomic
      constexpr int XR = D > 32 ? D - 32 : 1;
      constexpr int LPR = XR <= 1 ? 32 : XR <= 2 ? 16 : XR <= 4 ? 8 : XR <= 8 ? 4 : XR <= 16 ? 2 : 1;
      const int k = 32 + lane / LPR, q = lane % LPR, len = (P + LPR - 1) / LPR;
      const int t0 = q * len, t1 = min(P, t0 + len);
      if (k < D && t0 < t1) {
        double* row = reg + k * CS;
        double acc = 0.0;
        int prev = sk_s[t0 * D + k];
        for (int t = t0; t < t1; ++t) {
          const int k2 = sk_s[t * D + k];
          const double c = ccs[sp_s[t * D + k]];
          if (k2 != prev) atomicAdd(row + prev, acc), acc = 0.0, prev = k2;
          acc += c;
        }
        atomicAdd(row + prev, acc);
      }
    }
    __syncwarp();
    {  // W = C J_j (all of it in registers before the first store: W overwrites C)
      double Wacc[MTk][NTc][2];
#pragma unroll
      for (int m = 0; m < MTk; ++m)
#pragma unroll
        for (int n = 0; n < NTc; ++n) Wacc[m][n][0] = 0.0, Wacc[m][n][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        double bj[NTc];
#pragma unroll
        for (int n = 0; n < NTc; ++n) bj[n] = S::nz(kk, n) ? jfrag(gj, kk, n) : 0.0;
#pragma unroll
        for (int m = 0; m < MTk; ++m) {
          const double av = (8 * m + g < 4 * KS) ? reg[(8 * m + g) * CS + 4 * kk + l] : 0.0;
#pragma unroll
          for (int n = 0; n < NTc; ++n)
            if (S::nz(kk, n)) kmd_mma(Wacc[m][n][0], Wacc[m][n][1], av, bj[n]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < MTk; ++m)
#pragma unroll
        for (int n = 0; n < NTc; ++n)
          if (8 * m + g < 4 * KS) *reinterpret_cast<double2*>(reg + (8 * m + g) * WS + 8 * n + 2 * l) = make_double2(Wacc[m][n][0], Wacc[m][n][1]);
    }
    __syncwarp();

    // ---- pass 2: U, V of 8 permutations, rank-8 update of K -------------------------------------------------------
    double Kacc[NTc][NTc][2];
#pragma unroll
    for (int m = 0; m < NTc; ++m)
#pragma unroll
      for (int n = 0; n < NTc; ++n) Kacc[m][n][0] = 0.0, Kacc[m][n][1] = 0.0;
    double erow = 0.0, ecol = 0.0;  // lane e < 3N: -sum_p c_p v_p[e], +sum_p c_p u_p[e]
    for (int p0 = 0; p0 < Pp; p0 += 8) {
      const unsigned char* pr = pi_s + min(p0 + g, P - 1) * D;
      const unsigned char* ipr = ipi_s + min(p0 + g, P - 1) * D;
      double Ua[NTc][2], Va[NTc][2];
#pragma unroll
      for (int n = 0; n < NTc; ++n) Ua[n][0] = 0.0, Ua[n][1] = 0.0, Va[n][0] = 0.0, Va[n][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int k = 4 * kk + l;
        const bool kv = k < D;
        const double au = kv ? Xi[k] - Xj[pr[k]] : 0.0;
        const double av = kv ? Xi[ipr[k]] - Xj[k] : 0.0;
#pragma unroll
        for (int n = 0; n < NTc; ++n)
          if (S::nz(kk, n)) {
            kmd_mma(Ua[n][0], Ua[n][1], au, jfrag(gi, kk, n));
            kmd_mma(Va[n][0], Va[n][1], av, jfrag(gj, kk, n));
          }
      }
      __syncwarp();  // the previous chunk's U, V are consumed
#pragma unroll
      for (int n = 0; n < NTc; ++n) {
        *reinterpret_cast<double2*>(Uc + g * WS + 8 * n + 2 * l) = make_double2(Ua[n][0], Ua[n][1]);
        *reinterpret_cast<double2*>(Vc + g * WS + 8 * n + 2 * l) = make_double2(Va[n][0], Va[n][1]);
      }
      __syncwarp();
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int pl = 4 * k2 + l;
        const double f = fbs[p0 + pl];
        double ar[NTc], br[NTc];
#pragma unroll
        for (int m = 0; m < NTc; ++m) ar[m] = f * Uc[pl * WS + 8 * m + g], br[m] = Vc[pl * WS + 8 * m + g];
#pragma unroll
        for (int m = 0; m < NTc; ++m)
#pragma unroll
          for (int n = 0; n < NTc; ++n) kmd_mma(Kacc[m][n][0], Kacc[m][n][1], ar[m], br[n]);
      }
      if (a.use_E && lane < N3) {
#pragma unroll
        for (int q = 0; q < 8; ++q) erow = fma(-ccs[p0 + q], Vc[q * WS + lane], erow), ecol = fma(ccs[p0 + q], Uc[q * WS + lane], ecol);
      }
    }
    // K -= J_i^T W
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      double ai[NTc], bw[NTc];
#pragma unroll
      for (int m = 0; m < NTc; ++m) ai[m] = S::nz(kk, m) ? -jfrag(gi, kk, m) : 0.0, bw[m] = reg[(4 * kk + l) * WS + 8 * m + g];
#pragma unroll
      for (int m = 0; m < NTc; ++m)
#pragma unroll
        for (int n = 0; n < NTc; ++n)
          if (S::nz(kk, m)) kmd_mma(Kacc[m][n][0], Kacc[m][n][1], ai[m], bw[n]);
    }
    __syncwarp();  // W is consumed: the region becomes the output block
#pragma unroll
    for (int m = 0; m < NTc; ++m)
#pragma unroll
      for (int n = 0; n < NTc; ++n) {
        const int r = 8 * m + g, c = 8 * n + 2 * l;
        if (r < N3 && c < N3) reg[r * OS + c] = Kacc[m][n][0];
        if (r < N3 && c + 1 < N3) reg[r * OS + c + 1] = Kacc[m][n][1];
      }
    __syncwarp();
    double* Kij = a.K + ((size_t)i * N3) * a.ld + (size_t)jc * N3;
    double* Kji = a.K + ((size_t)j * N3) * a.ld + (size_t)i * N3;
    for (int e = lane; e < N3 * N3; e += 32) {
      const int r = e / N3, c = e - r * N3;
      Kij[(size_t)r * a.ld + c] = reg[r * OS + c];
      if (a.symmetric && i != j) Kji[(size_t)r * a.ld + c] = reg[c * OS + r];
    }
    if (a.use_E) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) eacc += __shfl_xor_sync(0xffffffffu, eacc, o);
      if (lane < N3) {
        a.K[(E0 + i) * a.ld + (size_t)j * N3 + lane] = erow;  // train.py:224-239
        a.K[((size_t)j * N3 + lane) * a.ld + E0 + i] = erow;  // train.py:262-287
        if (a.symmetric && i != j) {
          a.K[(E0 + j) * a.ld + (size_t)i * N3 + lane] = ecol;
          a.K[((size_t)i * N3 + lane) * a.ld + E0 + j] = ecol;
        }
      }
      if (lane == 0) {
        a.K[(E0 + j) * a.ld + E0 + i] = eacc;  // train.py:287-289
        if (a.symmetric && i != j) a.K[(E0 + i) * a.ld + E0 + j] = eacc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Matern-5/2 covariance of query structures vs. the permuted training rows (analysis/models.py:31-139).
// out[m][t P + p] = (1 + s d + 5 d^2 / (3 sig^2)) exp(-s d),  d = |x_m - X_t[pi_p]| = |x_m[pi_p^-1] - X_t|.
// One CTA per (query m, tile of kM52Rows training rows); thread <-> output (t, p), p fastest (coalesced).
// ---------------------------------------------------------------------------------------------------
constexpr int kM52Rows = 32;
constexpr int kM52Threads = 256;

struct Mat52Args {
  int n_atoms, D, T, P;
  double sig;
  const double* Q;        // queries: coordinates [n_q][n_atoms][3] (q_is_desc = 0) or descriptors [n_q][D]
  int q_is_desc;
  const double* X;        // training descriptors; X_t[e] at X[t * x_row_stride + e * x_elem_stride]
  long long x_row_stride;
  int x_elem_stride;
  const int* iperm;       // [P][D] pi_p^-1, or NULL for the identity (P must be 1)
  int stage_perms;        // 1: the P permuted copies of the query fit shared memory (mat52_smem_bytes)
  int rows_per_tile;      // training rows per CTA (>= kM52Rows; more for narrow descriptors / few permutations, so that a
                          // CTA has ~2000 outputs per query whatever the fragment size)
  double* out;            // [n_q][T * P]
};

// bytes of shared memory: query descriptor, row tile, and (when it fits) the P permuted copies of the query
__host__ __device__ inline size_t mat52_smem_bytes(int D, int P, bool stage_perms, int rows_per_tile = kM52Rows) {
  return ((size_t)D + (size_t)rows_per_tile * (D + 1) + (stage_perms ? (size_t)P * (D | 1) : 0)) * sizeof(double);
}
// rows per CTA: ~2048 outputs per query and CTA, the row tile at most 48 KB
__host__ inline int mat52_rows_per_tile(int D, int P, int T) {
  int rows = (2048 + P - 1) / P;
  rows = rows < kM52Rows ? kM52Rows : rows;
  const int cap = (48 * 1024) / ((D + 1) * 8);
  rows = rows > cap ? cap : rows;
  rows = rows < kM52Rows ? kM52Rows : rows;
  return rows > T ? (T > kM52Rows ? T : kM52Rows) : rows;
}

__global__ void __launch_bounds__(kM52Threads) k_mat52(const __grid_constant__ Mat52Args a) {
  extern __shared__ __align__(16) unsigned char m52_smem[];
  const int D = a.D, P = a.P, N = a.n_atoms;
  const int RPT = a.rows_per_tile;
  const int m = blockIdx.y, t0 = blockIdx.x * RPT;
  const int nrow = min(RPT, a.T - t0);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int DQ = D | 1;                              // odd pitch: lanes with consecutive p hit different banks
  double* xq = reinterpret_cast<double*>(m52_smem);  // [D]
  double* Xs = xq + D;                               // [RPT][D + 1]
  double* yq = Xs + RPT * (D + 1);                   // [P][DQ] permuted queries y_p[e] = x[pi_p^-1(e)] (stage_perms)
  for (int k = tid; k < D; k += nt) {
    if (a.q_is_desc) {
      xq[k] = a.Q[(size_t)m * D + k];
    } else {  // desc.py:79-158: x_k = 1 / |r_hi - r_lo|, un-fused in the reference's order
      int hi = 1;
      while ((hi + 1) * hi / 2 <= k) ++hi;
      const int lo = k - hi * (hi - 1) / 2;
      const double* r = a.Q + (size_t)m * N * 3;
      const double dx = __dsub_rn(r[hi * 3 + 0], r[lo * 3 + 0]), dy = __dsub_rn(r[hi * 3 + 1], r[lo * 3 + 1]),
                   dz = __dsub_rn(r[hi * 3 + 2], r[lo * 3 + 2]);
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      xq[k] = __ddiv_rn(1.0, __dsqrt_rn(d2));
    }
  }
  for (int idx = tid; idx < nrow * D; idx += nt) {
    const int tl = idx / D, e = idx - tl * D;
    Xs[tl * (D + 1) + e] = a.X[(size_t)(t0 + tl) * a.x_row_stride + (size_t)e * a.x_elem_stride];
  }
  __syncthreads();
  if (a.stage_perms) {
    for (int idx = tid; idx < P * D; idx += nt) {
      const int p = idx / D, e = idx - p * D;
      yq[p * DQ + e] = xq[a.iperm ? a.iperm[idx] : e];
    }
    __syncthreads();
  }
  const double sig = a.sig;
  const double s5 = sqrt(5.0) / sig, q2 = 5.0 / (3.0 * sig * sig);
  // thread <-> (4 consecutive training rows, permutation p): the permuted query element is read once for 4 outputs
  // (the kernel is bound by shared-memory wavefronts: ncu r2, 88 % of peak with one row per thread)
  constexpr int RB = 4;
  const int ngrp = (nrow + RB - 1) / RB;
  for (int idx = tid; idx < ngrp * P; idx += nt) {
    const int tg = idx / P, p = idx - tg * P;
    const int tl0 = tg * RB;
    const double* xr[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) xr[r] = Xs + min(tl0 + r, nrow - 1) * (D + 1);
    double s[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) s[r] = 0.0;
    if (a.stage_perms) {
      const double* yp = yq + p * DQ;
      for (int e = 0; e < D; ++e) {
        const double y = yp[e];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const double d = y - xr[r][e];
          s[r] = fma(d, d, s[r]);
        }
      }
    } else {
      const int* ip = a.iperm ? a.iperm + (size_t)p * D : nullptr;
      for (int e = 0; e < D; ++e) {
        const double y = xq[ip ? ip[e] : e];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const double d = y - xr[r][e];
          s[r] = fma(d, d, s[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (tl0 + r >= nrow) break;
      const double d = sqrt(s[r]);
      a.out[(size_t)m * a.T * P + (size_t)(t0 + tl0 + r) * P + p] = (1.0 + s5 * d + q2 * s[r]) * exp(-s5 * d);
    }
  }
}

}  // namespace mbg
