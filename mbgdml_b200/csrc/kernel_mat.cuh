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
// Permutations are processed in chunks of kKmPC so that shared memory does not grow with P.
#pragma once
#include "common.cuh"

namespace mbg {

constexpr int kKmPC = 16;        // permutations per chunk
constexpr int kKmThreads = 256;  // threads per CTA
constexpr int kKmMaxOut = 12;    // outputs per thread: ceil((3 * 18)^2 / 256)

struct KernelMatArgs {
  int n_atoms, D, T, P;
  double sig;
  int use_E;
  const double* X;   // [T][D]      R_desc
  const double* G;   // [T][D][3]   compressed Jacobian (desc.py:448-480)
  const int* pi;     // [P][D]      pi_p(k)
  const int* ipi;    // [P][D]      pi_p^-1(k)
  double* K;         // [rows][rows], rows = T 3N (+ T)
  long long ld;      // rows
};

__host__ __device__ inline size_t kernel_mat_smem_bytes(int n_atoms) {
  const int D = n_atoms * (n_atoms - 1) / 2, N3 = 3 * n_atoms;
  size_t doubles = 2 * (size_t)D            // Xi, Xj
                   + 2 * (size_t)D * 3      // gi, gj
                   + (size_t)kKmPC * D      // diff
                   + 4 * (size_t)kKmPC      // fb, c, kee, pad
                   + 2 * (size_t)kKmPC * N3 // u, v
                   + (size_t)D * N3;        // W
  return doubles * 8 + (size_t)D * 2 * sizeof(int);  // + pair table
}

template <int MAXOUT>
__global__ void __launch_bounds__(kKmThreads) k_kernel_mat(const KernelMatArgs a) {
  extern __shared__ __align__(16) unsigned char km_smem[];
  const int N = a.n_atoms, D = a.D, P = a.P, N3 = 3 * N;
  const int i = blockIdx.x, j = blockIdx.y;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  double* Xi = reinterpret_cast<double*>(km_smem);
  double* Xj = Xi + D;
  double* gi = Xj + D;            // [D][3]
  double* gj = gi + 3 * D;        // [D][3]
  double* diff = gj + 3 * D;      // [PC][D]
  double* fb = diff + kKmPC * D;  // [PC] 5 b_p
  double* cc = fb + kKmPC;        // [PC] c_p
  double* kee = cc + kKmPC;       // [PC]
  double* u = kee + 2 * kKmPC;    // [PC][3N]
  double* v = u + kKmPC * N3;     // [PC][3N]
  double* W = v + kKmPC * N3;     // [D][3N]
  int* pa = reinterpret_cast<int*>(W + (size_t)D * N3);  // [D] larger atom of pair k
  int* pb = pa + D;                                      // [D] smaller atom

  for (int k = tid; k < D; k += nt) {
    Xi[k] = a.X[(size_t)i * D + k];
    Xj[k] = a.X[(size_t)j * D + k];
    int hi = 1;
    while ((hi + 1) * hi / 2 <= k) ++hi;  // k = hi (hi - 1) / 2 + lo, lo < hi  (np.tril_indices order)
    pa[k] = hi, pb[k] = k - hi * (hi - 1) / 2;
  }
  for (int k = tid; k < 3 * D; k += nt) {
    gi[k] = a.G[(size_t)i * 3 * D + k];
    gj[k] = a.G[(size_t)j * 3 * D + k];
  }
  for (int k = tid; k < D * N3; k += nt) W[k] = 0.0;
  __syncthreads();

  const double sig = a.sig;
  const double sqrt5 = sqrt(5.0);
  double acc[MAXOUT];
#pragma unroll
  for (int o = 0; o < MAXOUT; ++o) acc[o] = 0.0;
  double erow = 0.0, eacc = 0.0;  // threads < 3N: -sum_p c_p v_p[tid]; thread 0: -sum_p kee_p

  for (int p0 = 0; p0 < P; p0 += kKmPC) {
    const int pc = min(kKmPC, P - p0);
    // A1: diff_p[k] = X_i[k] - X_j[pi_p(k)]
    for (int idx = tid; idx < pc * D; idx += nt) {
      const int pl = idx / D, k = idx - pl * D;
      diff[idx] = Xi[k] - Xj[a.pi[(size_t)(p0 + pl) * D + k]];
    }
    __syncthreads();
    // A2: norms and radial factors, one warp per permutation
    for (int pl = warp; pl < pc; pl += nw) {
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
    // A3: u_p = J_i^T diff_p, v_p = (J_j[pi_p])^T diff_p  -- one thread per (p, atom, axis)
    for (int idx = tid; idx < pc * N3; idx += nt) {
      const int pl = idx / N3, r = idx - pl * N3, at = r / 3, ax = r - at * 3;
      const int* ipi = a.ipi + (size_t)(p0 + pl) * D;
      const double* dp = diff + pl * D;
      double su = 0.0, sv = 0.0;
      for (int x = 0; x < N; ++x) {
        if (x == at) continue;
        const int hi = x > at ? x : at, lo = x > at ? at : x;
        const int k = hi * (hi - 1) / 2 + lo;
        const double sgn = (at == lo) ? 1.0 : -1.0;  // +g at the smaller atom of the pair, -g at the larger
        su = fma(sgn * gi[k * 3 + ax], dp[k], su);
        sv = fma(sgn * gj[k * 3 + ax], dp[ipi[k]], sv);  // row k of J_j sits at position pi_p^-1(k) of J_j[pi_p]
      }
      u[idx] = su, v[idx] = sv;
    }
    __syncthreads();
    // B1: W[k][c] += c_p J_j[pi_p(k)][c]  -- thread (k, axis) owns its row/axis: no conflicts
    for (int idx = tid; idx < D * 3; idx += nt) {
      const int k = idx / 3, ax = idx - k * 3;
      for (int pl = 0; pl < pc; ++pl) {
        const int k2 = a.pi[(size_t)(p0 + pl) * D + k];
        const double val = cc[pl] * gj[k2 * 3 + ax];
        W[k * N3 + pb[k2] * 3 + ax] += val;
        W[k * N3 + pa[k2] * 3 + ax] -= val;
      }
    }
    // B2: rank-1 part, outputs (r, c) strided over the CTA
#pragma unroll
    for (int o = 0; o < MAXOUT; ++o) {
      const int e = tid + o * kKmThreads;
      if (e < N3 * N3) {
        const int r = e / N3, c = e - r * N3;
        double s = acc[o];
        for (int pl = 0; pl < pc; ++pl) s = fma(fb[pl] * u[pl * N3 + r], v[pl * N3 + c], s);
        acc[o] = s;
      }
    }
    if (a.use_E) {
      if (tid < N3)
        for (int pl = 0; pl < pc; ++pl) erow = fma(-cc[pl], v[pl * N3 + tid], erow);
      if (tid == 0)
        for (int pl = 0; pl < pc; ++pl) eacc -= kee[pl];
    }
    __syncthreads();
  }

  // C: K_ij[r][c] = acc - sum_{k containing atom(r)} J_i[k][r] W[k][c]
  double* Kb = a.K + ((size_t)i * N3) * a.ld + (size_t)j * N3;
#pragma unroll
  for (int o = 0; o < MAXOUT; ++o) {
    const int e = tid + o * kKmThreads;
    if (e < N3 * N3) {
      const int r = e / N3, c = e - r * N3, at = r / 3, ax = r - at * 3;
      double s = acc[o];
      for (int x = 0; x < N; ++x) {
        if (x == at) continue;
        const int hi = x > at ? x : at, lo = x > at ? at : x;
        const int k = hi * (hi - 1) / 2 + lo;
        const double sgn = (at == lo) ? 1.0 : -1.0;
        s = fma(-sgn * gi[k * 3 + ax], W[k * N3 + c], s);
      }
      Kb[(size_t)r * a.ld + c] = s;
    }
  }
  if (a.use_E) {
    const size_t E0 = (size_t)a.T * N3;
    if (tid < N3) {
      a.K[(E0 + i) * a.ld + (size_t)j * N3 + tid] = erow;    // train.py:224-239
      a.K[((size_t)j * N3 + tid) * a.ld + E0 + i] = erow;    // train.py:262-287 (worker j' = T + i, block j)
    }
    if (tid == 0) a.K[(E0 + j) * a.ld + E0 + i] = eacc;      // train.py:287-289
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
  double* out;            // [n_q][T * P]
};

__global__ void __launch_bounds__(kM52Threads) k_mat52(const Mat52Args a) {
  extern __shared__ __align__(16) unsigned char m52_smem[];
  const int D = a.D, P = a.P, N = a.n_atoms;
  const int m = blockIdx.y, t0 = blockIdx.x * kM52Rows;
  const int nrow = min(kM52Rows, a.T - t0);
  const int tid = threadIdx.x, nt = blockDim.x;
  double* xq = reinterpret_cast<double*>(m52_smem);  // [D]
  double* Xs = xq + D;                               // [kM52Rows][D + 1]
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
  const double sig = a.sig;
  const double s5 = sqrt(5.0) / sig, q2 = 5.0 / (3.0 * sig * sig);
  for (int idx = tid; idx < nrow * P; idx += nt) {
    const int tl = idx / P, p = idx - tl * P;
    const int* ip = a.iperm ? a.iperm + (size_t)p * D : nullptr;
    const double* xr = Xs + tl * (D + 1);
    double s = 0.0;
    for (int e = 0; e < D; ++e) {
      const double d = xq[ip ? ip[e] : e] - xr[e];
      s = fma(d, d, s);
    }
    const double d = sqrt(s);
    a.out[(size_t)m * a.T * P + (size_t)(t0 + tl) * P + p] = (1.0 + s5 * d + q2 * s) * exp(-s5 * d);
  }
}

}  // namespace mbg
