// K3 + K4: Matern-5/2 gradient-domain contraction of accepted fragments against a GDML
// model, fused with the J^T projection, scaling, alchemical factor and the scatter into
// per-atom totals (or per-fragment rows).
//
// Reference: mbgdml/predictors/gdml_predict.py:94-142 (contraction), 151-163 (J^T F_d),
//            251-267 (x std, + c, alchemy, F[r_slice] += f, virial);
//            mbgdml/models/gdml_model.py:109-118 (permutation expansion).
//
// Design.  The reference expands the training set under all P permutations
// (rows j = (t, p): x_j[d] = X_t[pi_p(d)], A_j[d] = A_t[pi_p(d)]) and contracts one
// query descriptor x against T*P rows.  Here the *query* is permuted instead:
// y_p[e] = x[pi_p^-1(e)], so one thread owns one (fragment, permutation) query, keeps
// y_p and its accumulator G_p in registers, and streams the T un-permuted training
// rows (X_t, A_t interleaved as double2) from shared memory as warp-uniform broadcast
// loads.  Tiles of TR rows are staged by TMA bulk copies (cp.async.bulk + mbarrier),
// double buffered.  Per (query, row):
//     diff = y - X_t ; n = sqrt(5)|diff| ; a = diff . A_t
//     c1 = exp(-n/sig) 5/(3 sig^3) ; c2 = c1 (n + sig) ; w1 = (5/sig) a c1 [+ aE_t c2]
//     E += a c2 ;  s1 += w1 ;  G -= w1 X_t + c2 A_t            (G += s1 y at the end)
// which equals the reference's  F_d += w1 diff - c2 A_j  because sum_t w1 (y - X_t) =
// s1 y - sum_t w1 X_t.  Un-permuting, F_d[d] = sum_p G_p[pi_p(d)].
// All arithmetic is FP64 on the FMA pipe: 5 D-wide FP64 instructions per pair.
#pragma once
#include "geom.cuh"

namespace mbg {

struct MaternArgs {
  SysDev sys;
  const double* R;       // [n_frames][n_atoms][3]
  const double2* XA;     // [T_pad][D]: (X_t[e], A_t[e]); rows >= T are (0, 0)
  const double* alphaE;  // [T_pad] (zeros past T) or nullptr
  const int* iperm;      // [P][D]: pi_p^-1
  int T_pad, P;
  double sig, c, std;
  long long n_frag;
  const int* rec_frame;
  const int* rec_atoms;
  const double* rec_lambda;
  const long long* rec_slot;
  int n_slots_per_frame;  // decomposed mode: rows per frame
  int decomp;             // 0: scatter into E[frame], F[frame][atom]; 1: rows E[slot], F[slot][NA][3]
  int want_virial;
  double* E;
  double* F;
  double* virial;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Shared-memory load the compiler may not merge with an earlier identical load: the second pass over a
// training row re-reads it (2 x 16 B broadcast loads per dimension) instead of pinning 2*D registers.
__device__ __forceinline__ double2 lds_row(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

constexpr int kMaternThreads = 256;  // max queries per CTA
constexpr int kTileRows = 64;        // training rows per TMA tile

// Bytes of dynamic shared memory for a launch with `nt` threads and P permutations.
template <int NA>
__host__ __device__ constexpr size_t matern_smem_bytes(int nt, int P) {
  constexpr int D = NA * (NA - 1) / 2;
  const int maxf = (nt + P - 1) / P + 1;
  return (size_t)2 * kTileRows * D * 16 + 2 * kTileRows * 8 + (size_t)nt * D * 8 + (size_t)nt * 8 +
         (size_t)maxf * D * 8 + (size_t)maxf * NA * 3 * 8 + 64;
}

template <int NA, bool USE_E>
__global__ void __launch_bounds__(kMaternThreads, 1) k_matern(const MaternArgs a) {
  constexpr int D = NA * (NA - 1) / 2;
  constexpr int TR = kTileRows;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int P = a.P;
  const int maxf = (nt + P - 1) / P + 1;

  double2* tiles = reinterpret_cast<double2*>(smem_raw);  // [2][TR][D]
  double* aEs = reinterpret_cast<double*>(tiles + 2 * TR * D);  // [2][TR]
  double* Gs = aEs + 2 * TR;                                    // [nt][D]
  double* Es = Gs + (size_t)nt * D;                             // [nt]
  double* xs = Es + nt;                                         // [maxf][D]
  double* rs = xs + (size_t)maxf * D;                           // [maxf][NA][3]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rs + (size_t)maxf * NA * 3);

  const long long nq = a.n_frag * P;
  const long long q0 = (long long)blockIdx.x * nt;
  const long long q_end = (q0 + nt < nq) ? q0 + nt : nq;  // exclusive
  const long long f_lo = q0 / P;
  const int nf = (int)((q_end - 1) / P - f_lo) + 1;

  const int n_tiles = a.T_pad / TR;
  constexpr uint32_t kTileBytes = TR * D * 16;
  constexpr uint32_t kStageBytes = kTileBytes + (USE_E ? TR * 8 : 0);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (s < n_tiles) {
        mbar_expect_tx(&bars[s], kStageBytes);
        tma_load_1d(tiles + s * TR * D, a.XA + (size_t)s * TR * D, kTileBytes, &bars[s]);
        if (USE_E) tma_load_1d(aEs + s * TR, a.alphaE + (size_t)s * TR, TR * 8, &bars[s]);
      }
  }

  // ---- prologue: geometry of this CTA's fragments (same code as the cull kernel) ----
  for (int fl = tid; fl < nf; fl += nt) {
    const long long frag = f_lo + fl;
    const int* atoms = a.rec_atoms + frag * NA;
    const double* Rf = a.R + (long long)a.rec_frame[frag] * 3ll * a.sys.n_atoms;
    double r[NA][3];
    int at[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) at[k] = atoms[k];
    fragment_coords<NA>(a.sys, Rf, at, r);
    int k = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      rs[(fl * NA + i) * 3 + 0] = r[i][0], rs[(fl * NA + i) * 3 + 1] = r[i][1], rs[(fl * NA + i) * 3 + 2] = r[i][2];
#pragma unroll
      for (int j = 0; j < i; ++j, ++k)  // np.tril_indices(n, -1) order (desc.py:101-108)
        xs[fl * D + k] = __ddiv_rn(1.0, norm3_rn(__dsub_rn(r[i][0], r[j][0]), __dsub_rn(r[i][1], r[j][1]),
                                                  __dsub_rn(r[i][2], r[j][2])));
    }
  }
  __syncthreads();

  // ---- this thread's query: fragment fl_q under permutation perm ----
  const long long q = q0 + tid;
  const bool active = q < nq;
  const int fl_q = active ? (int)(q / P - f_lo) : 0;
  const int perm = active ? (int)(q % P) : 0;
  const int* ip = a.iperm + perm * D;
  double y[D], G[D];
#pragma unroll
  for (int e = 0; e < D; ++e) {
    y[e] = xs[fl_q * D + ip[e]];
    G[e] = 0.0;
  }
  double Eacc = 0.0, s1 = 0.0;
  const double sig = a.sig;
  const double sig_inv = 1.0 / sig;
  const double base_fact = 5.0 / (3.0 * sig * sig * sig);
  const double diag_fact = 5.0 / sig;
  const double sqrt5 = 2.23606797749978969641;
  const double third_sig_inv = 1.0 / (3.0 * sig);

  // ---- main loop over training tiles ----
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int s = tile & 1;
    mbar_wait(&bars[s], (tile >> 1) & 1);
    const uint32_t tb = smem_u32(tiles + s * TR * D);
    const double* ab = aEs + s * TR;
#pragma unroll 1
    for (int t = 0; t < TR; ++t) {
      const uint32_t row = tb + t * (D * 16);
      double n2a = 0.0, n2b = 0.0, aa = 0.0, ab2 = 0.0;
#pragma unroll
      for (int e = 0; e < D; ++e) {
        const double2 v = lds_row(row + e * 16);
        const double d = y[e] - v.x;
        if (e & 1) {
          n2b = fma(d, d, n2b);
          ab2 = fma(d, v.y, ab2);
        } else {
          n2a = fma(d, d, n2a);
          aa = fma(d, v.y, aa);
        }
      }
      const double adot = aa + ab2;
      const double nrm = sqrt5 * sqrt(n2a + n2b);
      const double ex = exp(-nrm * sig_inv);
      const double c1 = ex * base_fact;
      const double c2 = c1 * (nrm + sig);
      double w1 = adot * c1 * diag_fact;
      Eacc = fma(adot, c2, Eacc);
      if (USE_E) {  // gdml_predict.py:132-140
        const double aE = ab[t];
        w1 = fma(aE, c2, w1);
        Eacc = fma(aE * ex, 1.0 + (nrm * sig_inv) * (1.0 + nrm * third_sig_inv), Eacc);
      }
      s1 += w1;
#pragma unroll
      for (int e = 0; e < D; ++e) {
        const double2 v = lds_row(row + e * 16);
        G[e] = fma(-w1, v.x, G[e]);
        G[e] = fma(-c2, v.y, G[e]);
      }
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && tile + 2 < n_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[s], kStageBytes);
      tma_load_1d(tiles + s * TR * D, a.XA + (size_t)(tile + 2) * TR * D, kTileBytes, &bars[s]);
      if (USE_E) tma_load_1d(aEs + s * TR, a.alphaE + (size_t)(tile + 2) * TR, TR * 8, &bars[s]);
    }
  }

  // ---- epilogue: un-permute, reduce over permutations, project, scale, scatter ----
#pragma unroll
  for (int e = 0; e < D; ++e) Gs[tid * D + ip[e]] = active ? fma(s1, y[e], G[e]) : 0.0;
  Es[tid] = active ? Eacc : 0.0;
  __syncthreads();

  // F_d of fragment fl -> first row of its segment; E likewise.
  for (int idx = tid; idx < nf * (D + 1); idx += nt) {
    const int fl = idx / (D + 1), d = idx - fl * (D + 1);
    const long long fq0 = (f_lo + fl) * P;
    const int t_lo = (int)((fq0 > q0 ? fq0 : q0) - q0);
    const long long fq1 = fq0 + P;
    const int t_hi = (int)((fq1 < q_end ? fq1 : q_end) - q0);
    double sum = 0.0;
    if (d < D) {
      for (int t = t_lo; t < t_hi; ++t) sum += Gs[t * D + d];
      Gs[t_lo * D + d] = sum;
    } else {
      for (int t = t_lo; t < t_hi; ++t) sum += Es[t];
      Es[t_lo] = sum;
    }
  }
  __syncthreads();

  for (int idx = tid; idx < nf * (NA * 3 + 1); idx += nt) {
    const int fl = idx / (NA * 3 + 1), rem = idx - fl * (NA * 3 + 1);
    const long long frag = f_lo + fl;
    const long long fq0 = frag * P;
    const int t_lo = (int)((fq0 > q0 ? fq0 : q0) - q0);
    const double lam = a.rec_lambda[frag];
    const double* Fd = Gs + t_lo * D;
    const double* rf = rs + fl * NA * 3;
    const double* xf = xs + fl * D;
    if (rem == NA * 3) {
      // e = E_raw * std + c (c once per fragment: by the CTA that holds permutation 0), times lambda
      const double e = (Es[t_lo] * a.std + (fq0 >= q0 ? a.c : 0.0)) * lam;
      double* dst = a.decomp ? a.E + ((long long)a.rec_frame[frag] * a.n_slots_per_frame + a.rec_slot[frag])
                             : a.E + a.rec_frame[frag];
      atomicAdd(dst, e);
    } else {
      const int m = rem / 3, c = rem - m * 3;
      double f = 0.0;
#pragma unroll
      for (int b = 0; b < NA; ++b) {
        if (b == m) continue;
        const int i = b > m ? b : m, j = b > m ? m : b;
        const double inv = xf[i * (i - 1) / 2 + j];
        f = fma((rf[b * 3 + c] - rf[m * 3 + c]) * (inv * inv * inv), Fd[i * (i - 1) / 2 + j], f);
      }
      f *= a.std * lam;
      double* dst =
          a.decomp
              ? a.F + (((long long)a.rec_frame[frag] * a.n_slots_per_frame + a.rec_slot[frag]) * NA + m) * 3 + c
              : a.F + ((long long)a.rec_frame[frag] * a.sys.n_atoms + a.rec_atoms[frag * NA + m]) * 3 + c;
      atomicAdd(dst, f);
    }
  }

  if (a.want_virial) {  // stress.virial_atom_loop on the minimum-image coordinates (gdml_predict.py:266-267)
    for (int idx = tid; idx < nf * 9; idx += nt) {
      const int fl = idx / 9, ij = idx - fl * 9, vi = ij / 3, vj = ij - vi * 3;
      const long long frag = f_lo + fl;
      const long long fq0 = frag * P;
      const int t_lo = (int)((fq0 > q0 ? fq0 : q0) - q0);
      const double* Fd = Gs + t_lo * D;
      const double* rf = rs + fl * NA * 3;
      const double* xf = xs + fl * D;
      double w = 0.0;
      for (int m = 0; m < NA; ++m) {
        double f = 0.0;
        for (int b = 0; b < NA; ++b) {
          if (b == m) continue;
          const int i = b > m ? b : m, j = b > m ? m : b;
          const double inv = xf[i * (i - 1) / 2 + j];
          f = fma((rf[b * 3 + vj] - rf[m * 3 + vj]) * (inv * inv * inv), Fd[i * (i - 1) / 2 + j], f);
        }
        w = fma(rf[m * 3 + vi], f, w);
      }
      atomicAdd(a.virial + (long long)a.rec_frame[frag] * 9 + ij, w * a.std * a.rec_lambda[frag]);
    }
  }
}

}  // namespace mbg
