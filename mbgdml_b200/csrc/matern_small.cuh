// K3 + K4 for narrow descriptors (fragments of 2-4 atoms, D = 1, 3, 6; BASELINE config 2: water monomers).
//
// With D <= 6 the D-wide work of a (query, training row) pair is 5 D <= 30 FMA-pipe instructions against the 17 of the
// sqrt/exp chain: the pair is transcendental-bound and a GEMM recast only adds padding (an m8n8k4 tile is 2/3 empty for
// D = 3).  So: one THREAD per query, the difference form of the reference evaluated literally on the FMA pipe
// (diff = y - X_t; |diff|^2; diff.A_t; chain; G += w diff - u A_t), the whole model in shared memory as (X, A_s)
// pairs read as warp-uniform broadcast loads, the same 4096-entry exp table and cubic as the tensor-pipe kernel.
// Execution shape as k_matern_pw: a persistent grid that reads the fragment count on the device (no host round trip,
// graph-capturable); every CTA takes an equal contiguous share of the queries, 512 at a time.
//
// Reference: mbgdml/predictors/gdml_predict.py:94-142 (contraction), 151-163 (J^T F_d), 251-267 (scale, scatter, virial).
#pragma once
#include <type_traits>

#include "matern_pw.cuh"

namespace mbg {

constexpr int kSmallThreads = 512;  // queries per block of work == threads per CTA (one CTA of 16 warps per SM)

template <int NA>
__host__ __device__ constexpr size_t small_fixed_bytes(int P) {  // everything but the model rows
  constexpr int D = NA * (NA - 1) / 2;
  const int maxf = (kSmallThreads + P - 1) / P + 1;
  return (size_t)kMmaExpTable * 8 + (size_t)kSmallThreads * (D + 1) * 8 + (size_t)maxf * (D + NA * 3 + D + 1) * 8 +
         (size_t)maxf * 4 + sizeof(CellDev) + 64;
}

// rows of the model kept in shared memory per pass (all of them when they fit)
template <int NA>
__host__ __device__ inline int small_rows_per_pass(int T, int P, size_t smem_budget) {
  constexpr int D = NA * (NA - 1) / 2;
  const size_t avail = smem_budget > small_fixed_bytes<NA>(P) ? smem_budget - small_fixed_bytes<NA>(P) : 0;
  const long long fit = (long long)(avail / ((size_t)D * 16 + 8));
  return (int)(fit < T ? fit : T);
}

struct MaternSmallArgs {
  SysDev sys;
  const double* R;
  const double2* XA;         // [T_pad][D] (X_t[e], A_t[e]) (the FMA-pipe kernel's model layout, Dp = D for D <= 21)
  const double* alphaE;      // [T_pad] or nullptr
  const double* exp2_table;  // [kMmaExpTable]
  const int* iperm;          // [P][D]
  const int* perm;           // [P][D]
  int T, P, rows_per_pass;
  double sig, c, std;
  long long n_frag;
  const long long* n_frag_dev;
  const int* rec_frame;
  const int* rec_atoms;
  const double* rec_lambda;
  const long long* rec_slot;
  int n_slots_per_frame, decomp, want_virial;
  double* E;
  double* F;
  double* virial;
  int use_lat;
  double lat[9], lat_inv[9];
};

template <int NA, bool USE_E>
__global__ void __launch_bounds__(kSmallThreads, 1) k_matern_small(const __grid_constant__ MaternSmallArgs a) {
  constexpr int D = NA * (NA - 1) / 2, NT = kSmallThreads;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int tid = threadIdx.x, P = a.P;
  const int maxf = (NT + P - 1) / P + 1;
  const int RPP = a.rows_per_pass;
  double* etab = reinterpret_cast<double*>(sm_raw);                 // [4096]
  double2* rows = reinterpret_cast<double2*>(etab + kMmaExpTable);  // [RPP][D] (X, A_s = (5/sig) A)
  double* aEs = reinterpret_cast<double*>(rows + (size_t)RPP * D);  // [RPP]
  double* Gs = aEs + RPP;                                           // [NT][D + 1] per-query results
  double* xs = Gs + (size_t)NT * (D + 1);                           // [maxf][D]
  double* rs = xs + (size_t)maxf * D;                               // [maxf][NA][3]
  double* Fd = rs + (size_t)maxf * NA * 3;                          // [maxf][D + 1]
  int* bad = reinterpret_cast<int*>(Fd + (size_t)maxf * (D + 1));   // [maxf]
  CellDev* cell_s = reinterpret_cast<CellDev*>(reinterpret_cast<double*>(bad) + (maxf + 1) / 2);

  const long long n_frag = a.n_frag_dev ? min(a.n_frag, *a.n_frag_dev) : a.n_frag;
  const long long nq = n_frag * P;
  // Every CTA takes one contiguous share of the queries (a multiple of 32, processed NT at a time): all CTAs finish
  // together whatever the count -- with whole blocks of NT queries dealt round-robin, 391 blocks on 148 SMs are 3 rounds
  // for 2.64 rounds of work.
  const long long share = ((nq + gridDim.x - 1) / gridDim.x + 31) / 32 * 32;
  const long long q_lo = (long long)blockIdx.x * share, q_hi = min(q_lo + share, nq);
  if (q_lo >= nq) return;
  for (int k = tid; k < kMmaExpTable; k += NT) etab[k] = a.exp2_table[k];
  if (a.sys.periodic && a.sys.cell_stride == 0)
    for (int k = tid; k < (int)(sizeof(CellDev) / 8); k += NT)
      reinterpret_cast<double*>(cell_s)[k] = reinterpret_cast<const double*>(a.sys.cells)[k];
  const double sig = a.sig, diag = 5.0 / sig;
  const bool one_pass = RPP >= a.T;
  if (one_pass) {  // the whole model stays resident for every block of work of this CTA
    for (int k = tid; k < a.T * D; k += NT) rows[k] = make_double2(a.XA[k].x, diag * a.XA[k].y);
    if (USE_E)
      for (int k = tid; k < a.T; k += NT) aEs[k] = a.alphaE[k];
  }
  __syncthreads();

  const double k64 = -(double)kMmaExpTable * 1.4426950408889634 / sig;
  const double kc1 = sig / 5.0, kc0 = sig * sig / 5.0;
  const double s_max = (3900000.0 / k64) * (3900000.0 / k64);
  const int hi_max = __double2hiint(s_max);
  const double kMagic = 6755399441055744.0;
  const double base = 5.0 / (3.0 * sig * sig * sig);

  for (long long q0 = q_lo; q0 < q_hi; q0 += NT) {
    const long long q_end = (q0 + NT < q_hi) ? q0 + NT : q_hi;
    const long long f_lo = q0 / P;
    const int nf = (int)((q_end - 1) / P - f_lo) + 1;
    // ---- geometry + descriptors of this block's fragments (periodic.py:61-107, desc.py:42-158) ----
    for (int fl = tid; fl < nf; fl += NT) bad[fl] = 0;
    __syncthreads();
    for (int idx = tid; idx < nf * NA; idx += NT) {
      const int fl = idx / NA, k = idx - fl * NA;
      const long long frag = f_lo + fl;
      const int frame = a.rec_frame[frag];
      const double* Rf = a.R + (long long)frame * 3ll * a.sys.n_atoms;
      const double* p = Rf + 3ll * a.rec_atoms[frag * NA + k];
      double v[3] = {p[0], p[1], p[2]};
      if (a.sys.periodic) {
        const double* pf = Rf + 3ll * a.rec_atoms[frag * NA];
        const double d[3] = {__dsub_rn(v[0], pf[0]), __dsub_rn(v[1], pf[1]), __dsub_rn(v[2], pf[2])};
        const CellDev& cl = a.sys.cell_stride ? cell_of(a.sys, frame) : *cell_s;
        if (!(mic_naive(cl, d, v) < cl.half_min_len)) bad[fl] = 1;
      }
      rs[idx * 3 + 0] = v[0], rs[idx * 3 + 1] = v[1], rs[idx * 3 + 2] = v[2];
    }
    __syncthreads();
    if (a.sys.periodic) {  // find_mic falls back to the general algorithm for the whole fragment
      for (int idx = tid; idx < nf * NA; idx += NT) {
        const int fl = idx / NA, k = idx - fl * NA;
        if (!bad[fl]) continue;
        const long long frag = f_lo + fl;
        const int frame = a.rec_frame[frag];
        const double* Rf = a.R + (long long)frame * 3ll * a.sys.n_atoms;
        const double* p = Rf + 3ll * a.rec_atoms[frag * NA + k];
        const double* pf = Rf + 3ll * a.rec_atoms[frag * NA];
        const double d[3] = {__dsub_rn(p[0], pf[0]), __dsub_rn(p[1], pf[1]), __dsub_rn(p[2], pf[2])};
        double v[3];
        mic_general(a.sys.cell_stride ? cell_of(a.sys, frame) : *cell_s, d, v);
        rs[idx * 3 + 0] = v[0], rs[idx * 3 + 1] = v[1], rs[idx * 3 + 2] = v[2];
      }
      __syncthreads();
    }
    for (int idx = tid; idx < nf * D; idx += NT) {
      const int fl = idx / D, k = idx - fl * D;
      int i = 1;
#pragma unroll
      for (int ii = 1; ii < NA; ++ii)
        if (k >= ii * (ii + 1) / 2) i = ii + 1;
      const int j = k - i * (i - 1) / 2;
      const double* ri = rs + (fl * NA + i) * 3;
      const double* rj = rs + (fl * NA + j) * 3;
      double dv[3] = {__dsub_rn(ri[0], rj[0]), __dsub_rn(ri[1], rj[1]), __dsub_rn(ri[2], rj[2])};
      if (a.use_lat) pbc_diff(a.lat, a.lat_inv, dv);
      xs[idx] = __ddiv_rn(1.0, norm3_rn(dv[0], dv[1], dv[2]));
    }
    for (int idx = tid; idx < nf * (D + 1); idx += NT) Fd[idx] = 0.0;
    __syncthreads();

    // ---- this thread's query ----
    const long long q = q0 + tid;
    const bool act = q < q_end;
    const int flq = act ? (int)(q / P - f_lo) : 0;
    const int pq = act ? (int)(q % P) : 0;
    double y[D], G[D], Eq = 0.0, E2 = 0.0;
#pragma unroll
    for (int e = 0; e < D; ++e) y[e] = xs[flq * D + a.iperm[pq * D + e]], G[e] = 0.0;

    // ---- main loop: every training row against this query (gdml_predict.py:109-142) ----
    for (int t0 = 0; t0 < a.T; t0 += RPP) {
      const int nr = min(RPP, a.T - t0);
      if (!one_pass) {
        __syncthreads();
        for (int k = tid; k < nr * D; k += NT) rows[k] = make_double2(a.XA[(size_t)t0 * D + k].x, diag * a.XA[(size_t)t0 * D + k].y);
        if (USE_E)
          for (int k = tid; k < nr; k += NT) aEs[k] = a.alphaE[t0 + k];
        __syncthreads();
      }
      // RB rows per iteration: RB independent sqrt/exp chains per thread, advanced stage by stage so that dependent
      // FP64 instructions are RB apart; the remainder rows go one at a time
      auto block = [&](const int t, auto rb_tag) {
        constexpr int RB = decltype(rb_tag)::value;
        double df[RB][D], sq[RB], as[RB];
        double2 xa[RB][D];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
          for (int e = 0; e < D; ++e) xa[r][e] = rows[(t + r) * D + e];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          sq[r] = 0.0, as[r] = 0.0;
#pragma unroll
          for (int e = 0; e < D; ++e) {
            df[r][e] = y[e] - xa[r][e].x;
            sq[r] = fma(df[r][e], df[r][e], sq[r]);
            as[r] = fma(df[r][e], xa[r][e].y, as[r]);
          }
        }
        double g_[RB], h_[RB], r_[RB], t_[RB];
        int kk[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {  // 5|diff|^2, clamped to [1e-300, s_max] on the integer pipe
          sq[r] = 5.0 * sq[r];
          const int hi = min(max(__double2hiint(sq[r]), 0x01A56E1F), hi_max);
          sq[r] = __hiloint2double(hi, __double2loint(sq[r]));
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(t_[r]) : "d"(sq[r]));
#pragma unroll
        for (int r = 0; r < RB; ++r) g_[r] = sq[r] * t_[r];
#pragma unroll
        for (int r = 0; r < RB; ++r) h_[r] = __hiloint2double(__double2hiint(t_[r]) - 0x00100000, 0);  // t / 2
#pragma unroll
        for (int r = 0; r < RB; ++r) r_[r] = fma(-g_[r], h_[r], 0.5);
#pragma unroll
        for (int r = 0; r < RB; ++r) g_[r] = fma(g_[r], r_[r], g_[r]);
#pragma unroll
        for (int r = 0; r < RB; ++r) r_[r] = fma(-g_[r], g_[r], sq[r]);
#pragma unroll
        for (int r = 0; r < RB; ++r) g_[r] = fma(r_[r], h_[r], g_[r]);  // sqrt(5)|diff|
#pragma unroll
        for (int r = 0; r < RB; ++r) t_[r] = fma(g_[r], k64, kMagic);
#pragma unroll
        for (int r = 0; r < RB; ++r) sq[r] = fma(g_[r], kc1, kc0);  // (n + sig) sig/5
#pragma unroll
        for (int r = 0; r < RB; ++r) kk[r] = __double2loint(t_[r]);
#pragma unroll
        for (int r = 0; r < RB; ++r) t_[r] = t_[r] - kMagic;
#pragma unroll
        for (int r = 0; r < RB; ++r) r_[r] = fma(g_[r], k64, -t_[r]);
#pragma unroll
        for (int r = 0; r < RB; ++r) h_[r] = fma(r_[r], kMmaExpC3, kMmaExpC2);
#pragma unroll
        for (int r = 0; r < RB; ++r) t_[r] = etab[kk[r] & (kMmaExpTable - 1)];
#pragma unroll
        for (int r = 0; r < RB; ++r) h_[r] = fma(h_[r], r_[r], kMmaExpC1);
#pragma unroll
        for (int r = 0; r < RB; ++r) h_[r] = h_[r] * r_[r];
#pragma unroll
        for (int r = 0; r < RB; ++r) h_[r] = fma(t_[r], h_[r], t_[r]);
#pragma unroll
        for (int r = 0; r < RB; ++r)
          h_[r] = __hiloint2double(__double2hiint(h_[r]) + ((kk[r] >> kMmaExpBits) << 20), __double2loint(h_[r]));  // ex
#pragma unroll
        for (int r = 0; r < RB; ++r) r_[r] = as[r] * h_[r];  // w
#pragma unroll
        for (int r = 0; r < RB; ++r) sq[r] = sq[r] * h_[r];  // u = c2s
        if (USE_E) {  // gdml_predict.py:132-140
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const double aE = aEs[t + r], nos = g_[r] / sig;
            r_[r] = fma(aE * diag, sq[r], r_[r]);
            E2 = fma(aE * h_[r], 1.0 + nos * (1.0 + nos * (1.0 / 3.0)), E2);
          }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) Eq = fma(as[r], sq[r], Eq);
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
          for (int e = 0; e < D; ++e) G[e] = fma(r_[r], df[r][e], fma(-sq[r], xa[r][e].y, G[e]));
      };
      constexpr int RBM = 4;
      int t = 0;
#pragma unroll 1
      for (; t + RBM <= nr; t += RBM) block(t, std::integral_constant<int, RBM>());
#pragma unroll 1
      for (; t < nr; ++t) block(t, std::integral_constant<int, 1>());
    }

    // ---- epilogue: un-permute, reduce over the queries of a fragment, project, scatter ----
#pragma unroll
    for (int e = 0; e < D; ++e) Gs[tid * (D + 1) + e] = act ? base * G[e] : 0.0;
    Gs[tid * (D + 1) + D] = act ? fma(base, Eq, E2) : 0.0;
    __syncthreads();
    for (int idx = tid; idx < nf * (D + 1); idx += NT) {
      const int fl = idx / (D + 1), d = idx - fl * (D + 1);
      const long long fq0 = (f_lo + fl) * P, fq1 = fq0 + P;
      const int t_lo = (int)(max(fq0, q0) - q0), t_hi = (int)(min(fq1, q_end) - q0);
      double sum = 0.0;
      for (int t = t_lo; t < t_hi; ++t) {
        const int p = (int)((q0 + t) - fq0);
        sum += Gs[t * (D + 1) + (d < D ? a.perm[p * D + d] : D)];  // F_d[d] = sum_p G_p[pi_p(d)]
      }
      Fd[idx] = sum;
    }
    __syncthreads();
    for (int idx = tid; idx < nf * (NA * 3 + 1); idx += NT) {
      const int fl = idx / (NA * 3 + 1), rem = idx - fl * (NA * 3 + 1);
      const long long frag = f_lo + fl;
      const long long fq0 = frag * P;
      const double lam = a.rec_lambda[frag];
      const double* Fdf = Fd + fl * (D + 1);
      const double* rf = rs + fl * NA * 3;
      const double* xf = xs + fl * D;
      if (rem == NA * 3) {
        const double e = (Fdf[D] * a.std + (fq0 >= q0 ? a.c : 0.0)) * lam;
        double* dst = a.decomp ? a.E + ((long long)a.rec_frame[frag] * a.n_slots_per_frame + a.rec_slot[frag])
                               : a.E + a.rec_frame[frag];
        atomicAdd(dst, e);
      } else {
        const int m = rem / 3, c = rem - m * 3;
        double f = 0.0;
#pragma unroll
        for (int bb = 0; bb < NA; ++bb) {
          if (bb == m) continue;
          const int i = bb > m ? bb : m, j = bb > m ? m : bb;
          const double inv = xf[i * (i - 1) / 2 + j];
          double dc = rf[bb * 3 + c] - rf[m * 3 + c];
          if (a.use_lat) {
            double dv[3] = {rf[bb * 3] - rf[m * 3], rf[bb * 3 + 1] - rf[m * 3 + 1], rf[bb * 3 + 2] - rf[m * 3 + 2]};
            pbc_diff(a.lat, a.lat_inv, dv);
            dc = dv[c];
          }
          f = fma(dc * (inv * inv * inv), Fdf[i * (i - 1) / 2 + j], f);
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
      for (int idx = tid; idx < nf * 9; idx += NT) {
        const int fl = idx / 9, ij = idx - fl * 9, vi = ij / 3, vj = ij - vi * 3;
        const long long frag = f_lo + fl;
        const double* Fdf = Fd + fl * (D + 1);
        const double* rf = rs + fl * NA * 3;
        const double* xf = xs + fl * D;
        double wv = 0.0;
        for (int m = 0; m < NA; ++m) {
          double f = 0.0;
          for (int bb = 0; bb < NA; ++bb) {
            if (bb == m) continue;
            const int i = bb > m ? bb : m, j = bb > m ? m : bb;
            const double inv = xf[i * (i - 1) / 2 + j];
            double dc = rf[bb * 3 + vj] - rf[m * 3 + vj];
            if (a.use_lat) {
              double dv[3] = {rf[bb * 3] - rf[m * 3], rf[bb * 3 + 1] - rf[m * 3 + 1], rf[bb * 3 + 2] - rf[m * 3 + 2]};
              pbc_diff(a.lat, a.lat_inv, dv);
              dc = dv[vj];
            }
            f = fma(dc * (inv * inv * inv), Fdf[i * (i - 1) / 2 + j], f);
          }
          wv = fma(rf[m * 3 + vi], f, wv);
        }
        atomicAdd(a.virial + (long long)a.rec_frame[frag] * 9 + ij, wv * a.std * a.rec_lambda[frag]);
      }
    }
    __syncthreads();  // xs / rs / Fd / Gs are rewritten by the next block of work
  }
}

}  // namespace mbg
