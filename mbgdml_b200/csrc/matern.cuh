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
  const double2* XA;     // [T_pad][Dp]: (X_t[e], A_t[e]); rows >= T and entries e >= D are (0, 0)
  const double* alphaE;  // [T_pad] (zeros past T) or nullptr
  const double* exp2_table;  // [64] 2^(j/64) correctly rounded (host-computed)
  const int* iperm;      // [P][D]: pi_p^-1
  int T_pad, P;
  double sig, c, std;
  long long n_frag;
  const long long* n_frag_dev;  // optional: the accepted count still on the device (n_frag is then an upper bound)
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// Arithmetic wrapped in `asm volatile`: the compiler keeps volatile asm statements in program order, so
// the stages of the scalar chain stay interleaved with the (volatile) shared-memory loads of the other
// two streams exactly as written, instead of being re-serialised into one latency-bound run.
__device__ __forceinline__ double vfma(double a, double b, double c) {
  double r;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}
__device__ __forceinline__ double vmul(double a, double b) {
  double r;
  asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
__device__ __forceinline__ double vadd(double a, double b) {
  double r;
  asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
__device__ __forceinline__ double vmax(double a, double b) {
  double r;
  asm volatile("max.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}

constexpr int kExpTableBits = 6;      // 2^(j/64), j = 0..63, staged in shared memory
constexpr int kMaternThreads = 256;  // threads per CTA
constexpr int kTileRows = 64;        // training rows per TMA tile

// Register blocking of the contraction.  A "team" of L adjacent lanes owns Q queries; each lane keeps the
// slice [h*DL, (h+1)*DL) of the descriptor dimensions of all Q queries (y and G: 2*Q*DL doubles), so one
// 16-byte shared-memory load of (X_t[e], A_t[e]) feeds Q queries.  Partial |diff|^2 and diff.A are
// exchanged inside the team with warp shuffles; lane h evaluates the exp/sqrt chain of query h, so no
// transcendental is computed twice.  L = 1: no exchange.  L = 2 requires Q = 2.
template <int NA>
struct Blocking {
  static constexpr int D = NA * (NA - 1) / 2;
  static constexpr int L = (D > 21) ? 2 : 1;
  static constexpr int Q = (D <= 6) ? 4 : 2;
  static constexpr int DL = (D + L - 1) / L;  // dimensions per lane
  static constexpr int Dp = DL * L;           // padded row length of the device model (pad entries are 0)
  static constexpr int MinCtas = (D <= 6) ? 2 : 1;  // resident CTAs per SM the registers allow
  static_assert(L == 1 || (L == 2 && Q == 2), "unsupported team shape");
};

inline int model_row_pad(int na) {  // Dp for a fragment size (host side, model upload)
  const int D = na * (na - 1) / 2;
  const int L = (D > 21) ? 2 : 1;
  return (D + L - 1) / L * L;
}

// Queries per CTA for a launch with `nt` threads.
template <int NA>
__host__ __device__ constexpr int matern_queries_per_cta(int nt) {
  return nt / Blocking<NA>::L * Blocking<NA>::Q;
}

// Bytes of dynamic shared memory for a launch with `nt` threads and P permutations.
template <int NA>
__host__ __device__ constexpr size_t matern_smem_bytes(int nt, int P) {
  constexpr int D = Blocking<NA>::D, Dp = Blocking<NA>::Dp;
  const int qc = matern_queries_per_cta<NA>(nt);
  const int maxf = (qc + P - 1) / P + 1;
  return (size_t)2 * kTileRows * Dp * 16 + 2 * kTileRows * 8 + (size_t)qc * D * 8 + (size_t)qc * 8 +
         (size_t)maxf * D * 8 + (size_t)maxf * NA * 3 * 8 + (8 << kExpTableBits) + 64;  // incl. 4 mbarriers
}

// ---- branch-free FP64 math for the per-pair scalar chain --------------------------------------------
// libdevice's exp()/sqrt() carry slow-path branches that split the loop body into basic blocks and stop
// the scheduler from overlapping the (latency-bound) scalar chain of row t with the (throughput-bound)
// dot products of row t+1.  These versions are straight-line code, ~1 ulp.


// sqrt(x) for x >= 0 (x is clamped to 1e-300, so sqrt(0) returns 1e-150): rsqrt seed (MUFU.RSQ64H,
// ~2^-22) + one Goldschmidt iteration (~2^-43) + one residual correction (~1 ulp).
__device__ __forceinline__ double fast_sqrt(double x) {
  x = fmax(x, 1e-300);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, hh = 0.5 * y;
  const double r = fma(-g, hh, 0.5);
  g = fma(g, r, g);
  hh = fma(hh, r, hh);
  return fma(fma(-g, g, x), hh, g);
}

// exp(z) for z <= 0: z = (64 k + j) ln2/64 + r, exp(z) = 2^k 2^(j/64) (1 + expm1(r)), |r| <= ln2/128;
// expm1 by a degree-5 polynomial (truncation 3.5e-17), 2^(j/64) from a 64-entry shared-memory table.
__device__ __forceinline__ double fast_exp_nonpos(double z, const double* __restrict__ tab) {
  z = fmax(z, -708.0);  // below this the result underflows; it is then ~1e-308 instead of 0
  const double t = fma(z, 92.332482616893656, 6755399441055744.0);  // 64/ln2 ; 1.5 * 2^52 rounds to integer
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(kf, -0.01083042469326756, z);      // ln2/64 (high part)
  r = fma(kf, -2.9815858269852933e-12, r);          // ln2/64 (low part)
  double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
  p = fma(p, r, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;
  const double tj = tab[k & ((1 << kExpTableBits) - 1)];
  const double res = fma(tj, p, tj);
  // scale by 2^(k >> 6): add to the exponent field (result stays normal because z >= -708)
  return __hiloint2double(__double2hiint(res) + ((k >> kExpTableBits) << 20), __double2loint(res));
}

// One Matern-5/2 pair: from |diff|^2 (nsq) and diff.A (adot) to the weights of the accumulation.
struct PairWeights {
  double w1, c2, e;
};
struct PairConsts {
  double sig, neg_sig_inv, base_fact, diag_fact, third_sig_inv;
};
template <bool USE_E>
__device__ __forceinline__ PairWeights matern_pair(double nsq, double adot, double aE, const PairConsts& k,
                                                   const double* __restrict__ exp_tab) {
  const double nrm = fast_sqrt(5.0 * nsq);  // sqrt(5) |diff|
  const double ex = fast_exp_nonpos(nrm * k.neg_sig_inv, exp_tab);
  const double c1 = ex * k.base_fact;
  PairWeights r;
  r.c2 = c1 * (nrm + k.sig);
  r.w1 = adot * c1 * k.diag_fact;
  r.e = adot * r.c2;
  if (USE_E) {  // gdml_predict.py:132-140
    r.w1 = fma(aE, r.c2, r.w1);
    r.e = fma(aE * ex, 1.0 + (-nrm * k.neg_sig_inv) * (1.0 + nrm * k.third_sig_inv), r.e);
  }
  return r;
}

template <int NA, bool USE_E>
__global__ void __launch_bounds__(kMaternThreads, Blocking<NA>::MinCtas) k_matern(const __grid_constant__ MaternArgs a) {
  using B = Blocking<NA>;
  constexpr int D = B::D, L = B::L, Q = B::Q, DL = B::DL, Dp = B::Dp;
  constexpr int TR = kTileRows;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int P = a.P;
  const int qc = matern_queries_per_cta<NA>(nt);  // queries of this CTA
  const int maxf = (qc + P - 1) / P + 1;
  const int team = tid / L, h = tid % L;

  double2* tiles = reinterpret_cast<double2*>(smem_raw);         // [2][TR][Dp]
  double* aEs = reinterpret_cast<double*>(tiles + 2 * TR * Dp);  // [2][TR]
  double* Gs = aEs + 2 * TR;                                     // [qc][D]
  double* Es = Gs + (size_t)qc * D;                              // [qc]
  double* xs = Es + qc;                                          // [maxf][D]
  double* rs = xs + (size_t)maxf * D;                            // [maxf][NA][3]
  double* etab = rs + (size_t)maxf * NA * 3;                     // [64] 2^(j/64)
  uint64_t* bars = reinterpret_cast<uint64_t*>(etab + (1 << kExpTableBits));

  const long long n_frag = a.n_frag_dev ? min(a.n_frag, *a.n_frag_dev) : a.n_frag;
  const long long nq = n_frag * P;
  const long long q0 = (long long)blockIdx.x * qc;
  if (q0 >= nq) return;  // the grid was sized for an upper bound of the fragment count (whole CTA, before any barrier)
  const long long q_end = (q0 + qc < nq) ? q0 + qc : nq;  // exclusive
  const long long f_lo = q0 / P;
  const int nf = (int)((q_end - 1) / P - f_lo) + 1;

  const int n_tiles = a.T_pad / TR;
  constexpr uint32_t kTileBytes = TR * Dp * 16;
  constexpr uint32_t kStageBytes = kTileBytes + (USE_E ? TR * 8 : 0);

  // bars[0..1]: "full" (TMA landed a tile in stage s); bars[2..3]: "empty" (every warp is done with stage s).
  // There is no CTA-wide barrier in the main loop: warps drift apart freely, which keeps the FP64 pipe fed
  // by one warp's dot products while another warp sits in its latency-bound scalar chain.
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
        tma_load_1d(tiles + s * TR * Dp, a.XA + (size_t)s * TR * Dp, kTileBytes, &bars[s]);
        if (USE_E) tma_load_1d(aEs + s * TR, a.alphaE + (size_t)s * TR, TR * 8, &bars[s]);
      }
  }

  if (tid < (1 << kExpTableBits)) etab[tid] = a.exp2_table[tid];

  // ---- prologue: geometry of this CTA's fragments (same code as the cull kernel) ----
  for (int fl = tid; fl < nf; fl += nt) {
    const long long frag = f_lo + fl;
    const int* atoms = a.rec_atoms + frag * NA;
    const double* Rf = a.R + (long long)a.rec_frame[frag] * 3ll * a.sys.n_atoms;
    double r[NA][3];
    int at[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) at[k] = atoms[k];
    fragment_coords<NA>(a.sys, a.rec_frame[frag], Rf, at, r);
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

  // ---- this lane's Q queries (fragment, permutation) and its slice of their descriptors ----
  bool active[Q];
  const int* ip[Q];
  double y[Q][DL], G[Q][DL];
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    const long long q = q0 + (long long)team * Q + j;
    active[j] = q < nq;
    const int fl_q = active[j] ? (int)(q / P - f_lo) : 0;
    ip[j] = a.iperm + (active[j] ? (int)(q % P) : 0) * D;
#pragma unroll
    for (int e = 0; e < DL; ++e) {
      const int eg = h * DL + e;
      y[j][e] = (eg < D) ? xs[fl_q * D + ip[j][eg]] : 0.0;
      G[j][e] = 0.0;
    }
  }
  constexpr int OWN = (L == 1) ? Q : 1;  // queries whose exp/sqrt chain this lane evaluates
  double Eacc[OWN], s1[OWN];
#pragma unroll
  for (int j = 0; j < OWN; ++j) Eacc[j] = 0.0, s1[j] = 0.0;
  PairConsts pc;
  pc.sig = a.sig;
  pc.neg_sig_inv = -1.0 / a.sig;
  pc.base_fact = 5.0 / (3.0 * a.sig * a.sig * a.sig);
  pc.diag_fact = 5.0 / a.sig;
  pc.third_sig_inv = 1.0 / (3.0 * a.sig);

  // ---- main loop over training tiles ----
  // Per block of RB rows: (a) the scalar chains of the block (shuffle -> sqrt -> exp -> shuffle; latency
  // bound), (b) the dot products of the NEXT block (throughput bound, independent of (a)), (c) the
  // accumulation of the block.  Measured on B200 (profiles/r1_kernel_tuning.md): with 2 warps per SM
  // sub-partition the FP64 pipe idles when both warps sit in (a), so (a) is kept short (branch-free math,
  // RB interleaved chains) -- skewing the warps or dropping the per-tile barrier did not help.
  // Dimensions are processed in chunks of CH so that dependent FP64 instructions (difference -> FMA,
  // and the two FMAs into one accumulator) are CH*Q instructions apart, beyond the ~9-cycle DFMA latency.
  constexpr int CH = (DL % 3 == 0) ? 3 : (DL % 2 == 0) ? 2 : 1;
  auto dots = [&](uint32_t row, double (&n2)[Q], double (&aa)[Q]) {
#pragma unroll
    for (int j = 0; j < Q; ++j) n2[j] = 0.0, aa[j] = 0.0;
#pragma unroll
    for (int e0 = 0; e0 < DL; e0 += CH) {
      double2 v[CH];
      double d[CH][Q];
#pragma unroll
      for (int c = 0; c < CH; ++c) v[c] = lds_row(row + (e0 + c) * 16);
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int j = 0; j < Q; ++j) d[c][j] = y[j][e0 + c] - v[c].x;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
          n2[j] = fma(d[c][j], d[c][j], n2[j]);
          aa[j] = fma(d[c][j], v[c].y, aa[j]);
        }
    }
  };
  auto weights = [&](const double (&n2)[Q], const double (&aa)[Q], double aE, double (&w1)[Q], double (&c2)[Q]) {
    if constexpr (L == 1) {
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const PairWeights pw = matern_pair<USE_E>(n2[j], aa[j], aE, pc, etab);
        w1[j] = pw.w1, c2[j] = pw.c2;
        Eacc[j] += pw.e;
        s1[j] += pw.w1;
      }
    } else {  // L == 2, Q == 2: lane h owns query h; swap the other query's partial sums
      const double own_n = (h ? n2[1] : n2[0]) + __shfl_xor_sync(0xffffffffu, h ? n2[0] : n2[1], 1);
      const double own_a = (h ? aa[1] : aa[0]) + __shfl_xor_sync(0xffffffffu, h ? aa[0] : aa[1], 1);
      const PairWeights pw = matern_pair<USE_E>(own_n, own_a, aE, pc, etab);
      Eacc[0] += pw.e;
      s1[0] += pw.w1;
      const double w1x = __shfl_xor_sync(0xffffffffu, pw.w1, 1);
      const double c2x = __shfl_xor_sync(0xffffffffu, pw.c2, 1);
      w1[0] = h ? w1x : pw.w1, w1[1] = h ? pw.w1 : w1x;
      c2[0] = h ? c2x : pw.c2, c2[1] = h ? pw.c2 : c2x;
    }
  };
  auto accumulate = [&](uint32_t row, const double (&w1)[Q], const double (&c2)[Q]) {
#pragma unroll
    for (int e0 = 0; e0 < DL; e0 += CH) {
      double2 v[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) v[c] = lds_row(row + (e0 + c) * 16);
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int j = 0; j < Q; ++j) G[j][e0 + c] = fma(-w1[j], v[c].x, G[j][e0 + c]);
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int j = 0; j < Q; ++j) G[j][e0 + c] = fma(-c2[j], v[c].y, G[j][e0 + c]);
    }
  };
  // The loop handles RB rows per iteration so that every lane carries RB (x Q when L == 1) independent
  // scalar chains: the compiler interleaves them, and one chain latency is paid per RB rows instead of per row.
  constexpr int RB = 2;
  static_assert(TR % RB == 0, "tile rows must be a multiple of the row block");
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int s = tile & 1;
    mbar_wait(&bars[s], (tile >> 1) & 1);
    const uint32_t tb = smem_u32(tiles + s * TR * Dp) + h * (DL * 16);
    const double* ab = aEs + s * TR;
    double n2[RB][Q], aa[RB][Q];
#pragma unroll
    for (int r = 0; r < RB; ++r) dots(tb + r * (Dp * 16), n2[r], aa[r]);
#pragma unroll 1
    for (int t = 0; t < TR; t += RB) {
      const uint32_t row = tb + t * (Dp * 16);
      double w1[RB][Q], c2[RB][Q];
#pragma unroll
      for (int r = 0; r < RB; ++r) weights(n2[r], aa[r], USE_E ? ab[t + r] : 0.0, w1[r], c2[r]);  // (a) rows t..t+RB-1
      const uint32_t nxt = row + (t + RB < TR ? RB * Dp * 16 : 0);  // last block: recomputed, unused
#pragma unroll
      for (int r = 0; r < RB; ++r) dots(nxt + r * (Dp * 16), n2[r], aa[r]);                        // (b) next block
#pragma unroll
      for (int r = 0; r < RB; ++r) accumulate(row + r * (Dp * 16), w1[r], c2[r]);                  // (c) rows t..
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && tile + 2 < n_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[s], kStageBytes);
      tma_load_1d(tiles + s * TR * Dp, a.XA + (size_t)(tile + 2) * TR * Dp, kTileBytes, &bars[s]);
      if (USE_E) tma_load_1d(aEs + s * TR, a.alphaE + (size_t)(tile + 2) * TR, TR * 8, &bars[s]);
    }
  }

  // ---- epilogue: G += s1 y, un-permute, reduce over permutations, project, scale, scatter ----
  double s1q[Q];
  if constexpr (L == 1) {
#pragma unroll
    for (int j = 0; j < Q; ++j) s1q[j] = s1[j];
  } else {
    const double s1x = __shfl_xor_sync(0xffffffffu, s1[0], 1);
    s1q[0] = h ? s1x : s1[0], s1q[1] = h ? s1[0] : s1x;
  }
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    const int ql = team * Q + j;
#pragma unroll
    for (int e = 0; e < DL; ++e) {
      const int eg = h * DL + e;
      if (eg < D) Gs[ql * D + ip[j][eg]] = active[j] ? fma(s1q[j], y[j][e], G[j][e]) : 0.0;
    }
  }
  if constexpr (L == 1) {
#pragma unroll
    for (int j = 0; j < Q; ++j) Es[team * Q + j] = active[j] ? Eacc[j] : 0.0;
  } else {
    Es[team * Q + h] = (h ? active[1] : active[0]) ? Eacc[0] : 0.0;
  }
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
