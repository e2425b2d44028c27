// C ABI of mbgdml_b200 (see include/mbgdml_b200.h).  Host orchestration only: all
// prediction arithmetic runs in the kernels of enumerate.cuh / matern.cuh.  There is
// no CPU fallback: a request without a matching kernel returns MBG_ERR_UNSUPPORTED.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "common.cuh"
#include "enumerate.cuh"
#include "pairlist.cuh"
#include "hostmath.h"
#include "matern.cuh"
#include "matern_mma.cuh"
#include "matern_pw.cuh"
#include "matern_small.cuh"
#include "mbe_accum.cuh"
#include "kernel_mat.cuh"

using namespace mbg;

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(MBG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TRY(expr)          \
  do {                     \
    int _rc = (expr);      \
    if (_rc != MBG_OK) return _rc; \
  } while (0)

// ------------------------------------------------------------------------------------
// context / model objects
// ------------------------------------------------------------------------------------
struct DevBuf {  // growable device allocation
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return MBG_OK;
    if (p) cudaFree(p);
    p = nullptr, cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(MBG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return MBG_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr, cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kMaxPwLaunches = 1024;  // persistent-kernel launches per call whose counters / counts have their own slot
constexpr int kMaxJobsPerCall = 256;
constexpr int kMaxDeferred = 64;            // small jobs per call whose counts are read back once, at the end
constexpr long long kSmallJobCand = 65536;  // candidates (all frames) up to which a job is "small"

// Device tables of one (system, jobs) description: entity CSR, masses, cells, per-job entity sets / tuples / binomials.
struct Tables {
  DevBuf sys_ints, sys_masses, cells, job_ints;
  std::vector<DevBuf> job_binom;  // one per job (a captured step keeps all of them alive)
  std::vector<CellDev> h_cells;
  void release() {
    for (DevBuf* b : {&sys_ints, &sys_masses, &cells, &job_ints}) b->release();
    for (DevBuf& b : job_binom) b.release();
    job_binom.clear();
  }
};

// Scratch of the enumerate -> cull -> compact -> contract pipeline.  The context owns one set; a captured step
// (mbg_step_create) owns its own, because the graph's kernel nodes keep the addresses.
struct Scratch {
  DevBuf flags, blk_acc, blk_enum, blk_off;
  DevBuf rec_frame, rec_atoms, rec_lambda, rec_slot;
  DevBuf nl_adj, nl_deg, nl_cnt, nl_off;  // pair list of the job being culled (pairlist.cuh)
  // persistent-kernel path: per-launch device counters and counts (no host round trip inside a call)
  DevBuf pw_counters;   // [kMaxPwLaunches][kPwMaxSlices] dynamic work counters, zeroed at the start of a call
  DevBuf chunk_totals;  // [kMaxPwLaunches][2] (accepted, enumerated) of every cull chunk = n_frag_dev of its contraction
  DevBuf job_totals;    // [kMaxJobsPerCall][2] running sums per job (stats)
  int n_pw_launch = 0;
  int init();
  void release() {
    for (DevBuf* b : {&flags, &blk_acc, &blk_enum, &blk_off, &rec_frame, &rec_atoms, &rec_lambda, &rec_slot, &pw_counters,
                      &chunk_totals, &job_totals, &nl_adj, &nl_deg, &nl_cnt, &nl_off})
      b->release();
  }
};

struct mbg_context_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int sm_count = 0;
  int64_t launches = 0;
  int contraction = MBG_CONTRACT_AUTO;  // which contraction kernel mbg_predict uses
  int mma_cfg = 0;                      // developer builds: MT*1000 + threads of the trimer DMMA kernel
  int culling = MBG_CULL_AUTO;          // candidate generation of homogeneous 2-/3-body jobs
  // scratch
  Tables tab;
  Scratch scr;
  bool capturing = false;  // a step is being captured: no events, no launch records
  DevBuf R, outE, outF, outV, totals, mask_buf;
  DevBuf km_in, km_ints, km_ints2, km_jlist, km_out;  // kernel-matrix / covariance scratch (kernel_mat.cuh)
  DevBuf exp2_table;      // 2^(j/64), j = 0..63      (FMA-pipe kernel)
  DevBuf exp2_table_mma;  // 2^(j/4096), j = 0..4095  (DMMA kernel)
  std::vector<int> rec_chunk_slot;  // launch record -> chunk slot (fragment count resolved lazily), -1: count is exact
  long long* h_chunk = nullptr;     // pinned [kMaxPwLaunches][2]
  long long* h_totals = nullptr;  // pinned [2 + 2 * kMaxDeferred]
  DevBuf deferred_totals;         // [kMaxDeferred][2]: (accepted, enumerated) of jobs whose counts stay on the device
  int n_deferred = 0;
  int deferred_rec[64] = {0};     // launch record of each deferred job (its fragment count is patched at the end)
  // timing of the last call
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr;
  double ms_contract = 0, ms_other = 0;
  int64_t n_contract = 0;
  bool timing_pending = false;
  struct LaunchRec { int n_frag_atoms; long long n_frag; int n_train, n_perms; float ms; int kind; };
  std::vector<LaunchRec> launch_recs;
};

struct mbg_model_s {
  mbg_context_t ctx;
  int n_atoms, order, D, T, T_pad, P;
  double sig, c, std;
  bool use_E;
  bool use_lat = false;
  double lat[9], lat_inv[9];
  double2* XA = nullptr;     // FMA-pipe kernel: [T_pad][Dp] (X, A) interleaved
  double* tiles_mma = nullptr;  // DMMA kernel: packed tiles (matern_mma.cuh)
  double* mu = nullptr;         // DMMA kernel: centre of the training descriptors [D]
  double* alphaE = nullptr;
  int* iperm = nullptr;  // [P][D] pi_p^-1
  int* perm = nullptr;   // [P][D] pi_p
  CritDev crit;
};

int Scratch::init() {
  TRY(pw_counters.ensure((size_t)kMaxPwLaunches * kPwMaxSlices * sizeof(unsigned int)));
  TRY(chunk_totals.ensure((size_t)kMaxPwLaunches * 2 * sizeof(long long)));
  TRY(job_totals.ensure((size_t)kMaxJobsPerCall * 2 * sizeof(long long)));
  CUDA_TRY(cudaMemset(pw_counters.p, 0, (size_t)kMaxPwLaunches * kPwMaxSlices * sizeof(unsigned int)));
  return MBG_OK;
}

static inline int mbg_model_rows_padded(int n_row_blocks) { return (n_row_blocks * 8 + kTileRows - 1) / kTileRows * kTileRows; }

static cudaEvent_t next_event(mbg_context_t ctx) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev_pool.push_back(e);
  }
  return ctx->ev_pool[ctx->ev_used++];
}

// ------------------------------------------------------------------------------------
// kernel dispatch tables (one instantiation per supported fragment size)
// ------------------------------------------------------------------------------------
// fragment sizes with kernels: culling + tensor-pipe contraction for 2..18 atoms (water/methanol trimers reach 18),
// FMA-pipe contraction for 2..9
#define MBG_FOR_EACH_NA_FMA(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)
#define MBG_FOR_EACH_NA(X) MBG_FOR_EACH_NA_FMA(X) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18)

static bool na_supported(int na) {
  switch (na) {
#define X(N) case N:
    MBG_FOR_EACH_NA(X)
#undef X
    return true;
    default:
      return false;
  }
}

static int launch_cull(mbg_context_t ctx, int na, int blocks, const CullArgs& a, const PairListDev* nl = nullptr) {
  switch (na) {
#define X(N) \
  case N:    \
    if (nl) k_cull_nl<N><<<blocks, kCullThreads, 0, ctx->stream>>>(a, *nl); \
    else k_cull<N><<<blocks, kCullThreads, 0, ctx->stream>>>(a); \
    break;
    MBG_FOR_EACH_NA(X)
#undef X
    default:
      return fail(MBG_ERR_UNSUPPORTED, "no cull kernel for fragments of %d atoms", na);
  }
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return MBG_OK;
}

template <int NA, bool USE_E>
static int launch_matern_t(mbg_context_t ctx, const MaternArgs& a) {
  // largest CTA (multiple of 32 queries) whose shared memory fits
  int dev_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  int nt = kMaternThreads;
  while (nt > 32 && matern_smem_bytes<NA>(nt, a.P) > (size_t)dev_smem) nt -= 32;
  const size_t smem = matern_smem_bytes<NA>(nt, a.P);
  if (smem > (size_t)dev_smem) return fail(MBG_ERR_UNSUPPORTED, "contraction tile does not fit shared memory");
  CUDA_TRY(cudaFuncSetAttribute(k_matern<NA, USE_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nq = a.n_frag * a.P;
  const int qc = matern_queries_per_cta<NA>(nt);
  const long long grid = (nq + qc - 1) / qc;
  if (grid > 0x7fffffffLL) return fail(MBG_ERR_UNSUPPORTED, "too many fragments in one launch");
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  cudaEventRecord(e0, ctx->stream);
  k_matern<NA, USE_E><<<(unsigned)grid, nt, smem, ctx->stream>>>(a);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches++;
  ctx->n_contract++;
  ctx->launch_recs.push_back({NA, a.n_frag, a.T_pad, a.P, 0.f, MBG_CONTRACT_FMA});
  CUDA_TRY(cudaGetLastError());
  return MBG_OK;
}

template <int NA, int MT, int NTHR, int MINB, bool USE_E, bool YS = false>
static int launch_matern_mma_t(mbg_context_t ctx, const MaternMmaArgs& a_in) {
  MaternMmaArgs a = a_in;
  int dev_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  const long long nq = a.n_frag * a.P;
  // Launch shape.  Large launches: NTHR threads per CTA, one row slice.  Small launches (single MD frames, small
  // clusters, the minor models of a mixed system) would leave SMs idle and pay the full latency of a T-row loop:
  // pick the CTA size (the kernel reads its shape from blockDim/gridDim; NTHR is only the launch bound) and the
  // number of training-row slices (gridDim.y; partial sums meet in the output atomics) that minimise a simple
  // cost model -- waves x row blocks per slice x max(latency floor, resident warps per SM) + a per-wave overhead.
  auto kern = k_matern_mma<NA, MT, NTHR, MINB, USE_E, YS>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)std::min<size_t>(mma_smem_bytes<NA, MT, YS>(NTHR, a.P), (size_t)dev_smem)));
  const int n_tiles = a.T_pad / mma_tile_rows(NA), bpt = mma_tile_rows(NA) / 8;
  int nt = NTHR, slices = 1;
  {
    double best = 1e300;
    const int min_nt = ctx->mma_cfg == 2 ? NTHR : 32;  // MBG_MMA_CFG=2 (developer): always the full-size shape
    for (int cand_nt = NTHR; cand_nt >= min_nt; cand_nt -= 32) {
      const size_t sm = mma_smem_bytes<NA, MT, YS>(cand_nt, a.P);
      if (sm > (size_t)dev_smem) continue;
      static thread_local std::map<std::pair<int, size_t>, int> occ_cache;  // per instantiation: (threads, smem) -> CTAs/SM
      int occ = 0;
      auto hit = occ_cache.find({cand_nt, sm});
      if (hit != occ_cache.end()) {
        occ = hit->second;
      } else {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, cand_nt, sm) != cudaSuccess) occ = 0;
        occ_cache[{cand_nt, sm}] = occ;
      }
      if (occ < 1) continue;
      const int qc_c = mma_queries_per_cta<NA, MT>(cand_nt);
      const long long g = (nq + qc_c - 1) / qc_c;
      for (int sl = 1; sl <= (ctx->mma_cfg == 2 ? 1 : std::max(1, n_tiles / 2)); sl *= 2) {
        const int tps = (n_tiles + sl - 1) / sl;
        const int sl_eff = std::min((n_tiles + tps - 1) / tps, (a.n_row_blocks + tps * bpt - 1) / (tps * bpt));
        const long long ctas = g * sl_eff;
        const long long waves = (ctas + (long long)ctx->sm_count * occ - 1) / ((long long)ctx->sm_count * occ);
        const double resident = (double)std::min<long long>(occ, (ctas + ctx->sm_count - 1) / ctx->sm_count) * (cand_nt / 32);
        const double cost = (double)waves * ((double)std::min(a.n_row_blocks, tps * bpt) * std::max(5.0, resident) + 40.0);
        if (cost < best * 0.97) best = cost, nt = cand_nt, slices = sl_eff, a.tiles_per_slice = tps;
        if (ctas >= 4ll * ctx->sm_count * occ) break;  // more slices only add waves
      }
      if (g >= 4ll * ctx->sm_count) break;  // large launch: keep the full-size CTA
    }
    if (best == 1e300) return fail(MBG_ERR_UNSUPPORTED, "contraction tile does not fit shared memory");
  }
  const int qc = mma_queries_per_cta<NA, MT>(nt);
  const long long grid = (nq + qc - 1) / qc;
  const size_t smem = mma_smem_bytes<NA, MT, YS>(nt, a.P);
  // Tail splitting for launches of several waves whose last wave is less than half full (one MD frame of the
  // 140-H2O box: 24.07 waves): the tail batches are cut into row slices that together fill the SMs once.
  a.main_ctas = (int)std::min<long long>(grid, 0x7fffffffLL), a.tail_slices = 1, a.tail_tiles_per_slice = a.tiles_per_slice;
  long long grid_x = grid;
  if (slices == 1 && ctx->mma_cfg != 2) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem) != cudaSuccess || occ < 1) occ = 1;
    const long long per_wave = (long long)ctx->sm_count * occ, rem = grid % per_wave;
    if (grid > per_wave && rem > 0 && 2 * rem <= per_wave) {
      int ts = (int)std::min<long long>(per_wave / rem, std::max(1, n_tiles / 2));
      const int tps = (n_tiles + ts - 1) / ts;
      ts = std::min((n_tiles + tps - 1) / tps, (a.n_row_blocks + tps * bpt - 1) / (tps * bpt));  // slices that hold data
      if (ts > 1) a.main_ctas = (int)(grid - rem), a.tail_slices = ts, a.tail_tiles_per_slice = tps, grid_x = grid - rem + rem * ts;
    }
  }
  if (grid_x > 0x7fffffffLL) return fail(MBG_ERR_UNSUPPORTED, "too many fragments in one launch");
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  cudaEventRecord(e0, ctx->stream);
  kern<<<dim3((unsigned)grid_x, (unsigned)slices), nt, smem, ctx->stream>>>(a);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches++;
  ctx->n_contract++;
  ctx->launch_recs.push_back({NA, a.n_frag, a.T_pad, a.P, 0.f, MBG_CONTRACT_DMMA});
  CUDA_TRY(cudaGetLastError());
  return MBG_OK;
}

static int launch_matern_mma(mbg_context_t ctx, int na, bool use_E, const MaternMmaArgs& a) {
#ifdef MBG_TUNING_VARIANTS  // developer build: extra shapes of the trimer kernel, MBG_MMA_CFG = MT*10000 + threads*10 + CTAs/SM
  if (na == 9 && !use_E) {
    switch (ctx->mma_cfg) {
#define V(MT, NT, MB) case MT * 10000 + NT * 10 + MB: return launch_matern_mma_t<9, MT, NT, MB, false>(ctx, a);
      V(2, 384, 1) V(2, 512, 1) V(4, 256, 1) V(3, 256, 1) V(2, 256, 1)
      V(3, 128, 2) V(2, 128, 3) V(4, 128, 2) V(2, 128, 2) V(2, 192, 2) V(1, 256, 2) V(1, 128, 4) V(1, 128, 5)
#undef V
      default: break;
    }
  }
#endif
  // Shape per fragment size (measured on B200, profiles/r1_dmma_tuning.md): as many m-tiles per warp as the
  // register file allows without spilling (3 for D = 28, 36; 4 below).  Trimers of 9 atoms (the headline shape)
  // run 12 warps per SM with the GEMM-1 query operands in shared memory (168 registers): three warps per
  // scheduler keep the FP64 pipe fed while one of them sits in its sqrt/exp chain.
  // (models with alphas_E need more registers per thread and keep the 8-warp shape)
  // Measured and rejected for the other sizes: 12 warps + smem operands for 6-atom fragments (0.95 vs 0.85 ms on
  // the box140 dimers), two 8-warp CTAs per SM for 3-atom fragments (0.83 vs 0.80 ms on 100 000 monomers).
  if (na == 9 && !use_E && ctx->mma_cfg != 1) return launch_matern_mma_t<9, 3, 384, 1, false, true>(ctx, a);
  switch (na) {
#define X(N)                                                                                        \
  case N: {                                                                                         \
    constexpr int MT = (N >= 13) ? 1 : (N >= 10) ? 2 : (N >= 8) ? 3 : 4;                            \
    return use_E ? launch_matern_mma_t<N, MT, 256, 1, true>(ctx, a) : launch_matern_mma_t<N, MT, 256, 1, false>(ctx, a); \
  }
    MBG_FOR_EACH_NA(X)
#undef X
    default:
      return fail(MBG_ERR_UNSUPPORTED, "no contraction kernel for fragments of %d atoms", na);
  }
}


constexpr long long kCtaMaxCandidates = 1 << 20;  // run_job_pw: 9- to 12-atom launches up to this many candidates use the CTA kernel
constexpr long long kSmallMinQueries = 48 * 1024;  // k_matern_small from this many queries (fragments x permutations) on

// Persistent per-warp form of the tensor-pipe contraction (matern_pw.cuh): always one CTA per SM; the kernel reads
// the fragment count on the device, so the launch shape never depends on it.
template <int NA, int MT, int NW, bool USE_E, bool YS>
static int launch_matern_pw_t(mbg_context_t ctx, Scratch& sc, const MaternPwArgs& a_in, int chunk_slot) {
  MaternPwArgs a = a_in;
  int dev_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  int nw = NW;  // fewer warps when the per-warp regions (many fragments per item: few permutations) do not fit
  while (nw > 1 && pw_smem_bytes<NA, MT, NW, YS>(a.P, nw) > (size_t)dev_smem) nw -= (nw > 4 ? 4 : 1);
  const size_t smem = pw_smem_bytes<NA, MT, NW, YS>(a.P, nw);
  if (smem > (size_t)dev_smem) return fail(MBG_ERR_UNSUPPORTED, "contraction tile does not fit shared memory");
  auto kern = k_matern_pw<NA, MT, NW, USE_E, YS>;
  static thread_local size_t attr_smem = 0;  // per instantiation
  if (smem > attr_smem) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int n_tiles_data = (a.n_row_blocks + mma_tile_rows(NA) / 8 - 1) / (mma_tile_rows(NA) / 8);
  a.max_slices = 1;
  while (a.max_slices * 2 <= kPwMaxSlices && a.max_slices * 4 <= n_tiles_data) a.max_slices *= 2;  // >= 2 tiles per slice
  if (ctx->mma_cfg == 2) a.max_slices = 1;  // developer: never slice
  a.work_counters = sc.pw_counters.as<unsigned int>() + (size_t)(sc.n_pw_launch % kMaxPwLaunches) * kPwMaxSlices;
  if (sc.n_pw_launch >= kMaxPwLaunches)  // slots wrap around: re-zero this one (stream-ordered after its last user)
    CUDA_TRY(cudaMemsetAsync(a.work_counters, 0, kPwMaxSlices * sizeof(unsigned int), ctx->stream));
  sc.n_pw_launch++;
  if (ctx->capturing) {  // inside a captured step: no timing events, no records
    kern<<<ctx->sm_count, nw * 32, smem, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return MBG_OK;
  }
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  cudaEventRecord(e0, ctx->stream);
  kern<<<ctx->sm_count, nw * 32, smem, ctx->stream>>>(a);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches++;
  ctx->n_contract++;
  ctx->launch_recs.push_back({NA, a.n_frag, (mbg_model_rows_padded(a.n_row_blocks)), a.P, 0.f, MBG_CONTRACT_DMMA});
  ctx->rec_chunk_slot.resize(ctx->launch_recs.size(), -1);
  ctx->rec_chunk_slot.back() = chunk_slot;
  CUDA_TRY(cudaGetLastError());
  return MBG_OK;
}

static int launch_matern_pw(mbg_context_t ctx, Scratch& sc, int na, bool use_E, const MaternPwArgs& a, int chunk_slot) {
  // Shapes as for the wave-launched kernel (profiles/r1_dmma_tuning.md): 9-atom fragments run 12 warps with the
  // GEMM-1 query operands in shared memory, everything else 8 warps with the operands in registers.
  if (na == 9 && !use_E && ctx->mma_cfg != 1) return launch_matern_pw_t<9, 3, 12, false, true>(ctx, sc, a, chunk_slot);
  switch (na) {
#define X(N)                                                                                        \
  case N: {                                                                                         \
    constexpr int MT = (N >= 13) ? 1 : (N >= 10) ? 2 : (N >= 8) ? 3 : 4;                            \
    return use_E ? launch_matern_pw_t<N, MT, 8, true, false>(ctx, sc, a, chunk_slot)                    \
                 : launch_matern_pw_t<N, MT, 8, false, false>(ctx, sc, a, chunk_slot);                  \
  }
    MBG_FOR_EACH_NA(X)
#undef X
    default:
      return fail(MBG_ERR_UNSUPPORTED, "no contraction kernel for fragments of %d atoms", na);
  }
}

// Wide fragments (13-18 atoms: one m-tile per warp) behind the sync-free API: the CTA-cooperative kernel in its
// persistent mode (matern_mma.cuh).  With one m-tile per warp and two warps per scheduler a per-warp worker cannot
// hide its own prologue / epilogue (profiles/r2_pw_tuning.md section 5: k_matern_pw runs 5-8 % behind here), a CTA
// that does them with all its threads can.
template <int NA, int MT, int NTHR, bool USE_E, bool YS>
static int launch_matern_mma_persistent_t(mbg_context_t ctx, Scratch& sc, const MaternMmaArgs& a_in, int chunk_slot) {
  MaternMmaArgs a = a_in;
  int dev_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  auto kern = k_matern_mma<NA, MT, NTHR, 1, USE_E, YS>;
  const size_t smem = mma_smem_bytes<NA, MT, YS>(NTHR, a.P);
  if (smem > (size_t)dev_smem) return fail(MBG_ERR_UNSUPPORTED, "contraction tile does not fit shared memory");
  static thread_local size_t attr_smem = 0;  // per instantiation
  static thread_local int occ = 0;
  if (smem > attr_smem) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHR, smem) != cudaSuccess || occ < 1) occ = 1;
  }
  const int n_tiles_data = (a.n_row_blocks + mma_tile_rows(NA) / 8 - 1) / (mma_tile_rows(NA) / 8);
  a.persistent = 1, a.max_slices = 1;
  while (a.max_slices * 2 <= 32 && a.max_slices * 4 <= n_tiles_data) a.max_slices *= 2;  // >= 2 tiles per slice
  a.tiles_per_slice = a.T_pad / mma_tile_rows(NA), a.main_ctas = 0x7fffffff, a.tail_slices = 1, a.tail_tiles_per_slice = a.tiles_per_slice;
  sc.n_pw_launch++;  // keeps the chunk slot of this launch its own
  const unsigned grid = (unsigned)(ctx->sm_count * occ);
  if (ctx->capturing) {  // inside a captured step: no timing events, no records
    kern<<<grid, NTHR, smem, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return MBG_OK;
  }
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  cudaEventRecord(e0, ctx->stream);
  kern<<<grid, NTHR, smem, ctx->stream>>>(a);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches++;
  ctx->n_contract++;
  ctx->launch_recs.push_back({NA, a.n_frag, a.T_pad, a.P, 0.f, MBG_CONTRACT_DMMA});
  ctx->rec_chunk_slot.resize(ctx->launch_recs.size(), -1);
  ctx->rec_chunk_slot.back() = chunk_slot;
  CUDA_TRY(cudaGetLastError());
  return MBG_OK;
}

static int launch_matern_mma_persistent(mbg_context_t ctx, Scratch& sc, int na, bool use_E, const MaternMmaArgs& a, int chunk_slot) {
  // the shapes of the wave-launched kernel (launch_matern_mma): 12 warps with y' in shared memory for 9 atoms
  if (na == 9 && !use_E) return launch_matern_mma_persistent_t<9, 3, 384, false, true>(ctx, sc, a, chunk_slot);
  switch (na) {
#define X(N)                                                                                             \
  case N: {                                                                                              \
    constexpr int MT = (N >= 13) ? 1 : (N >= 10) ? 2 : 3;                                                \
    return use_E ? launch_matern_mma_persistent_t<N, MT, 256, true, false>(ctx, sc, a, chunk_slot)       \
                 : launch_matern_mma_persistent_t<N, MT, 256, false, false>(ctx, sc, a, chunk_slot);     \
  }
    X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18)
#undef X
    default:
      return fail(MBG_ERR_UNSUPPORTED, "no persistent CTA kernel for fragments of %d atoms", na);
  }
}

// Thread-per-query FMA-pipe kernel for narrow descriptors (matern_small.cuh): fragments of 2-4 atoms.
template <int NA, bool USE_E>
static int launch_matern_small_t(mbg_context_t ctx, const MaternSmallArgs& a_in, int chunk_slot) {
  MaternSmallArgs a = a_in;
  int dev_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  a.rows_per_pass = small_rows_per_pass<NA>(a.T, a.P, (size_t)dev_smem);
  if (a.rows_per_pass < 1) return fail(MBG_ERR_UNSUPPORTED, "model does not fit the narrow-descriptor kernel");
  constexpr int D = NA * (NA - 1) / 2;
  const size_t smem = small_fixed_bytes<NA>(a.P) + (size_t)a.rows_per_pass * (D * 16 + 8);
  auto kern = k_matern_small<NA, USE_E>;
  static thread_local size_t attr_smem = 0;
  if (smem > attr_smem) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const long long nq_max = a.n_frag * a.P;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ctx->sm_count, (nq_max + kSmallThreads - 1) / kSmallThreads));
  if (ctx->capturing) {
    kern<<<grid, kSmallThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return MBG_OK;
  }
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  cudaEventRecord(e0, ctx->stream);
  kern<<<grid, kSmallThreads, smem, ctx->stream>>>(a);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches++;
  ctx->n_contract++;
  ctx->launch_recs.push_back({NA, a.n_frag, (a.T + kTileRows - 1) / kTileRows * kTileRows, a.P, 0.f, MBG_CONTRACT_DMMA});
  ctx->rec_chunk_slot.resize(ctx->launch_recs.size(), -1);
  ctx->rec_chunk_slot.back() = chunk_slot;
  CUDA_TRY(cudaGetLastError());
  return MBG_OK;
}

static int launch_matern_small(mbg_context_t ctx, int na, bool use_E, const MaternSmallArgs& a, int chunk_slot) {
  switch (na) {
#define X(N) \
  case N:    \
    return use_E ? launch_matern_small_t<N, true>(ctx, a, chunk_slot) : launch_matern_small_t<N, false>(ctx, a, chunk_slot);
    X(2) X(3) X(4)
#undef X
    default:
      return fail(MBG_ERR_UNSUPPORTED, "no narrow-descriptor kernel for fragments of %d atoms", na);
  }
}

static int launch_matern(mbg_context_t ctx, int na, bool use_E, const MaternArgs& a) {
  switch (na) {
#define X(N) \
  case N:    \
    return use_E ? launch_matern_t<N, true>(ctx, a) : launch_matern_t<N, false>(ctx, a);
    MBG_FOR_EACH_NA_FMA(X)
#undef X
    default:
      return fail(MBG_ERR_UNSUPPORTED, "no FMA-pipe contraction kernel for fragments of %d atoms (use MBG_CONTRACT_DMMA / AUTO)", na);
  }
}

// ------------------------------------------------------------------------------------
// small kernels used by the API layer
// ------------------------------------------------------------------------------------
__global__ void k_fill(double* p, long long n, double v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// zero the decomposed output rows of accepted fragments (the rest stay NaN)
__global__ void k_zero_rows(const long long* rec_slot, long long n_rec, int row_len, double* E, double* F,
                            const long long* n_rec_dev = nullptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_rec_dev) n_rec = min(n_rec, *n_rec_dev);
  if (i >= n_rec * (row_len + 1)) return;
  const long long r = i / (row_len + 1);
  const int k = (int)(i - r * (row_len + 1));
  if (k == row_len)
    E[rec_slot[r]] = 0.0;
  else
    F[rec_slot[r] * row_len + k] = 0.0;
}

template <int CHAINS>
__global__ void k_dfma_peak(double* out, int iters, double seed) {
  double acc[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) acc[k] = seed + k + threadIdx.x * 1e-3;
  const double m = 1.0 + seed * 1e-9, b = seed * 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) acc[k] = fma(acc[k], m, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;  // never true; keeps the loop alive
}

template <int CHAINS>
__global__ void k_dmma_peak(double* out, int iters, double seed) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) c0[k] = seed + k, c1[k] = seed - k;
  const double av = 1.0 + threadIdx.x * 1e-6, bv = 1.0 - threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) dmma884_v(c0[k], c1[k], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += c0[k] + c1[k];
  if (s == 123.456) out[0] = s;  // never true; keeps the loop alive
}

// exp(-c sqrt(x)) with the CUDA math library, 4 independent chains per thread: the "exp . sqrt rate" SURVEY 8(d) asks
// the descriptor + kernel microbenchmark (config 2, D = 3: transcendental-bound) to be reported against
__global__ void k_exp_sqrt_rate(double* out, int iters, double seed) {
  double x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = seed * 0.1 + k * 0.01 + threadIdx.x * 1e-4;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = exp(-2.23606797749979 * sqrt(x[k]) * 0.01) + 0.25;
  }
  const double s = x[0] + x[1] + x[2] + x[3];
  if (s == 123.456) out[0] = s;  // never true; keeps the loop alive
}

// generic (runtime-n) geometry kernels for the small API entry points
__global__ void k_criteria_eval(int kind, int n, const double* masses, const int* group, int n_groups,
                                const double* R, long long n_R, double* out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_R) return;
  const double* r = R + s * 3ll * n;
  if (kind == MBG_CRIT_MAX_ATOM_PAIR_DIST) {
    double mx = 0.0;
    for (int i = 0; i < n; ++i)
      for (int k = i + 1; k < n; ++k) {
        const double d = norm3_rn(__dsub_rn(r[3 * i], r[3 * k]), __dsub_rn(r[3 * i + 1], r[3 * k + 1]),
                                  __dsub_rn(r[3 * i + 2], r[3 * k + 2]));
        mx = d > mx ? d : mx;
      }
    out[s] = mx;
    return;
  }
  double num[3] = {0, 0, 0}, den = 0;
  for (int a = 0; a < n; ++a) {
    for (int k = 0; k < 3; ++k) num[k] = __dadd_rn(num[k], __dmul_rn(r[3 * a + k], masses[a]));
    den = __dadd_rn(den, masses[a]);
  }
  double dsum = 0;
  for (int g = 0; g < n_groups; ++g) {
    double gn[3] = {0, 0, 0}, gd = 0;
    for (int a = 0; a < n; ++a)
      if (group[a] == g) {
        for (int k = 0; k < 3; ++k) gn[k] = __dadd_rn(gn[k], __dmul_rn(r[3 * a + k], masses[a]));
        gd = __dadd_rn(gd, masses[a]);
      }
    dsum = __dadd_rn(dsum, norm3_rn(__dsub_rn(__ddiv_rn(num[0], den), __ddiv_rn(gn[0], gd)),
                                    __dsub_rn(__ddiv_rn(num[1], den), __ddiv_rn(gn[1], gd)),
                                    __dsub_rn(__ddiv_rn(num[2], den), __ddiv_rn(gn[2], gd))));
  }
  out[s] = dsum;
}

// _from_r (_gdml/desc.py:203-233): one thread per (structure, atom pair)
__global__ void k_from_r(int n, int use_lat, CellDev lat /* .cell = lattice, .cell_inv = inverse */, const double* R,
                         long long n_R, double* x, double* J) {
  const int D = n * (n - 1) / 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_R * D) return;
  const long long s = t / D;
  const int k = (int)(t - s * D);
  int i = 1;
  while ((i + 1) * i / 2 <= k) ++i;
  const int j = k - i * (i - 1) / 2;
  const double* ri = R + (s * n + i) * 3;
  const double* rj = R + (s * n + j) * 3;
  double d[3] = {__dsub_rn(ri[0], rj[0]), __dsub_rn(ri[1], rj[1]), __dsub_rn(ri[2], rj[2])};
  if (use_lat) pbc_diff(lat.cell, lat.cell_inv, d);
  const double dist = norm3_rn(d[0], d[1], d[2]);
  x[t] = __ddiv_rn(1.0, dist);
  const double d3 = dist * dist * dist;
  J[t * 3] = d[0] / d3, J[t * 3 + 1] = d[1] / d3, J[t * 3 + 2] = d[2] / d3;
}

// Cell.d_mic (periodic.py:61-84): minimum image of n displacement vectors; the cut-off test runs over the pair
// distances of the vectors themselves (pdist(d_periodic)), not against an origin atom.
__global__ void k_d_mic(CellDev s, const double* d_in, int n, int check_cutoff, double* out, int* accepted) {
  if (threadIdx.x || blockIdx.x) return;
  bool naive_ok = true;
  for (int a = 0; a < n; ++a) {
    const double d[3] = {d_in[3 * a], d_in[3 * a + 1], d_in[3 * a + 2]};
    double v[3];
    naive_ok = (mic_naive(s, d, v) < s.half_min_len) && naive_ok;
    out[3 * a] = v[0], out[3 * a + 1] = v[1], out[3 * a + 2] = v[2];
  }
  if (!naive_ok)
    for (int a = 0; a < n; ++a) {
      const double d[3] = {d_in[3 * a], d_in[3 * a + 1], d_in[3 * a + 2]};
      double v[3];
      mic_general(s, d, v);
      out[3 * a] = v[0], out[3 * a + 1] = v[1], out[3 * a + 2] = v[2];
    }
  int ok = 1;
  if (check_cutoff)
    for (int i = 0; i < n; ++i)
      for (int k = i + 1; k < n; ++k)
        if (norm3_rn(__dsub_rn(out[3 * i], out[3 * k]), __dsub_rn(out[3 * i + 1], out[3 * k + 1]),
                     __dsub_rn(out[3 * i + 2], out[3 * k + 2])) >= s.cutoff)
          ok = 0;
  *accepted = ok;
}

__global__ void k_r_mic(CellDev s, const double* r, int n, double* out, int* accepted) {
  if (threadIdx.x || blockIdx.x) return;
  bool naive_ok = true;
  for (int a = 0; a < n; ++a) {
    const double d[3] = {__dsub_rn(r[3 * a], r[0]), __dsub_rn(r[3 * a + 1], r[1]), __dsub_rn(r[3 * a + 2], r[2])};
    double v[3];
    naive_ok = (mic_naive(s, d, v) < s.half_min_len) && naive_ok;
    out[3 * a] = v[0], out[3 * a + 1] = v[1], out[3 * a + 2] = v[2];
  }
  if (!naive_ok)
    for (int a = 0; a < n; ++a) {
      const double d[3] = {__dsub_rn(r[3 * a], r[0]), __dsub_rn(r[3 * a + 1], r[1]), __dsub_rn(r[3 * a + 2], r[2])};
      double v[3];
      mic_general(s, d, v);
      out[3 * a] = v[0], out[3 * a + 1] = v[1], out[3 * a + 2] = v[2];
    }
  int ok = 1;
  for (int i = 0; i < n; ++i)
    for (int k = i + 1; k < n; ++k)
      if (norm3_rn(__dsub_rn(out[3 * i], out[3 * k]), __dsub_rn(out[3 * i + 1], out[3 * k + 1]),
                   __dsub_rn(out[3 * i + 2], out[3 * k + 2])) >= s.cutoff)
        ok = 0;
  *accepted = ok;
}

// ------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------
static int fill_cell(CellDev& s, const double* cell, double cutoff, const int32_t* pbc = nullptr) {
  memset(&s, 0, sizeof(s));
  s.cutoff = cutoff;
  for (int i = 0; i < 3; ++i) s.pbc[i] = (!pbc || pbc[i]) ? 1 : 0;
  const bool full = s.pbc[0] && s.pbc[1] && s.pbc[2];
  memcpy(s.cell, cell, 9 * sizeof(double));
  if (!hostmath::inverse3(s.cell, s.cell_inv)) return fail(MBG_ERR_ARG, "cell is singular");
  double mn = std::numeric_limits<double>::infinity();
  for (int i = 0; i < 3; ++i) {
    const double* v = cell + 3 * i;
    mn = std::min(mn, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
  }
  // find_mic tries the naive image only when all three directions are periodic (dim == 3)
  s.half_min_len = full ? 0.5 * mn : -1.0;
  if (!hostmath::minkowski_reduce_pbc(s.cell, s.pbc, s.rcell)) return fail(MBG_ERR_ARG, "Minkowski reduction of the cell failed");
  if (!hostmath::inverse3(s.rcell, s.rcell_inv)) return fail(MBG_ERR_ARG, "reduced cell is singular");
  const bool ortho = cell[1] == 0 && cell[2] == 0 && cell[3] == 0 && cell[5] == 0 && cell[6] == 0 && cell[7] == 0;
  s.fast_reject = (full && ortho && cutoff <= s.half_min_len) ? 1 : 0;
  s.ortho = (full && ortho) ? 1 : 0;
  return MBG_OK;
}

// Cell table of a call: one CellDev for all frames (sys->cell) or one per frame (sys->frame_cells).
static int build_cells(const mbg_system_desc* sys, int n_frames, std::vector<CellDev>& cells, int* stride) {
  cells.clear();
  *stride = 0;
  if (sys->frame_cells) {
    if (!sys->frame_cutoffs) return fail(MBG_ERR_ARG, "frame_cells without frame_cutoffs");
    cells.resize(std::max(n_frames, 0));
    for (int f = 0; f < n_frames; ++f) TRY(fill_cell(cells[f], sys->frame_cells + 9 * (size_t)f, sys->frame_cutoffs[f], sys->pbc));
    *stride = 1;
  } else if (sys->cell) {
    cells.resize(1);
    TRY(fill_cell(cells[0], sys->cell, sys->cell_cutoff, sys->pbc));
  }
  return MBG_OK;
}

static int upload_system(mbg_context_t ctx, Tables& tb, const mbg_system_desc* sys, int n_frames, SysDev& s) {
  if (!sys || sys->n_atoms <= 0 || sys->n_entities <= 0 || !sys->entity_offsets || !sys->entity_atoms)
    return fail(MBG_ERR_ARG, "invalid system description");
  if (sys->entity_offsets[0] != 0 || sys->entity_offsets[sys->n_entities] > sys->n_atoms)
    return fail(MBG_ERR_ARG, "entity_offsets must start at 0 and end at <= n_atoms");
  const int n_vals = sys->entity_offsets[sys->n_entities];
  for (int e = 0; e < sys->n_entities; ++e)
    if (sys->entity_offsets[e + 1] < sys->entity_offsets[e]) return fail(MBG_ERR_ARG, "entity_offsets not monotone");
  for (int k = 0; k < n_vals; ++k)
    if (sys->entity_atoms[k] < 0 || sys->entity_atoms[k] >= sys->n_atoms)
      return fail(MBG_ERR_ARG, "entity_atoms[%d] out of range", k);
  const size_t n_int = (size_t)sys->n_entities + 1 + n_vals;
  TRY(tb.sys_ints.ensure(n_int * sizeof(int)));
  TRY(tb.sys_masses.ensure((size_t)sys->n_atoms * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(tb.sys_ints.p, sys->entity_offsets, ((size_t)sys->n_entities + 1) * sizeof(int),
                           cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(tb.sys_ints.as<int>() + sys->n_entities + 1, sys->entity_atoms, (size_t)n_vals * sizeof(int),
                           cudaMemcpyHostToDevice, ctx->stream));
  std::vector<double> ones;
  const double* masses = sys->masses;
  if (!masses) {
    ones.assign(sys->n_atoms, 1.0);
    masses = ones.data();
  }
  CUDA_TRY(cudaMemcpyAsync(tb.sys_masses.p, masses, (size_t)sys->n_atoms * sizeof(double), cudaMemcpyHostToDevice,
                           ctx->stream));
  if (!sys->masses) CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // `ones` goes out of scope
  s.n_atoms = sys->n_atoms;
  s.n_entities = sys->n_entities;
  s.ent_off = tb.sys_ints.as<int>();
  s.ent_atoms = tb.sys_ints.as<int>() + sys->n_entities + 1;
  s.masses = tb.sys_masses.as<double>();
  std::vector<CellDev>& cells = tb.h_cells;  // stays alive until the next call
  TRY(build_cells(sys, n_frames, cells, &s.cell_stride));
  s.periodic = cells.empty() ? 0 : 1;
  s.cells = nullptr;
  if (!cells.empty()) {
    TRY(tb.cells.ensure(cells.size() * sizeof(CellDev)));
    CUDA_TRY(cudaMemcpyAsync(tb.cells.p, cells.data(), cells.size() * sizeof(CellDev), cudaMemcpyHostToDevice, ctx->stream));
    s.cells = tb.cells.as<CellDev>();
  }
  return MBG_OK;
}

// Validate a job against the system and build its device description.
static int build_job(mbg_context_t ctx, Tables& tb, const mbg_system_desc* sys, const mbg_job* job, size_t ints_offset,
                     JobDev& j, size_t* ints_used, int job_index = 0) {
  if (!job || !job->model) return fail(MBG_ERR_ARG, "job without model");
  const mbg_model_t m = job->model;
  if (m->ctx != ctx) return fail(MBG_ERR_ARG, "model belongs to another context");
  memset(&j, 0, sizeof(j));
  j.order = m->order;
  j.n_frag_atoms = m->n_atoms;
  auto ent_size = [&](int e) { return sys->entity_offsets[e + 1] - sys->entity_offsets[e]; };
  int* d_ints = tb.job_ints.as<int>() + ints_offset;
  if (job->combs) {
    if (job->n_combs < 0) return fail(MBG_ERR_ARG, "n_combs < 0");
    j.explicit_combs = 1;
    j.cand_per_frame = job->n_combs;
    for (long long c = 0; c < job->n_combs; ++c) {
      int na = 0;
      for (int s = 0; s < m->order; ++s) {
        const int e = job->combs[c * m->order + s];
        if (e < 0 || e >= sys->n_entities) return fail(MBG_ERR_ARG, "comb %lld: entity id %d out of range", c, e);
        na += ent_size(e);
      }
      if (na != m->n_atoms)
        return fail(MBG_ERR_ARG, "comb %lld has %d atoms but the model expects %d", c, na, m->n_atoms);
    }
    const size_t n = (size_t)job->n_combs * m->order;
    TRY(tb.job_ints.ensure((ints_offset + n + 16) * sizeof(int)));
    d_ints = tb.job_ints.as<int>() + ints_offset;
    if (n) CUDA_TRY(cudaMemcpyAsync(d_ints, job->combs, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    j.combs = d_ints;
    *ints_used = n;
  } else {
    if (!job->set_offsets || !job->set_entities) return fail(MBG_ERR_ARG, "job has neither sets nor combs");
    if (job->set_offsets[0] != 0) return fail(MBG_ERR_ARG, "set_offsets[0] must be 0");
    long long prod = 1;
    int na = 0;
    j.sets_ascending = 1;
    for (int s = 0; s < m->order; ++s) {
      const int lo = job->set_offsets[s], hi = job->set_offsets[s + 1];
      if (hi < lo) return fail(MBG_ERR_ARG, "set_offsets not monotone");
      j.set_off[s] = lo, j.set_off[s + 1] = hi;
      int slot_atoms = -1;
      for (int k = lo; k < hi; ++k) {
        const int e = job->set_entities[k];
        if (k > lo && e <= job->set_entities[k - 1]) j.sets_ascending = 0;
        if (e < 0 || e >= sys->n_entities) return fail(MBG_ERR_ARG, "set entity id %d out of range", e);
        if (slot_atoms < 0) slot_atoms = ent_size(e);
        if (ent_size(e) != slot_atoms)
          return fail(MBG_ERR_ARG, "entities of slot %d have different atom counts", s);
      }
      na += slot_atoms < 0 ? 0 : slot_atoms;
      j.slot_atoms[s] = slot_atoms < 0 ? 0 : slot_atoms;
      const long long n_s = hi - lo;
      if (n_s == 0) prod = 0;
      else if (prod > (1ll << 60) / n_s) return fail(MBG_ERR_UNSUPPORTED, "candidate space overflows");
      else prod *= n_s;
    }
    if (prod > 0 && na != m->n_atoms)
      return fail(MBG_ERR_ARG, "fragments have %d atoms but the model expects %d", na, m->n_atoms);
    j.cand_per_frame = prod;
    const size_t n = (size_t)job->set_offsets[m->order];
    TRY(tb.job_ints.ensure((ints_offset + n + 16) * sizeof(int)));
    d_ints = tb.job_ints.as<int>() + ints_offset;
    if (n) CUDA_TRY(cudaMemcpyAsync(d_ints, job->set_entities, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    j.set_ent = d_ints;
    *ints_used = n;
    // Homogeneous sets (e.g. [h2o, h2o, h2o]): the ascending tuples of the product are exactly the C(n, order)
    // combinations in lexicographic order, so the device unranks them directly instead of filtering n^order tuples.
    if (m->order >= 2 && prod > 0) {
      const int n0 = job->set_offsets[1] - job->set_offsets[0];
      bool homo = true;
      for (int s = 0; s < m->order && homo; ++s) {
        if (job->set_offsets[s + 1] - job->set_offsets[s] != n0) homo = false;
        for (int k = 0; k < n0 && homo; ++k) {
          const int e = job->set_entities[job->set_offsets[s] + k];
          if (e != job->set_entities[k] || (k > 0 && e <= job->set_entities[k - 1])) homo = false;
        }
      }
      const size_t w = (size_t)m->order + 1, tbl = ((size_t)n0 + 1) * w;
      if (homo && tbl * sizeof(long long) <= (64u << 20)) {
        std::vector<long long> binom(tbl, 0);
        const long long kSat = 1ll << 61;
        for (int a = 0; a <= n0; ++a) {
          binom[(size_t)a * w] = 1;
          for (int r = 1; r <= m->order; ++r) {
            const long long up = a ? binom[(size_t)(a - 1) * w + r] : 0, ul = a ? binom[(size_t)(a - 1) * w + r - 1] : 0;
            binom[(size_t)a * w + r] = std::min(kSat, up + ul);
          }
        }
        const long long n_comb = binom[(size_t)n0 * w + m->order];
        if (n_comb < kSat) {
          if ((int)tb.job_binom.size() <= job_index) tb.job_binom.resize(job_index + 1);
          DevBuf& bb = tb.job_binom[job_index];
          TRY(bb.ensure(tbl * sizeof(long long)));
          CUDA_TRY(cudaMemcpyAsync(bb.p, binom.data(), tbl * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
          j.homogeneous = 1;
          j.n_set = n0;
          j.binom = bb.as<long long>();
          j.cand_per_frame = n_comb;
        }
      }
    }
  }
  if (job->n_alchemy < 0 || job->n_alchemy > kMaxAlchemy)
    return fail(MBG_ERR_UNSUPPORTED, "at most %d alchemical scalers per model", kMaxAlchemy);
  j.n_alch = job->n_alchemy;
  for (int a = 0; a < job->n_alchemy; ++a) {
    const double f = job->alchemy_factor[a];
    if (!(f >= 0.0 && f <= 1.0)) return fail(MBG_ERR_ARG, "alchemical factor must be in [0, 1] (switching.py:47)");
    j.alch_ent[a] = job->alchemy_entity[a];
    j.alch_fac[a] = f;
  }
  return MBG_OK;
}

struct Outputs {
  double* E;
  double* F;
  double* V;
};


// Shard geometry (enumerate.cuh: global_candidate): granules of kShardGranule candidates dealt round-robin to the ranks;
// a rank's own space is its granules back to back, processed in blocks of kCullThreads.
static long long shard_local_granules(long long n_cand_total, int shard_rank, int shard_count) {
  const long long n_g = (n_cand_total + kShardGranule - 1) / kShardGranule;
  return n_g > shard_rank ? (n_g - shard_rank + shard_count - 1) / shard_count : 0;
}
static long long shard_blocks(long long n_cand_total, int shard_rank, int shard_count) {
  constexpr int gpb = kCullThreads / kShardGranule;
  return (shard_local_granules(n_cand_total, shard_rank, shard_count) + gpb - 1) / gpb;
}
// Candidates (< n_cand_total) covered by blocks [b0, b0 + blocks) of one shard: every granule is full except the last
// granule of the whole space.
static long long shard_candidates(long long b0, long long blocks, int shard_rank, int shard_count, long long n_cand_total) {
  constexpr int gpb = kCullThreads / kShardGranule;
  const long long n_loc = shard_local_granules(n_cand_total, shard_rank, shard_count);
  const long long lo = std::min(b0 * gpb, n_loc), hi = std::min((b0 + blocks) * gpb, n_loc);
  if (hi <= lo) return 0;
  long long n = (hi - lo) * kShardGranule;
  const long long n_g = (n_cand_total + kShardGranule - 1) / kShardGranule;
  if ((n_g - 1) % shard_count == shard_rank) {  // the partial last granule belongs to this shard
    const long long loc = (n_g - 1 - shard_rank) / shard_count;
    if (loc >= lo && loc < hi) n -= n_g * kShardGranule - n_cand_total;
  }
  return n;
}

// Pair-list candidate generation (pairlist.cuh): does it apply to this job, and with which pair condition?
constexpr long long kPairListMinCand = 1ll << 21;  // MBG_CULL_AUTO: candidates per job (all frames) from which it is used
static bool pair_list_applies(mbg_context_t ctx, const SysDev& s, const JobDev& j, const mbg_model_t m, int n_frames,
                              long long n_cand_total, int* mode, double* thr) {
  if (ctx->culling == MBG_CULL_FULL) return false;
  // 2- and 3-body jobs over entity sets; homogeneous sets up to 6-body (C(deg, order - 1) stays far below 2^61)
  if (j.explicit_combs || j.order < 2 || (j.order > 3 && !(j.homogeneous && j.order <= 6 && j.n_set <= 4096))) return false;
  if (ctx->culling == MBG_CULL_AUTO && n_cand_total < kPairListMinCand) return false;
  const int n0 = j.homogeneous ? j.n_set : j.set_off[1] - j.set_off[0];
  if ((long long)n_frames * n0 >= (1ll << 30)) return false;
  if (!j.homogeneous) {  // a row's reduced candidates are addressed with 32-bit neighbour indices
    if (!j.sets_ascending) return false;  // k_count_ascending searches the sets
    for (int s = 1; s < j.order; ++s)
      if (j.set_off[s + 1] - j.set_off[s] > (1 << 24)) return false;
  }
  const CritDev& c = m->crit;
  const bool upper = c.kind != MBG_CRIT_NONE && (c.bound == MBG_BOUND_UPPER || c.bound == MBG_BOUND_RANGE);
  *thr = std::numeric_limits<double>::infinity();
  if (s.periodic) {
    *mode = 0;
    if (upper && c.kind == MBG_CRIT_MAX_ATOM_PAIR_DIST) *thr = c.hi;
    return true;
  }
  if (!upper) return false;
  *thr = c.hi;
  if (c.kind == MBG_CRIT_MAX_ATOM_PAIR_DIST) {
    *mode = 2;
    return true;
  }
  // com_distance_sum: the criterion's groups must be the fragment's entities (one group per slot)
  if (c.n_groups != j.order) return false;
  int first[kMaxOrder + 1] = {0};
  for (int p = 0; p < j.order; ++p) first[p + 1] = first[p] + j.slot_atoms[p];
  if (first[j.order] != j.n_frag_atoms) return false;
  for (int p = 0; p < j.order; ++p)
    for (int a = first[p]; a < first[p + 1]; ++a)
      if (c.group[a] != c.group[first[p]]) return false;
  for (int p = 0; p < j.order; ++p)
    for (int q = p + 1; q < j.order; ++q)
      if (c.group[first[p]] == c.group[first[q]]) return false;
  *mode = 1;
  return true;
}

// Build the pair list of a job on the context's stream.  With want_count a plain call brings the reduced candidate
// count to the host (the one synchronisation of this path: it sizes the cull launches and the record buffers by what
// can be accepted -- worth it from kPairListSyncCand candidates on, where launches over the full space would cost more
// than the round trip); otherwise, and always inside a prepared step (captured graph), the count stays on the device,
// *n_reduced = -1, and the launches that follow keep the full space as their bound.
constexpr long long kPairListSyncCand = 1ll << 23;  // 512 H2O (22 M): 0.47 ms with the count, 0.78 ms without; box140 x 8 (3.6 M): equal
static int build_pair_list(mbg_context_t ctx, Scratch& sc, const SysDev& s, const JobDev& j, const double* dR, int n_frames,
                           int mode, double thr, PairListDev& nl, long long* n_reduced, bool want_count = true) {
  nl.hetero = j.homogeneous ? 0 : 1;
  nl.n0 = j.homogeneous ? j.n_set : j.set_off[1] - j.set_off[0];
  nl.W = ((j.homogeneous ? j.n_set : j.set_off[2] - j.set_off[1]) + 63) / 64;
  nl.W2 = (nl.hetero && j.order == 3) ? (j.set_off[3] - j.set_off[2] + 63) / 64 : 0;
  nl.n_rows = n_frames * nl.n0;
  nl.mode = mode, nl.thr = thr;
  TRY(sc.nl_adj.ensure((size_t)nl.n_rows * (nl.W + nl.W2) * sizeof(unsigned long long)));
  TRY(sc.nl_deg.ensure((size_t)nl.n_rows * 2 * sizeof(int)));
  TRY(sc.nl_cnt.ensure((size_t)nl.n_rows * sizeof(long long)));
  TRY(sc.nl_off.ensure(((size_t)nl.n_rows + 1) * sizeof(long long)));
  nl.adj = sc.nl_adj.as<unsigned long long>(), nl.deg = sc.nl_deg.as<int>();
  nl.adj2 = nl.adj + (size_t)nl.n_rows * nl.W, nl.deg2 = nl.deg + nl.n_rows;
  nl.row_cnt = sc.nl_cnt.as<long long>(), nl.row_off = sc.nl_off.as<long long>();
  k_pair_bits<<<(nl.n_rows + kPairWarps - 1) / kPairWarps, kPairWarps * 32, 0, ctx->stream>>>(s, j, nl, dR);
  k_scan_rows<<<1, 1024, 0, ctx->stream>>>(nl.row_cnt, nl.n_rows, nl.row_off);
  ctx->launches += 2;
  CUDA_TRY(cudaGetLastError());
  *n_reduced = -1;
  if (ctx->capturing || !want_count) return MBG_OK;
  CUDA_TRY(cudaMemcpyAsync(ctx->h_totals, nl.row_off + nl.n_rows, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *n_reduced = ctx->h_totals[0];
  return MBG_OK;
}

// Enumerate + cull + compact + contract one job with the persistent contraction kernel: no host synchronisation.
// Records and flags are sized for the candidates of a chunk (an upper bound of what it accepts); every kernel
// downstream of the scan reads the accepted count from the device.  The counts for `stats` stay on the device
// (job_totals[job_slot]) until the caller asks for them.
static int run_job_pw(mbg_context_t ctx, Scratch& sc, const SysDev& s, const JobDev& j, const mbg_model_t m, const double* dR,
                      int n_frames, int shard_rank, int shard_count, bool want_virial, bool decomp, const Outputs& out,
                      int job_slot, long long* n_cand_shard_out) {
  long long n_cand_total = (long long)n_frames * j.cand_per_frame;
  long long my_granules = shard_blocks(n_cand_total, shard_rank, shard_count);  // blocks of this shard
  const int NA = j.n_frag_atoms;
  *n_cand_shard_out = shard_candidates(0, my_granules, shard_rank, shard_count, n_cand_total);
  if (my_granules == 0) return MBG_OK;
  // Large homogeneous jobs: cull the pair-list candidates instead of all C(n, order) tuples.  From here on the
  // "candidate space" that is sharded, chunked and flagged is the reduced one; records come out the same.
  PairListDev nl;
  memset(&nl, 0, sizeof(nl));
  int nl_mode = 0;
  double nl_thr = 0;
  const bool use_nl = !decomp && pair_list_applies(ctx, s, j, m, n_frames, n_cand_total, &nl_mode, &nl_thr);
  bool reduced_known = false;  // the chunks below are sized by the pair list's count instead of the full space
  if (use_nl) {
    long long n_reduced = 0;
    TRY(build_pair_list(ctx, sc, s, j, dR, n_frames, nl_mode, nl_thr, nl, &n_reduced, *n_cand_shard_out >= kPairListSyncCand));
    reduced_known = n_reduced >= 0;
    long long* enum_tot = sc.job_totals.as<long long>() + 2 * (size_t)job_slot + 1;
    if (j.homogeneous) {  // every tuple is ascending: the enumerated count is this shard's candidate count
      k_add_total<<<1, 1, 0, ctx->stream>>>(enum_tot, *n_cand_shard_out);
      ctx->launches++;
    } else {  // the ascending tuples of the product space, counted on the device; this shard's share of them
      unsigned long long* cnt = reinterpret_cast<unsigned long long*>(nl.row_cnt);  // row counts are consumed by now
      const long long pairs = (long long)(j.set_off[1] - j.set_off[0]) * (j.set_off[2] - j.set_off[1]);
      CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream));
      k_count_ascending<<<(unsigned)((pairs + 255) / 256), 256, 0, ctx->stream>>>(j, cnt);
      k_add_share<<<1, 1, 0, ctx->stream>>>(enum_tot, cnt, n_frames, shard_rank, shard_count);
      ctx->launches += 2;
    }
    if (n_reduced >= 0) {
      n_cand_total = n_reduced;
      my_granules = shard_blocks(n_cand_total, shard_rank, shard_count);
      if (my_granules == 0) return MBG_OK;
    }
  }
  const size_t rec_bytes = 4 + 4 * (size_t)NA + 8 + 8;
  long long chunk = (long long)((1ull << 30) / (kCullThreads * rec_bytes));  // <= 1 GiB of records per chunk
  chunk = std::min(std::max(chunk, 1ll << 10), std::min(1ll << 18, my_granules));
  const size_t n_rec_max = (size_t)chunk * kCullThreads;
  TRY(sc.flags.ensure(n_rec_max));
  TRY(sc.blk_acc.ensure((size_t)chunk * sizeof(int)));
  TRY(sc.blk_enum.ensure((size_t)chunk * sizeof(int)));
  TRY(sc.blk_off.ensure((size_t)chunk * sizeof(long long)));
  TRY(sc.rec_frame.ensure(n_rec_max * sizeof(int)));
  TRY(sc.rec_atoms.ensure(n_rec_max * NA * sizeof(int)));
  TRY(sc.rec_lambda.ensure(n_rec_max * sizeof(double)));
  TRY(sc.rec_slot.ensure(n_rec_max * sizeof(long long)));
  long long* job_tot = sc.job_totals.as<long long>() + 2 * (size_t)job_slot;
  for (long long g0 = 0; g0 < my_granules; g0 += chunk) {
    const int blocks = (int)std::min(chunk, my_granules - g0);
    const int slot = sc.n_pw_launch % kMaxPwLaunches;  // the contraction launched below takes the same slot
    long long* totals = sc.chunk_totals.as<long long>() + 2 * (size_t)slot;
    CullArgs ca;
    ca.sys = s, ca.crit = m->crit, ca.job = j, ca.R = dR;
    ca.n_cand_total = n_cand_total, ca.granule0 = g0;
    ca.shard_rank = shard_rank, ca.shard_count = shard_count;
    ca.flags = sc.flags.as<unsigned char>();
    ca.block_accept = sc.blk_acc.as<int>(), ca.block_enum = sc.blk_enum.as<int>();
    TRY(launch_cull(ctx, NA, blocks, ca, use_nl ? &nl : nullptr));
    k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(sc.blk_acc.as<int>(), sc.blk_enum.as<int>(), blocks,
                                               sc.blk_off.as<long long>(), totals, job_tot);
    CompactArgs pa;
    pa.sys = s, pa.job = j;
    pa.n_cand_total = n_cand_total, pa.granule0 = g0;
    pa.shard_rank = shard_rank, pa.shard_count = shard_count;
    pa.flags = sc.flags.as<unsigned char>();
    pa.block_offsets = sc.blk_off.as<long long>();
    pa.rec_base = 0;
    pa.rec_frame = sc.rec_frame.as<int>(), pa.rec_atoms = sc.rec_atoms.as<int>();
    pa.rec_lambda = sc.rec_lambda.as<double>(), pa.rec_slot = sc.rec_slot.as<long long>();
    if (use_nl) k_compact_nl<<<blocks, kCullThreads, 0, ctx->stream>>>(pa, nl);
    else k_compact<<<blocks, kCullThreads, 0, ctx->stream>>>(pa);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    const long long n_up = shard_candidates(g0, blocks, shard_rank, shard_count, n_cand_total);
    if (decomp) {
      const long long n = n_up * (NA * 3 + 1);
      k_zero_rows<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sc.rec_slot.as<long long>(), n_up, NA * 3,
                                                                         out.E, out.F, totals);
      ctx->launches++;
    }
    // Narrow descriptors: thread-per-query FMA-pipe kernel (MBG_MMA_CFG=6: off).  A thread walks all T P training rows
    // for its query, a latency floor of ~0.19 ms (T P = 2000) that only pays once there are enough queries to fill
    // the machine: below kSmallMinQueries (upper bound of this launch) the persistent tensor-pipe kernel is quicker
    // (16 monomers: 0.03 ms; measured crossover ~47 k queries).
    if (NA <= 4 && ctx->mma_cfg != 6 && (n_up * m->P >= kSmallMinQueries || ctx->mma_cfg == 8)) {  // 8 (developer): always
      MaternSmallArgs q;
      q.sys = s, q.R = dR;
      q.XA = m->XA, q.alphaE = m->alphaE, q.exp2_table = ctx->exp2_table_mma.as<double>();
      q.iperm = m->iperm, q.perm = m->perm;
      q.T = m->T, q.P = m->P, q.rows_per_pass = 0;
      q.sig = m->sig, q.c = m->c, q.std = m->std;
      q.n_frag = n_up, q.n_frag_dev = totals;
      q.rec_frame = sc.rec_frame.as<int>(), q.rec_atoms = sc.rec_atoms.as<int>();
      q.rec_lambda = sc.rec_lambda.as<double>(), q.rec_slot = sc.rec_slot.as<long long>();
      q.n_slots_per_frame = (int)j.cand_per_frame, q.decomp = decomp ? 1 : 0, q.want_virial = want_virial ? 1 : 0;
      q.E = out.E, q.F = out.F, q.virial = out.V;
      q.use_lat = m->use_lat ? 1 : 0;
      if (m->use_lat) memcpy(q.lat, m->lat, sizeof(q.lat)), memcpy(q.lat_inv, m->lat_inv, sizeof(q.lat_inv));
      sc.n_pw_launch++;  // keeps the chunk slot of this launch its own
      TRY(launch_matern_small(ctx, NA, m->use_E, q, slot));
      continue;
    }
    // CTA-cooperative persistent kernel (MBG_MMA_CFG=9: off): always for fragments of 10-18 atoms, and for 9
    // atoms when the launch can only be mid-size -- at most kCtaMaxCandidates candidates (a frame or two of an MD
    // box), at least four rounds of batches if all were accepted: there its device-side tail splitting is worth 2.7 %
    // (one 140-H2O frame 11.03 -> 10.73 ms), while large launches (-1.7 %) and small ones stay on the per-warp kernel.
    // Without a criteria and without a cell nothing is culled (heterogeneous product spaces aside) and the candidates
    // ARE the count: the window is then 8 .. 96 rounds of batches.
    // 10- to 12-atom fragments (two m-tiles per warp, 8 warps) behave like the wide ones: the CTA form is quicker at
    // every size measured (29 000 / 116 200 [h2o, h2o, meoh] fragments: 12.80 -> 12.43 ms / 50.53 -> 49.44 ms, 0.92 of
    // the DMMA peak; water trimers of the same runs: 44.15 per-warp vs 44.91 ms).  When the pair list sized this
    // chunk, n_up counts reduced candidates, of which ~5 x more are accepted than of all tuples.
    const long long bound_batches = n_up * m->P / (NA == 9 ? 288 : 128);
    const bool culled = m->crit.kind != 0 || s.periodic || !(j.homogeneous || j.explicit_combs);
    const long long cand_equiv = reduced_known ? 5 * n_up : n_up;
    const bool mid_size = NA >= 9 && (culled ? cand_equiv <= kCtaMaxCandidates && bound_batches >= 4ll * ctx->sm_count
                                             : bound_batches >= 8ll * ctx->sm_count && bound_batches <= 96ll * ctx->sm_count);
    if ((NA >= 10 || mid_size || (NA >= 9 && ctx->mma_cfg == 10)) && !m->use_lat && ctx->mma_cfg != 9 && ctx->mma_cfg != 1) {  // 10 (developer): always from 9 atoms on
      MaternMmaArgs w;
      memset(&w, 0, sizeof(w));
      w.sys = s, w.R = dR;
      w.tiles = m->tiles_mma, w.mu = m->mu, w.iperm = m->iperm;
      w.exp2_table = ctx->exp2_table_mma.as<double>();
      w.T_pad = m->T_pad, w.P = m->P, w.n_row_blocks = (m->T + 7) / 8;
      w.sig = m->sig, w.c = m->c, w.std = m->std;
      w.n_frag = n_up, w.n_frag_dev = totals;
      w.rec_frame = sc.rec_frame.as<int>(), w.rec_atoms = sc.rec_atoms.as<int>();
      w.rec_lambda = sc.rec_lambda.as<double>(), w.rec_slot = sc.rec_slot.as<long long>();
      w.n_slots_per_frame = (int)j.cand_per_frame, w.decomp = decomp ? 1 : 0, w.want_virial = want_virial ? 1 : 0;
      w.E = out.E, w.F = out.F, w.virial = out.V;
      TRY(launch_matern_mma_persistent(ctx, sc, NA, m->use_E, w, slot));
      continue;
    }
    MaternPwArgs a;
    a.sys = s;
    a.R = dR;
    a.tiles = m->tiles_mma, a.mu = m->mu, a.iperm = m->iperm, a.perm = m->perm;
    a.exp2_table = ctx->exp2_table_mma.as<double>();
    a.P = m->P;
    a.n_row_blocks = (m->T + 7) / 8;
    a.max_slices = 1;
    a.sig = m->sig, a.c = m->c, a.std = m->std;
    a.n_frag = n_up, a.n_frag_dev = totals;
    a.rec_frame = sc.rec_frame.as<int>(), a.rec_atoms = sc.rec_atoms.as<int>();
    a.rec_lambda = sc.rec_lambda.as<double>(), a.rec_slot = sc.rec_slot.as<long long>();
    a.n_slots_per_frame = (int)j.cand_per_frame;
    a.decomp = decomp ? 1 : 0;
    a.want_virial = want_virial ? 1 : 0;
    a.E = out.E, a.F = out.F, a.virial = out.V;
    a.work_counters = nullptr;
    a.use_lat = m->use_lat ? 1 : 0;
    if (m->use_lat) memcpy(a.lat, m->lat, sizeof(a.lat)), memcpy(a.lat_inv, m->lat_inv, sizeof(a.lat_inv));
    TRY(launch_matern_pw(ctx, sc, NA, m->use_E, a, slot));
  }
  return MBG_OK;
}

// Enumerate + cull + compact + contract one job over n_frames frames (wave-launched and FMA-pipe kernels: the
// accepted count of large jobs makes a host round trip to size the grid).
static int run_job(mbg_context_t ctx, const SysDev& s, const JobDev& j, const mbg_model_t m, const double* dR,
                   int n_frames, int shard_rank, int shard_count, bool want_virial, bool decomp, const Outputs& out,
                   mbg_job_stats* stats, int* deferred_slot = nullptr) {
  if (m->use_lat)
    return fail(MBG_ERR_UNSUPPORTED, "models carrying a lattice run on the MBG_CONTRACT_DMMA / AUTO kernel only");
  const long long n_cand_total = (long long)n_frames * j.cand_per_frame;
  long long n_cand_shard = 0, n_enum = 0, n_acc = 0;
  const long long my_granules = shard_blocks(n_cand_total, shard_rank, shard_count);  // blocks of this shard
  const long long kChunk = 1ll << 18;  // blocks per launch (67 M candidates, 64 MiB of flags)
  const long long kMaxRec = 1ll << 22; // contract at the latest when this many records are pending
  const int NA = j.n_frag_atoms;

  long long rec_pending = 0;
  const long long* n_frag_dev = nullptr;  // set by the small-job path below
  auto flush = [&]() -> int {
    if (rec_pending == 0) return MBG_OK;
    if (decomp) {
      const long long n = rec_pending * (NA * 3 + 1);
      k_zero_rows<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->scr.rec_slot.as<long long>(), rec_pending,
                                                                         NA * 3, out.E, out.F);
      ctx->launches++;
    }
    if (ctx->contraction != MBG_CONTRACT_FMA) {  // AUTO = DMMA for every fragment size
      MaternMmaArgs a;
      a.sys = s;
      a.R = dR;
      a.tiles = m->tiles_mma, a.mu = m->mu, a.iperm = m->iperm;
      a.exp2_table = ctx->exp2_table_mma.as<double>();
      a.T_pad = m->T_pad, a.P = m->P;
      a.n_row_blocks = (m->T + 7) / 8;
      a.sig = m->sig, a.c = m->c, a.std = m->std;
      a.n_frag = rec_pending, a.n_frag_dev = n_frag_dev;
      a.rec_frame = ctx->scr.rec_frame.as<int>(), a.rec_atoms = ctx->scr.rec_atoms.as<int>();
      a.rec_lambda = ctx->scr.rec_lambda.as<double>(), a.rec_slot = ctx->scr.rec_slot.as<long long>();
      a.n_slots_per_frame = (int)j.cand_per_frame;
      a.decomp = decomp ? 1 : 0;
      a.want_virial = want_virial ? 1 : 0;
      a.E = out.E, a.F = out.F, a.virial = out.V;
      a.persistent = 0, a.max_slices = 1;
      TRY(launch_matern_mma(ctx, NA, m->use_E, a));
      rec_pending = 0;
      return MBG_OK;
    }
    MaternArgs a;
    a.sys = s;
    a.R = dR;
    a.XA = m->XA, a.alphaE = m->alphaE, a.iperm = m->iperm;
    a.exp2_table = ctx->exp2_table.as<double>();
    a.T_pad = m->T_pad, a.P = m->P;
    a.sig = m->sig, a.c = m->c, a.std = m->std;
    a.n_frag = rec_pending, a.n_frag_dev = n_frag_dev;
    a.rec_frame = ctx->scr.rec_frame.as<int>(), a.rec_atoms = ctx->scr.rec_atoms.as<int>();
    a.rec_lambda = ctx->scr.rec_lambda.as<double>(), a.rec_slot = ctx->scr.rec_slot.as<long long>();
    a.n_slots_per_frame = (int)j.cand_per_frame;
    a.decomp = decomp ? 1 : 0;
    a.want_virial = want_virial ? 1 : 0;
    a.E = out.E, a.F = out.F, a.virial = out.V;
    TRY(launch_matern(ctx, NA, m->use_E, a));
    rec_pending = 0;
    return MBG_OK;
  };

  // Small jobs (single MD frames of small systems, the minor models of a mixed system): no host round trip for
  // the accepted count.  Records and the contraction grid are sized for the upper bound (every candidate of this
  // shard accepted); the kernels read the true count from the device and surplus CTAs exit at once.  The counts
  // for `stats` are read back once, at the end of the call (mbg_predict).
  if (deferred_slot && !decomp && my_granules > 0 && my_granules <= kChunk && n_cand_total <= kSmallJobCand &&
      ctx->n_deferred < kMaxDeferred && ctx->mma_cfg != 3) {
    const int blocks = (int)my_granules, slot = ctx->n_deferred++;
    TRY(ctx->scr.flags.ensure((size_t)blocks * kCullThreads));
    TRY(ctx->scr.blk_acc.ensure((size_t)blocks * sizeof(int)));
    TRY(ctx->scr.blk_enum.ensure((size_t)blocks * sizeof(int)));
    TRY(ctx->scr.blk_off.ensure((size_t)blocks * sizeof(long long)));
    TRY(ctx->deferred_totals.ensure((size_t)kMaxDeferred * 2 * sizeof(long long)));
    long long* totals = ctx->deferred_totals.as<long long>() + 2 * slot;
    CullArgs ca;
    ca.sys = s, ca.crit = m->crit, ca.job = j, ca.R = dR;
    ca.n_cand_total = n_cand_total, ca.granule0 = 0;
    ca.shard_rank = shard_rank, ca.shard_count = shard_count;
    ca.flags = ctx->scr.flags.as<unsigned char>();
    ca.block_accept = ctx->scr.blk_acc.as<int>(), ca.block_enum = ctx->scr.blk_enum.as<int>();
    TRY(launch_cull(ctx, NA, blocks, ca));
    k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->scr.blk_acc.as<int>(), ctx->scr.blk_enum.as<int>(), blocks,
                                               ctx->scr.blk_off.as<long long>(), totals);
    ctx->launches++;
    n_cand_shard += shard_candidates(0, blocks, shard_rank, shard_count, n_cand_total);
    TRY(ctx->scr.rec_frame.ensure((size_t)n_cand_shard * sizeof(int)));
    TRY(ctx->scr.rec_atoms.ensure((size_t)n_cand_shard * NA * sizeof(int)));
    TRY(ctx->scr.rec_lambda.ensure((size_t)n_cand_shard * sizeof(double)));
    TRY(ctx->scr.rec_slot.ensure((size_t)n_cand_shard * sizeof(long long)));
    CompactArgs pa;
    pa.sys = s, pa.job = j;
    pa.n_cand_total = n_cand_total, pa.granule0 = 0;
    pa.shard_rank = shard_rank, pa.shard_count = shard_count;
    pa.flags = ctx->scr.flags.as<unsigned char>();
    pa.block_offsets = ctx->scr.blk_off.as<long long>();
    pa.rec_base = 0;
    pa.rec_frame = ctx->scr.rec_frame.as<int>(), pa.rec_atoms = ctx->scr.rec_atoms.as<int>();
    pa.rec_lambda = ctx->scr.rec_lambda.as<double>(), pa.rec_slot = ctx->scr.rec_slot.as<long long>();
    k_compact<<<blocks, kCullThreads, 0, ctx->stream>>>(pa);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    rec_pending = n_cand_shard, n_frag_dev = totals;
    TRY(flush());
    ctx->deferred_rec[slot] = (int)ctx->launch_recs.size() - 1;
    *deferred_slot = slot;
    if (stats) stats->n_candidates = n_cand_shard, stats->n_enumerated = 0, stats->n_accepted = 0;  // patched by the caller
    return MBG_OK;
  }
  for (long long g0 = 0; g0 < my_granules; g0 += kChunk) {
    const int blocks = (int)std::min(kChunk, my_granules - g0);
    TRY(ctx->scr.flags.ensure((size_t)blocks * kCullThreads));
    TRY(ctx->scr.blk_acc.ensure((size_t)blocks * sizeof(int)));
    TRY(ctx->scr.blk_enum.ensure((size_t)blocks * sizeof(int)));
    TRY(ctx->scr.blk_off.ensure((size_t)blocks * sizeof(long long)));
    TRY(ctx->totals.ensure(2 * sizeof(long long)));
    CullArgs ca;
    ca.sys = s, ca.crit = m->crit, ca.job = j, ca.R = dR;
    ca.n_cand_total = n_cand_total, ca.granule0 = g0;
    ca.shard_rank = shard_rank, ca.shard_count = shard_count;
    ca.flags = ctx->scr.flags.as<unsigned char>();
    ca.block_accept = ctx->scr.blk_acc.as<int>(), ca.block_enum = ctx->scr.blk_enum.as<int>();
    TRY(launch_cull(ctx, NA, blocks, ca));
    k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->scr.blk_acc.as<int>(), ctx->scr.blk_enum.as<int>(), blocks,
                                               ctx->scr.blk_off.as<long long>(), ctx->totals.as<long long>());
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const long long acc = ctx->h_totals[0];
    n_enum += ctx->h_totals[1];
    n_acc += acc;
    n_cand_shard += shard_candidates(g0, blocks, shard_rank, shard_count, n_cand_total);
    if (acc == 0) continue;
    const long long need = rec_pending + acc;
    const bool fits = (size_t)need * sizeof(int) <= ctx->scr.rec_frame.cap && (size_t)need * NA * sizeof(int) <= ctx->scr.rec_atoms.cap &&
                      (size_t)need * sizeof(double) <= ctx->scr.rec_lambda.cap && (size_t)need * sizeof(long long) <= ctx->scr.rec_slot.cap;
    // growing a record buffer would drop pending records: contract them first
    if (rec_pending > 0 && (need > kMaxRec || !fits)) TRY(flush());
    const long long need2 = rec_pending + acc;
    TRY(ctx->scr.rec_frame.ensure((size_t)need2 * sizeof(int)));
    TRY(ctx->scr.rec_atoms.ensure((size_t)need2 * NA * sizeof(int)));
    TRY(ctx->scr.rec_lambda.ensure((size_t)need2 * sizeof(double)));
    TRY(ctx->scr.rec_slot.ensure((size_t)need2 * sizeof(long long)));
    CompactArgs pa;
    pa.sys = s, pa.job = j;
    pa.n_cand_total = n_cand_total, pa.granule0 = g0;
    pa.shard_rank = shard_rank, pa.shard_count = shard_count;
    pa.flags = ctx->scr.flags.as<unsigned char>();
    pa.block_offsets = ctx->scr.blk_off.as<long long>();
    pa.rec_base = rec_pending;
    pa.rec_frame = ctx->scr.rec_frame.as<int>(), pa.rec_atoms = ctx->scr.rec_atoms.as<int>();
    pa.rec_lambda = ctx->scr.rec_lambda.as<double>(), pa.rec_slot = ctx->scr.rec_slot.as<long long>();
    k_compact<<<blocks, kCullThreads, 0, ctx->stream>>>(pa);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    rec_pending += acc;
  }
  TRY(flush());
  if (stats) {
    stats->n_candidates = n_cand_shard;
    stats->n_enumerated = n_enum;
    stats->n_accepted = n_acc;
  }
  return MBG_OK;
}

static bool uses_pw(mbg_context_t ctx) {
  return ctx->contraction == MBG_CONTRACT_AUTO || ctx->contraction == MBG_CONTRACT_DMMA;
}

static void begin_call(mbg_context_t ctx) {
  ctx->ev_used = 0;
  ctx->n_contract = 0;
  ctx->launch_recs.clear();
  ctx->rec_chunk_slot.clear();
  if (ctx->scr.n_pw_launch > 0) {  // re-zero the work counters the previous call used
    const size_t used = (size_t)std::min(ctx->scr.n_pw_launch, kMaxPwLaunches) * kPwMaxSlices * sizeof(unsigned int);
    cudaMemsetAsync(ctx->scr.pw_counters.p, 0, used, ctx->stream);
  }
  ctx->scr.n_pw_launch = 0;
  cudaEventRecord(ctx->ev_call0, ctx->stream);
}

// Marks the end of a call on the stream; the elapsed times are read lazily (resolve_timing).
static void end_call_timing(mbg_context_t ctx) {
  cudaEventRecord(ctx->ev_call1, ctx->stream);
  ctx->timing_pending = true;
}

static void resolve_timing(mbg_context_t ctx) {
  if (!ctx->timing_pending) return;
  float total = 0;
  cudaEventSynchronize(ctx->ev_call1);
  cudaEventElapsedTime(&total, ctx->ev_call0, ctx->ev_call1);
  double contract = 0;
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]);
    if (i / 2 < ctx->launch_recs.size()) ctx->launch_recs[i / 2].ms = ms;
    contract += ms;
  }
  ctx->ms_contract = contract;
  ctx->ms_other = total - contract;
  ctx->timing_pending = false;
  // fragment counts of the persistent-kernel launches (they never came to the host during the call)
  if (!ctx->rec_chunk_slot.empty() && ctx->scr.n_pw_launch > 0 && ctx->scr.n_pw_launch <= kMaxPwLaunches &&
      cudaMemcpy(ctx->h_chunk, ctx->scr.chunk_totals.p, (size_t)ctx->scr.n_pw_launch * 2 * sizeof(long long),
                 cudaMemcpyDeviceToHost) == cudaSuccess)
    for (size_t i = 0; i < ctx->launch_recs.size() && i < ctx->rec_chunk_slot.size(); ++i)
      if (ctx->rec_chunk_slot[i] >= 0) ctx->launch_recs[i].n_frag = ctx->h_chunk[2 * ctx->rec_chunk_slot[i]];
}


// ------------------------------------------------------------------------------------
// Prepared steps: fixed (system, jobs, frame count), device-resident tables, the whole evaluation in one CUDA graph
// ------------------------------------------------------------------------------------
constexpr size_t kStepDirectCopyBytes = 1u << 20;  // mbg_step_run: coordinate sets from this size on skip the pinned mirrors

struct mbg_step_s {
  mbg_context_t ctx = nullptr;
  int n_atoms = 0, n_frames = 0, n_jobs = 0, n_cells = 0;
  bool want_virial = false;
  Tables tab;
  Scratch scr;
  SysDev sys;
  std::vector<JobDev> jobs;
  std::vector<mbg_model_t> models;
  std::vector<long long> n_candidates;
  DevBuf dR, dOut;             // [n_frames][n_atoms][3]; [E | F | virial]
  double* h_R = nullptr;       // pinned mirrors: the graph's copy nodes keep these addresses
  double* h_out = nullptr;
  CellDev* h_cells = nullptr;
  long long* h_tot = nullptr;  // [n_jobs][2]
  size_t n_out = 0;
  int shard_rank = 0, shard_count = 1;
  int32_t pbc[3] = {1, 1, 1};  // Cell.pbc of the system the step was created for (new cells keep it)
  cudaGraph_t graph = nullptr, graph_dev = nullptr;  // with / without the host <-> device copies
  cudaGraphExec_t exec = nullptr, exec_dev = nullptr;
};

// Everything one evaluation enqueues on the context's stream -- no host synchronisation, no allocation after the
// first pass (sizes depend only on what mbg_step_create fixed).
static int step_enqueue(mbg_step_t st, bool host_io) {
  mbg_context_t ctx = st->ctx;
  const size_t nR = (size_t)st->n_frames * st->n_atoms * 3;
  if (host_io) CUDA_TRY(cudaMemcpyAsync(st->dR.p, st->h_R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (st->n_cells)
    CUDA_TRY(cudaMemcpyAsync(st->tab.cells.p, st->h_cells, (size_t)st->n_cells * sizeof(CellDev), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemsetAsync(st->dOut.p, 0, st->n_out * sizeof(double), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(st->scr.job_totals.p, 0, (size_t)std::max(st->n_jobs, 1) * 2 * sizeof(long long), ctx->stream));
  if (st->scr.n_pw_launch > 0)
    CUDA_TRY(cudaMemsetAsync(st->scr.pw_counters.p, 0,
                             (size_t)std::min(st->scr.n_pw_launch, kMaxPwLaunches) * kPwMaxSlices * sizeof(unsigned int), ctx->stream));
  st->scr.n_pw_launch = 0;
  Outputs out;
  out.E = st->dOut.as<double>(), out.F = out.E + st->n_frames, out.V = out.F + nR;
  for (int k = 0; k < st->n_jobs; ++k) {
    if (st->jobs[k].cand_per_frame <= 0) continue;
    long long ncs = 0;
    TRY(run_job_pw(ctx, st->scr, st->sys, st->jobs[k], st->models[k], st->dR.as<double>(), st->n_frames, st->shard_rank,
                   st->shard_count, st->want_virial, false, out, k, &ncs));
    st->n_candidates[k] = ncs;
  }
  if (host_io) CUDA_TRY(cudaMemcpyAsync(st->h_out, st->dOut.p, st->n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(st->h_tot, st->scr.job_totals.p, (size_t)std::max(st->n_jobs, 1) * 2 * sizeof(long long),
                           cudaMemcpyDeviceToHost, ctx->stream));
  return MBG_OK;
}

extern "C" int mbg_step_destroy(mbg_step_t st) {
  if (!st) return MBG_OK;
  cudaSetDevice(st->ctx->device);
  cudaStreamSynchronize(st->ctx->stream);
  if (st->exec) cudaGraphExecDestroy(st->exec);
  if (st->graph) cudaGraphDestroy(st->graph);
  if (st->exec_dev) cudaGraphExecDestroy(st->exec_dev);
  if (st->graph_dev) cudaGraphDestroy(st->graph_dev);
  st->tab.release();
  st->scr.release();
  st->dR.release();
  st->dOut.release();
  if (st->h_R) cudaFreeHost(st->h_R);
  if (st->h_out) cudaFreeHost(st->h_out);
  if (st->h_cells) cudaFreeHost(st->h_cells);
  if (st->h_tot) cudaFreeHost(st->h_tot);
  delete st;
  return MBG_OK;
}

extern "C" int mbg_step_create(mbg_context_t ctx, const mbg_system_desc* sys, int32_t n_frames, const mbg_job* jobs,
                               int32_t n_jobs, uint32_t flags, int32_t shard_rank, int32_t shard_count, mbg_step_t* out_step) {
  if (!ctx || !sys || !jobs || !out_step) return fail(MBG_ERR_ARG, "NULL argument");
  if (n_frames < 1 || n_jobs < 1 || n_jobs > kMaxJobsPerCall) return fail(MBG_ERR_ARG, "bad frame / job count");
  if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(MBG_ERR_ARG, "bad shard spec");
  if (flags & ~(uint32_t)MBG_FLAG_VIRIAL) return fail(MBG_ERR_UNSUPPORTED, "flag not supported by mbg_step_create");
  if (!uses_pw(ctx)) return fail(MBG_ERR_UNSUPPORTED, "prepared steps need the persistent contraction kernel (MBG_CONTRACT_DMMA / AUTO)");
  const bool want_virial = (flags & MBG_FLAG_VIRIAL) != 0;
  if (want_virial && !sys->cell && !sys->frame_cells) return fail(MBG_ERR_ARG, "virial requires a periodic cell (gdml_predict.py:204-205)");
  CUDA_TRY(cudaSetDevice(ctx->device));
  mbg_step_t st = new mbg_step_s();
  st->ctx = ctx;
  st->n_atoms = sys->n_atoms, st->n_frames = n_frames, st->n_jobs = n_jobs, st->want_virial = want_virial;
  st->shard_rank = shard_rank, st->shard_count = shard_count;
  auto bail = [&](int rc) {
    ctx->capturing = false;
    mbg_step_destroy(st);
    return rc;
  };
  int rc = st->scr.init();
  if (rc != MBG_OK) return bail(rc);
  memset(&st->sys, 0, sizeof(st->sys));
  if ((rc = upload_system(ctx, st->tab, sys, n_frames, st->sys)) != MBG_OK) return bail(rc);
  st->n_cells = (int)st->tab.h_cells.size();
  if (sys->pbc)
    for (int i = 0; i < 3; ++i) st->pbc[i] = sys->pbc[i] ? 1 : 0;
  size_t total = 0;
  for (int q = 0; q < n_jobs; ++q) {
    const mbg_job& jq = jobs[q];
    if (!jq.model) return bail(fail(MBG_ERR_ARG, "job %d without model", q));
    total += jq.combs ? (size_t)std::max<int64_t>(jq.n_combs, 0) * jq.model->order
                      : (jq.set_offsets ? (size_t)jq.set_offsets[jq.model->order] : 0);
    total += 64;
  }
  if ((rc = st->tab.job_ints.ensure(total * sizeof(int))) != MBG_OK) return bail(rc);
  st->jobs.resize(n_jobs), st->models.resize(n_jobs), st->n_candidates.assign(n_jobs, 0);
  size_t ints_off = 0;
  for (int k = 0; k < n_jobs; ++k) {
    size_t used = 0;
    if ((rc = build_job(ctx, st->tab, sys, &jobs[k], ints_off, st->jobs[k], &used, k)) != MBG_OK) return bail(rc);
    ints_off += (used + 15) / 16 * 16 + 16;
    st->models[k] = jobs[k].model;
  }
  const size_t nR = (size_t)n_frames * sys->n_atoms * 3;
  st->n_out = (size_t)n_frames + nR + (size_t)n_frames * 9;
  if ((rc = st->dR.ensure(nR * sizeof(double) + 8)) != MBG_OK || (rc = st->dOut.ensure(st->n_out * sizeof(double) + 8)) != MBG_OK)
    return bail(rc);
  if (cudaMallocHost(&st->h_R, nR * sizeof(double) + 8) != cudaSuccess ||
      cudaMallocHost(&st->h_out, st->n_out * sizeof(double) + 8) != cudaSuccess ||
      cudaMallocHost(&st->h_cells, (size_t)std::max(st->n_cells, 1) * sizeof(CellDev)) != cudaSuccess ||
      cudaMallocHost(&st->h_tot, (size_t)n_jobs * 2 * sizeof(long long)) != cudaSuccess)
    return bail(fail(MBG_ERR_NOMEM, "cudaMallocHost for the step failed"));
  memset(st->h_R, 0, nR * sizeof(double));
  for (int c = 0; c < st->n_cells; ++c) st->h_cells[c] = st->tab.h_cells[c];
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return bail(fail(MBG_ERR_CUDA, "table upload failed"));
  // pass 1 (not captured): sizes every scratch buffer; pass 2: the same sequence recorded into a graph
  ctx->capturing = true;
  if ((rc = step_enqueue(st, true)) != MBG_OK) return bail(rc);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return bail(fail(MBG_ERR_CUDA, "step warm-up failed: %s", cudaGetErrorString(cudaGetLastError())));
  for (int pass = 0; pass < 2; ++pass) {  // with and without the host <-> device copies of R and the results
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess)
      return bail(fail(MBG_ERR_CUDA, "cudaStreamBeginCapture failed (is the context's stream the legacy default stream?)"));
    rc = step_enqueue(st, pass == 0);
    cudaGraph_t& g = pass == 0 ? st->graph : st->graph_dev;
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
    if (rc != MBG_OK) return bail(rc);
    if (ce != cudaSuccess || !g) return bail(fail(MBG_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(ce)));
    if (cudaGraphInstantiate(pass == 0 ? &st->exec : &st->exec_dev, g, 0) != cudaSuccess)
      return bail(fail(MBG_ERR_CUDA, "cudaGraphInstantiate failed"));
  }
  ctx->capturing = false;
  *out_step = st;
  return MBG_OK;
}

extern "C" int mbg_step_run(mbg_step_t st, const double* R, const double* cells, const double* cutoffs, double* E, double* F,
                            double* virial, mbg_job_stats* stats) {
  if (!st || !R || !E || !F) return fail(MBG_ERR_ARG, "NULL argument");
  if (st->want_virial && !virial) return fail(MBG_ERR_ARG, "the step computes the virial: pass a virial buffer");
  mbg_context_t ctx = st->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t nR = (size_t)st->n_frames * st->n_atoms * 3;
  // Small systems (MD frames) go through the step's own pinned mirrors inside ONE graph launch (lowest latency);
  // large coordinate sets are copied straight between the caller's buffers and the device around the device-only
  // graph: two host memcpy passes over megabytes cost more than the extra stream operations.
  const bool direct = nR * sizeof(double) >= kStepDirectCopyBytes;
  if (!direct) memcpy(st->h_R, R, nR * sizeof(double));
  if (cells) {
    if (!st->n_cells) return fail(MBG_ERR_ARG, "the step was created without a periodic cell");
    if (!cutoffs) return fail(MBG_ERR_ARG, "cells without cutoffs");
    for (int c = 0; c < st->n_cells; ++c)
      if (memcmp(st->h_cells[c].cell, cells + 9 * (size_t)c, 9 * sizeof(double)) != 0 || st->h_cells[c].cutoff != cutoffs[c])
        TRY(fill_cell(st->h_cells[c], cells + 9 * (size_t)c, cutoffs[c], st->pbc));
  }
  if (direct) {
    const double* o = st->dOut.as<double>();
    CUDA_TRY(cudaMemcpyAsync(st->dR.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaGraphLaunch(st->exec_dev, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(E, o, (size_t)st->n_frames * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(F, o + st->n_frames, nR * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (st->want_virial)
      CUDA_TRY(cudaMemcpyAsync(virial, o + st->n_frames + nR, (size_t)st->n_frames * 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  } else {
    CUDA_TRY(cudaGraphLaunch(st->exec, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(E, st->h_out, (size_t)st->n_frames * sizeof(double));
    memcpy(F, st->h_out + st->n_frames, nR * sizeof(double));
    if (st->want_virial) memcpy(virial, st->h_out + st->n_frames + nR, (size_t)st->n_frames * 9 * sizeof(double));
  }
  if (stats)
    for (int k = 0; k < st->n_jobs; ++k) {
      stats[k].n_candidates = st->n_candidates[k];
      stats[k].n_accepted = st->h_tot[2 * k], stats[k].n_enumerated = st->h_tot[2 * k + 1];
    }
  return MBG_OK;
}

// Device-resident variant: dR, dE, dF, dV are device pointers; everything is enqueued on the context's stream and the
// call returns without synchronising (cells, when given, are host arrays as in mbg_step_run).
extern "C" int mbg_step_run_device(mbg_step_t st, const double* dR, const double* cells, const double* cutoffs, double* dE,
                                   double* dF, double* dV) {
  if (!st || !dR || !dE || !dF) return fail(MBG_ERR_ARG, "NULL argument");
  if (st->want_virial && !dV) return fail(MBG_ERR_ARG, "the step computes the virial: pass a virial buffer");
  mbg_context_t ctx = st->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t nR = (size_t)st->n_frames * st->n_atoms * 3;
  if (cells) {
    if (!st->n_cells) return fail(MBG_ERR_ARG, "the step was created without a periodic cell");
    if (!cutoffs) return fail(MBG_ERR_ARG, "cells without cutoffs");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the pinned cell table may still be read by the previous replay
    for (int c = 0; c < st->n_cells; ++c)
      if (memcmp(st->h_cells[c].cell, cells + 9 * (size_t)c, 9 * sizeof(double)) != 0 || st->h_cells[c].cutoff != cutoffs[c])
        TRY(fill_cell(st->h_cells[c], cells + 9 * (size_t)c, cutoffs[c], st->pbc));
  }
  CUDA_TRY(cudaMemcpyAsync(st->dR.p, dR, nR * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(cudaGraphLaunch(st->exec_dev, ctx->stream));
  const double* o = st->dOut.as<double>();
  CUDA_TRY(cudaMemcpyAsync(dE, o, (size_t)st->n_frames * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(dF, o + st->n_frames, nR * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (st->want_virial)
    CUDA_TRY(cudaMemcpyAsync(dV, o + st->n_frames + nR, (size_t)st->n_frames * 9 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return MBG_OK;
}

// Counters of the most recent replay of a step (valid after the stream has been synchronised).
extern "C" int mbg_step_stats(mbg_step_t st, mbg_job_stats* stats) {
  if (!st || !stats) return fail(MBG_ERR_ARG, "NULL argument");
  for (int k = 0; k < st->n_jobs; ++k) {
    stats[k].n_candidates = st->n_candidates[k];
    stats[k].n_accepted = st->h_tot[2 * k], stats[k].n_enumerated = st->h_tot[2 * k + 1];
  }
  return MBG_OK;
}

// ------------------------------------------------------------------------------------
// exported functions
// ------------------------------------------------------------------------------------
extern "C" {

static int context_init(mbg_context_t ctx, int device, int sm_count);

const char* mbg_last_error(void) { return g_last_error.c_str(); }
int mbg_abi_version(void) { return MBG_ABI_VERSION; }

int mbg_device_count(int* count) {
  if (!count) return fail(MBG_ERR_ARG, "count is NULL");
  CUDA_TRY(cudaGetDeviceCount(count));
  return MBG_OK;
}

int mbg_context_create(int device, mbg_context_t* out) {
  if (!out) return fail(MBG_ERR_ARG, "ctx is NULL");
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(MBG_ERR_ARG, "device %d out of range (%d visible)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MBG_ERR_UNSUPPORTED, "this build targets sm_100a (B200); device %d is sm_%d%d", device, prop.major,
                prop.minor);
  mbg_context_t ctx = new mbg_context_s();
  const int rc = context_init(ctx, device, prop.multiProcessorCount);
  if (rc != MBG_OK) {
    mbg_context_destroy(ctx);  // releases whatever was allocated before the failure
    return rc;
  }
  *out = ctx;
  return MBG_OK;
}

static int context_init(mbg_context_t ctx, int device, int sm_count) {
  ctx->device = device;
  ctx->sm_count = sm_count;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  CUDA_TRY(cudaMallocHost(&ctx->h_totals, (2 + 2 * kMaxDeferred) * sizeof(long long)));
  CUDA_TRY(cudaMallocHost(&ctx->h_chunk, (size_t)(kMaxPwLaunches + kMaxJobsPerCall) * 2 * sizeof(long long)));
  TRY(ctx->scr.init());
  {
    double tab[1 << kExpTableBits];
    for (int j = 0; j < (1 << kExpTableBits); ++j) tab[j] = (double)exp2l((long double)j / (1 << kExpTableBits));
    TRY(ctx->exp2_table.ensure(sizeof(tab)));
    CUDA_TRY(cudaMemcpy(ctx->exp2_table.p, tab, sizeof(tab), cudaMemcpyHostToDevice));
    double tab2[kMmaExpTable];
    for (int j = 0; j < kMmaExpTable; ++j) tab2[j] = (double)exp2l((long double)j / kMmaExpTable);
    TRY(ctx->exp2_table_mma.ensure(sizeof(tab2)));
    CUDA_TRY(cudaMemcpy(ctx->exp2_table_mma.p, tab2, sizeof(tab2), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaEventCreate(&ctx->ev_call0));
  CUDA_TRY(cudaEventCreate(&ctx->ev_call1));
  if (const char* v = getenv("MBG_CONTRACT")) {  // developer override; mbg_context_set_contraction wins
    if (!strcmp(v, "fma")) ctx->contraction = MBG_CONTRACT_FMA;
    if (!strcmp(v, "dmma")) ctx->contraction = MBG_CONTRACT_DMMA;
    if (!strcmp(v, "dmma_wave")) ctx->contraction = MBG_CONTRACT_DMMA_WAVE;
  }
  if (const char* v = getenv("MBG_MMA_CFG")) ctx->mma_cfg = atoi(v);
  if (const char* v = getenv("MBG_CULL")) {  // developer override; mbg_context_set_culling wins
    if (!strcmp(v, "full")) ctx->culling = MBG_CULL_FULL;
    if (!strcmp(v, "pairlist")) ctx->culling = MBG_CULL_PAIRLIST;
  }
  return MBG_OK;
}

int mbg_context_destroy(mbg_context_t ctx) {
  if (!ctx) return MBG_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->tab.release();
  ctx->scr.release();
  for (DevBuf* b : {&ctx->R, &ctx->outE, &ctx->outF, &ctx->outV, &ctx->totals, &ctx->exp2_table, &ctx->exp2_table_mma, &ctx->km_in,
                    &ctx->deferred_totals, &ctx->mask_buf, &ctx->km_ints, &ctx->km_ints2, &ctx->km_jlist, &ctx->km_out})
    b->release();
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->ev_call0) cudaEventDestroy(ctx->ev_call0);
  if (ctx->ev_call1) cudaEventDestroy(ctx->ev_call1);
  if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
  if (ctx->h_chunk) cudaFreeHost(ctx->h_chunk);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return MBG_OK;
}

int mbg_context_set_stream(mbg_context_t ctx, void* cuda_stream) {
  if (!ctx) return fail(MBG_ERR_ARG, "ctx is NULL");
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return MBG_OK;
}

int mbg_context_synchronize(mbg_context_t ctx) {
  if (!ctx) return fail(MBG_ERR_ARG, "ctx is NULL");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MBG_OK;
}

int mbg_context_set_contraction(mbg_context_t ctx, int kind) {
  if (!ctx) return fail(MBG_ERR_ARG, "ctx is NULL");
  if (kind != MBG_CONTRACT_AUTO && kind != MBG_CONTRACT_FMA && kind != MBG_CONTRACT_DMMA && kind != MBG_CONTRACT_DMMA_WAVE)
    return fail(MBG_ERR_ARG, "unknown contraction kind %d", kind);
  ctx->contraction = kind;
  return MBG_OK;
}

int mbg_context_set_culling(mbg_context_t ctx, int kind) {
  if (!ctx) return fail(MBG_ERR_ARG, "ctx is NULL");
  if (kind != MBG_CULL_AUTO && kind != MBG_CULL_FULL && kind != MBG_CULL_PAIRLIST)
    return fail(MBG_ERR_ARG, "unknown culling kind %d", kind);
  ctx->culling = kind;
  return MBG_OK;
}

int mbg_context_launch_count(mbg_context_t ctx, int64_t* n) {
  if (!ctx || !n) return fail(MBG_ERR_ARG, "NULL argument");
  *n = ctx->launches;
  return MBG_OK;
}

int mbg_context_last_timing(mbg_context_t ctx, double* ms_contract, double* ms_other, int64_t* n_contract) {
  if (!ctx) return fail(MBG_ERR_ARG, "ctx is NULL");
  CUDA_TRY(cudaSetDevice(ctx->device));
  resolve_timing(ctx);
  if (ms_contract) *ms_contract = ctx->ms_contract;
  if (ms_other) *ms_other = ctx->ms_other;
  if (n_contract) *n_contract = ctx->n_contract;
  return MBG_OK;
}

int mbg_context_last_launches(mbg_context_t ctx, int32_t max_records, int32_t* n_records, int32_t* n_frag_atoms,
                              int64_t* n_fragments, int32_t* n_train_padded, int32_t* n_perms, double* ms) {
  if (!ctx || !n_records) return fail(MBG_ERR_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  resolve_timing(ctx);
  const int n = (int)ctx->launch_recs.size();
  *n_records = n;
  for (int i = 0; i < n && i < max_records; ++i) {
    const auto& r = ctx->launch_recs[i];
    if (n_frag_atoms) n_frag_atoms[i] = r.n_frag_atoms;
    if (n_fragments) n_fragments[i] = r.n_frag;
    if (n_train_padded) n_train_padded[i] = r.n_train;
    if (n_perms) n_perms[i] = r.n_perms;
    if (ms) ms[i] = r.ms;
  }
  return MBG_OK;
}

int mbg_model_create(mbg_context_t ctx, const mbg_model_desc* d, mbg_model_t* out) {
  if (!ctx || !d || !out) return fail(MBG_ERR_ARG, "NULL argument");
  if (d->n_atoms < 2 || d->n_atoms > kMaxFragAtoms) return fail(MBG_ERR_ARG, "n_atoms=%d out of range", d->n_atoms);
  if (!na_supported(d->n_atoms))
    return fail(MBG_ERR_UNSUPPORTED, "no kernels for fragments of %d atoms in this build", d->n_atoms);
  if (d->order < 1 || d->order > kMaxOrder) return fail(MBG_ERR_ARG, "order=%d out of range", d->order);
  if (d->n_train < 1 || d->n_perms < 1) return fail(MBG_ERR_ARG, "n_train and n_perms must be positive");
  if (!d->R_desc || !d->R_d_desc_alpha || !d->tril_perms_lin) return fail(MBG_ERR_ARG, "model arrays are NULL");
  if (!(d->sig > 0)) return fail(MBG_ERR_ARG, "sig must be positive");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int D = d->n_atoms * (d->n_atoms - 1) / 2, T = d->n_train, P = d->n_perms;
  // pi_p(dd) = tril_perms_lin[dd*P + p] mod D (gdml_model.py:109-118); must be a permutation of 0..D-1
  std::vector<int> iperm((size_t)P * D), fperm((size_t)P * D);
  for (int p = 0; p < P; ++p) {
    std::vector<char> seen(D, 0);
    for (int dd = 0; dd < D; ++dd) {
      const int64_t lin = d->tril_perms_lin[(size_t)dd * P + p];
      if (lin < 0 || lin >= (int64_t)P * D) return fail(MBG_ERR_ARG, "tril_perms_lin[%d] out of range", dd * P + p);
      const int e = (int)(lin % D);
      if (seen[e]) return fail(MBG_ERR_UNSUPPORTED, "tril_perms_lin row %d is not a permutation", p);
      seen[e] = 1;
      iperm[(size_t)p * D + e] = dd;  // pi_p^-1(e) = dd
      fperm[(size_t)p * D + dd] = e;  // pi_p(dd) = e
    }
  }
  const int T_pad = (T + kTileRows - 1) / kTileRows * kTileRows;
  const int Dp = model_row_pad(d->n_atoms);  // rows padded to the kernel's lane blocking
  std::vector<double2> xa((size_t)T_pad * Dp, make_double2(0.0, 0.0));
  for (int t = 0; t < T; ++t)
    for (int e = 0; e < D; ++e)
      xa[(size_t)t * Dp + e] = make_double2(d->R_desc[(size_t)e * T + t], d->R_d_desc_alpha[(size_t)t * D + e]);
  mbg_model_t m = new mbg_model_s();
  m->ctx = ctx;
  m->n_atoms = d->n_atoms, m->order = d->order, m->D = D, m->T = T, m->T_pad = T_pad, m->P = P;
  m->sig = d->sig, m->c = d->c, m->std = d->std;
  m->use_E = d->alphas_E != nullptr;
  auto cleanup = [&]() { mbg_model_destroy(m); };
#define MODEL_TRY(expr)                                                                                   \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      cleanup();                                                                                          \
      return fail(MBG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                                     \
  } while (0)
  if (d->lattice) {
    m->use_lat = true;
    memcpy(m->lat, d->lattice, 9 * sizeof(double));
    if (!hostmath::inverse3(m->lat, m->lat_inv)) {
      cleanup();
      return fail(MBG_ERR_ARG, "model lattice is singular");
    }
  }
  if (cudaMalloc(&m->XA, xa.size() * sizeof(double2)) != cudaSuccess ||
      cudaMalloc(&m->iperm, iperm.size() * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&m->perm, fperm.size() * sizeof(int)) != cudaSuccess) {
    cleanup();
    return fail(MBG_ERR_NOMEM, "cudaMalloc for model failed");
  }
  MODEL_TRY(cudaMemcpy(m->XA, xa.data(), xa.size() * sizeof(double2), cudaMemcpyHostToDevice));
  MODEL_TRY(cudaMemcpy(m->iperm, iperm.data(), iperm.size() * sizeof(int), cudaMemcpyHostToDevice));
  MODEL_TRY(cudaMemcpy(m->perm, fperm.data(), fperm.size() * sizeof(int), cudaMemcpyHostToDevice));
  {
    // DMMA layout (matern_mma.cuh): centred descriptors, A scaled by 5/sig, per-row -|X'|^2/2 and -X'.A_s,
    // packed per tile of 64 rows; side arrays in column order (column j of a block <-> row tau(j)).
    static_assert(kTileRows % 32 == 0, "T_pad (multiple of kTileRows) must be whole DMMA tiles of 16 or 32 rows");
    const int Dq = mma_row_pad(d->n_atoms), TILE = mma_tile_doubles(d->n_atoms), TRm = mma_tile_rows(d->n_atoms);
    std::vector<double> mu(D, 0.0);
    for (int e = 0; e < D; ++e) {
      long double acc = 0;
      for (int t = 0; t < T; ++t) acc += d->R_desc[(size_t)e * T + t];
      mu[e] = (double)(acc / T);
    }
    const double diag = 5.0 / d->sig;
    std::vector<double> tl((size_t)(T_pad / TRm) * TILE, 0.0);
    for (int t = 0; t < T; ++t) {
      double* tile = tl.data() + (size_t)(t / TRm) * TILE;
      const int r = t % TRm;
      double* xrow = tile + (size_t)r * Dq;
      double* arow = tile + (size_t)TRm * Dq + (size_t)r * Dq;
      long double xx = 0, xa = 0;
      for (int e = 0; e < D; ++e) {
        const double xv = d->R_desc[(size_t)e * T + t] - mu[e];
        const double av = diag * d->R_d_desc_alpha[(size_t)t * D + e];
        xrow[e] = xv, arow[e] = av;
        xx += (long double)xv * xv, xa += (long double)xv * av;
      }
      xrow[D] = 1.0;  // spare column of GEMM 2: Gacc[q][D] = sum_t w = s1
      // position of row r in column order: r = 8 b + tau(j)  ->  8 b + j  (tau is an involution)
      const int pos = (r / 8) * 8 + mma_tau(r % 8);
      double* side = tile + (size_t)2 * TRm * Dq;
      side[2 * pos] = (double)(-0.5L * xx), side[2 * pos + 1] = (double)(-xa);  // seeds of the GEMM-1 accumulators
      if (d->alphas_E) side[2 * TRm + pos] = d->alphas_E[t];
    }
    if (cudaMalloc(&m->tiles_mma, tl.size() * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&m->mu, (size_t)D * sizeof(double)) != cudaSuccess) {
      cleanup();
      return fail(MBG_ERR_NOMEM, "cudaMalloc for model failed");
    }
    MODEL_TRY(cudaMemcpy(m->tiles_mma, tl.data(), tl.size() * sizeof(double), cudaMemcpyHostToDevice));
    MODEL_TRY(cudaMemcpy(m->mu, mu.data(), (size_t)D * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (m->use_E) {
    std::vector<double> aE(T_pad, 0.0);
    for (int t = 0; t < T; ++t) aE[t] = d->alphas_E[t];
    if (cudaMalloc(&m->alphaE, aE.size() * sizeof(double)) != cudaSuccess) {
      cleanup();
      return fail(MBG_ERR_NOMEM, "cudaMalloc for model failed");
    }
    MODEL_TRY(cudaMemcpy(m->alphaE, aE.data(), aE.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // criteria
  CritDev& c = m->crit;
  memset(&c, 0, sizeof(c));
  c.kind = d->criteria_kind, c.bound = d->criteria_bound, c.lo = d->criteria_lo, c.hi = d->criteria_hi;
  if (c.kind < MBG_CRIT_NONE || c.kind > MBG_CRIT_MAX_ATOM_PAIR_DIST || c.bound < 0 || c.bound > 2) {
    cleanup();
    return fail(MBG_ERR_ARG, "unknown criteria kind/bound");
  }
  if (c.kind == MBG_CRIT_COM_DISTANCE_SUM) {
    if (!d->criteria_entity_ids) {
      cleanup();
      return fail(MBG_ERR_ARG, "com_distance_sum needs criteria_entity_ids");
    }
    // relabel to 0..k-1 in ascending label order (set() of small ints iterates ascending)
    std::vector<int> labels(d->criteria_entity_ids, d->criteria_entity_ids + d->n_atoms);
    std::vector<int> uniq = labels;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    c.n_groups = (int)uniq.size();
    for (int a = 0; a < d->n_atoms; ++a)
      c.group[a] = (int)(std::lower_bound(uniq.begin(), uniq.end(), labels[a]) - uniq.begin());
  }
  *out = m;
  return MBG_OK;
#undef MODEL_TRY
}

int mbg_model_destroy(mbg_model_t m) {
  if (!m) return MBG_OK;
  cudaSetDevice(m->ctx->device);
  if (m->XA) cudaFree(m->XA);
  if (m->tiles_mma) cudaFree(m->tiles_mma);
  if (m->mu) cudaFree(m->mu);
  if (m->alphaE) cudaFree(m->alphaE);
  if (m->iperm) cudaFree(m->iperm);
  if (m->perm) cudaFree(m->perm);
  delete m;
  return MBG_OK;
}

int mbg_predict(mbg_context_t ctx, const mbg_system_desc* sys, const double* R, int32_t n_frames, const mbg_job* jobs,
                int32_t n_jobs, uint32_t flags, int32_t shard_rank, int32_t shard_count, double* E, double* F,
                double* virial, mbg_job_stats* stats) {
  if (!ctx || !sys || !R || !jobs || !E || !F) return fail(MBG_ERR_ARG, "NULL argument");
  if (n_frames < 0 || n_jobs < 0) return fail(MBG_ERR_ARG, "negative count");
  if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(MBG_ERR_ARG, "bad shard spec");
  const bool want_virial = (flags & MBG_FLAG_VIRIAL) != 0;
  if (want_virial && !virial) return fail(MBG_ERR_ARG, "MBG_FLAG_VIRIAL without a virial buffer");
  if (want_virial && !sys->cell && !sys->frame_cells)
    return fail(MBG_ERR_ARG, "virial requires a periodic cell (gdml_predict.py:204-205)");
  CUDA_TRY(cudaSetDevice(ctx->device));
  begin_call(ctx);
  SysDev s;
  memset(&s, 0, sizeof(s));
  TRY(upload_system(ctx, ctx->tab, sys, n_frames, s));
  const size_t nR = (size_t)n_frames * sys->n_atoms * 3;
  const double* dR = R;
  if (!(flags & MBG_FLAG_R_ON_DEVICE)) {
    TRY(ctx->R.ensure(nR * sizeof(double) + 8));
    CUDA_TRY(cudaMemcpyAsync(ctx->R.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dR = ctx->R.as<double>();
  }
  Outputs out;
  const bool out_dev = (flags & MBG_FLAG_OUT_ON_DEVICE) != 0;
  if (out_dev) {
    out.E = E, out.F = F, out.V = virial;
  } else {
    TRY(ctx->outE.ensure((size_t)n_frames * sizeof(double) + 8));
    TRY(ctx->outF.ensure(nR * sizeof(double) + 8));
    TRY(ctx->outV.ensure((size_t)n_frames * 9 * sizeof(double) + 8));
    out.E = ctx->outE.as<double>(), out.F = ctx->outF.as<double>(), out.V = ctx->outV.as<double>();
  }
  if (!(out_dev && (flags & MBG_FLAG_ACCUMULATE))) {
    CUDA_TRY(cudaMemsetAsync(out.E, 0, (size_t)n_frames * sizeof(double), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(out.F, 0, nR * sizeof(double), ctx->stream));
    if (want_virial) CUDA_TRY(cudaMemsetAsync(out.V, 0, (size_t)n_frames * 9 * sizeof(double), ctx->stream));
  }
  size_t ints_off = 0;
  ctx->n_deferred = 0;
  std::vector<std::pair<int, int>> deferred;  // (job, slot of its device-side counts)
  const bool pw = uses_pw(ctx);
  if (pw) {
    if (n_jobs > kMaxJobsPerCall) return fail(MBG_ERR_UNSUPPORTED, "at most %d jobs per call", kMaxJobsPerCall);
    if (n_jobs > 0) CUDA_TRY(cudaMemsetAsync(ctx->scr.job_totals.p, 0, (size_t)n_jobs * 2 * sizeof(long long), ctx->stream));
  }
  for (int k = 0; k < n_jobs && n_frames > 0; ++k) {
    JobDev j;
    size_t used = 0;
    // job tables are consumed by kernels enqueued below; give every job its own slice, and never
    // reallocate mid-call: pre-size for all jobs on first use
    if (k == 0) {
      size_t total = 0;
      for (int q = 0; q < n_jobs; ++q) {
        const mbg_job& jq = jobs[q];
        if (!jq.model) return fail(MBG_ERR_ARG, "job %d without model", q);
        total += jq.combs ? (size_t)std::max<int64_t>(jq.n_combs, 0) * jq.model->order
                          : (jq.set_offsets ? (size_t)jq.set_offsets[jq.model->order] : 0);
        total += 64;
      }
      TRY(ctx->tab.job_ints.ensure(total * sizeof(int)));
    }
    TRY(build_job(ctx, ctx->tab, sys, &jobs[k], ints_off, j, &used, k));
    ints_off += (used + 15) / 16 * 16 + 16;  // keep slices 64-byte aligned
    mbg_job_stats st = {0, 0, 0};
    int slot = -1;
    if (j.cand_per_frame > 0) {
      if (pw) {
        long long n_cand_shard = 0;
        TRY(run_job_pw(ctx, ctx->scr, s, j, jobs[k].model, dR, n_frames, shard_rank, shard_count, want_virial, false, out, k,
                       &n_cand_shard));
        st.n_candidates = n_cand_shard;
      } else {
        TRY(run_job(ctx, s, j, jobs[k].model, dR, n_frames, shard_rank, shard_count, want_virial, false, out, &st, &slot));
      }
    }
    if (stats) stats[k] = st;
    if (slot >= 0) deferred.push_back({k, slot});
  }
  long long* h_job = ctx->h_chunk + 2 * (size_t)kMaxPwLaunches;
  const bool want_job_totals = pw && stats && n_jobs > 0 && n_frames > 0;
  if (want_job_totals)
    CUDA_TRY(cudaMemcpyAsync(h_job, ctx->scr.job_totals.p, (size_t)n_jobs * 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (!deferred.empty())  // counts of the small jobs: one copy, resolved by the synchronisation below
    CUDA_TRY(cudaMemcpyAsync(ctx->h_totals + 2, ctx->deferred_totals.p, (size_t)ctx->n_deferred * 2 * sizeof(long long),
                             cudaMemcpyDeviceToHost, ctx->stream));
  if (!out_dev) {
    CUDA_TRY(cudaMemcpyAsync(E, out.E, (size_t)n_frames * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(F, out.F, nR * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (want_virial)
      CUDA_TRY(cudaMemcpyAsync(virial, out.V, (size_t)n_frames * 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  } else if ((!deferred.empty() && stats) || want_job_totals) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // device outputs stay enqueued otherwise; the caller asked for counts
  }
  if (want_job_totals)
    for (int k = 0; k < n_jobs; ++k) stats[k].n_accepted = h_job[2 * k], stats[k].n_enumerated = h_job[2 * k + 1];
  if (stats)
    for (const auto& d : deferred) {
      stats[d.first].n_accepted = ctx->h_totals[2 + 2 * d.second];
      stats[d.first].n_enumerated = ctx->h_totals[2 + 2 * d.second + 1];
      const int rec = ctx->deferred_rec[d.second];
      if (rec >= 0 && rec < (int)ctx->launch_recs.size()) ctx->launch_recs[rec].n_frag = stats[d.first].n_accepted;
    }
  end_call_timing(ctx);
  return MBG_OK;
}

int mbg_predict_decomp(mbg_context_t ctx, const mbg_system_desc* sys, const double* R, const mbg_job* job,
                       uint32_t flags, double* E, double* F, mbg_job_stats* stats) {
  if (!ctx || !sys || !R || !job || !E || !F) return fail(MBG_ERR_ARG, "NULL argument");
  if (!job->combs) return fail(MBG_ERR_ARG, "mbg_predict_decomp needs explicit combs (mbe.py:936-951)");
  if (flags & (MBG_FLAG_OUT_ON_DEVICE | MBG_FLAG_VIRIAL | MBG_FLAG_ACCUMULATE))
    return fail(MBG_ERR_UNSUPPORTED, "flag not supported by mbg_predict_decomp");
  CUDA_TRY(cudaSetDevice(ctx->device));
  begin_call(ctx);
  SysDev s;
  memset(&s, 0, sizeof(s));
  TRY(upload_system(ctx, ctx->tab, sys, 1, s));
  const size_t nR = (size_t)sys->n_atoms * 3;
  const double* dR = R;
  if (!(flags & MBG_FLAG_R_ON_DEVICE)) {
    TRY(ctx->R.ensure(nR * sizeof(double) + 8));
    CUDA_TRY(cudaMemcpyAsync(ctx->R.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dR = ctx->R.as<double>();
  }
  const mbg_model_t m = job->model;
  if (!m) return fail(MBG_ERR_ARG, "job without model");
  const long long n = job->n_combs, row = (long long)m->n_atoms * 3;
  mbg_job_stats st = {0, 0, 0};
  if (n > 0) {
    TRY(ctx->tab.job_ints.ensure(((size_t)n * m->order + 32) * sizeof(int)));
    JobDev j;
    size_t used = 0;
    TRY(build_job(ctx, ctx->tab, sys, job, 0, j, &used));
    TRY(ctx->outE.ensure((size_t)n * sizeof(double) + 8));
    TRY(ctx->outF.ensure((size_t)(n * row) * sizeof(double) + 8));
    const double nan = std::numeric_limits<double>::quiet_NaN();
    k_fill<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->outE.as<double>(), n, nan);
    k_fill<<<(unsigned)((n * row + 255) / 256), 256, 0, ctx->stream>>>(ctx->outF.as<double>(), n * row, nan);
    ctx->launches += 2;
    Outputs out = {ctx->outE.as<double>(), ctx->outF.as<double>(), nullptr};
    if (uses_pw(ctx)) {
      long long n_cand_shard = 0;
      CUDA_TRY(cudaMemsetAsync(ctx->scr.job_totals.p, 0, 2 * sizeof(long long), ctx->stream));
      TRY(run_job_pw(ctx, ctx->scr, s, j, m, dR, 1, 0, 1, false, true, out, 0, &n_cand_shard));
      st.n_candidates = n_cand_shard;
      CUDA_TRY(cudaMemcpyAsync(ctx->h_chunk + 2 * (size_t)kMaxPwLaunches, ctx->scr.job_totals.p, 2 * sizeof(long long),
                               cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      TRY(run_job(ctx, s, j, m, dR, 1, 0, 1, false, true, out, &st));
    }
    CUDA_TRY(cudaMemcpyAsync(E, out.E, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(F, out.F, (size_t)(n * row) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (n > 0 && uses_pw(ctx)) {
    st.n_accepted = ctx->h_chunk[2 * (size_t)kMaxPwLaunches], st.n_enumerated = ctx->h_chunk[2 * (size_t)kMaxPwLaunches + 1];
  }
  end_call_timing(ctx);
  if (stats) *stats = st;
  return MBG_OK;
}

int mbg_accept_mask(mbg_context_t ctx, const mbg_system_desc* sys, const double* R, int32_t n_frames, const mbg_job* job,
                    uint32_t flags, uint8_t* mask, int64_t* n_per_frame) {
  if (!ctx || !sys || !R || !job || !n_per_frame) return fail(MBG_ERR_ARG, "NULL argument");
  if (n_frames < 0) return fail(MBG_ERR_ARG, "negative count");
  if (flags & ~(uint32_t)MBG_FLAG_R_ON_DEVICE) return fail(MBG_ERR_UNSUPPORTED, "flag not supported by mbg_accept_mask");
  if (!job->model) return fail(MBG_ERR_ARG, "job without model");
  CUDA_TRY(cudaSetDevice(ctx->device));
  SysDev s;
  memset(&s, 0, sizeof(s));
  TRY(upload_system(ctx, ctx->tab, sys, n_frames, s));
  const mbg_model_t m = job->model;
  const size_t n_ints = job->combs ? (size_t)std::max<int64_t>(job->n_combs, 0) * m->order
                                   : (job->set_offsets ? (size_t)job->set_offsets[m->order] : 0);
  TRY(ctx->tab.job_ints.ensure((n_ints + 64) * sizeof(int)));
  JobDev j;
  size_t used = 0;
  TRY(build_job(ctx, ctx->tab, sys, job, 0, j, &used));
  *n_per_frame = j.cand_per_frame;
  if (!mask || n_frames == 0 || j.cand_per_frame == 0) return MBG_OK;
  const size_t nR = (size_t)n_frames * sys->n_atoms * 3;
  const double* dR = R;
  if (!(flags & MBG_FLAG_R_ON_DEVICE)) {
    TRY(ctx->R.ensure(nR * sizeof(double) + 8));
    CUDA_TRY(cudaMemcpyAsync(ctx->R.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dR = ctx->R.as<double>();
  }
  const long long n_cand_total = (long long)n_frames * j.cand_per_frame;
  const long long kChunk = 1ll << 18;
  int nl_mode = 0;
  double nl_thr = 0;
  if (n_cand_total <= (1ll << 31) && pair_list_applies(ctx, s, j, m, n_frames, n_cand_total, &nl_mode, &nl_thr)) {
    // Through the pair list: every tuple of a homogeneous job is ascending (flag 2); the accepted ones of the reduced
    // space set their byte of the full-space mask by their combination rank.
    PairListDev nl;
    memset(&nl, 0, sizeof(nl));
    long long n_reduced = 0;
    TRY(build_pair_list(ctx, ctx->scr, s, j, dR, n_frames, nl_mode, nl_thr, nl, &n_reduced));
    TRY(ctx->mask_buf.ensure((size_t)n_cand_total));
    if (j.homogeneous) {
      CUDA_TRY(cudaMemsetAsync(ctx->mask_buf.p, 2, (size_t)n_cand_total, ctx->stream));
    } else {
      k_mask_init<<<(unsigned)((n_cand_total + kCullThreads - 1) / kCullThreads), kCullThreads, 0, ctx->stream>>>(
          j, n_cand_total, ctx->mask_buf.as<unsigned char>());
      ctx->launches++;
    }
    const long long n_blocks = (n_reduced + kCullThreads - 1) / kCullThreads;
    for (long long g0 = 0; g0 < n_blocks; g0 += kChunk) {
      const int blocks = (int)std::min(kChunk, n_blocks - g0);
      TRY(ctx->scr.flags.ensure((size_t)blocks * kCullThreads));
      TRY(ctx->scr.blk_acc.ensure((size_t)blocks * sizeof(int)));
      TRY(ctx->scr.blk_enum.ensure((size_t)blocks * sizeof(int)));
      CullArgs ca;
      ca.sys = s, ca.crit = m->crit, ca.job = j, ca.R = dR;
      ca.n_cand_total = n_reduced, ca.granule0 = g0;
      ca.shard_rank = 0, ca.shard_count = 1;
      ca.flags = ctx->scr.flags.as<unsigned char>();
      ca.block_accept = ctx->scr.blk_acc.as<int>(), ca.block_enum = ctx->scr.blk_enum.as<int>();
      TRY(launch_cull(ctx, m->n_atoms, blocks, ca, &nl));
      k_mask_nl<<<blocks, kCullThreads, 0, ctx->stream>>>(ca, nl, ctx->mask_buf.as<unsigned char>(), 0, n_cand_total);
      ctx->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(mask, ctx->mask_buf.p, (size_t)n_cand_total, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return MBG_OK;
  }
  const long long n_granules = (n_cand_total + kCullThreads - 1) / kCullThreads;
  for (long long g0 = 0; g0 < n_granules; g0 += kChunk) {
    const int blocks = (int)std::min(kChunk, n_granules - g0);
    TRY(ctx->scr.flags.ensure((size_t)blocks * kCullThreads));
    TRY(ctx->scr.blk_acc.ensure((size_t)blocks * sizeof(int)));
    TRY(ctx->scr.blk_enum.ensure((size_t)blocks * sizeof(int)));
    CullArgs ca;
    ca.sys = s, ca.crit = m->crit, ca.job = j, ca.R = dR;
    ca.n_cand_total = n_cand_total, ca.granule0 = g0;
    ca.shard_rank = 0, ca.shard_count = 1;
    ca.flags = ctx->scr.flags.as<unsigned char>();
    ca.block_accept = ctx->scr.blk_acc.as<int>(), ca.block_enum = ctx->scr.blk_enum.as<int>();
    TRY(launch_cull(ctx, m->n_atoms, blocks, ca));
    const long long lo = g0 * kCullThreads, n = std::min<long long>((long long)blocks * kCullThreads, n_cand_total - lo);
    CUDA_TRY(cudaMemcpyAsync(mask + lo, ctx->scr.flags.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return MBG_OK;
}

int mbg_enumerate(mbg_context_t ctx, int32_t order, const int32_t* set_offsets, const int32_t* set_entities,
                  int32_t* combs, int64_t* n_combs) {
  if (!ctx || !set_offsets || !set_entities || !n_combs) return fail(MBG_ERR_ARG, "NULL argument");
  if (order < 1 || order > kMaxOrder) return fail(MBG_ERR_ARG, "order=%d out of range", order);
  CUDA_TRY(cudaSetDevice(ctx->device));
  JobDev j;
  memset(&j, 0, sizeof(j));
  j.order = order;
  long long prod = 1;
  for (int s = 0; s < order; ++s) {
    j.set_off[s] = set_offsets[s], j.set_off[s + 1] = set_offsets[s + 1];
    const long long n_s = set_offsets[s + 1] - set_offsets[s];
    if (n_s < 0) return fail(MBG_ERR_ARG, "set_offsets not monotone");
    if (n_s == 0) prod = 0;
    else if (prod > (1ll << 60) / n_s) return fail(MBG_ERR_UNSUPPORTED, "candidate space overflows");
    else prod *= n_s;
  }
  const size_t n_int = (size_t)set_offsets[order];
  TRY(ctx->tab.job_ints.ensure((n_int + 16) * sizeof(int)));
  if (n_int) CUDA_TRY(cudaMemcpyAsync(ctx->tab.job_ints.p, set_entities, n_int * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  j.set_ent = ctx->tab.job_ints.as<int>();
  j.cand_per_frame = prod;
  const long long capacity = combs ? *n_combs : 0;
  long long total = 0;
  const long long kChunk = 1ll << 18;
  const long long n_blocks = (prod + kCullThreads - 1) / kCullThreads;
  for (long long b0 = 0; b0 < n_blocks; b0 += kChunk) {
    const int blocks = (int)std::min(kChunk, n_blocks - b0);
    TRY(ctx->scr.flags.ensure((size_t)blocks * kCullThreads));
    TRY(ctx->scr.blk_acc.ensure((size_t)blocks * sizeof(int)));
    TRY(ctx->scr.blk_off.ensure((size_t)blocks * sizeof(long long)));
    TRY(ctx->totals.ensure(2 * sizeof(long long)));
    const long long cand0 = b0 * kCullThreads;
    k_flag_ascending<<<blocks, kCullThreads, 0, ctx->stream>>>(j, cand0, prod, ctx->scr.flags.as<unsigned char>(),
                                                               ctx->scr.blk_acc.as<int>());
    k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->scr.blk_acc.as<int>(), nullptr, blocks, ctx->scr.blk_off.as<long long>(),
                                               ctx->totals.as<long long>());
    ctx->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const long long cnt = ctx->h_totals[0];
    if (combs && cnt > 0) {
      if (total + cnt > capacity) return fail(MBG_ERR_ARG, "combs buffer too small");
      TRY(ctx->scr.rec_atoms.ensure((size_t)cnt * order * sizeof(int)));
      k_write_combs<<<blocks, kCullThreads, 0, ctx->stream>>>(j, cand0, ctx->scr.flags.as<unsigned char>(),
                                                              ctx->scr.blk_off.as<long long>(), 0, ctx->scr.rec_atoms.as<int>());
      ctx->launches++;
      CUDA_TRY(cudaMemcpyAsync(combs + total * order, ctx->scr.rec_atoms.p, (size_t)cnt * order * sizeof(int),
                               cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    total += cnt;
  }
  *n_combs = total;
  return MBG_OK;
}

int mbg_criteria_eval(mbg_context_t ctx, int32_t kind, int32_t n_atoms, const double* masses, const int32_t* entity_ids,
                      const double* R, int64_t n_R, double* values) {
  if (!ctx || !R || !values) return fail(MBG_ERR_ARG, "NULL argument");
  if (n_atoms < 1 || n_R < 0) return fail(MBG_ERR_ARG, "bad sizes");
  if (kind != MBG_CRIT_COM_DISTANCE_SUM && kind != MBG_CRIT_MAX_ATOM_PAIR_DIST)
    return fail(MBG_ERR_UNSUPPORTED, "unknown descriptor kind %d", kind);
  if (kind == MBG_CRIT_COM_DISTANCE_SUM && (!masses || !entity_ids))
    return fail(MBG_ERR_ARG, "com_distance_sum needs masses and entity_ids");
  if (n_R == 0) return MBG_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<int> group(n_atoms, 0);
  int n_groups = 0;
  std::vector<double> m(n_atoms, 1.0);
  if (kind == MBG_CRIT_COM_DISTANCE_SUM) {
    std::vector<int> uniq(entity_ids, entity_ids + n_atoms);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    n_groups = (int)uniq.size();
    for (int a = 0; a < n_atoms; ++a) {
      group[a] = (int)(std::lower_bound(uniq.begin(), uniq.end(), entity_ids[a]) - uniq.begin());
      m[a] = masses[a];
    }
  }
  const size_t nR = (size_t)n_R * n_atoms * 3;
  TRY(ctx->R.ensure(nR * sizeof(double)));
  TRY(ctx->outE.ensure((size_t)n_R * sizeof(double)));
  TRY(ctx->tab.sys_masses.ensure((size_t)n_atoms * sizeof(double)));
  TRY(ctx->tab.sys_ints.ensure((size_t)n_atoms * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(ctx->R.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->tab.sys_masses.p, m.data(), (size_t)n_atoms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->tab.sys_ints.p, group.data(), (size_t)n_atoms * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  k_criteria_eval<<<(unsigned)((n_R + 127) / 128), 128, 0, ctx->stream>>>(kind, n_atoms, ctx->tab.sys_masses.as<double>(),
                                                                          ctx->tab.sys_ints.as<int>(), n_groups,
                                                                          ctx->R.as<double>(), n_R, ctx->outE.as<double>());
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(values, ctx->outE.p, (size_t)n_R * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MBG_OK;
}

int mbg_d_mic(mbg_context_t ctx, const double* cell, double cutoff, const double* d, int32_t n, int32_t check_cutoff,
              double* out, int32_t* accepted, const int32_t* pbc) {
  if (!ctx || !cell || !d || !out || !accepted) return fail(MBG_ERR_ARG, "NULL argument");
  if (n < 1) return fail(MBG_ERR_ARG, "n < 1");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CellDev s;
  TRY(fill_cell(s, cell, cutoff, pbc));
  TRY(ctx->R.ensure((size_t)n * 3 * sizeof(double)));
  TRY(ctx->outF.ensure((size_t)n * 3 * sizeof(double)));
  TRY(ctx->totals.ensure(2 * sizeof(long long)));
  CUDA_TRY(cudaMemcpyAsync(ctx->R.p, d, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_d_mic<<<1, 32, 0, ctx->stream>>>(s, ctx->R.as<double>(), n, check_cutoff ? 1 : 0, ctx->outF.as<double>(), ctx->totals.as<int>());
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, ctx->outF.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *accepted = *reinterpret_cast<int*>(ctx->h_totals);
  return MBG_OK;
}

int mbg_from_r(mbg_context_t ctx, int32_t n_atoms, const double* lattice, const double* R, int64_t n_R, double* x, double* J) {
  if (!ctx || !R || !x || !J) return fail(MBG_ERR_ARG, "NULL argument");
  if (n_atoms < 2 || n_R < 0) return fail(MBG_ERR_ARG, "bad sizes");
  if (n_R == 0) return MBG_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CellDev lat;
  memset(&lat, 0, sizeof(lat));
  if (lattice) {
    memcpy(lat.cell, lattice, 9 * sizeof(double));
    if (!hostmath::inverse3(lat.cell, lat.cell_inv)) return fail(MBG_ERR_ARG, "lattice is singular");
  }
  const long long D = (long long)n_atoms * (n_atoms - 1) / 2, nt = n_R * D;
  const size_t nR = (size_t)n_R * n_atoms * 3;
  TRY(ctx->R.ensure(nR * sizeof(double)));
  TRY(ctx->outE.ensure((size_t)nt * sizeof(double)));
  TRY(ctx->outF.ensure((size_t)nt * 3 * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(ctx->R.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_from_r<<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(n_atoms, lattice ? 1 : 0, lat, ctx->R.as<double>(), n_R,
                                                                ctx->outE.as<double>(), ctx->outF.as<double>());
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(x, ctx->outE.p, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(J, ctx->outF.p, (size_t)nt * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MBG_OK;
}

int mbg_r_mic(mbg_context_t ctx, const double* cell, double cutoff, const double* r, int32_t n, double* out,
              int32_t* accepted, const int32_t* pbc) {
  if (!ctx || !cell || !r || !out || !accepted) return fail(MBG_ERR_ARG, "NULL argument");
  if (n < 1) return fail(MBG_ERR_ARG, "n < 1");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CellDev s;
  TRY(fill_cell(s, cell, cutoff, pbc));
  s.fast_reject = 0;  // the caller wants the coordinates even when the fragment is rejected
  TRY(ctx->R.ensure((size_t)n * 3 * sizeof(double)));
  TRY(ctx->outF.ensure((size_t)n * 3 * sizeof(double)));
  TRY(ctx->totals.ensure(2 * sizeof(long long)));
  CUDA_TRY(cudaMemcpyAsync(ctx->R.p, r, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_r_mic<<<1, 32, 0, ctx->stream>>>(s, ctx->R.as<double>(), n, ctx->outF.as<double>(), ctx->totals.as<int>());
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, ctx->outF.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *accepted = *reinterpret_cast<int*>(ctx->h_totals);
  return MBG_OK;
}

int mbg_measure_fp64_peak(mbg_context_t ctx, double* tflops, double* sm_clock_mhz_est) {
  if (!ctx || !tflops) return fail(MBG_ERR_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(ctx->totals.ensure(2 * sizeof(long long)));
  constexpr int kChains = 8;
  const int threads = 512, blocks = ctx->sm_count * 4, iters = 1 << 16;
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    k_dfma_peak<kChains><<<blocks, threads, 0, ctx->stream>>>(ctx->totals.as<double>(), iters, 1.0 + rep);
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    ctx->launches++;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * kChains * (double)iters * threads * blocks;
    if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  // 64 FP64 FMA lanes per SM per clock on sm_100
  if (sm_clock_mhz_est) *sm_clock_mhz_est = best * 1e12 / (2.0 * 64 * ctx->sm_count) / 1e6;
  return MBG_OK;
}

int mbg_mbe_accumulate(mbg_context_t ctx, int64_t n_ref, const int64_t* ref_rows, const int64_t* pair_offsets,
                       const int64_t* lower_idx, const int32_t* slot_entity, int32_t n_slots, int32_t n_atoms_ref,
                       const int32_t* atom_entity, const int32_t* atom_rank, int32_t n_atoms_lower,
                       const int32_t* lower_off, const int32_t* lower_atoms, const double* E_lower,
                       const double* Deriv_lower, int64_t n_lower, double sign, double* E, double* Deriv,
                       int64_t n_total) {
  if (!ctx || !ref_rows || !pair_offsets || !atom_entity || !atom_rank || !lower_off || !lower_atoms || !E || !Deriv)
    return fail(MBG_ERR_ARG, "NULL argument");
  if (n_ref < 0 || n_slots < 1 || n_atoms_ref < 1 || n_atoms_lower < 1 || n_lower < 0 || n_total < 0)
    return fail(MBG_ERR_ARG, "bad sizes");
  if (!(sign == 1.0 || sign == -1.0)) return fail(MBG_ERR_ARG, "sign must be +1 (add) or -1 (remove)");
  if (n_ref == 0) return MBG_OK;
  const int64_t n_pairs = pair_offsets[n_ref];
  if (pair_offsets[0] != 0 || n_pairs < 0) return fail(MBG_ERR_ARG, "pair_offsets must start at 0");
  if (n_pairs > 0 && (!lower_idx || !slot_entity || !E_lower || !Deriv_lower)) return fail(MBG_ERR_ARG, "NULL argument");
  for (int64_t i = 0; i < n_ref; ++i) {
    if (pair_offsets[i + 1] < pair_offsets[i]) return fail(MBG_ERR_ARG, "pair_offsets not monotone");
    if (ref_rows[i] < 0 || ref_rows[i] >= n_total) return fail(MBG_ERR_ARG, "ref_rows[%lld] out of range", (long long)i);
  }
  for (int64_t p = 0; p < n_pairs; ++p)
    if (lower_idx[p] < 0 || lower_idx[p] >= n_lower) return fail(MBG_ERR_ARG, "lower_idx[%lld] out of range", (long long)p);
  if (lower_off[0] != 0 || lower_off[n_slots] > n_atoms_lower) return fail(MBG_ERR_ARG, "bad lower_off");
  {  // entity sizes must agree: Deriv[i][i_z] += deriv_lower[i_z_lower] (mbe.py:193) needs len(i_z) == len(i_z_lower)
    int n_ent = 0;
    for (int at = 0; at < n_atoms_ref; ++at) {
      if (atom_entity[at] < 0 || atom_rank[at] < 0) return fail(MBG_ERR_ARG, "atom_entity / atom_rank must be >= 0");
      n_ent = std::max(n_ent, atom_entity[at] + 1);
    }
    std::vector<int> ent_size(n_ent, 0);
    for (int at = 0; at < n_atoms_ref; ++at) ent_size[atom_entity[at]]++;
    for (int64_t p = 0; p < n_pairs; ++p)
      for (int k = 0; k < n_slots; ++k) {
        const int e = slot_entity[p * n_slots + k];
        if (e < 0 || e >= n_ent) return fail(MBG_ERR_ARG, "slot_entity[%lld][%d] out of range", (long long)p, k);
        if (ent_size[e] != lower_off[k + 1] - lower_off[k])
          return fail(MBG_ERR_ARG, "entity %d has %d atoms but slot %d of the lower structure has %d", e, ent_size[e], k,
                      lower_off[k + 1] - lower_off[k]);
      }
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  // one staging allocation: [int64 block | int32 block | double block]
  const size_t n64 = (size_t)n_ref + (size_t)n_ref + 1 + (size_t)n_pairs;
  const size_t n32 = (size_t)n_pairs * n_slots + 2 * (size_t)n_atoms_ref + (size_t)n_slots + 1 + (size_t)lower_off[n_slots];
  const size_t nd_low = (size_t)n_lower * (1 + 3 * (size_t)n_atoms_lower);
  const size_t nd_out = (size_t)n_total * (1 + 3 * (size_t)n_atoms_ref);
  const size_t bytes = n64 * 8 + ((n32 * 4 + 7) / 8) * 8 + (nd_low + nd_out) * 8;
  TRY(ctx->scr.flags.ensure(bytes + 64));
  char* base = ctx->scr.flags.as<char>();
  long long* d64 = reinterpret_cast<long long*>(base);
  int* d32 = reinterpret_cast<int*>(base + n64 * 8);
  double* dd = reinterpret_cast<double*>(base + n64 * 8 + ((n32 * 4 + 7) / 8) * 8);
  auto up = [&](void* dst, const void* src, size_t nbytes) {
    return nbytes ? cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
  };
  MbeAccumArgs a;
  a.n_ref = n_ref, a.n_atoms_ref = n_atoms_ref, a.n_atoms_lower = n_atoms_lower, a.n_slots = n_slots;
  long long* q = d64;
  a.ref_rows = q;
  CUDA_TRY(up(q, ref_rows, (size_t)n_ref * 8));
  q += n_ref;
  a.pair_offsets = q;
  CUDA_TRY(up(q, pair_offsets, ((size_t)n_ref + 1) * 8));
  q += n_ref + 1;
  a.lower_idx = q;
  CUDA_TRY(up(q, lower_idx, (size_t)n_pairs * 8));
  int* r = d32;
  a.slot_entity = r;
  CUDA_TRY(up(r, slot_entity, (size_t)n_pairs * n_slots * 4));
  r += (size_t)n_pairs * n_slots;
  a.atom_entity = r;
  CUDA_TRY(up(r, atom_entity, (size_t)n_atoms_ref * 4));
  r += n_atoms_ref;
  a.atom_rank = r;
  CUDA_TRY(up(r, atom_rank, (size_t)n_atoms_ref * 4));
  r += n_atoms_ref;
  a.lower_off = r;
  CUDA_TRY(up(r, lower_off, ((size_t)n_slots + 1) * 4));
  r += n_slots + 1;
  a.lower_atoms = r;
  CUDA_TRY(up(r, lower_atoms, (size_t)lower_off[n_slots] * 4));
  double* w = dd;
  a.E_lower = w;
  CUDA_TRY(up(w, E_lower, (size_t)n_lower * 8));
  w += n_lower;
  a.D_lower = w;
  CUDA_TRY(up(w, Deriv_lower, (size_t)n_lower * n_atoms_lower * 3 * 8));
  w += (size_t)n_lower * n_atoms_lower * 3;
  a.E = w;
  CUDA_TRY(up(w, E, (size_t)n_total * 8));
  w += n_total;
  a.D = w;
  CUDA_TRY(up(w, Deriv, (size_t)n_total * n_atoms_ref * 3 * 8));
  a.sign = sign;
  const long long n_thr = n_ref * ((long long)n_atoms_ref * 3 + 1);
  k_mbe_accumulate<<<(unsigned)((n_thr + 255) / 256), 256, 0, ctx->stream>>>(a);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(E, a.E, (size_t)n_total * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(Deriv, a.D, (size_t)n_total * n_atoms_ref * 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MBG_OK;
}

// ------------------------------------------------------------------------------------
// training-side kernels (kernel_mat.cuh)
// ------------------------------------------------------------------------------------
// One timed launch record for mbg_context_last_launches (kind 2 = kernel matrix, 3 = covariance).
static void record_aux_launch(mbg_context_t ctx, cudaEvent_t e0, cudaEvent_t e1, int na, long long units, int T, int P,
                              int kind) {
  (void)e0, (void)e1;
  ctx->launch_recs.push_back({na, units, T, P, 0.f, kind});
  ctx->n_contract++;
  ctx->launches++;
}

} // extern "C" (reopened below: the helper is a template)
template <int NA>
struct KmdPick {  // the tensor-pipe kernel-matrix kernel exists for fragments of up to 10 atoms (3N <= 32 lanes)
  static void (*kern())(KernelMatArgs, long long) {
    if constexpr (NA <= 10) return k_kernel_mat_d<NA>;
    else return nullptr;
  }
  static size_t bytes(int nw, int P) {
    if constexpr (NA <= 10) return KmdShape<NA>::supported(P) ? KmdShape<NA>::bytes(nw, P) : ~(size_t)0;
    else return ~(size_t)0;
  }
};
extern "C" {
static int kernel_mat_impl(mbg_context_t ctx, int32_t n_atoms, int32_t n_train, int32_t n_perms, const double* R_desc,
                           const double* R_d_desc, const int64_t* tril_perms_lin, double sig, int32_t use_E_cstr,
                           const int32_t* j_list, int32_t n_j, uint32_t flags, double* K) {
  if (!ctx || !R_desc || !R_d_desc || !tril_perms_lin || !K) return fail(MBG_ERR_ARG, "NULL argument");
  if (j_list) {
    if (use_E_cstr) return fail(MBG_ERR_UNSUPPORTED, "kernel matrix: column subsets with energy constraints");
    if (n_j < 1) return fail(MBG_ERR_ARG, "empty column-block list");
    for (int k = 0; k < n_j; ++k)
      if (j_list[k] < 0 || j_list[k] >= n_train) return fail(MBG_ERR_ARG, "column block %d out of range", (int)j_list[k]);
  }
  if (n_atoms < 2 || n_atoms > 18) return fail(MBG_ERR_UNSUPPORTED, "kernel matrix: n_atoms=%d outside 2..18", n_atoms);
  if (n_train < 1 || n_perms < 1) return fail(MBG_ERR_ARG, "n_train and n_perms must be positive");
  if (!(sig > 0)) return fail(MBG_ERR_ARG, "sig must be positive");
  if (flags & ~(uint32_t)MBG_FLAG_OUT_ON_DEVICE) return fail(MBG_ERR_UNSUPPORTED, "flag not supported by mbg_kernel_mat");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int N = n_atoms, D = N * (N - 1) / 2, T = n_train, P = n_perms, N3 = 3 * N;
  // pi_p(k) = tril_perms_lin[k P + p] mod D (train.py:163-167); must be a permutation
  std::vector<int> tab((size_t)2 * P * D);
  int* pi = tab.data();
  int* ipi = tab.data() + (size_t)P * D;
  for (int p = 0; p < P; ++p) {
    std::vector<char> seen(D, 0);
    for (int k = 0; k < D; ++k) {
      const int64_t lin = tril_perms_lin[(size_t)k * P + p];
      if (lin < 0 || lin >= (int64_t)P * D) return fail(MBG_ERR_ARG, "tril_perms_lin[%d] out of range", k * P + p);
      const int e = (int)(lin % D);
      if (seen[e]) return fail(MBG_ERR_UNSUPPORTED, "tril_perms_lin column %d is not a permutation", p);
      seen[e] = 1;
      pi[(size_t)p * D + k] = e, ipi[(size_t)p * D + e] = k;
    }
  }
  const size_t rows = (size_t)T * N3 + (use_E_cstr ? T : 0);
  const size_t cols = j_list ? (size_t)n_j * N3 : rows;
  const size_t nX = (size_t)T * D, nG = (size_t)T * D * 3;
  begin_call(ctx);
  TRY(ctx->km_in.ensure((nX + nG) * sizeof(double)));
  TRY(ctx->km_ints.ensure(tab.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(ctx->km_in.p, R_desc, nX * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->km_in.as<double>() + nX, R_d_desc, nG * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->km_ints.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  double* dK = K;
  const bool out_dev = (flags & MBG_FLAG_OUT_ON_DEVICE) != 0;
  if (!out_dev) {
    TRY(ctx->km_out.ensure(rows * cols * sizeof(double)));
    dK = ctx->km_out.as<double>();
  }
  // closed under inversion?  (pi_q == pi_p^-1 for some q, for every p)
  bool closed = true;
  for (int p = 0; p < P && closed; ++p) {
    bool found = false;
    for (int q = 0; q < P && !found; ++q) found = std::equal(ipi + (size_t)p * D, ipi + (size_t)(p + 1) * D, pi + (size_t)q * D);
    closed = found;
  }
  KernelMatArgs a;
  if (j_list) closed = false;  // every (i, j in list) block is computed on its own
  a.T = T, a.P = P, a.sig = sig, a.use_E = use_E_cstr ? 1 : 0, a.symmetric = closed ? 1 : 0;
  a.X = ctx->km_in.as<double>(), a.G = ctx->km_in.as<double>() + nX;
  a.pi = ctx->km_ints.as<int>(), a.ipi = ctx->km_ints.as<int>() + (size_t)P * D;
  a.K = dK, a.ld = (long long)cols;
  a.srt_p = nullptr, a.srt_k = nullptr;
  a.j_list = nullptr, a.n_j = 0;
  if (j_list) {
    TRY(ctx->km_jlist.ensure((size_t)n_j * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(ctx->km_jlist.p, j_list, (size_t)n_j * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    a.j_list = ctx->km_jlist.as<int>(), a.n_j = n_j;
  }
  const long long n_blocks = closed ? (long long)T * (T + 1) / 2 : (long long)T * (j_list ? n_j : T);
  if (n_blocks > 0x7fffffffLL) return fail(MBG_ERR_UNSUPPORTED, "kernel matrix: too many blocks for one launch");
  // warp-per-pair kernel for narrow descriptors (D <= 15: measured 4.3x / 1.24x faster than CTA-per-pair for water
  // monomers / dimers at T = 1000, 0.93x for trimers, where its 12 warps per SM are too few to hide the latency of
  // the shared-memory chains: profiles/r2_train_kernels.md), else CTA per pair
  int dev_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  void (*kern)(KernelMatArgs) = nullptr;
  void (*kern_w)(KernelMatArgs, long long) = nullptr;
  void (*kern_d)(KernelMatArgs, long long) = nullptr;
  size_t smem = 0, smem_w4 = 0, smem_w = 0, smem_d = 0;
  int nw = 0, nw_d = 0;
  switch (N) {
#define X(NA_)                                                                                   \
  case NA_:                                                                                      \
    kern = k_kernel_mat<NA_>, smem = KmShape<NA_>::kBytes;                                       \
    kern_w = k_kernel_mat_w<NA_>, smem_w4 = KmwShape<NA_>::bytes(4, P);                          \
    for (nw = kKmwMaxWarps; nw > 4 && KmwShape<NA_>::bytes(nw, P) > (size_t)dev_smem; --nw) {}   \
    smem_w = KmwShape<NA_>::bytes(nw, P);                                                        \
    kern_d = KmdPick<NA_>::kern();                                                               \
    for (nw_d = kKmdMaxWarps; nw_d > 0 && KmdPick<NA_>::bytes(nw_d, P) > (size_t)dev_smem; --nw_d) {} \
    smem_d = KmdPick<NA_>::bytes(nw_d, P);                                                       \
    break;
    MBG_FOR_EACH_NA(X)
#undef X
    default: return fail(MBG_ERR_UNSUPPORTED, "kernel matrix: no kernel for %d atoms", N);
  }
  // MBG_MMA_CFG=4 / 5 (developer): always CTA per pair / always warp per pair when it fits
  // tensor-pipe kernel (warp per pair, every sum a DMMA chain) for wider descriptors, up to 10 atoms; MBG_MMA_CFG=7
  // (developer): whenever it exists
  const bool tensor = kern_d && nw_d >= 4 && ctx->mma_cfg != 4 && ctx->mma_cfg != 5 && (D >= 10 || ctx->mma_cfg == 7);
  const bool per_warp = !tensor && smem_w4 <= (size_t)dev_smem && ctx->mma_cfg != 4 && (D <= 15 || ctx->mma_cfg == 5);
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  if (tensor) {
    // per descriptor k: the permutations ordered by their image pi_p(k) (stable), for the scatter-free build of C
    std::vector<int> srt((size_t)2 * P * D);
    std::vector<int> order(P);
    for (int k = 0; k < D; ++k) {
      for (int q = 0; q < P; ++q) order[q] = q;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pi[(size_t)x * D + k] < pi[(size_t)y * D + k]; });
      for (int t = 0; t < P; ++t) srt[(size_t)t * D + k] = order[t], srt[(size_t)P * D + (size_t)t * D + k] = pi[(size_t)order[t] * D + k];
    }
    TRY(ctx->km_ints2.ensure(srt.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(ctx->km_ints2.p, srt.data(), srt.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // srt is a local
    a.srt_p = ctx->km_ints2.as<int>(), a.srt_k = ctx->km_ints2.as<int>() + (size_t)P * D;
    CUDA_TRY(cudaFuncSetAttribute(kern_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
    const long long ctas = std::min<long long>((n_blocks + nw_d - 1) / nw_d, (long long)ctx->sm_count);
    cudaEventRecord(e0, ctx->stream);
    kern_d<<<(unsigned)ctas, nw_d * 32, smem_d, ctx->stream>>>(a, n_blocks);
    cudaEventRecord(e1, ctx->stream);
  } else if (per_warp) {
    CUDA_TRY(cudaFuncSetAttribute(kern_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    const long long ctas = std::min<long long>((n_blocks + nw - 1) / nw, (long long)ctx->sm_count * 4);
    cudaEventRecord(e0, ctx->stream);
    kern_w<<<(unsigned)ctas, nw * 32, smem_w, ctx->stream>>>(a, n_blocks);
    cudaEventRecord(e1, ctx->stream);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEventRecord(e0, ctx->stream);
    kern<<<(unsigned)n_blocks, kKmThreads, smem, ctx->stream>>>(a);
    cudaEventRecord(e1, ctx->stream);
  }
  CUDA_TRY(cudaGetLastError());
  record_aux_launch(ctx, e0, e1, N, (long long)T * (j_list ? n_j : T), T, P, 2);
  if (!out_dev) {
    CUDA_TRY(cudaMemcpyAsync(K, dK, rows * cols * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  } else if (j_list) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // j_list is the caller's memory
  }
  end_call_timing(ctx);
  return MBG_OK;
}

int mbg_kernel_mat(mbg_context_t ctx, int32_t n_atoms, int32_t n_train, int32_t n_perms, const double* R_desc,
                   const double* R_d_desc, const int64_t* tril_perms_lin, double sig, int32_t use_E_cstr,
                   uint32_t flags, double* K) {
  return kernel_mat_impl(ctx, n_atoms, n_train, n_perms, R_desc, R_d_desc, tril_perms_lin, sig, use_E_cstr, nullptr, 0, flags, K);
}

int mbg_kernel_mat_cols(mbg_context_t ctx, int32_t n_atoms, int32_t n_train, int32_t n_perms, const double* R_desc,
                        const double* R_d_desc, const int64_t* tril_perms_lin, double sig, const int32_t* col_blocks,
                        int32_t n_col_blocks, uint32_t flags, double* K) {
  if (!col_blocks) return fail(MBG_ERR_ARG, "NULL argument");
  return kernel_mat_impl(ctx, n_atoms, n_train, n_perms, R_desc, R_d_desc, tril_perms_lin, sig, 0, col_blocks, n_col_blocks, flags, K);
}

static int launch_mat52(mbg_context_t ctx, Mat52Args a, int64_t n_q, int na_rec) {
  if (n_q > 65535) return fail(MBG_ERR_UNSUPPORTED, "covariance: more than 65535 query structures per call");
  a.rows_per_tile = mat52_rows_per_tile(a.D, a.P, a.T);
  a.stage_perms = mat52_smem_bytes(a.D, a.P, true, a.rows_per_tile) <= 96 * 1024 ? 1 : 0;  // two CTAs per SM at least
  const size_t smem = mat52_smem_bytes(a.D, a.P, a.stage_perms != 0, a.rows_per_tile);
  if (smem > 200 * 1024) return fail(MBG_ERR_UNSUPPORTED, "covariance: descriptor dimension %d too large", a.D);
  CUDA_TRY(cudaFuncSetAttribute(k_mat52, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx);
  cudaEventRecord(e0, ctx->stream);
  k_mat52<<<dim3((a.T + a.rows_per_tile - 1) / a.rows_per_tile, (unsigned)n_q), kM52Threads, smem, ctx->stream>>>(a);
  cudaEventRecord(e1, ctx->stream);
  CUDA_TRY(cudaGetLastError());
  record_aux_launch(ctx, e0, e1, na_rec, (long long)n_q * a.T * a.P, a.T, a.P, 3);
  return MBG_OK;
}

int mbg_mat52(mbg_context_t ctx, mbg_model_t m, const double* R, int64_t n_R, uint32_t flags, double* out) {
  if (!ctx || !m || !R || !out) return fail(MBG_ERR_ARG, "NULL argument");
  if (n_R < 0) return fail(MBG_ERR_ARG, "negative count");
  if (flags & ~(uint32_t)(MBG_FLAG_R_ON_DEVICE | MBG_FLAG_OUT_ON_DEVICE))
    return fail(MBG_ERR_UNSUPPORTED, "flag not supported by mbg_mat52");
  if (n_R == 0) return MBG_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  begin_call(ctx);
  const size_t nR = (size_t)n_R * m->n_atoms * 3, nOut = (size_t)n_R * m->T * m->P;
  const double* dR = R;
  if (!(flags & MBG_FLAG_R_ON_DEVICE)) {
    TRY(ctx->km_in.ensure(nR * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(ctx->km_in.p, R, nR * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dR = ctx->km_in.as<double>();
  }
  double* dOut = out;
  const bool out_dev = (flags & MBG_FLAG_OUT_ON_DEVICE) != 0;
  if (!out_dev) {
    TRY(ctx->km_out.ensure(nOut * sizeof(double)));
    dOut = ctx->km_out.as<double>();
  }
  Mat52Args a;
  a.n_atoms = m->n_atoms, a.D = m->D, a.T = m->T, a.P = m->P, a.sig = m->sig;
  a.Q = dR, a.q_is_desc = 0;
  a.X = reinterpret_cast<const double*>(m->XA), a.x_row_stride = 2ll * model_row_pad(m->n_atoms), a.x_elem_stride = 2;
  a.iperm = m->iperm, a.out = dOut;
  TRY(launch_mat52(ctx, a, n_R, m->n_atoms));
  if (!out_dev) {
    CUDA_TRY(cudaMemcpyAsync(out, dOut, nOut * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  end_call_timing(ctx);
  return MBG_OK;
}

int mbg_mat52_rows(mbg_context_t ctx, int32_t dim_d, const double* r_desc, int64_t n_q, double sig, const double* rows,
                   int64_t n_rows, double* out) {
  if (!ctx || !r_desc || !rows || !out) return fail(MBG_ERR_ARG, "NULL argument");
  if (dim_d < 1 || n_q < 0 || n_rows < 0 || n_rows > 0x7fffffff) return fail(MBG_ERR_ARG, "bad sizes");
  if (!(sig > 0)) return fail(MBG_ERR_ARG, "sig must be positive");
  if (n_q == 0 || n_rows == 0) return MBG_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  begin_call(ctx);
  const size_t nQ = (size_t)n_q * dim_d, nX = (size_t)n_rows * dim_d, nOut = (size_t)n_q * n_rows;
  TRY(ctx->km_in.ensure((nQ + nX) * sizeof(double)));
  TRY(ctx->km_out.ensure(nOut * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(ctx->km_in.p, r_desc, nQ * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->km_in.as<double>() + nQ, rows, nX * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  Mat52Args a;
  a.n_atoms = 0, a.D = dim_d, a.T = (int)n_rows, a.P = 1, a.sig = sig;
  a.Q = ctx->km_in.as<double>(), a.q_is_desc = 1;
  a.X = ctx->km_in.as<double>() + nQ, a.x_row_stride = dim_d, a.x_elem_stride = 1;
  a.iperm = nullptr, a.out = ctx->km_out.as<double>();
  TRY(launch_mat52(ctx, a, n_q, 0));
  CUDA_TRY(cudaMemcpyAsync(out, a.out, nOut * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  end_call_timing(ctx);
  return MBG_OK;
}

int mbg_measure_exp_sqrt_rate(mbg_context_t ctx, double* ops_per_s) {
  if (!ctx || !ops_per_s) return fail(MBG_ERR_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(ctx->totals.ensure(2 * sizeof(long long)));
  const int threads = 256, blocks = ctx->sm_count * 4, iters = 1 << 12;
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    k_exp_sqrt_rate<<<blocks, threads, 0, ctx->stream>>>(ctx->totals.as<double>(), iters, 1.0 + rep);
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    ctx->launches++;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::max(best, 4.0 * iters * threads * blocks / (ms * 1e-3));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ops_per_s = best;
  return MBG_OK;
}

int mbg_measure_fp64_tensor_peak(mbg_context_t ctx, double* tflops) {
  if (!ctx || !tflops) return fail(MBG_ERR_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(ctx->totals.ensure(2 * sizeof(long long)));
  constexpr int kChains = 8;
  const int threads = 256, blocks = ctx->sm_count * 2, iters = 1 << 14;
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    k_dmma_peak<kChains><<<blocks, threads, 0, ctx->stream>>>(ctx->totals.as<double>(), iters, 1.0 + rep);
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    ctx->launches++;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 DMMA = 8 x 8 x 4 multiply-adds per warp
    const double flop = 2.0 * 256 * kChains * (double)iters * (threads / 32) * blocks;
    if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return MBG_OK;
}

}  // extern "C"
