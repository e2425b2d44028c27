// Standalone timing harness for k_matern_mma (csrc/matern_mma.cuh): synthetic trimers against a random model,
// CUDA-event timed, so that kernel experiments compile in seconds instead of rebuilding the whole library.
// Results are NOT checked here (parity lives in tests/); this only reports time and algorithmic TFLOP/s.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo [-DEXP_...] -o tools/mma_harness tools/mma_harness.cu
//   tools/mma_harness [n_frag] [T]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef H_NA
#define H_NA 9   // atoms per fragment
#define H_P 48   // permutations
#define H_MT 3   // m-tiles per warp
#endif
#include "../mbgdml_b200/csrc/common.cuh"
#include "../mbgdml_b200/csrc/geom.cuh"
#include "../mbgdml_b200/csrc/matern_mma.cuh"
#include "../mbgdml_b200/csrc/matern_pw.cuh"

using namespace mbg;

#define CK(x)                                                                                    \
  do {                                                                                           \
    cudaError_t e = (x);                                                                         \
    if (e != cudaSuccess) {                                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);             \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

static double frand() { return rand() / (double)RAND_MAX; }

template <int MT, int NTHR, bool YS, int MINB = 1>
static int run(const char* name, MaternMmaArgs a, int reps, bool persistent = false) {
  constexpr int NA = H_NA, D = NA * (NA - 1) / 2;
  auto kern = k_matern_mma<NA, MT, NTHR, MINB, false, YS>;
  const size_t smem = mma_smem_bytes<NA, MT, YS>(NTHR, a.P);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nq = a.n_frag * a.P;
  const int qc = mma_queries_per_cta<NA, MT>(NTHR);
  unsigned grid = (unsigned)((nq + qc - 1) / qc);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHR, smem);
  if (persistent) {  // resident CTAs walk the batches; the kernel picks the row slices
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    grid = (unsigned)(sms * occ), a.persistent = 1, a.max_slices = 16;
  }
  CK(cudaMemset(a.E, 0, 8));
  printf("[occupancy %d CTA/SM] ", occ);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  kern<<<grid, NTHR, smem>>>(a);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    kern<<<grid, NTHR, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flop = (double)a.n_frag * 1000.0 * a.P * (9 * D + 10);
  double E_sum;
  CK(cudaMemcpy(&E_sum, a.E, 8, cudaMemcpyDeviceToHost));
  printf("%-28s grid %6u  smem %6zu  %8.3f ms  %6.2f TFLOP/s (algorithmic, T=1000 formula scaled by rows %d)  E(%d launches) = %.9e\n", name, grid,
         smem, best, flop * (a.n_row_blocks * 8 / 1000.0) / best / 1e9, a.n_row_blocks * 8, reps + 1, E_sum);
  return 0;
}

// persistent per-warp kernel (matern_pw.cuh): one CTA per SM, work counters zeroed before every launch
template <int MT, int NW, bool YS>
static int run_pw(const char* name, const MaternMmaArgs& m, int reps, int sm_count, int nw_launch = NW) {
  constexpr int NA = H_NA, D = NA * (NA - 1) / 2;
  MaternPwArgs a;
  memset(&a, 0, sizeof(a));
  a.sys = m.sys, a.R = m.R, a.tiles = m.tiles, a.mu = m.mu, a.exp2_table = m.exp2_table, a.iperm = m.iperm;
  {  // forward permutations from the inverse table
    std::vector<int> ip((size_t)m.P * D), fp((size_t)m.P * D);
    CK(cudaMemcpy(ip.data(), m.iperm, ip.size() * 4, cudaMemcpyDeviceToHost));
    for (int p = 0; p < m.P; ++p)
      for (int e = 0; e < D; ++e) fp[p * D + ip[p * D + e]] = e;
    int* dfp;
    CK(cudaMalloc(&dfp, fp.size() * 4));
    CK(cudaMemcpy(dfp, fp.data(), fp.size() * 4, cudaMemcpyHostToDevice));
    a.perm = dfp;
  }
  a.P = m.P, a.n_row_blocks = m.n_row_blocks, a.max_slices = 1;
#ifdef EXP_SLICES
  a.max_slices = EXP_SLICES;
#endif
  a.sig = m.sig, a.c = m.c, a.std = m.std, a.n_frag = m.n_frag, a.n_frag_dev = nullptr;
  a.rec_frame = m.rec_frame, a.rec_atoms = m.rec_atoms, a.rec_lambda = m.rec_lambda, a.rec_slot = m.rec_slot;
  a.n_slots_per_frame = m.n_slots_per_frame, a.decomp = 0, a.want_virial = 0;
  a.E = m.E, a.F = m.F, a.virial = nullptr;
  unsigned int* ctr;
  CK(cudaMalloc(&ctr, kPwMaxSlices * 4));
  a.work_counters = ctr;
#ifdef PW_PROFILE
  long long* dprof;
  CK(cudaMalloc(&dprof, (size_t)sm_count * NW * 8 * 8));
  CK(cudaMemset(dprof, 0, (size_t)sm_count * NW * 8 * 8));
  a.prof = dprof;
#endif
  auto kern = k_matern_pw<NA, MT, NW, false, YS>;
  const size_t smem = pw_smem_bytes<NA, MT, NW, YS>(a.P, nw_launch);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  CK(cudaMemset(m.E, 0, 8));
  CK(cudaMemset(ctr, 0, kPwMaxSlices * 4));
  kern<<<sm_count, nw_launch * 32, smem>>>(a);
  CK(cudaDeviceSynchronize());
  double E;
  CK(cudaMemcpy(&E, m.E, 8, cudaMemcpyDeviceToHost));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemset(ctr, 0, kPwMaxSlices * 4));
    cudaEventRecord(e0);
    kern<<<sm_count, nw_launch * 32, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flop = (double)a.n_frag * 1000.0 * a.P * (9 * D + 10);
  printf("%-28s grid %6d  smem %6zu  %8.3f ms  %6.2f TFLOP/s   E(one launch) = %.9e\n", name, sm_count, smem, best,
         flop * (a.n_row_blocks * 8 / 1000.0) / best / 1e9, E);
#ifdef PW_PROFILE
  {
    std::vector<long long> hp((size_t)sm_count * NW * 8);
    CK(cudaMemcpy(hp.data(), dprof, hp.size() * 8, cudaMemcpyDeviceToHost));
    double tot = 0, pro = 0, loop = 0, epi = 0, acq = 0, items = 0, epi1 = 0;
    int n = 0;
    for (int c = 0; c < sm_count; ++c)
      for (int w = 0; w < nw_launch; ++w) {
        const long long* o = hp.data() + ((size_t)c * NW + w) * 8;
        tot += o[0], pro += o[1], loop += o[2], epi += o[3], acq += o[4], items += o[5], epi1 += o[6], ++n;
      }
    printf("   per warp (cycles, mean over %d warps): total %.0f  prologue %.0f  loop %.0f  epilogue %.0f  [acquire waits inside: %.0f]  items %.1f\n",
           n, tot / n, pro / n, loop / n, epi / n, acq / n, items / n);
    printf("   per item: prologue %.0f  loop %.0f  epilogue %.0f (stage+reduce %.0f)  acquire %.0f cycles\n", pro / items, loop / items, epi / items, epi1 / items, acq / items);
    const long long* o = hp.data();
    for (int w = 0; w < nw_launch; ++w)
      printf("   CTA 0 warp %2d: total %lld pro %lld loop %lld epi %lld acq %lld items %lld\n", w, o[w * 8], o[w * 8 + 1], o[w * 8 + 2], o[w * 8 + 3], o[w * 8 + 4], o[w * 8 + 5]);
  }
#endif
  cudaFree(ctr);
  return 0;
}

int main(int argc, char** argv) {
  const long long n_frag = argc > 1 ? atoll(argv[1]) : 41664;
  const int T = argc > 2 ? atoi(argv[2]) : 1000;
  constexpr int NA = H_NA, D = NA * (NA - 1) / 2, P = H_P, NM = NA / 3;
  const int n_atoms = 192;
  srand(1);
  std::vector<double> R((size_t)n_atoms * 3);
  for (int m = 0; m < 64; ++m) {  // 4 x 4 x 4 grid of "molecules", 3 atoms each ~1 A apart
    const double c[3] = {3.1 * (m % 4), 3.1 * ((m / 4) % 4), 3.1 * (m / 16)};
    for (int k = 0; k < 3; ++k)
      for (int x = 0; x < 3; ++x) R[(m * 3 + k) * 3 + x] = c[x] + (k == x ? 0.96 : 0.0) + 0.1 * frand();
  }
  std::vector<int> rec_frame(n_frag, 0), rec_atoms((size_t)n_frag * NA);
  std::vector<double> rec_lambda(n_frag, 1.0);
  std::vector<long long> rec_slot(n_frag);
  for (long long f = 0; f < n_frag; ++f) {
    int m[NM];
    for (int k = 0; k < NM; ++k) {
      bool dup;
      do {
        m[k] = rand() % 64, dup = false;
        for (int j = 0; j < k; ++j) dup |= m[j] == m[k];
      } while (dup);
    }
    for (int k = 0; k < NM; ++k)
      for (int x = 0; x < 3; ++x) rec_atoms[f * NA + k * 3 + x] = m[k] * 3 + x;
    rec_slot[f] = f;
  }
  const int TR = mma_tile_rows(NA), Dp = mma_row_pad(NA), TILE = mma_tile_doubles(NA);
  const int T_pad = (T + 63) / 64 * 64;
  std::vector<double> tiles((size_t)(T_pad / TR) * TILE, 0.0), mu(D, 0.3);
  for (int t = 0; t < T; ++t) {
    double* tile = tiles.data() + (size_t)(t / TR) * TILE;
    const int r = t % TR;
    double xx = 0, xa = 0;
    for (int e = 0; e < D; ++e) {
      const double xv = 0.2 * (frand() - 0.5), av = frand() - 0.5;
      tile[r * Dp + e] = xv, tile[TR * Dp + r * Dp + e] = av;
      xx += xv * xv, xa += xv * av;
    }
    tile[r * Dp + D] = 1.0;
    const int pos = (r / 8) * 8 + mma_tau(r % 8);
    tile[2 * TR * Dp + 2 * pos] = -0.5 * xx, tile[2 * TR * Dp + 2 * pos + 1] = -xa;
  }
  std::vector<double> etab(kMmaExpTable);
  for (int j = 0; j < kMmaExpTable; ++j) etab[j] = exp2((double)j / kMmaExpTable);
  std::vector<int> iperm((size_t)P * D);
  for (int p = 0; p < P; ++p)
    for (int e = 0; e < D; ++e) iperm[p * D + e] = (e + p) % D;  // some permutation per p

  MaternMmaArgs a;
  memset(&a, 0, sizeof(a));
  a.sys.n_atoms = n_atoms, a.sys.periodic = 0;
  double *dR, *dT, *dMu, *dE, *dF, *dTab, *dLam;
  int *dIp, *dFr, *dAt;
  long long* dSl;
  CK(cudaMalloc(&dR, R.size() * 8)); CK(cudaMalloc(&dT, tiles.size() * 8)); CK(cudaMalloc(&dMu, D * 8));
  CK(cudaMalloc(&dE, 8)); CK(cudaMalloc(&dF, R.size() * 8)); CK(cudaMalloc(&dTab, etab.size() * 8));
  CK(cudaMalloc(&dLam, n_frag * 8)); CK(cudaMalloc(&dIp, iperm.size() * 4)); CK(cudaMalloc(&dFr, n_frag * 4));
  CK(cudaMalloc(&dAt, rec_atoms.size() * 4)); CK(cudaMalloc(&dSl, n_frag * 8));
  CK(cudaMemcpy(dR, R.data(), R.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dT, tiles.data(), tiles.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dMu, mu.data(), D * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dTab, etab.data(), etab.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dLam, rec_lambda.data(), n_frag * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dIp, iperm.data(), iperm.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dFr, rec_frame.data(), n_frag * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAt, rec_atoms.data(), rec_atoms.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dSl, rec_slot.data(), n_frag * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dE, 0, 8)); CK(cudaMemset(dF, 0, R.size() * 8));
  a.R = dR, a.tiles = dT, a.mu = dMu, a.exp2_table = dTab, a.iperm = dIp;
  a.T_pad = T_pad, a.P = P, a.n_row_blocks = (T + 7) / 8, a.tiles_per_slice = T_pad / TR, a.main_ctas = 0x7fffffff, a.tail_slices = 1, a.tail_tiles_per_slice = T_pad / TR;
  a.sig = 100.0, a.c = 0.0, a.std = 1.0, a.n_frag = n_frag;
  a.rec_frame = dFr, a.rec_atoms = dAt, a.rec_lambda = dLam, a.rec_slot = dSl;
  a.n_slots_per_frame = (int)n_frag, a.decomp = 0, a.want_virial = 0;
  a.E = dE, a.F = dF, a.virial = nullptr;
#ifdef EXP_NAME
  printf("experiment: %s\n", EXP_NAME);
#endif
#if H_NA != 9
  if (run<H_MT, 256, false>("wave 256 thr", a, 5)) return 1;
  if (run<H_MT, 256, false>("wave 256 thr, persistent", a, 5, true)) return 1;
  int sm_count_w = 0;
  cudaDeviceGetAttribute(&sm_count_w, cudaDevAttrMultiProcessorCount, 0);
  if (run_pw<H_MT, 8, false>("persistent 8 warps", a, 5, sm_count_w)) return 1;
  return 0;
#elif defined(EXP_TWO_CTAS)
  if (run<3, 192, true, 2>("MT=3 192 thr x 2 CTA/SM, y' smem", a, 5)) return 1;
#else
  if (run<3, 256, false>("MT=3 256 thr", a, 5)) return 1;
  CK(cudaMemset(dE, 0, 8));
  if (run<3, 384, true>("MT=3 384 thr, y' in smem", a, 0)) return 1;
  {
    double E1;
    CK(cudaMemcpy(&E1, dE, 8, cudaMemcpyDeviceToHost));
    printf("wave kernel                  E(one launch) = %.9e\n", E1);
  }
  if (run<3, 384, true>("MT=3 384 thr, y' in smem", a, 5)) return 1;
#endif
  int sm_count = 0;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, 0);
  if (run<3, 384, true>("CTA kernel, persistent mode", a, 5, true)) return 1;
  if (run_pw<3, 12, true>("persistent 12 warps, y' smem", a, 5, sm_count)) return 1;
#ifdef EXP_16W
  if (run_pw<2, 16, true>("persistent 16 warps MT=2 ys", a, 5, sm_count)) return 1;
  if (run<2, 512, true>("wave MT=2 512 thr, y' smem", a, 5)) return 1;
#endif
  if (run_pw<3, 8, false>("persistent 8 warps", a, 5, sm_count)) return 1;
  double E;
  CK(cudaMemcpy(&E, dE, 8, cudaMemcpyDeviceToHost));
  printf("E (sum over all launches) = %.6e\n", E);
  return 0;
}
