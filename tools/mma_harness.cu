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

#include "../mbgdml_b200/csrc/common.cuh"
#include "../mbgdml_b200/csrc/geom.cuh"
#include "../mbgdml_b200/csrc/matern_mma.cuh"

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
static int run(const char* name, MaternMmaArgs a, int reps) {
  constexpr int NA = 9;
  auto kern = k_matern_mma<NA, MT, NTHR, MINB, false, YS>;
  const size_t smem = mma_smem_bytes<NA, MT, YS>(NTHR, a.P);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nq = a.n_frag * a.P;
  const int qc = mma_queries_per_cta<NA, MT>(NTHR);
  const unsigned grid = (unsigned)((nq + qc - 1) / qc);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHR, smem);
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
  const double flop = (double)a.n_frag * 1000.0 * a.P * (9 * 36 + 10);
  printf("%-28s grid %6u  smem %6zu  %8.3f ms  %6.2f TFLOP/s (algorithmic, T=1000 formula scaled by rows %d)\n", name, grid,
         smem, best, flop * (a.n_row_blocks * 8 / 1000.0) / best / 1e9, a.n_row_blocks * 8);
  return 0;
}

int main(int argc, char** argv) {
  const long long n_frag = argc > 1 ? atoll(argv[1]) : 41664;
  const int T = argc > 2 ? atoi(argv[2]) : 1000;
  constexpr int NA = 9, D = 36, P = 48;
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
    int m[3];
    do {
      m[0] = rand() % 64, m[1] = rand() % 64, m[2] = rand() % 64;
    } while (m[0] == m[1] || m[0] == m[2] || m[1] == m[2]);
    for (int k = 0; k < 3; ++k)
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
#ifdef EXP_TWO_CTAS
  if (run<3, 192, true, 2>("MT=3 192 thr x 2 CTA/SM, y' smem", a, 5)) return 1;
#else
  if (run<3, 256, false>("MT=3 256 thr", a, 5)) return 1;
  if (run<3, 384, true>("MT=3 384 thr, y' in smem", a, 5)) return 1;
#endif
  double E;
  CK(cudaMemcpy(&E, dE, 8, cudaMemcpyDeviceToHost));
  printf("E (sum over all launches) = %.6e\n", E);
  return 0;
}
