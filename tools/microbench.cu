// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) vs mixed, exp/sqrt cost.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Evidence for DESIGN.md ("tensor cores only if they beat the FMA pipe").
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int CH>
__global__ void k_dfma(double* out, int iters, double seed) {
  double acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = seed + k + threadIdx.x * 1e-3;
  const double m = 1.0 + seed * 1e-9, b = seed * 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = fma(acc[k], m, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void k_dadd(double* out, int iters, double seed) {
  double acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = seed + k + threadIdx.x * 1e-3;
  const double b = seed * 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = __dadd_rn(acc[k], b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per thread
template <int CH>
__global__ void k_dmma884(double* out, int iters, double seed) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) c0[k] = seed + k, c1[k] = seed - k;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c0[k] + c1[k];
  if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
template <int CH>
__global__ void k_dmma16816(double* out, int iters, double seed) {
  double c[CH][4];
#pragma unroll
  for (int k = 0; k < CH; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[k][j] = seed + k + j;
  double a[8], b[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 1.0 + (threadIdx.x + j) * 1e-6;
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 1.0 - (threadIdx.x + j) * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
          "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
          : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
          : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
            "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  if (s == 123.456) out[0] = s;
}

// mixed: per iteration CH DMMA m8n8k4 + FM DFMA (independent chains) in the same warp
template <int CH, int FM>
__global__ void k_mixed(double* out, int iters, double seed) {
  double c0[CH], c1[CH], acc[FM];
#pragma unroll
  for (int k = 0; k < CH; ++k) c0[k] = seed + k, c1[k] = seed - k;
#pragma unroll
  for (int k = 0; k < FM; ++k) acc[k] = seed + k;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  const double m = 1.0 + seed * 1e-9, bb = seed * 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
#pragma unroll
    for (int k = 0; k < FM; ++k) acc[k] = fma(acc[k], m, bb);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c0[k] + c1[k];
#pragma unroll
  for (int k = 0; k < FM; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;
}

template <int MODE>  // 0: exp, 1: sqrt, 2: exp+sqrt chain as in the Matern kernel
__global__ void k_transc(double* out, int iters, double seed) {
  double x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = seed * 0.1 + k * 0.01 + threadIdx.x * 1e-4;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (MODE == 0) x[k] = exp(-x[k]) + 0.25;
      if (MODE == 1) x[k] = sqrt(x[k]) + 0.25;
      if (MODE == 2) x[k] = exp(-2.23606797749979 * sqrt(x[k]) * 0.01) + 0.25;
    }
  }
  double s = x[0] + x[1] + x[2] + x[3];
  if (s == 123.456) out[0] = s;
}

template <class F>
double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s, %d SMs, cc %d.%d, L2 %d MB, smem/block optin %zu\n", p.name, p.multiProcessorCount, p.major,
         p.minor, p.l2CacheSize >> 20, p.sharedMemPerBlockOptin);
  double* d;
  CK(cudaMalloc(&d, 64));
  const int sm = p.multiProcessorCount;
  const int iters = 1 << 15;
  auto rep = [&](const char* name, double flop, double ms) {
    printf("%-44s %9.3f ms  %8.2f TFLOP/s (or Gop/s x1e-3)\n", name, ms, flop / (ms * 1e-3) / 1e12);
  };
  for (int threads : {128, 256, 512, 1024}) {
    const int blocks = sm * (2048 / threads);
    char nm[96];
    double ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(d, iters, 1.0); });
    snprintf(nm, sizeof nm, "DFMA 8 chains, %4d thr x %d blk/SM", threads, 2048 / threads);
    rep(nm, 2.0 * 8 * iters * threads * blocks, ms);
  }
  {
    const int threads = 256, blocks = sm;  // the contraction kernel's occupancy: 8 warps/SM
    rep("DFMA 8 chains, 256 thr x 1 blk/SM (8 warps)", 2.0 * 8 * iters * threads * blocks,
        time_ms([&] { k_dfma<8><<<blocks, threads>>>(d, iters, 1.0); }));
    rep("DFMA 4 chains, 256 thr x 1 blk/SM", 2.0 * 4 * iters * threads * blocks,
        time_ms([&] { k_dfma<4><<<blocks, threads>>>(d, iters, 1.0); }));
    rep("DFMA 2 chains, 256 thr x 1 blk/SM", 2.0 * 2 * iters * threads * blocks,
        time_ms([&] { k_dfma<2><<<blocks, threads>>>(d, iters, 1.0); }));
    rep("DFMA 1 chain , 256 thr x 1 blk/SM (latency)", 2.0 * 1 * iters * threads * blocks,
        time_ms([&] { k_dfma<1><<<blocks, threads>>>(d, iters, 1.0); }));
    rep("DADD 8 chains, 256 thr x 1 blk/SM (1 flop)", 1.0 * 8 * iters * threads * blocks,
        time_ms([&] { k_dadd<8><<<blocks, threads>>>(d, iters, 1.0); }));
  }
  for (int threads : {256, 1024}) {
    const int blocks = sm * (2048 / threads) / (threads == 256 ? 8 : 1);
    char nm[96];
    const double warps = (double)threads / 32 * blocks;
    snprintf(nm, sizeof nm, "DMMA m8n8k4 8 acc, %d thr, %d blk", threads, blocks);
    rep(nm, 512.0 * 8 * iters * warps, time_ms([&] { k_dmma884<8><<<blocks, threads>>>(d, iters, 1.0); }));
    snprintf(nm, sizeof nm, "DMMA m16n8k16 4 acc, %d thr, %d blk", threads, blocks);
    rep(nm, 4096.0 * 4 * iters * warps, time_ms([&] { k_dmma16816<4><<<blocks, threads>>>(d, iters, 1.0); }));
    snprintf(nm, sizeof nm, "mixed 4 DMMA884 + 16 DFMA, %d thr, %d blk", threads, blocks);
    rep(nm, (512.0 * 4 + 64.0 * 16) * iters * warps, time_ms([&] { k_mixed<4, 16><<<blocks, threads>>>(d, iters, 1.0); }));
    snprintf(nm, sizeof nm, "mixed 4 DMMA884 + 4 DFMA, %d thr, %d blk", threads, blocks);
    rep(nm, (512.0 * 4 + 64.0 * 4) * iters * warps, time_ms([&] { k_mixed<4, 4><<<blocks, threads>>>(d, iters, 1.0); }));
  }
  {  // DMMA latency: 1 warp per SM sub-partition, CH independent accumulators per warp
    const int threads = 128, blocks = sm;
    const double warps = (double)threads / 32 * blocks;
    char nm[96];
#define LAT(CH)                                                                                            \
    {                                                                                                      \
      const double ms = time_ms([&] { k_dmma884<CH><<<blocks, threads>>>(d, iters, 1.0); });              \
      snprintf(nm, sizeof nm, "DMMA m8n8k4 %d acc, 1 warp/subpart: %.1f ns/DMMA/warp", CH, ms * 1e6 / (iters * CH)); \
      rep(nm, 512.0 * CH * iters * warps, ms);                                                             \
    }
    LAT(1) LAT(2) LAT(3) LAT(4) LAT(6) LAT(8)
#undef LAT
  }
  {
    const int threads = 256, blocks = sm * 4, it = 1 << 12;
    const double n = 4.0 * it * threads * blocks;
    rep("exp(double)   [Gop/s x1e-3]", n, time_ms([&] { k_transc<0><<<blocks, threads>>>(d, it, 1.0); }));
    rep("sqrt(double)  [Gop/s x1e-3]", n, time_ms([&] { k_transc<1><<<blocks, threads>>>(d, it, 1.0); }));
    rep("exp(-c*sqrt)  [Gop/s x1e-3]", n, time_ms([&] { k_transc<2><<<blocks, threads>>>(d, it, 1.0); }));
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
