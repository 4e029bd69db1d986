// Microbenchmark: MUFU.EX2 issue rate per warp as a function of warps per SM sub-partition, with and without
// independent FMA-pipe work interleaved (what attention4's softmax warps do).  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = -0.001f * (threadIdx.x + i);
  float acc0 = 0.f, acc1 = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {           // 64 back-to-back independent MUFU
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
    } else if (MODE == 1) {    // MUFU interleaved with 2 independent FFMA each
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc0) : "f"(1.0001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc1) : "f"(0.9999f));
      }
    } else {                   // three separate passes: 32 FFMA2-ish, 64 MUFU, 64 FADD (attention4 today)
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(0.999f), "f"(-0.01f));
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc0) : "f"(r[i]));
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = r[i] * -0.5f;   // keep arguments small and negative
  }
  long long t1 = clock64();
  float s = acc0 + acc1;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 8);
  const int iters = 200;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps_per_smsp = 1; warps_per_smsp <= 6; ++warps_per_smsp) {
      const int threads = warps_per_smsp * 4 * 32;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_iter = double(c) / iters;
      printf("mode %d  warps/SMSP %d : %.0f clk per 64-MUFU iteration per warp -> %.2f clk/MUFU/warp, XU util %.0f%%\n",
             mode, warps_per_smsp, per_iter, per_iter / 64, 100.0 * warps_per_smsp * 64 * 8 / per_iter);
    }
  }
  return 0;
}
