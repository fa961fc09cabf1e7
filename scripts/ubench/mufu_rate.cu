// MUFU throughput probe: ex2.approx.ftz.f32 vs ex2.approx.f16x2 (2 MUFU.EX2.F16 per instruction) vs tanh.approx.f16x2.
// Prints results per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + 0.001f * (threadIdx.x + i), h[i] = 0x38003800u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters, 0.5f);
      if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters, 0.5f);
      if (mode == 2) k<2><<<148 * 2, 1024>>>(out, iters, 0.5f);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
    }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double instr = 148.0 * 2 * 1024 * (double)iters * 8;
    const double results = instr * (mode == 0 ? 1 : 2);
    printf("mode %d (%s): %.3f ms  %.2f T instr/s  %.2f T results/s  (%.1f results/clk/SM at 1.9 GHz)\n", mode,
           mode == 0 ? "ex2.f32" : mode == 1 ? "ex2.f16x2" : "tanh.f16x2", ms, instr / ms / 1e9, results / ms / 1e9,
           results / (ms * 1e-3) / 148 / 1.9e9);
    if (cudaGetLastError() != cudaSuccess) printf("error\n");
  }
  return 0;
}
