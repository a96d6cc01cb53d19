// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition as a function of resident warps (1..4 per scheduler).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu ; ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] * 0.25f - 0.3f;   // keeps the values bounded; FMA pipe, 1 per MUFU
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps += 4) {
    k<<<148, warps * 32>>>(out, 100, 0.1f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148, warps * 32>>>(out, iters, 0.1f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ex = double(iters) * 16 * warps * 32 * 148;
    printf("warps/SM %2d (per scheduler %d): %.3f ms  %.1f Gex2/s  = %.2f ex2/clk/SM at 1.9 GHz\n", warps, warps / 4, ms, ex / ms / 1e6,
           ex / ms / 1e6 / 148 / 1.9);
  }
  return 0;
}
