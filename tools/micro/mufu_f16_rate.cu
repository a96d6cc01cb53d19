// Micro-benchmark: is the 16-bit MUFU exponential (MUFU.EX2.F16 / .BF16, what ex2.approx.f16x2 / .bf16x2 lower to: two per
// packed pair) issued at a higher rate than the fp32 one?  Reports exponentials per clock and SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_f16_rate mufu_f16_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int V>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0x00010001u + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (V == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
      if (V == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      if (V == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (v[i] & 0x3bff3bffu) | 0x30003000u;   // keeps the packed values small and finite (ALU)
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
void run(const char* name, int per, uint32_t* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    k<V><<<148, warps * 32>>>(out, 100, 0x30003000u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<148, warps * 32>>>(out, iters, 0x30003000u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ex = double(iters) * 16 * per * warps * 32 * 148;
    printf("%-28s warps/scheduler %d: %.2f exponentials/clk/SM (1.9 GHz)\n", name, warps / 4, ex / ms / 1e6 / 148 / 1.9);
  }
}
int main() {
  uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
  run<0>("ex2.approx.ftz.f32", 1, out);
  run<1>("ex2.approx.f16x2", 2, out);
  run<2>("ex2.approx.ftz.bf16x2", 2, out);
  return 0;
}
