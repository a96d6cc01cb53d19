// Micro-benchmark: does the f32 -> packed 16-bit conversion share the MUFU (XU) pipe?  One warp per scheduler (and two),
// 16 ex2 per iteration plus 8 packs of the variant under test; reports ex2 per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_pack_rate mufu_pack_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
template <int V>
__global__ void k(uint32_t* out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      uint32_t pk = 0;
      if (V == 1) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[i + 1]), "f"(v[i]));
      if (V == 2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[i + 1]), "f"(v[i]));
      if (V == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[i + 1]), "f"(v[i]));
      if (V == 4) {   // bf16 by integer ops: round-to-nearest-even-free (add half ulp), take the high halves with one PRMT
        const uint32_t a = __float_as_uint(v[i]) + 0x8000u, b = __float_as_uint(v[i + 1]) + 0x8000u;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk) : "r"(a), "r"(b));
      }
      if (V == 5) {   // fp16 by FMA-pipe / ALU ops for values in [0, 65504]: rebias the exponent with one FMUL, round, shift, PRMT
        const uint32_t a = (__float_as_uint(v[i] * 1.925929944e-34f) + 0x1000u) >> 13;        // * 2^-112
        const uint32_t b = (__float_as_uint(v[i + 1] * 1.925929944e-34f) + 0x1000u) >> 13;
        asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(pk) : "r"(a), "r"(b));
      }
      acc ^= pk;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] * 0.25f - 0.3f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(s);
}
template <int V>
void run(const char* name, uint32_t* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8}) {
    k<V><<<148, warps * 32>>>(out, 100, 0.1f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<148, warps * 32>>>(out, iters, 0.1f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ex = double(iters) * 16 * warps * 32 * 148;
    printf("%-44s warps/scheduler %d: %.2f ex2/clk/SM (1.9 GHz)\n", name, warps / 4, ex / ms / 1e6 / 148 / 1.9);
  }
}
int main() {
  uint32_t* out; cudaMalloc(&out, 148 * 1024 * 4);
  run<0>("ex2 only", out);
  run<1>("ex2 + cvt.rn.satfinite.f16x2.f32 per pair", out);
  run<2>("ex2 + cvt.rn.f16x2.f32 per pair", out);
  run<3>("ex2 + cvt.rn.bf16x2.f32 per pair", out);
  run<4>("ex2 + integer bf16 pack (2 IADD + PRMT)", out);
  run<5>("ex2 + FMUL/IADD/SHF/PRMT fp16 pack", out);
  return 0;
}
