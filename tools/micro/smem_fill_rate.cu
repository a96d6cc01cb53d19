// Micro-benchmark: shared-memory fill rate of one SM from L2 / HBM with bulk async copies (the TMA engine), as a
// function of the bytes kept in flight (stages x chunk) and of how many SMs stream at once.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_fill_rate smem_fill_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const char* src, size_t span, int chunk, int stages, int iters, long long* clk) {
  extern __shared__ __align__(128) char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  char* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const char* base = src + (size_t(blockIdx.x) * 7919u * 4096u) % (span - size_t(chunk) * 64);
    long long t0 = clock64();
    size_t off = 0;
    for (int i = 0; i < iters + stages; ++i) {
      const int s = i % stages;
      const uint32_t ph = uint32_t((i / stages - 1) & 1);
      if (i >= stages) {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar + s)), "r"(ph) : "memory");
      }
      if (i < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + size_t(s) * chunk)), "l"(base + off), "r"(chunk), "r"(s32(bar + s)) : "memory");
        off += chunk;
        if (off + chunk > size_t(chunk) * 64) off = 0;     // a 64-chunk window per CTA: L2 resident after the first pass
      }
    }
    clk[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const size_t span = 64ull << 20;
  char* src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
  long long* clk; cudaMalloc(&clk, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 4000;
  printf("%6s %7s %7s %9s %12s %12s\n", "SMs", "chunkKB", "stages", "inflKB", "GB/s per SM", "aggregate TB/s");
  for (int sms : {1, 148})
    for (int chunk : {8192, 16384, 32768})
      for (int stages : {1, 2, 4, 6, 8, 12}) {
        if (size_t(chunk) * stages > 200 * 1024) continue;
        k<<<sms, 32, 1024 + size_t(chunk) * stages>>>(src, span, chunk, stages, 200, clk);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<sms, 32, 1024 + size_t(chunk) * stages>>>(src, span, chunk, stages, iters, clk);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double per = double(chunk) * iters / (ms * 1e-3) / 1e9;
        printf("%6d %7d %7d %9d %12.1f %12.2f\n", sms, chunk / 1024, stages, chunk * stages / 1024, per, per * sms / 1e3);
      }
  return 0;
}
