// Register <-> (lane, column) mapping of tcgen05.ld .16x256b and tcgen05.st .16x128b, decoded against the .32x32b shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_layout tmem_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = slot;
  const uint32_t row = warp * 32 + lane;
  const uint32_t mine = tb + ((uint32_t)(warp * 32) << 16);
  // 1. write value = row * 256 + col through .32x32b (thread = lane, register = column)
  for (int c = 0; c < 32; c += 4) {
    uint32_t v0 = row * 256 + c, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + c), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  // 2. read 16 lanes x 16 columns through .16x256b.x2 (8 registers), lanes 0..15 of this warp's quarter, then 16..31
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tb + ((uint32_t)(warp * 32 + half * 16) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + lane) * 8 + i] = r[i];
  }
  __syncthreads();
  // 3. write through .16x128b.x2 (4 registers; value = thread * 16 + register) into columns 32.., read back through .32x32b
  for (int half = 0; half < 2; ++half) {
    uint32_t v[4];
    for (int i = 0; i < 4; ++i) v[i] = 0x10000u * (half + 1) + lane * 16 + i;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb + 32 + ((uint32_t)(warp * 32 + half * 16) << 16)),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(mine + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 8; ++i) out[8 * 32 * 8 + row * 8 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4 * (8 * 32 * 8 + 128 * 8));
  k<<<1, 128>>>(d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  static uint32_t h[8 * 32 * 8 + 128 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("== ld .16x256b.x2, warp 1 (lanes 32..63): thread: reg -> (row, col)\n");
  for (int half = 0; half < 2; ++half)
    for (int t = 0; t < 32; ++t) {
      printf("half %d t%2d:", half, t);
      for (int i = 0; i < 8; ++i) { uint32_t v = h[((1 * 2 + half) * 32 + t) * 8 + i]; printf(" (%u,%u)", v / 256, v % 256); }
      printf("\n");
    }
  printf("== st .16x128b.x2 read back by .32x32b, warp 1: row: col -> (half, thread, reg)\n");
  for (int row = 32; row < 64; ++row) {
    printf("row %2d:", row);
    for (int i = 0; i < 8; ++i) { uint32_t v = h[8 * 32 * 8 + row * 8 + i]; printf(" (%u,%u,%u)", v >> 16, (v & 0xffff) / 16, v % 16); }
    printf("\n");
  }
  return 0;
}
