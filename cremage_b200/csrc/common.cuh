// Shared device helpers for the sm_100a kernels of cremage_b200: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st / fences) and UMMA descriptor construction.
// Everything here is inline PTX; there is no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#define CB_DEVINL __device__ __forceinline__

namespace cb {

// ----------------------------------------------------------------------------------------------
// 16-bit activation / weight type: fp16 (-DCB_FP16, the reference's own GPU precision: model.half() + autocast,
// sd/image_generator.py:489,748) or bf16 (default build flag absent).  Both feed tcgen05 kind::f16 at the same rate;
// accumulation and all statistics are fp32 either way.
// ----------------------------------------------------------------------------------------------
#ifdef CB_FP16
using act_t = __half;
using act_t2 = __half2;
#define CB_MMA_FMT 0u
#define CB_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define CB_ACT_DTYPE_ID 1
#else
using act_t = __nv_bfloat16;
using act_t2 = __nv_bfloat162;
#define CB_MMA_FMT 1u
#define CB_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define CB_ACT_DTYPE_ID 2
#endif

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
enum : int {
  CB_OK = 0,
  CB_ERR_INVALID = -1,   // bad argument / unsupported shape
  CB_ERR_CUDA = -2,      // a CUDA runtime / driver call failed
  CB_ERR_NODRIVER = -3,  // cuTensorMapEncodeTiled could not be resolved (no driver)
};
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;  // kernels launched by this library (cb_launch_count)
#define CB_LAUNCHED(n) cb::g_launches.fetch_add((n), std::memory_order_relaxed)
int cuda_fail(cudaError_t e, const char* what);

#define CB_CHECK_CUDA(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return cb::cuda_fail(_e, #expr);   \
  } while (0)

#define CB_REQUIRE(cond, ...)           \
  do {                                  \
    if (!(cond)) {                      \
      cb::set_error(__VA_ARGS__);       \
      return cb::CB_ERR_INVALID;        \
    }                                   \
  } while (0)

// Encode a tiled act_t (fp16 / bf16) tensor map (rank 2..5), SWIZZLE_128B, zero OOB fill.
// dims / strides are in ELEMENTS (strides[0] is implied = 1); box in elements.
// swizzle_bytes: 128 (operand tiles, 64-element inner box) or 64 (epilogue panels, 32-element inner box).
int make_tmap_act(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_elems, const uint32_t* box, int swizzle_bytes = 128);

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (CB_PDL=0 turns it off), signals `launch_dependents` as its first
// instruction and executes `griddepcontrol.wait` before its first access to global memory.  The next kernel of the
// stream -- also inside a captured CUDA graph -- is then scheduled onto SMs as they drain and runs its prologue
// (barrier init, tensor-map prefetch, TMEM allocation, index arithmetic) under the tail of this one; `wait` returns
// once the whole preceding grid has completed and its writes are visible, so data dependencies are unchanged.
// What it buys is per-node latency: the small-batch regime (UNet batch 2: ~350 nodes of a few microseconds each).
// ----------------------------------------------------------------------------------------------
bool pdl_enabled();   // runtime.cu
// Per-DEVICE state (runtime.cu).  cudaFuncSetAttribute opt-ins and the SM count belong to the current device, not to the
// process or the calling thread: a host that drives several GPUs from one process gets each of them configured.
int sm_count();                          // multiprocessors of the current device (148 on a B200; cached per device)
struct DeviceOnce { unsigned char done[64]; };
bool device_once_needed(DeviceOnce& o);  // true until device_once_done() was called for the current device
void device_once_done(DeviceOnce& o);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// simple kernels: both at the top
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }
#endif

// ----------------------------------------------------------------------------------------------
// small device utilities
// ----------------------------------------------------------------------------------------------
CB_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

CB_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t.reg .b32 R1;\n\t"
      "elect.sync R1|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
CB_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
CB_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CB_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
CB_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
CB_DEVINL uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must not hang the GPU box; after ~2 s of spinning the kernel traps.
CB_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3FFFu) == 0 && (clock64() - t0) > 4000000000LL) __trap();
  }
}

// ----------------------------------------------------------------------------------------------
// TMA loads (tile mode) into shared memory, completion on an mbarrier
// ----------------------------------------------------------------------------------------------
CB_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
CB_DEVINL void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
CB_DEVINL void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
CB_DEVINL void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store (tile mode) shared -> global, bulk-group completion
CB_DEVINL void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
CB_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING their shared-memory source
template <int N>
CB_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA spans both SMs' tensor cores; each CTA stages its own A
// rows and HALF of the B tile, so the shared-memory fill per flop drops by a third to a half.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address -> same offset in the pair's even (leader) CTA
CB_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
CB_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes land on the LEADER's mbarrier
CB_DEVINL void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
CB_DEVINL void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
CB_DEVINL void tmem_alloc_pair(uint32_t smem_result_addr, uint32_t ncols) {  // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols)
               : "memory");
}
CB_DEVINL void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
CB_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's 128 rows] * B[N/2 rows from each CTA]; issued by one thread of the leader
CB_DEVINL void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once every previously issued MMA of this thread has completed
CB_DEVINL void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// arrive on the mbarrier at this offset in the pair's leader CTA (from either CTA)
CB_DEVINL void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (UMMA reads of a tile written with st.shared)
CB_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, loads/stores, fences
// ----------------------------------------------------------------------------------------------
CB_DEVINL void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols)
               : "memory");
}
CB_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
CB_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
CB_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CB_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate. One thread issues.
CB_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from tensor memory (16-bit values packed two per 32-bit cell, row = lane)
CB_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on an mbarrier once every previously issued tcgen05.mma of this thread has completed
CB_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

CB_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
CB_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (base_lane + t), columns [col, col+32)
CB_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
CB_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp InstrDescriptor / SmemDescriptor)
// ----------------------------------------------------------------------------------------------
// Instruction descriptor, kind::f16, A/B = act_t (fp16 or bf16), D = fp32.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (CB_MMA_FMT << 7)               // a_format (0 = F16, 1 = BF16)
         | (CB_MMA_FMT << 10)              // b_format
         | (uint32_t(a_mn_major) << 15)    // a_major (0 = K)
         | (uint32_t(b_mn_major) << 16)    // b_major (0 = K)
         | (uint32_t(N >> 3) << 17)        // n_dim
         | (uint32_t(M >> 4) << 24);       // m_dim
}
// Shared-memory matrix descriptor, SWIZZLE_128B. `lbo_bytes`/`sbo_bytes` per the canonical layouts:
//   K-major : rows of 128 B, 8-row atoms of 1024 B -> SBO = 1024, LBO unused (1)
//   MN-major: 64-element (128 B) MN runs, 8 K-rows per 1024 B atom -> SBO = 1024 (next 8 K rows),
//             LBO = byte distance between consecutive 64-element MN panels
__host__ __device__ constexpr uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((smem_addr & 0x3FFFFu) >> 4)        // start address
         | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16)
         | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32)
         | (uint64_t(1) << 46)                          // descriptor version (Blackwell)
         | (uint64_t(2) << 61);                         // layout type = SWIZZLE_128B
}

// ----------------------------------------------------------------------------------------------
// numerics helpers
// ----------------------------------------------------------------------------------------------
#ifdef CB_FP16
CB_DEVINL float sat_f16(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }  // finite saturation, NaN passes
CB_DEVINL uint32_t pack_act2(float lo, float hi) {
  uint32_t r;  // one F2FP.SATFINITE: converts, saturates to +-65504 and packs (first source -> upper half)
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
CB_DEVINL float2 unpack_act2(uint32_t u) {
  act_t2 v = *reinterpret_cast<act_t2*>(&u);
  return __half22float2(v);
}
CB_DEVINL act_t to_act(float x) { return __float2half_rn(sat_f16(x)); }
CB_DEVINL float from_act(act_t x) { return __half2float(x); }
#else
CB_DEVINL uint32_t pack_act2(float lo, float hi) {
  act_t2 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
CB_DEVINL float2 unpack_act2(uint32_t u) {
  act_t2 v = *reinterpret_cast<act_t2*>(&u);
  return __bfloat1622float2(v);
}
CB_DEVINL act_t to_act(float x) { return __float2bfloat16(x); }
CB_DEVINL float from_act(act_t x) { return __bfloat162float(x); }
#endif
CB_DEVINL float silu_f(float x) { return x / (1.f + __expf(-x)); }
CB_DEVINL float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, far below the 16-bit output rounding): one MUFU.RCP, one
// MUFU.EX2 and 8 FMA-pipe instructions instead of libdevice erff's two-branch polynomial -- the GEGLU epilogue applies
// it to 84 M elements per top-level feed-forward.
CB_DEVINL float gelu_fast_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float erf_abs = fmaf(-p * t, e, 1.f);             // erf(|x|/sqrt2)
  const float hx = 0.5f * x;
  return fmaf(copysignf(erf_abs, x), hx, hx);              // 0.5 x (1 + erf)
}

// f[e] *= gelu(g[e]) for eight gates at once, written stage by stage so that the eight dependency chains are
// interleaved by construction (two epilogue warps per scheduler cannot hide a serial Horner chain).
// gelu(g) = 0.5 g (1 + erf(g / sqrt2)) = relu(g) - 0.5 |g| erfc(|g| / sqrt2); erfc by Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7) in the variable w = |g| sqrt(log2(e) / 2), so that exp(-g^2 / 2) = 2^(-w^2); the -0.5 |g| / w
// factor is folded into the polynomial coefficients.  14 instructions per gate, two of them MUFU.
CB_DEVINL void geglu_mul8(float (&f)[8], const float (&g)[8]) {
  float w[8], t[8], ex[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] = fabsf(g[e]) * 0.8493218003f;
#pragma unroll
  for (int e = 0; e < 8; ++e) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t[e]) : "f"(fmaf(w[e], 0.2727374809f, 1.f)));
#pragma unroll
  for (int e = 0; e < 8; ++e) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex[e]) : "f"(-w[e] * w[e]));
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] = fmaf(-0.6248546950f, t[e], 0.8554778804f);
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] = fmaf(q[e], t[e], -0.8367933924f);
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] = fmaf(q[e], t[e], 0.1674846542f);
#pragma unroll
  for (int e = 0; e < 8; ++e) q[e] = fmaf(q[e], t[e], -0.1500194578f);
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] *= ex[e];
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] *= t[e];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] *= fmaf(q[e], w[e], fmaxf(g[e], 0.f));
}

CB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CB_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace cb
