// Host-side runtime of libcremage_b200: error reporting, launch counting and TMA tensor-map encoding.
// cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so the library has no link-time
// dependency on libcuda (it must load on a GPU-less host for the symbol tests).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CB_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
  return (dev >= 0 && dev < 64) ? dev : 0;
}

int sm_count() {
  static std::atomic<int> cache[64];
  const int slot = current_device_slot();
  int v = cache[slot].load(std::memory_order_relaxed);
  if (v == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, slot) != cudaSuccess || sms <= 0) {
      (void)cudaGetLastError();
      sms = 148;   // B200 (also what the planner assumes on a GPU-less host)
    }
    cache[slot].store(sms, std::memory_order_relaxed);
    v = sms;
  }
  return v;
}

// The opt-ins guarded by these flags are idempotent, so two threads racing through the first call on a device only
// repeat them; the flag is a plain byte per device.
bool device_once_needed(DeviceOnce& o) { return reinterpret_cast<std::atomic<unsigned char>&>(o.done[current_device_slot()]).load(std::memory_order_acquire) == 0; }
void device_once_done(DeviceOnce& o) { reinterpret_cast<std::atomic<unsigned char>&>(o.done[current_device_slot()]).store(1, std::memory_order_release); }

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return CB_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else (void)cudaGetLastError();
  });
  return fn;
}

int make_tmap_act(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver): the CUDA path cannot run on this host");
    return CB_ERR_NODRIVER;
  }
  if (rank < 2 || rank > 5) {
    set_error("make_tmap_act: rank %d unsupported", rank);
    return CB_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
    set_error("make_tmap_act: base pointer %p is not 16-byte aligned", base);
    return CB_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) {
      set_error("make_tmap_act: box[%d] = %u out of range", i, box[i]);
      return CB_ERR_INVALID;
    }
    if (i > 0) {
      gstr[i - 1] = strides_elems[i] * 2;  // bytes
      if (gstr[i - 1] % 16 != 0) {
        set_error("make_tmap_act: stride[%d] = %llu bytes is not a multiple of 16", i, (unsigned long long)gstr[i - 1]);
        return CB_ERR_INVALID;
      }
    }
  }
  CUresult r = fn(out, CB_TMAP_DTYPE, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu,... box %u,%u,...)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return CB_ERR_CUDA;
  }
  return CB_OK;
}

}  // namespace cb

extern "C" const char* cb_last_error(void) { return cb::g_err; }
extern "C" int cb_version(void) { return 100; }
extern "C" int cb_act_dtype(void) { return CB_ACT_DTYPE_ID; }
extern "C" int64_t cb_launch_count(void) { return (int64_t)cb::g_launches.load(); }
