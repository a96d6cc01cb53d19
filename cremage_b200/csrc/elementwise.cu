// Layout conversions and small elementwise kernels of the denoising path (all HBM/latency bound).
#include <cuda_fp16.h>

#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

// NCHW (fp32 / fp16 / bf16) -> NHWC bf16, zero padding channels [c, c_pad). One thread per (n, pixel).
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, long long n, int c, long long hw, int c_pad, float scale,
                                    act_t* __restrict__ dst) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * hw) return;
  const long long b = i / hw, p = i - b * hw;
  const T* s = src + b * c * hw + p;
  act_t* d = dst + i * c_pad;
  for (int ch = 0; ch < c_pad; ++ch) {
    float v = ch < c ? float(s[(long long)ch * hw]) * scale : 0.f;
    d[ch] = to_act(v);
  }
}

// 1x1 channel mix (tiny c, cout <= 16) fused with NCHW fp32 -> NHWC bf16. One thread per (n, pixel).
__global__ void pointwise_nchw_to_nhwc_kernel(const float* __restrict__ src, long long n, int c, long long hw,
                                              const float* __restrict__ w, const float* __restrict__ b, int cout,
                                              int c_pad, float scale, act_t* __restrict__ dst) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * hw) return;
  const long long bi = i / hw, p = i - bi * hw;
  const float* s = src + bi * c * hw + p;
  float xin[16];
  for (int ci = 0; ci < c; ++ci) xin[ci] = s[(long long)ci * hw] * scale;
  act_t* d = dst + i * c_pad;
  for (int co = 0; co < c_pad; ++co) {
    float acc = 0.f;
    if (co < cout) {
      acc = b ? b[co] : 0.f;
      for (int ci = 0; ci < c; ++ci) acc = fmaf(w[co * c + ci], xin[ci], acc);
    }
    d[co] = to_act(acc);
  }
}

// ControlNet residual injection (cldm/cldm.py:59-66: `h += control.pop()`, `hs.pop() + control.pop()`):
// dst[n][p][ch] = base[n][p][ch] + ctrl[n][ch][p], the control tensor arriving NCHW from the (out-of-scope) ControlNet
template <typename T>
__global__ void add_nchw_to_nhwc_kernel(const act_t* __restrict__ base, const T* __restrict__ ctrl, int c, long long hw,
                                        act_t* __restrict__ dst) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j;
    const long long p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (ch < c && p < hw) ? float(ctrl[(b * c + ch) * hw + p]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int ch = c0 + threadIdx.x;
    if (p < hw && ch < c) {
      const long long o = (b * hw + p) * c + ch;
      dst[o] = to_act(from_act(base[o]) + tile[threadIdx.x][j]);
    }
  }
}

// generic tiled transpose for wide channel counts: [n][c][hw] -> [n][hw][c_pad]
template <typename T>
__global__ void nchw_to_nhwc_tiled_kernel(const T* __restrict__ src, int c, long long hw, int c_pad, float scale,
                                          act_t* __restrict__ dst) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j;
    const long long p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (ch < c && p < hw) ? float(src[(b * c + ch) * hw + p]) * scale : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int ch = c0 + threadIdx.x;
    if (p < hw && ch < c_pad) dst[(b * hw + p) * c_pad + ch] = to_act(tile[threadIdx.x][j]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int c, long long hw, long long c_ld,
                                    float* __restrict__ dst) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int ch = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < hw && ch < c) ? float(src[(b * hw + p) * c_ld + ch]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ch = c0 + j;
    const long long p = p0 + threadIdx.x;
    if (ch < c && p < hw) dst[(b * c + ch) * hw + p] = tile[threadIdx.x][j];
  }
}

// nearest 2x upsample, one thread per 16-byte channel vector of an OUTPUT pixel
__global__ void upsample2x_kernel(const uint4* __restrict__ src, long long n, int h, int w, int cv,
                                  uint4* __restrict__ dst) {
  pdl_prologue();
  const long long total = n * (2LL * h) * (2LL * w) * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % cv);
    long long t = i / cv;
    const int ox = int(t % (2 * w));
    t /= (2 * w);
    const int oy = int(t % (2 * h));
    const long long b = t / (2 * h);
    dst[i] = src[((b * h + (oy >> 1)) * w + (ox >> 1)) * cv + v];
  }
}

// parity split: dst[2*ph+pw][n][h/2][w/2][c] = src[n][2y+ph][2x+pw][c]
__global__ void parity_split_kernel(const uint4* __restrict__ src, long long n, int h, int w, int cv,
                                    uint4* __restrict__ dst) {
  pdl_prologue();
  const long long total = n * h * w * cv;
  const int h2 = h >> 1, w2 = w >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % cv);
    long long t = i / cv;
    const int x = int(t % w);
    t /= w;
    const int y = int(t % h);
    const long long b = t / h;
    const int plane = ((y & 1) << 1) | (x & 1);
    dst[(((plane * n + b) * h2 + (y >> 1)) * w2 + (x >> 1)) * cv + v] = src[i];
  }
}

// sinusoidal timestep embedding, out[n][dim] = [cos(t*f) | sin(t*f)]; the frequency table f (fp32 [dim/2]) is built
// on the host with the reference's own expression so it is bit-identical to the reference's.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, long long n, int dim,
                                          const float* __restrict__ freqs, act_t* __restrict__ out) {
  pdl_prologue();
  const int half = dim / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const long long b = i / half;
  const int k = int(i - b * half);
  const float arg = t[b] * freqs[k];
  out[b * dim + k] = to_act(cosf(arg));
  out[b * dim + half + k] = to_act(sinf(arg));
  if ((dim & 1) && k == 0) out[b * dim + dim - 1] = to_act(0.f);
}

// direct 3x3 conv (pad 1, stride 1) for tiny cin; wgt fp32 [3][3][cin][cout]; thread = (pixel, 8 output channels)
template <int CIN>
__global__ void conv3x3_small_cin_kernel(const act_t* __restrict__ src, long long n, int h, int w, int cin_ld,
                                         const float* __restrict__ wgt, const float* __restrict__ bias, int cout,
                                         act_t* __restrict__ out) {
  pdl_prologue();
  const int cg = cout >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * h * w * cg) return;
  const int g = int(i % cg);
  long long t = i / cg;
  const int x = int(t % w);
  t /= w;
  const int y = int(t % h);
  const long long b = t / h;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = bias ? bias[g * 8 + e] : 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      if (xx < 0 || xx >= w) continue;
      const act_t* s = src + ((b * h + yy) * w + xx) * cin_ld;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float a = from_act(s[ci]);
        const float* wp = wgt + ((ky * 3 + kx) * CIN + ci) * cout + g * 8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
        acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
        acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
        acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
        acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
      }
    }
  }
  act_t* o = out + ((b * h + y) * w + x) * (long long)cout + g * 8;
  *reinterpret_cast<uint4*>(o) = make_uint4(pack_act2(acc[0], acc[1]), pack_act2(acc[2], acc[3]),
                                            pack_act2(acc[4], acc[5]), pack_act2(acc[6], acc[7]));
}

__global__ void silu_add_kernel(const act_t* __restrict__ x, const act_t* __restrict__ add,
                                long long count, act_t* __restrict__ out) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v = from_act(x[i]);
  if (add) v += from_act(add[i]);
  out[i] = to_act(silu_f(v));
}

// DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py:24-37) from the encoder's moments
// [n][2c][hw] fp32 NCHW: mean = first c channels, logvar = clamp(last c, -30, 20), std = exp(0.5 logvar),
// sample = mean + std * noise  (noise = the caller's torch.randn draw; null -> sample = mean, the mode)
__global__ void diag_gaussian_kernel(const float* __restrict__ moments, const float* __restrict__ noise, long long n,
                                     long long c, long long hw, float scale, float* __restrict__ mean_out,
                                     float* __restrict__ std_out, float* __restrict__ sample_out) {
  pdl_prologue();
  const long long total = n * c * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / (c * hw), rem = i - img * (c * hw);
    const float mean = moments[img * 2 * c * hw + rem];
    const float logvar = fminf(fmaxf(moments[img * 2 * c * hw + c * hw + rem], -30.f), 20.f);
    const float sd = expf(0.5f * logvar);
    if (mean_out) mean_out[i] = mean;
    if (std_out) std_out[i] = sd;
    if (sample_out) sample_out[i] = scale * (noise ? mean + sd * noise[i] : mean);
  }
}

// bilinear upsample by an integer factor, align_corners=False (F.interpolate semantics: src = (dst + 0.5) / scale - 0.5,
// clamped at 0; neighbours clamped at the last index), fp32 planes [planes][h][w] -> [planes][h*f][w*f]
__global__ void bilinear_upsample_kernel(const float* __restrict__ src, long long planes, int h, int w, int f,
                                         float* __restrict__ dst) {
  pdl_prologue();
  const int oh = h * f, ow = w * f;
  const long long total = planes * oh * ow;
  const float rs = 1.f / float(f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = int(i % ow);
    long long t = i / ow;
    const int oy = int(t % oh);
    const long long pl = t / oh;
    float sy = (float(oy) + 0.5f) * rs - 0.5f;
    float sx = (float(ox) + 0.5f) * rs - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = int(sy), x0 = int(sx);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - float(y0), lx = sx - float(x0);
    const float* sp = src + pl * h * w;
    const float v00 = sp[y0 * w + x0], v01 = sp[y0 * w + x1], v10 = sp[y1 * w + x0], v11 = sp[y1 * w + x1];
    dst[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

__global__ void image_to_u8_kernel(const float* __restrict__ src, long long npix, long long c_ld,
                                   uint8_t* __restrict__ dst) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = (src[i * c_ld + ch] + 1.0f) / 2.0f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    dst[i * 3 + ch] = (uint8_t)(255.f * v);  // truncating, like numpy astype(uint8)
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_nchw_to_nhwc(const void* src, int src_dtype, int64_t n, int64_t c, int64_t hw, int64_t c_pad,
                               float scale, void* dst, cudaStream_t stream) {
  CB_REQUIRE(src && dst && n > 0 && c > 0 && hw > 0 && c_pad >= c, "cb_nchw_to_nhwc: bad arguments");
  CB_REQUIRE(src_dtype >= 0 && src_dtype <= 2, "cb_nchw_to_nhwc: src_dtype must be 0 (f32), 1 (f16) or 2 (bf16)");
  auto D = (act_t*)dst;
  if (c_pad <= 16) {
    const long long total = n * hw;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (src_dtype == 0) (void)cb::launch_k(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(256), (size_t)(0), stream, (const float*)src, n, (int)c, hw, (int)c_pad, scale, D);
    else if (src_dtype == 1) (void)cb::launch_k(nchw_to_nhwc_kernel<__half>, dim3(grid), dim3(256), (size_t)(0), stream, (const __half*)src, n, (int)c, hw, (int)c_pad, scale, D);
    else (void)cb::launch_k(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(grid), dim3(256), (size_t)(0), stream, (const __nv_bfloat16*)src, n, (int)c, hw, (int)c_pad, scale, D);
  } else {
    dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c_pad + 31) / 32), (unsigned)n), block(32, 8);
    if (src_dtype == 0) (void)cb::launch_k(nchw_to_nhwc_tiled_kernel<float>, dim3(grid), dim3(block), (size_t)(0), stream, (const float*)src, (int)c, hw, (int)c_pad, scale, D);
    else if (src_dtype == 1) (void)cb::launch_k(nchw_to_nhwc_tiled_kernel<__half>, dim3(grid), dim3(block), (size_t)(0), stream, (const __half*)src, (int)c, hw, (int)c_pad, scale, D);
    else (void)cb::launch_k(nchw_to_nhwc_tiled_kernel<__nv_bfloat16>, dim3(grid), dim3(block), (size_t)(0), stream, (const __nv_bfloat16*)src, (int)c, hw, (int)c_pad, scale, D);
  }
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_add_nchw_to_nhwc(const void* base, const void* ctrl, int ctrl_dtype, int64_t n, int64_t c, int64_t hw,
                                   void* dst, cudaStream_t stream) {
  CB_REQUIRE(base && ctrl && dst && n > 0 && c > 0 && hw > 0 && n <= 65535, "cb_add_nchw_to_nhwc: bad arguments");
  CB_REQUIRE(ctrl_dtype >= 0 && ctrl_dtype <= 2, "cb_add_nchw_to_nhwc: ctrl_dtype must be 0 (f32), 1 (f16) or 2 (bf16)");
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n), block(32, 8);
  auto B = (const act_t*)base;
  auto D = (act_t*)dst;
  if (ctrl_dtype == 0) (void)cb::launch_k(add_nchw_to_nhwc_kernel<float>, dim3(grid), dim3(block), (size_t)(0), stream, B, (const float*)ctrl, (int)c, hw, D);
  else if (ctrl_dtype == 1) (void)cb::launch_k(add_nchw_to_nhwc_kernel<__half>, dim3(grid), dim3(block), (size_t)(0), stream, B, (const __half*)ctrl, (int)c, hw, D);
  else (void)cb::launch_k(add_nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(grid), dim3(block), (size_t)(0), stream, B, (const __nv_bfloat16*)ctrl, (int)c, hw, D);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_pointwise_nchw_to_nhwc(const float* src, int64_t n, int64_t c, int64_t hw, const float* w,
                                         const float* b, int64_t cout, int64_t c_pad, float scale, void* dst,
                                         cudaStream_t stream) {
  CB_REQUIRE(src && w && dst && n > 0 && hw > 0, "cb_pointwise_nchw_to_nhwc: bad arguments");
  CB_REQUIRE(c > 0 && c <= 16 && cout > 0 && cout <= c_pad && c_pad <= 16, "cb_pointwise_nchw_to_nhwc: c, cout, c_pad must be <= 16");
  const long long total = n * hw;
  (void)cb::launch_k(pointwise_nchw_to_nhwc_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), stream, src, n, (int)c, hw, w, b, (int)cout, (int)c_pad, scale, (act_t*)dst);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_nhwc_to_nchw_f32(const void* src, int src_f32, int64_t n, int64_t c, int64_t hw, int64_t c_ld,
                                   float* dst, cudaStream_t stream) {
  CB_REQUIRE(src && dst && n > 0 && c > 0 && hw > 0 && c_ld >= c, "cb_nhwc_to_nchw_f32: bad arguments");
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n), block(32, 8);
  if (src_f32) (void)cb::launch_k(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(block), (size_t)(0), stream, (const float*)src, (int)c, hw, c_ld, dst);
  else (void)cb::launch_k(nhwc_to_nchw_kernel<act_t>, dim3(grid), dim3(block), (size_t)(0), stream, (const act_t*)src, (int)c, hw, c_ld, dst);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

static unsigned grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

extern "C" int cb_upsample2x_nhwc(const void* src, int64_t n, int64_t h, int64_t w, int64_t c, void* dst,
                                  cudaStream_t stream) {
  CB_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "cb_upsample2x_nhwc: bad arguments");
  const long long total = n * 4 * h * w * (c / 8);
  (void)cb::launch_k(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), stream, (const uint4*)src, n, (int)h, (int)w, (int)(c / 8), (uint4*)dst);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_parity_split_nhwc(const void* src, int64_t n, int64_t h, int64_t w, int64_t c, void* dst,
                                    cudaStream_t stream) {
  CB_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && h % 2 == 0 && w % 2 == 0,
             "cb_parity_split_nhwc: bad arguments (even h, w; c %% 8 == 0)");
  const long long total = n * h * w * (c / 8);
  (void)cb::launch_k(parity_split_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), stream, (const uint4*)src, n, (int)h, (int)w, (int)(c / 8), (uint4*)dst);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_timestep_embedding(const float* t, int64_t n, int dim, const float* freqs, void* out,
                                     cudaStream_t stream) {
  CB_REQUIRE(t && freqs && out && n > 0 && dim >= 2, "cb_timestep_embedding: bad arguments");
  const long long total = n * (dim / 2);
  (void)cb::launch_k(timestep_embedding_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), (size_t)(0), stream, t, n, dim, freqs, (act_t*)out);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_conv3x3_small_cin(const void* src, int64_t n, int64_t h, int64_t w, int cin, int64_t cin_ld,
                                    const float* wgt, const float* bias, int64_t cout, void* out, cudaStream_t stream) {
  CB_REQUIRE(src && wgt && out && n > 0 && h > 0 && w > 0, "cb_conv3x3_small_cin: bad arguments");
  CB_REQUIRE(cout % 8 == 0, "cb_conv3x3_small_cin: cout must be a multiple of 8");
  CB_REQUIRE(cin_ld >= cin, "cb_conv3x3_small_cin: cin_ld < cin");
  const long long total = n * h * w * (cout / 8);
  const unsigned grid = (unsigned)((total + 255) / 256);
  auto S = (const act_t*)src;
  auto O = (act_t*)out;
  switch (cin) {
    case 3: (void)cb::launch_k(conv3x3_small_cin_kernel<3>, dim3(grid), dim3(256), (size_t)(0), stream, S, n, (int)h, (int)w, (int)cin_ld, wgt, bias, (int)cout, O); break;
    case 4: (void)cb::launch_k(conv3x3_small_cin_kernel<4>, dim3(grid), dim3(256), (size_t)(0), stream, S, n, (int)h, (int)w, (int)cin_ld, wgt, bias, (int)cout, O); break;
    case 8: (void)cb::launch_k(conv3x3_small_cin_kernel<8>, dim3(grid), dim3(256), (size_t)(0), stream, S, n, (int)h, (int)w, (int)cin_ld, wgt, bias, (int)cout, O); break;
    case 9: (void)cb::launch_k(conv3x3_small_cin_kernel<9>, dim3(grid), dim3(256), (size_t)(0), stream, S, n, (int)h, (int)w, (int)cin_ld, wgt, bias, (int)cout, O); break;
    default: CB_REQUIRE(false, "cb_conv3x3_small_cin: cin %d unsupported (3, 4, 8, 9)", cin);
  }
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_silu_add(const void* x, const void* add, int64_t count, void* out, cudaStream_t stream) {
  CB_REQUIRE(x && out && count > 0, "cb_silu_add: bad arguments");
  (void)cb::launch_k(silu_add_kernel, dim3((unsigned)((count + 255) / 256)), dim3(256), (size_t)(0), stream, (const act_t*)x, (const act_t*)add, count, (act_t*)out);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_diag_gaussian(const float* moments, const float* noise, int64_t n, int64_t c, int64_t hw, float scale,
                                float* mean_out, float* std_out, float* sample_out, cudaStream_t stream) {
  CB_REQUIRE(moments && n > 0 && c > 0 && hw > 0 && (mean_out || std_out || sample_out), "cb_diag_gaussian: bad arguments");
  (void)cb::launch_k(diag_gaussian_kernel, dim3(grid_for(n * c * hw, 256)), dim3(256), (size_t)(0), stream, moments, noise, n, c, hw, scale == 0.f ? 1.f : scale,
                                                                     mean_out, std_out, sample_out);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_bilinear_upsample_f32(const float* src, int64_t planes, int64_t h, int64_t w, int factor, float* dst,
                                       cudaStream_t stream) {
  CB_REQUIRE(src && dst && planes > 0 && h > 0 && w > 0 && factor >= 1, "cb_bilinear_upsample_f32: bad arguments");
  const long long total = planes * h * factor * w * factor;
  (void)cb::launch_k(bilinear_upsample_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), stream, src, planes, (int)h, (int)w, factor, dst);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_image_to_u8(const void* src, int64_t n, int64_t hw, int64_t c_ld, uint8_t* dst, cudaStream_t stream) {
  CB_REQUIRE(src && dst && n > 0 && hw > 0 && c_ld >= 3, "cb_image_to_u8: bad arguments");
  const long long npix = n * hw;
  (void)cb::launch_k(image_to_u8_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), (size_t)(0), stream, (const float*)src, npix, c_ld, dst);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// weight repacking: fp32 [cout][c0 + c1][taps] (OIHW with taps = kh*kw contiguous) -> 16-bit [cout][taps][pad64(c0) | pad64(c1)]
// ----------------------------------------------------------------------------------------------------------------
namespace cb {
__global__ void pack_weight_kernel(const float* __restrict__ w, long long cout, int c0, int c1, int taps, act_t* __restrict__ out) {
  pdl_prologue();
  const int p0 = (c0 + 63) / 64 * 64, p1 = (c1 + 63) / 64 * 64;
  const long long kcols = (long long)taps * (p0 + p1), total = cout * kcols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / kcols;
    const int k = int(i - o * kcols), t = k / (p0 + p1), c = k - t * (p0 + p1);
    int ci = -1;
    if (c < p0) { if (c < c0) ci = c; }
    else if (c - p0 < c1) ci = c0 + (c - p0);
    out[i] = to_act(ci < 0 ? 0.f : w[(o * (c0 + c1) + ci) * taps + t]);
  }
}
}  // namespace cb

extern "C" int cb_pack_weight(const float* w, int64_t cout, int64_t c0, int64_t c1, int taps, void* out, cudaStream_t stream) {
  CB_REQUIRE(w && out && cout > 0 && c0 > 0 && c1 >= 0 && taps >= 1 && taps <= 9, "cb_pack_weight: bad arguments");
  const long long total = cout * (long long)taps * ((c0 + 63) / 64 * 64 + (c1 + 63) / 64 * 64);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  (void)cb::launch_k(cb::pack_weight_kernel, dim3((unsigned)(blocks < 148 * 16 ? blocks : 148 * 16)), dim3(threads), (size_t)(0), stream, w, cout, (int)c0, (int)c1, taps,
                                                                                                    (cb::act_t*)out);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return cb::CB_OK;
}
