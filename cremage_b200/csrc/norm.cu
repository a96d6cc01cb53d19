// Bandwidth-bound normalisation kernels: GroupNorm(+SiLU) over NHWC bf16 (two-source capable), LayerNorm,
// row softmax.  All statistics in fp32; 16-byte vector loads/stores; warp-shuffle / smem reductions.
#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

struct alignas(16) Vec8 { uint32_t u[4]; };

CB_DEVINL Vec8 ld_vec8(const act_t* p) {
  Vec8 v;
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v.u[0] = t.x; v.u[1] = t.y; v.u[2] = t.z; v.u[3] = t.w;
  return v;
}
CB_DEVINL uint4 ld_shared_v4_u(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
CB_DEVINL void st_vec8(act_t* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]), pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
}
CB_DEVINL void unpack8(const Vec8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = unpack_act2(v.u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// y[i] <- y[i] * sigmoid(y[i]) for 8 values with 8 MUFU.EX2 + 2 MUFU.RCP: one reciprocal serves four denominators
// (1/a0 = a1 * (a2*a3) / (a0*a1*a2*a3) ...), which keeps the XU pipe (16 MUFU/clk/SM) and the issue slots of the
// streaming GroupNorm+SiLU pass below the HBM bound.  (Measured alternative: part of the exponentials as FMA-pipe
// polynomials -- faster in short bursts, slower under the sustained power cap; profiles/r1_groupnorm_fused_stats.md.)
// The exponent is clamped at 2^30 so the 4-way product stays finite; silu(x) for x < -20.8 is below the smallest
// 16-bit subnormal either way.
CB_DEVINL void silu8(float (&y)[8]) {
#pragma unroll
  for (int h = 0; h < 8; h += 4) {
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(-1.4426950408889634f * y[h + i], 30.f)));
      a[i] = 1.f + e;
    }
    const float p01 = a[0] * a[1], p23 = a[2] * a[3];
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    const float r01 = r * p23, r23 = r * p01;
    y[h + 0] *= r01 * a[1];
    y[h + 1] *= r01 * a[0];
    y[h + 2] *= r23 * a[3];
    y[h + 3] *= r23 * a[2];
  }
}

// ------------------------------------------------------------------------------------------------------------
// GroupNorm pass 1: per-(image, group) sum and sum of squares -- DETERMINISTIC (no floating-point atomics):
// grid = (splits, n); block = CV * P threads where CV = C/8 channel vectors; a thread keeps a FIXED channel vector
// and walks pixels p0 + lane_p, +P, ... accumulating 8 per-channel sums in registers; the CTA reduces them per group
// in a fixed order and writes one partial per (image, split, group); the last CTA of an image to finish (ticket
// counter) folds the partials in split order into stats[n][group][2].
// ------------------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const act_t* __restrict__ x0, int c0, const act_t* __restrict__ x1,
                                int c1, long long hw, int groups, int P, long long pix_per_cta,
                                float* __restrict__ stats, float* __restrict__ partials,
                                unsigned int* __restrict__ counters) {
  pdl_prologue();
  extern __shared__ float s_part[];  // [2][P][C]
  __shared__ int s_is_last;
  const int C = c0 + c1;
  const int CV = C >> 3;
  const int cv = threadIdx.x % CV;
  const int lp = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int splits = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;  // blockDim % 32 == 0

  const int c = cv << 3;
  const act_t* src;
  long long ld;
  if (c < c0) { src = x0 + (long long)n * hw * c0 + c; ld = c0; }
  else        { src = x1 + (long long)n * hw * c1 + (c - c0); ld = c1; }

  const long long p_begin = (long long)blockIdx.x * pix_per_cta;
  long long p_end = p_begin + pix_per_cta;
  if (p_end > hw) p_end = hw;
  if (lp >= P) p_end = p_begin;  // block is padded to whole warps; the padding threads only help reduce

  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  long long p = p_begin + lp;
  // 4 independent 16-byte loads in flight per thread
  for (; p + 3LL * P < p_end; p += 4LL * P) {
    Vec8 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_vec8(src + (p + (long long)u * P) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
  for (; p < p_end; p += P) {
    float f[8];
    unpack8(ld_vec8(src + p * ld), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
  }
  if (lp < P) {
    float* sp_s = s_part + (long long)lp * C + c;
    float* sp_q = s_part + (long long)(P + lp) * C + c;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sp_s[i] = s[i]; sp_q[i] = q[i]; }
  }
  __syncthreads();
  // per group: a warp sums the gs*P per-channel partials lane-strided, then a fixed-order shuffle tree
  const int gs = C / groups;
  float* my_partials = partials + ((long long)n * splits + blockIdx.x) * groups * 2;
  for (int g = warp; g < groups; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int e = lane; e < gs * P; e += 32) {
      const int l = e / gs, ch = g * gs + (e - l * gs);
      a += s_part[(long long)l * C + ch];
      b += s_part[(long long)(P + l) * C + ch];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { my_partials[g * 2] = a; my_partials[g * 2 + 1] = b; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&counters[n], 1u);
    s_is_last = (ticket == (unsigned int)splits - 1u);
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  const float* img_partials = partials + (long long)n * splits * groups * 2;
  for (int g = warp; g < groups; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int sp = lane; sp < splits; sp += 32) {
      a += __ldcg(img_partials + ((long long)sp * groups + g) * 2);
      b += __ldcg(img_partials + ((long long)sp * groups + g) * 2 + 1);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      stats[((long long)n * groups + g) * 2 + 0] = a;
      stats[((long long)n * groups + g) * 2 + 1] = b;
    }
  }
  if (threadIdx.x == 0) counters[n] = 0u;  // self-reset for the next call
}

// GroupNorm pass 2 (streaming): y = silu?((x - mean) * rstd * gamma + beta) -> 16-bit [n][hw][C]; one read + one
// write per element.  The per-(image, group) statistics come either as folded sums stats[n][groups][2] (stand-alone
// pass 1) or as S-row partial tables part_s[n][S_s][2][c_s/2] (sum | sum of squares per channel pair, written by the
// producing cb_igemm launches, reduced to <= GN_PART_ROWS rows by gn_fold1_kernel): the CTA folds them in a fixed
// order in fp64 in its prologue.
constexpr int GN_PART_ROWS = 32;
constexpr int GN_APPLY_UNROLL = 4;   // 16-byte loads in flight per thread of the apply pass
constexpr int GN_PART_MAXC = 2560;   // widest (concatenated) input of the from-partials path

// per-thread affine coefficients of channels [c, c+8) of image n: y = a * x + b
template <bool PARTS>
CB_DEVINL void gn_coefficients(int c0, int c1, long long hw, int groups, float eps, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ stats,
                               const float* __restrict__ part0, int S0, const float* __restrict__ part1, int S1, int n,
                               int c, bool active, float (&a)[8], float (&b)[8]) {
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float s_col[PARTS ? 2 * GN_PART_MAXC / 2 : 1];   // [plane][pair of the concat]: column totals of the tables
  __shared__ double s_gs[PARTS ? 128 : 1];                    // [plane][group]
  const int C = c0 + c1;
  const int gs = C / groups;
  const float inv_cnt = 1.f / (float(gs) * float(hw));
  if (PARTS) {
    // 1. every thread totals whole columns of the S-row tables: independent, coalesced loads (8 in flight)
    const int hc = C >> 1, h0 = c0 >> 1;
    for (int idx = threadIdx.x; idx < C; idx += blockDim.x) {
      const int plane = idx >= hc, pc = idx - plane * hc;
      const float* q;
      int S, W;
      if (pc < h0) { S = S0; W = c0; q = part0 + (long long)n * S0 * c0 + plane * h0 + pc; }
      else         { S = S1; W = c1; q = part1 + (long long)n * S1 * c1 + plane * (c1 >> 1) + (pc - h0); }
      double acc = 0.0;
      int r = 0;
      for (; r + 8 <= S; r += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(q + (long long)(r + u) * W);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += (double)v[u];
      }
      for (; r < S; ++r) acc += (double)__ldg(q + (long long)r * W);
      s_col[idx] = (float)acc;
    }
    __syncthreads();
    // 2. per (plane, group): the group's pairs in channel order
    if ((int)threadIdx.x < 2 * groups) {
      const int plane = threadIdx.x / groups, g = threadIdx.x - plane * groups;
      const float* q = s_col + plane * hc + ((g * gs) >> 1);
      double acc = 0.0;
      for (int i = 0; i < (gs >> 1); ++i) acc += (double)q[i];
      s_gs[threadIdx.x] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
      const double mean = s_gs[threadIdx.x] * (double)inv_cnt;
      double var = s_gs[groups + threadIdx.x] * (double)inv_cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[threadIdx.x] = (float)mean;
      s_rstd[threadIdx.x] = rsqrtf((float)var + eps);
    }
    __syncthreads();
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c + i) / gs;
    float mean, rstd;
    if (PARTS) {
      mean = s_mean[g];
      rstd = s_rstd[g];
    } else {
      const float sum = stats[((long long)n * groups + g) * 2 + 0];
      const float sq = stats[((long long)n * groups + g) * 2 + 1];
      mean = sum * inv_cnt;
      rstd = rsqrtf(fmaxf(sq * inv_cnt - mean * mean, 0.f) + eps);
    }
    a[i] = rstd * gamma[c + i];
    b[i] = beta[c + i] - mean * a[i];
  }
}

template <bool SILU>
CB_DEVINL void gn_affine_store(const Vec8& v, const float (&a)[8], const float (&b)[8], act_t* dst) {
  float f[8];
  unpack8(v, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], a[k], b[k]);
  if (SILU) silu8(f);
  st_vec8(dst, f);
}

// register-path apply (tensors of a few chunks per CTA: launch-latency bound anyway)
template <bool SILU, bool PARTS>
__global__ void gn_apply_kernel(const act_t* __restrict__ x0, int c0, const act_t* __restrict__ x1,
                                int c1, long long hw, int groups, int P, float eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ stats, const float* __restrict__ part0, int S0,
                                const float* __restrict__ part1, int S1, act_t* __restrict__ out) {
  pdl_prologue();
  const int C = c0 + c1;
  const int CV = C >> 3;
  const int cv = threadIdx.x % CV;
  const int lp = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int c = cv << 3;
  float a[8], b[8];
  gn_coefficients<PARTS>(c0, c1, hw, groups, eps, gamma, beta, stats, part0, S0, part1, S1, n, c, lp < P, a, b);
  if (lp >= P) return;  // padding threads of the warp-rounded block
  const act_t* src;
  long long ld;
  if (c < c0) { src = x0 + (long long)n * hw * c0 + c; ld = c0; }
  else        { src = x1 + (long long)n * hw * c1 + (c - c0); ld = c1; }
  act_t* dst = out + (long long)n * hw * C + c;
  // chunks of U*P consecutive pixels, interleaved over the CTAs of the image: at any moment the resident CTAs stream one
  // narrow window of the tensor, not gridDim.x far-apart ranges
  constexpr int U = GN_APPLY_UNROLL;
  const long long chunk = (long long)U * P;
  for (long long base = (long long)blockIdx.x * chunk; base < hw; base += (long long)gridDim.x * chunk) {
    const long long p = base + lp;
    if (base + chunk <= hw) {
      Vec8 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ld_vec8(src + (p + (long long)u * P) * ld);
#pragma unroll
      for (int u = 0; u < U; ++u) gn_affine_store<SILU>(v[u], a, b, dst + (p + (long long)u * P) * C);
    } else {
      for (long long pp = p; pp < hw; pp += P) gn_affine_store<SILU>(ld_vec8(src + pp * ld), a, b, dst + pp * C);
    }
  }
}

// Streaming apply for tensors far larger than the L2 (the VAE decoder): the input moves HBM -> shared memory with 1-D
// bulk copies (cp.async.bulk + mbarrier complete_tx) through a GN_TMA_STAGES-deep ring per CTA, so the bytes in flight
// per SM are set by shared memory (3 CTAs x 3 stages x 16 KiB), not by the registers of the loading threads; the
// threads read their 16-byte vectors from the ring, normalise and store straight to global.  A chunk of U*P pixels is
// one contiguous block per source.
constexpr int GN_TMA_STAGES = 3;
constexpr int GN_TMA_THREADS = 256;

CB_DEVINL void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <bool SILU, bool PARTS>
__global__ void __launch_bounds__(GN_TMA_THREADS, 3)
gn_apply_tma_kernel(const act_t* __restrict__ x0, int c0, const act_t* __restrict__ x1, int c1, long long hw, int groups,
                    int P, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ stats, const float* __restrict__ part0, int S0,
                    const float* __restrict__ part1, int S1, act_t* __restrict__ out) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t gn_ring[];   // [stages][chunk of source 0 | chunk of source 1]
  __shared__ __align__(8) uint64_t s_full[GN_TMA_STAGES];
  constexpr int U = GN_APPLY_UNROLL;
  const int C = c0 + c1;
  const int CV = C >> 3;
  const int cv = threadIdx.x % CV;
  const int lp = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int c = cv << 3;
  const long long chunk = (long long)U * P;                  // pixels per chunk
  const long long nfull = hw / chunk;
  const uint32_t bytes0 = (uint32_t)(chunk * c0 * 2), bytes1 = (uint32_t)(chunk * c1 * 2);
  const uint32_t stage_bytes = bytes0 + bytes1;
  const uint32_t ring = smem_u32(gn_ring);
  const act_t* g0 = x0 + (long long)n * hw * c0;
  const act_t* g1 = c1 ? x1 + (long long)n * hw * c1 : nullptr;
  auto issue = [&](long long i, int s) {                      // one thread: chunk i -> stage s
    const uint32_t bar = smem_u32(&s_full[s]);
    mbar_expect_tx(bar, stage_bytes);
    bulk_load_1d(ring + (uint32_t)s * stage_bytes, g0 + i * chunk * c0, bytes0, bar);
    if (bytes1) bulk_load_1d(ring + (uint32_t)s * stage_bytes + bytes0, g1 + i * chunk * c1, bytes1, bar);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < GN_TMA_STAGES; ++s) mbar_init(smem_u32(&s_full[s]), 1);
    fence_mbar_init();
    for (int s = 0; s < GN_TMA_STAGES; ++s) {
      const long long i = blockIdx.x + (long long)s * gridDim.x;
      if (i < nfull) issue(i, s);                             // the ring fills while the prologue folds the statistics
    }
  }
  float a[8], b[8];
  gn_coefficients<PARTS>(c0, c1, hw, groups, eps, gamma, beta, stats, part0, S0, part1, S1, n, c, lp < P, a, b);
  __syncthreads();   // barriers initialised (PARTS = false has no barrier inside gn_coefficients)
  const bool active = lp < P;
  // this thread's vector inside a stage: pixel u*P + lp of the chunk, channels [c, c+8) of its source
  const uint32_t voff = c < c0 ? (uint32_t)(lp * c0 + c) * 2u : bytes0 + (uint32_t)(lp * c1 + (c - c0)) * 2u;
  const uint32_t vstep = (uint32_t)P * (uint32_t)(c < c0 ? c0 : c1) * 2u;
  act_t* dst = out + (long long)n * hw * C + c;
  int s = 0;
  uint32_t parity = 0;
  for (long long i = blockIdx.x; i < nfull; i += gridDim.x) {
    mbar_wait(smem_u32(&s_full[s]), parity);
    Vec8 v[U];
    if (active) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint4 t = ld_shared_v4_u(ring + (uint32_t)s * stage_bytes + voff + (uint32_t)u * vstep);
        v[u].u[0] = t.x; v[u].u[1] = t.y; v[u].u[2] = t.z; v[u].u[3] = t.w;
      }
    }
    // The refill below is an ASYNC-proxy write to shared memory that these GENERIC-proxy loads have just read: the
    // CTA barrier orders the loads against the other threads, not against the bulk-copy engine, so every thread
    // fences its reads into the async proxy first.  (Without it one 256-byte pixel of a stage was, about once in a
    // million chunks, read after the next chunk had started to land: one wrong pixel in ~3 % of the 8 x 512 x 512 x 128
    // calls, tools/gn_repro.py.)
    fence_proxy_async_smem();
    __syncthreads();   // every thread holds its vectors: the stage may be refilled
    if (threadIdx.x == 0) {
      const long long nx = i + (long long)GN_TMA_STAGES * gridDim.x;
      if (nx < nfull) issue(nx, s);
    }
    if (active) {
      const long long p = i * chunk + lp;
#pragma unroll
      for (int u = 0; u < U; ++u) gn_affine_store<SILU>(v[u], a, b, dst + (p + (long long)u * P) * C);
    }
    if (++s == GN_TMA_STAGES) { s = 0; parity ^= 1u; }
  }
  // the image's last, partial chunk
  if (blockIdx.x == gridDim.x - 1 && active) {
    const act_t* src = c < c0 ? g0 + c : g1 + (c - c0);
    const long long ld = c < c0 ? c0 : c1;
    for (long long pp = nfull * chunk + lp; pp < hw; pp += P) gn_affine_store<SILU>(ld_vec8(src + pp * ld), a, b, dst + pp * C);
  }
}

// First-level fold of the producers' partial tables: in[n][bpi][W] (W = 2 * (c/2) floats per M tile) ->
// out[n][S][W], row s = sum of blocks [s*bpi/S, (s+1)*bpi/S) in block order, fp64 accumulation (deterministic).
// grid = (S, n); a thread owns column t % W and every (256/W)-th block of the range; fully coalesced reads.
constexpr int GN_FOLD_THREADS = 256;

__global__ void __launch_bounds__(GN_FOLD_THREADS)
gn_fold1_kernel(const float* __restrict__ in, int bpi, int W, int S, float* __restrict__ out) {
  pdl_prologue();
  __shared__ double s_red[GN_FOLD_THREADS];
  const int s = blockIdx.x, n = blockIdx.y;
  const int b_lo = int((long long)s * bpi / S), b_hi = int((long long)(s + 1) * bpi / S);
  const int wcols = W < GN_FOLD_THREADS ? W : GN_FOLD_THREADS;      // columns handled per sweep
  const int lanes = GN_FOLD_THREADS / wcols;                         // block-lanes per column
  const int col_in = threadIdx.x % wcols, bl = threadIdx.x / wcols;
  const float* base = in + (long long)n * bpi * W;
  float* dst = out + ((long long)n * S + s) * W;
  for (int col0 = 0; col0 < W; col0 += wcols) {
    const int col = col0 + col_in;
    double acc = 0.0;
    if (col < W && bl < lanes) {
      int b = b_lo + bl;
      for (; b + 3 * lanes < b_hi; b += 4 * lanes) {
        const float v0 = __ldg(base + (long long)b * W + col);
        const float v1 = __ldg(base + (long long)(b + lanes) * W + col);
        const float v2 = __ldg(base + (long long)(b + 2 * lanes) * W + col);
        const float v3 = __ldg(base + (long long)(b + 3 * lanes) * W + col);
        acc += (double)v0; acc += (double)v1; acc += (double)v2; acc += (double)v3;
      }
      for (; b < b_hi; b += lanes) acc += (double)__ldg(base + (long long)b * W + col);
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    if (bl == 0 && col < W) {
      double t = s_red[col_in];
      for (int l = 1; l < lanes; ++l) t += s_red[l * wcols + col_in];
      dst[col] = (float)t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// GroupNorm, single pass for tensors whose per-image group slab fits the shared memory of one thread-block cluster
// (every GroupNorm of the UNet): a cluster of 8 CTAs owns (image n, a set of `gset` groups); CTA r loads pixels
// [r*ppc, (r+1)*ppc) of those channels ONCE into shared memory while accumulating statistics, the 8 partials are
// exchanged through distributed shared memory (fixed order -> deterministic), then each CTA normalises its slab from
// shared memory and writes it.  HBM traffic = one read + one write, the algorithmic minimum.
// ------------------------------------------------------------------------------------------------------------
constexpr int GN_CLUSTER = 8;

__global__ void gn_cluster_kernel(const act_t* __restrict__ x0, int c0, const act_t* __restrict__ x1, int c1,
                                  long long hw, int groups, int gset, int P, int ppc, float eps,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                  act_t* __restrict__ out) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t gn_smem[];
  const int C = c0 + c1;
  const int gs = C / groups;
  const int nch = gset * gs;               // channels of this cluster (multiple of 8)
  const int cvs = nch >> 3;
  uint4* slab = reinterpret_cast<uint4*>(gn_smem);                                   // [ppc][cvs] 16-byte vectors
  float* s_part = reinterpret_cast<float*>(gn_smem + (size_t)ppc * cvs * 16);        // [2][P][nch]
  float* s_grp = s_part + 2 * P * nch;                                               // [gset][2] this CTA's partial
  float* s_stat = s_grp + 2 * gset;                                                  // [gset][2] mean, rstd
  const unsigned rank = blockIdx.x;  // cluster spans gridDim.x == GN_CLUSTER
  const int n = blockIdx.z;
  const int c_lo = blockIdx.y * nch;
  const int cv = threadIdx.x % cvs;
  const int lp = threadIdx.x / cvs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int c = c_lo + (cv << 3);
  const act_t* src;
  long long ld;
  if (c < c0) { src = x0 + (long long)n * hw * c0 + c; ld = c0; }
  else        { src = x1 + (long long)n * hw * c1 + (c - c0); ld = c1; }
  const long long p_lo = (long long)rank * ppc;
  long long p_hi = p_lo + ppc;
  if (p_hi > hw) p_hi = hw;

  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (lp < P) {
    long long pp = p_lo + lp;
    for (; pp + 3LL * P < p_hi; pp += 4LL * P) {   // four independent 16-byte loads in flight
      uint4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const uint4*>(src + (pp + (long long)u * P) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        slab[(pp + (long long)u * P - p_lo) * cvs + cv] = t[u];
        Vec8 v; v.u[0] = t[u].x; v.u[1] = t[u].y; v.u[2] = t[u].z; v.u[3] = t[u].w;
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    for (; pp < p_hi; pp += P) {
      const uint4 t = *reinterpret_cast<const uint4*>(src + pp * ld);
      slab[(pp - p_lo) * cvs + cv] = t;
      Vec8 v; v.u[0] = t.x; v.u[1] = t.y; v.u[2] = t.z; v.u[3] = t.w;
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
    float* ps = s_part + (long long)lp * nch + (cv << 3);
    float* pq = s_part + (long long)(P + lp) * nch + (cv << 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) { ps[i] = s[i]; pq[i] = q[i]; }
  }
  __syncthreads();
  for (int g = warp; g < gset; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int e = lane; e < gs * P; e += 32) {
      const int l = e / gs, ch = g * gs + (e - l * gs);
      a += s_part[(long long)l * nch + ch];
      b += s_part[(long long)(P + l) * nch + ch];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { s_grp[2 * g] = a; s_grp[2 * g + 1] = b; }
  }
  // cluster barrier #1: every CTA's partials are written
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if ((int)threadIdx.x < gset) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    const uint32_t local = smem_u32(s_grp + 2 * g);
    for (unsigned r = 0; r < GN_CLUSTER; ++r) {   // fixed rank order: deterministic
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      float ra, rb;
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(ra) : "r"(remote) : "memory");
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(rb) : "r"(remote + 4u) : "memory");
      a += ra;
      b += rb;
    }
    const float inv_cnt = 1.f / (float(gs) * float(hw));
    const float mean = a * inv_cnt;
    const float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
    s_stat[2 * g] = mean;
    s_stat[2 * g + 1] = rsqrtf(var + eps);
  }
  // cluster barrier #2: all remote reads are done (a CTA may exit afterwards) and s_stat is visible CTA-wide
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (lp >= P) return;
  float ga[8], gb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = ((cv << 3) + i) / gs;
    const float mean = s_stat[2 * g], rstd = s_stat[2 * g + 1];
    ga[i] = rstd * gamma[c + i];
    gb[i] = beta[c + i] - mean * ga[i];
  }
  act_t* dst = out + (long long)n * hw * C + c;
  for (long long pp = p_lo + lp; pp < p_hi; pp += P) {
    const uint4 t = slab[(pp - p_lo) * cvs + cv];
    Vec8 v; v.u[0] = t.x; v.u[1] = t.y; v.u[2] = t.z; v.u[3] = t.w;
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], ga[i], gb[i]);
    if (silu) silu8(f);
    st_vec8(dst + pp * C, f);
  }
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (two-pass mean / variance, exact in fp32).
// ------------------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void layernorm_kernel(const act_t* __restrict__ x, long long rows, int C, float eps,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 act_t* __restrict__ out) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int CV = C >> 3;
  const act_t* src = x + row * C;
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < CV) {
      unpack8(ld_vec8(src + v * 8), f[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += f[i][e];
    }
  }
  sum = warp_sum(sum);
  const float mean = sum / float(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < CV) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = f[i][e] - mean; sq = fmaf(d, d, sq); }
    }
  }
  sq = warp_sum(sq);
  const float rstd = rsqrtf(sq / float(C) + eps);
  act_t* dst = out + row * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < CV) {
      float y[8];
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = fmaf((f[i][e] - mean) * rstd, gg[e], bb[e]);
      st_vec8(dst + v * 8, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Row softmax: dst[r][:] = softmax(scale * src[r][:]), src fp32 or bf16, dst bf16 (may alias a bf16 src), fp32 math;
// one CTA per row, the row is staged in shared memory.
// ------------------------------------------------------------------------------------------------------------
template <bool SRC_F32>
__global__ void softmax_rows_kernel(const void* __restrict__ src, long long src_ld, act_t* __restrict__ dst,
                                    long long dst_ld, long long cols, float scale) {
  pdl_prologue();
  extern __shared__ float s_row[];
  __shared__ float red[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  float m = -INFINITY;
  if (SRC_F32) {
    const float* row = reinterpret_cast<const float*>(src) + (long long)blockIdx.x * src_ld;
    for (long long v = tid; v < (cols >> 2); v += blockDim.x) {
      const float4 t = *reinterpret_cast<const float4*>(row + v * 4);
      const float f[4] = {t.x * scale, t.y * scale, t.z * scale, t.w * scale};
#pragma unroll
      for (int e = 0; e < 4; ++e) { s_row[v * 4 + e] = f[e]; m = fmaxf(m, f[e]); }
    }
  } else {
    const act_t* row = reinterpret_cast<const act_t*>(src) + (long long)blockIdx.x * src_ld;
    for (long long v = tid; v < (cols >> 3); v += blockDim.x) {
      float f[8];
      unpack8(ld_vec8(row + v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { f[e] *= scale; s_row[v * 8 + e] = f[e]; m = fmaxf(m, f[e]); }
    }
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (long long i = tid; i < cols; i += blockDim.x) {
    const float e = __expf(s_row[i] - m);
    s_row[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < nw; ++i) sum += red[i];
  const float inv = 1.f / sum;
  act_t* orow = dst + (long long)blockIdx.x * dst_ld;
  for (long long v = tid; v < (cols >> 3); v += blockDim.x) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = s_row[v * 8 + e] * inv;
    st_vec8(orow + v * 8, f);
  }
}

}  // namespace cb

using namespace cb;

namespace {
struct GnPlan { int CV, P, threads; long long splits, pix_per_cta; size_t smem; };
GnPlan gn_plan(int64_t n, int64_t hw, int64_t C) {
  GnPlan g;
  g.CV = int(C / 8);
  g.P = 384 / g.CV;
  if (g.P < 1) g.P = 1;
  if ((int64_t)g.P > hw) g.P = (int)hw;
  g.threads = (g.CV * g.P + 31) / 32 * 32;  // whole warps: the reductions shuffle with a full mask
  // enough CTAs to fill 148 SMs a few times over, at least ~8 pixels per thread-lane when the image is large
  long long splits = (148LL * 4 + n - 1) / n;
  long long max_splits = (hw + (long long)g.P * 8 - 1) / ((long long)g.P * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long ppc = (hw + splits - 1) / splits;
  ppc = ((ppc + g.P - 1) / g.P) * g.P;
  g.splits = (hw + ppc - 1) / ppc;
  g.pix_per_cta = ppc;
  g.smem = sizeof(float) * 2 * (size_t)g.P * (size_t)C;
  return g;
}
// CTAs per image of the streaming apply pass: exactly the resident slots of the device (chunks are interleaved over the
// CTAs, so any count balances; a partial second wave would cost a whole extra pass)
// block shape of the apply pass: ~256 threads = CV channel vectors x P pixels (at ~60 registers four such CTAs fill an
// SM; 384-thread CTAs would leave a third of the register file idle)
int gn_apply_shape(int64_t C, int64_t hw, int* P) {
  const int CV = int(C / 8);
  int p = 256 / CV;
  if (p < 1) p = 1;
  if ((int64_t)p > hw) p = (int)hw;
  *P = p;
  return (CV * p + 31) / 32 * 32;
}
template <typename K>
unsigned gn_apply_ctas(K kernel, int threads, size_t smem, int64_t n, int64_t hw, int P) {
  const int sms = sm_count();
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long ctas = (long long)sms * per_sm / n;
  const long long chunks = (hw + (long long)GN_APPLY_UNROLL * P - 1) / ((long long)GN_APPLY_UNROLL * P);
  if (ctas > chunks) ctas = chunks;
  if (ctas < 1) ctas = 1;
  return (unsigned)ctas;
}

// the streaming pass of both GroupNorm forms: bulk-copy ring for tensors beyond the L2, register path otherwise
template <bool PARTS>
int launch_gn_apply(const void* x0, int64_t c0, const void* x1, int64_t c1, int64_t n, int64_t hw, int groups, float eps,
                    const float* gamma, const float* beta, int silu, const float* stats, const float* p0, int S0,
                    const float* p1, int S1, void* out, cudaStream_t stream) {
  const int64_t C = c0 + c1;
  const int CV = int(C / 8);
  static const long long tma_min_bytes = [] {
    const char* e = getenv("CB_GN_TMA_MIN_BYTES");
    return e ? atoll(e) : 0LL;   // every tensor with at least four chunks per CTA row takes the ring
  }();
  if (CV <= GN_TMA_THREADS && 2 * n * hw * C >= tma_min_bytes && hw >= 4LL * GN_APPLY_UNROLL * (GN_TMA_THREADS / CV)) {
    const int P = GN_TMA_THREADS / CV;
    const size_t smem = (size_t)GN_TMA_STAGES * GN_APPLY_UNROLL * P * C * 2;
    auto k = silu ? gn_apply_tma_kernel<true, PARTS> : gn_apply_tma_kernel<false, PARTS>;
    static DeviceOnce cfg{};
    if (device_once_needed(cfg)) {
      CB_CHECK_CUDA(cudaFuncSetAttribute(gn_apply_tma_kernel<true, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      CB_CHECK_CUDA(cudaFuncSetAttribute(gn_apply_tma_kernel<false, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      device_once_done(cfg);
    }
    dim3 grid(gn_apply_ctas(k, GN_TMA_THREADS, smem, n, hw, P), (unsigned)n);
    (void)cb::launch_k(k, dim3(grid), dim3(GN_TMA_THREADS), (size_t)(smem), stream, (const act_t*)x0, (int)c0, (const act_t*)x1, (int)c1, hw, groups, P, eps, gamma,
                                              beta, stats, p0, S0, p1, S1, (act_t*)out);
  } else {
    int P = 1;
    const int threads = gn_apply_shape(C, hw, &P);
    auto k = silu ? gn_apply_kernel<true, PARTS> : gn_apply_kernel<false, PARTS>;
    dim3 grid(gn_apply_ctas(k, threads, 0, n, hw, P), (unsigned)n);
    (void)cb::launch_k(k, dim3(grid), dim3(threads), (size_t)(0), stream, (const act_t*)x0, (int)c0, (const act_t*)x1, (int)c1, hw, groups, P, eps, gamma, beta, stats,
                                    p0, S0, p1, S1, (act_t*)out);
  }
  CB_CHECK_CUDA(cudaGetLastError());
  return CB_OK;
}
size_t gn_ws_floats(const GnPlan& g, int64_t n, int groups) {
  return (size_t)n * groups * 2 + (size_t)n * g.splits * groups * 2 + (size_t)n;
}
}  // namespace

extern "C" int64_t cb_groupnorm_workspace_bytes(int64_t c, int64_t n, int64_t hw, int groups) {
  if (c <= 0 || n <= 0 || hw <= 0 || groups <= 0 || c % 8) return 0;
  return (int64_t)(gn_ws_floats(gn_plan(n, hw, c), n, groups) * sizeof(float));
}

extern "C" int cb_groupnorm_nhwc(const void* x0, int64_t c0, const void* x1, int64_t c1, int64_t n, int64_t hw,
                                 int groups, float eps, const float* gamma, const float* beta, int silu, void* out,
                                 float* stats, cudaStream_t stream) {
  CB_REQUIRE(x0 && out && stats && gamma && beta, "cb_groupnorm_nhwc: null pointer");
  const int64_t C = c0 + c1;
  CB_REQUIRE(c0 > 0 && c0 % 8 == 0 && c1 >= 0 && c1 % 8 == 0, "cb_groupnorm_nhwc: channels must be multiples of 8");
  CB_REQUIRE(c1 == 0 || x1, "cb_groupnorm_nhwc: c1 > 0 but x1 is null");
  CB_REQUIRE(groups > 0 && C % groups == 0, "cb_groupnorm_nhwc: %lld channels not divisible into %d groups", (long long)C, groups);
  CB_REQUIRE(n > 0 && hw > 0, "cb_groupnorm_nhwc: empty input");
  CB_REQUIRE(C / 8 <= 1024, "cb_groupnorm_nhwc: more than 8192 channels unsupported");
  // ---- single-pass cluster kernel when the per-image slab of a group set fits 8 CTAs' shared memory
  {
    const int gs = int(C / groups);
    const int ppc = int((hw + GN_CLUSTER - 1) / GN_CLUSTER);
    int gset = 0;
    // largest group set (longest contiguous per-pixel run, fewest clusters) whose slab fits and that still yields
    // >= 256 CTAs; measured better than many tiny clusters (cluster launch + two cluster barriers per CTA dominate)
    for (int cand = groups; cand >= 1; cand >>= 1) {
      const long long nch = (long long)cand * gs;
      if (groups % cand || nch % 8 || nch * 2 < 64) continue;
      if ((long long)ppc * nch * 2 <= 96 * 1024 && (long long)n * (groups / cand) * GN_CLUSTER >= 256) { gset = cand; break; }
    }
    if (gset == 0) {  // small batch: take the largest set that fits
      for (int cand = groups; cand >= 1; cand >>= 1) {
        const long long nch = (long long)cand * gs;
        if (groups % cand || nch % 8 || nch * 2 < 64) continue;
        if ((long long)ppc * nch * 2 <= 96 * 1024) { gset = cand; break; }
      }
    }
    if (gset > 0 && hw >= GN_CLUSTER && n <= 65535) {
      const int nch = gset * gs, cvs = nch / 8;
      int P = 256 / cvs;
      if (P < 1) P = 1;
      if (P > ppc) P = ppc;
      const int threads = (cvs * P + 31) / 32 * 32;
      if (threads <= 1024) {
        const size_t smem = (size_t)ppc * nch * 2 + sizeof(float) * (2 * (size_t)P * nch + 4 * (size_t)gset) + 16;
        static DeviceOnce cfg{};
        if (device_once_needed(cfg)) {
          CB_CHECK_CUDA(cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          device_once_done(cfg);
        }
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(GN_CLUSTER, (unsigned)(groups / gset), (unsigned)n);
        lc.blockDim = dim3((unsigned)threads);
        lc.dynamicSmemBytes = smem;
        lc.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = GN_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = pdl_enabled() ? 2 : 1;
        CB_CHECK_CUDA(cudaLaunchKernelEx(&lc, gn_cluster_kernel, (const act_t*)x0, (int)c0, (const act_t*)x1, (int)c1,
                                         (long long)hw, groups, gset, P, ppc, eps, gamma, beta, silu, (act_t*)out));
        CB_LAUNCHED(1);
        return CB_OK;
      }
    }
  }
  const GnPlan g = gn_plan(n, hw, C);
  CB_REQUIRE(g.smem <= 160 * 1024, "cb_groupnorm_nhwc: needs %zu bytes of shared memory", g.smem);
  // workspace: [n][groups][2] final sums | [n][splits][groups][2] partials | [n] ticket counters
  float* partials = stats + (size_t)n * groups * 2;
  unsigned int* counters = reinterpret_cast<unsigned int*>(partials + (size_t)n * g.splits * groups * 2);
  CB_CHECK_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * n, stream));
  static DeviceOnce configured{};
  if (device_once_needed(configured)) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    device_once_done(configured);
  }
  dim3 grid((unsigned)g.splits, (unsigned)n);
  (void)cb::launch_k(gn_stats_kernel, dim3(grid), dim3(g.threads), (size_t)(g.smem), stream, (const act_t*)x0, (int)c0, (const act_t*)x1,
                                                       (int)c1, hw, groups, g.P, g.pix_per_cta, stats, partials, counters);
  CB_CHECK_CUDA(cudaGetLastError());
  {
    const int rc = launch_gn_apply<false>(x0, c0, x1, c1, n, hw, groups, eps, gamma, beta, silu, stats, nullptr, 0, nullptr, 0, out, stream);
    if (rc) return rc;
  }
  CB_LAUNCHED(2);
  return CB_OK;
}

extern "C" int64_t cb_gn_partial_blocks(int64_t h, int64_t w, int tw, int th) {
  if (h <= 0 || w <= 0 || tw <= 0 || th <= 0 || (tw * th) % 32 != 0) return 0;
  return ((w + tw - 1) / tw) * ((h + th - 1) / th);   // one row per M tile of the image
}

extern "C" int cb_groupnorm_from_partials(const void* x0, int64_t c0, const float* part0, int64_t bpi0, const void* x1,
                                          int64_t c1, const float* part1, int64_t bpi1, int64_t n, int64_t hw, int groups,
                                          float eps, const float* gamma, const float* beta, int silu, void* out,
                                          float* stats, cudaStream_t stream) {
  CB_REQUIRE(x0 && part0 && out && stats && gamma && beta, "cb_groupnorm_from_partials: null pointer");
  const int64_t C = c0 + c1;
  CB_REQUIRE(c0 > 0 && c0 % 8 == 0 && c1 >= 0 && c1 % 8 == 0, "cb_groupnorm_from_partials: channels must be multiples of 8");
  CB_REQUIRE(c1 == 0 || (x1 && part1 && bpi1 > 0), "cb_groupnorm_from_partials: c1 > 0 needs x1 and its partials");
  CB_REQUIRE(groups > 0 && C % groups == 0 && (C / groups) % 2 == 0, "cb_groupnorm_from_partials: %lld channels / %d groups must be an even group size", (long long)C, groups);
  CB_REQUIRE(n > 0 && n <= 65535 && hw > 0 && bpi0 > 0, "cb_groupnorm_from_partials: empty input");
  CB_REQUIRE(C <= GN_PART_MAXC, "cb_groupnorm_from_partials: more than %d channels unsupported", GN_PART_MAXC);
  CB_REQUIRE(groups <= 64, "cb_groupnorm_from_partials: at most 64 groups");
  // first-level fold of a long partial table into <= GN_PART_ROWS rows (workspace: `stats`, both sources back to back)
  const float* p0 = part0; const float* p1 = part1;
  int S0 = (int)bpi0, S1 = (int)bpi1, launches = 1;
  float* ws = stats;
  if (bpi0 > GN_PART_ROWS) {
    (void)cb::launch_k(gn_fold1_kernel, dim3(GN_PART_ROWS, (unsigned)n), dim3(GN_FOLD_THREADS), (size_t)0, stream, part0, (int)bpi0, (int)c0, GN_PART_ROWS, ws);
    CB_CHECK_CUDA(cudaGetLastError());
    p0 = ws; S0 = GN_PART_ROWS; ws += (size_t)n * GN_PART_ROWS * c0; ++launches;
  }
  if (c1 > 0 && bpi1 > GN_PART_ROWS) {
    (void)cb::launch_k(gn_fold1_kernel, dim3(GN_PART_ROWS, (unsigned)n), dim3(GN_FOLD_THREADS), (size_t)0, stream, part1, (int)bpi1, (int)c1, GN_PART_ROWS, ws);
    CB_CHECK_CUDA(cudaGetLastError());
    p1 = ws; S1 = GN_PART_ROWS; ++launches;
  }
  {
    const int rc = launch_gn_apply<true>(x0, c0, x1, c1, n, hw, groups, eps, gamma, beta, silu, nullptr, p0, S0,
                                         c1 > 0 ? p1 : nullptr, S1, out, stream);
    if (rc) return rc;
  }
  CB_LAUNCHED(launches);
  return CB_OK;
}

extern "C" int cb_layernorm(const void* x, int64_t rows, int64_t c, float eps, const float* gamma, const float* beta,
                            void* out, cudaStream_t stream) {
  CB_REQUIRE(x && out && gamma && beta, "cb_layernorm: null pointer");
  CB_REQUIRE(c > 0 && c % 8 == 0 && c <= 4096, "cb_layernorm: width %lld unsupported (multiple of 8, <= 4096)", (long long)c);
  CB_REQUIRE(rows > 0, "cb_layernorm: empty input");
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  const int maxv = int((c / 8 + 31) / 32);
  auto X = (const act_t*)x;
  auto O = (act_t*)out;
  if (maxv <= 2)       (void)cb::launch_k(layernorm_kernel<2>, dim3(grid), dim3(warps * 32), (size_t)(0), stream, X, rows, (int)c, eps, gamma, beta, O);
  else if (maxv <= 4)  (void)cb::launch_k(layernorm_kernel<4>, dim3(grid), dim3(warps * 32), (size_t)(0), stream, X, rows, (int)c, eps, gamma, beta, O);
  else if (maxv <= 8)  (void)cb::launch_k(layernorm_kernel<8>, dim3(grid), dim3(warps * 32), (size_t)(0), stream, X, rows, (int)c, eps, gamma, beta, O);
  else                 (void)cb::launch_k(layernorm_kernel<16>, dim3(grid), dim3(warps * 32), (size_t)(0), stream, X, rows, (int)c, eps, gamma, beta, O);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

extern "C" int cb_softmax_rows(const void* src, int src_f32, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                               int64_t cols, float scale, cudaStream_t stream) {
  CB_REQUIRE(src && dst, "cb_softmax_rows: null pointer");
  CB_REQUIRE(cols > 0 && cols % 8 == 0 && cols <= 48 * 1024, "cb_softmax_rows: cols %lld unsupported (multiple of 8, <= 49152)", (long long)cols);
  CB_REQUIRE(src_ld % 8 == 0 && dst_ld % 8 == 0 && src_ld >= cols && dst_ld >= cols && rows > 0, "cb_softmax_rows: bad ld / rows");
  const size_t smem = sizeof(float) * cols;
  static DeviceOnce configured{};
  if (device_once_needed(configured)) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(softmax_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CHECK_CUDA(cudaFuncSetAttribute(softmax_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    device_once_done(configured);
  }
  if (src_f32) (void)cb::launch_k(softmax_rows_kernel<true>, dim3((unsigned)rows), dim3(256), (size_t)(smem), stream, src, src_ld, (act_t*)dst, dst_ld, cols, scale);
  else (void)cb::launch_k(softmax_rows_kernel<false>, dim3((unsigned)rows), dim3(256), (size_t)(smem), stream, src, src_ld, (act_t*)dst, dst_ld, cols, scale);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}
