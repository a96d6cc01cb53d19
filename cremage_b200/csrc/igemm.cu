// Implicit-GEMM kernel for sm_100a: every contraction of the denoising path that is a "rows x K  *  K x Cout"
// product -- conv3x3 (stride 1, stride 2 via parity planes), conv1x1, nn.Linear -- runs through this one kernel.
//
//   D[row, co] = sum_{tap, c} A_src(c)[pixel(row) + shift(tap), c] * Wt[co, tap, c]      (fp32 accumulate in TMEM)
//
// * A is read straight from the NHWC bf16 activation tensor(s) with 4-D TMA boxes {64 ch, TW, TH, TN}
//   (TW*TH*TN = 128 output pixels); a 3x3 tap is just a shifted box and the conv zero padding is TMA's
//   out-of-bounds zero fill.  No im2col buffer exists.  Two A sources give the UNet skip-concat for free
//   (reference: torch.cat([h, hs.pop()], 1) at ldm/modules/diffusionmodules/openaimodel.py:808).
// * B (weights, repacked [Cout][tap][Cin] bf16, K-major) is read with 2-D TMA boxes {64, BN}.
// * tcgen05.mma (cta_group::1, M=128, N=BN, K=16) issued by one thread, accumulator in TMEM.
// * Warp roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> bias / per-image bias / residual / SiLU / GEGLU / per-head scatter -> global).
// * Non-persistent grid, 2 CTAs per SM co-resident so one CTA's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 64;           // K elements per pipeline stage (= one 128-byte swizzle row)
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int NUM_THREADS = 192;

struct IGemmKParams {
  // rows
  int n_img, H, W;       // output pixel grid
  int TW, TH, TN;        // tile decomposition, TW*TH*TN == 128
  int tiles_w, tiles_h;  // tiles along w / h (tiles along n = gridDim.y / (tiles_w*tiles_h))
  // K loop
  int taps, chunks0, chunks1, num_k;
  int tap_dw[9], tap_dh[9], tap_dn[9];
  // N
  int cout, bn, stages;
  uint32_t idesc, tmem_cols;
  // epilogue
  int mode, act, out_f32;
  const float* bias;
  int bias_len;          // entries of `bias`
  const float* rowbias;
  long long rowbias_ld;
  int rowbias_vec;       // row-bias rows are 16-byte aligned -> float4 loads
  const act_t* residual;
  long long res_ld;
  void* out;
  long long out_ld;
  float out_scale;
  // heads mode
  int hd, hdpad, hheads, htokens;
  long long hwhich_stride;
};

__device__ __forceinline__ float apply_act(float v, int act) { return act == CB_ACT_SILU ? silu_f(v) : v; }

__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
             const __grid_constant__ CUtensorMap mapB, const IGemmKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16K][B bn*128] (1024-aligned), then barriers
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_stage_bytes = uint32_t(p.bn) * 128u;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const uint32_t bar_base = smem_base + uint32_t(p.stages) * stage_bytes;
  // barriers: full[s] at bar_base + 8*s ; empty[s] at bar_base + 8*(stages+s); accum at 8*(2*stages); tmem ptr after
  auto full_bar = [&](int s) { return bar_base + 8u * uint32_t(s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * uint32_t(p.stages + s); };
  const uint32_t accum_bar = bar_base + 8u * uint32_t(2 * p.stages);
  const uint32_t tmem_slot = accum_bar + 8u;
  const uint32_t bias_smem = (tmem_slot + 4u + 15u) & ~15u;  // float[bn]: this tile's bias slice
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int nt = blockIdx.x;  // N tile
  const int mt = blockIdx.y;  // M tile
  const int tw = mt % p.tiles_w;
  const int th = (mt / p.tiles_w) % p.tiles_h;
  const int tn = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.chunks1 > 0) tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int cpt = p.chunks0 + p.chunks1;
      int stage = 0;
      uint32_t phase = 0;
      int tap = 0, ch = 0;
      for (int kt = 0; kt < p.num_k; ++kt) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
        const uint32_t sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(full_bar(stage), stage_bytes);
        const int cw = w0 + p.tap_dw[tap], chh = h0 + p.tap_dh[tap], cn = n0 + p.tap_dn[tap];
        if (ch < p.chunks0) tma_load_4d(sa, &mapA0, full_bar(stage), ch * BK, cw, chh, cn);
        else                tma_load_4d(sa, &mapA1, full_bar(stage), (ch - p.chunks0) * BK, cw, chh, cn);
        tma_load_2d(sb, &mapB, full_bar(stage), kt * BK, nt * p.bn);
        if (++ch == cpt) { ch = 0; ++tap; }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = 0; kt < p.num_k; ++kt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint64_t adesc = make_sdesc_sw128(sa, 16, 1024);
        const uint64_t bdesc = make_sdesc_sw128(sb, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr >> 4) units
          umma_bf16(tmem_acc, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), p.idesc, (kt | k) != 0);
        }
        umma_commit(empty_bar(stage));  // frees this smem stage once the MMAs above have read it
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;        // row of the tile
    const int rn = r / (p.TW * p.TH);
    const int rh = (r / p.TW) % p.TH;
    const int rw = r % p.TW;
    const int n = n0 + rn, h = h0 + rh, w = w0 + rw;
    const bool row_ok = (n < p.n_img) && (h < p.H) && (w < p.W);
    const long long row = (static_cast<long long>(n) * p.H + h) * p.W + w;

    // stage this tile's bias slice in shared memory while the main loop runs; the epilogue reads it as broadcasts
    const float* sbias = reinterpret_cast<const float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));
    {
      float* sb = const_cast<float*>(sbias);
      for (int i = threadIdx.x - 64; i < p.bn; i += 128) {
        const int bc = nt * p.bn + i;
        sb[i] = (p.bias != nullptr && bc < p.bias_len) ? __ldg(p.bias + bc) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
    }

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_acc + (uint32_t(q * 32) << 16);

    if (p.mode == CB_EPI_GEGLU) {
      // tile columns [0, bn/2) hold x, [bn/2, bn) hold the gate (weights were interleaved per tile on the host)
      const int half = p.bn >> 1;
      for (int c = 0; c < half; c += 32) {
        uint32_t xv[32], gv[32];
        tmem_ld32(trow + uint32_t(c), xv);
        tmem_ld32(trow + uint32_t(half + c), gv);
        tmem_ld_wait();
        if (row_ok) {
          const int ocol0 = nt * half + c;  // output column
          act_t* optr = reinterpret_cast<act_t*>(p.out) + row * p.out_ld + ocol0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (ocol0 + j < p.cout) {  // cout here = number of OUTPUT columns (inner dim), multiple of 8
              uint32_t packed[4];
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const int tc0 = c + j + e;  // tile-relative column in the permuted weight/bias order
                float x0 = __uint_as_float(xv[j + e]) + sbias[tc0];
                float x1 = __uint_as_float(xv[j + e + 1]) + sbias[tc0 + 1];
                float g0 = __uint_as_float(gv[j + e]) + sbias[tc0 + half];
                float g1 = __uint_as_float(gv[j + e + 1]) + sbias[tc0 + half + 1];
                packed[e >> 1] = pack_act2(x0 * gelu_erf_f(g0), x1 * gelu_erf_f(g1));
              }
              *reinterpret_cast<uint4*>(optr + j) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
        }
      }
    } else {
      for (int c = 0; c < p.bn; c += 32) {
        uint32_t v[32];
        tmem_ld32(trow + uint32_t(c), v);
        tmem_ld_wait();
        const int col0 = nt * p.bn + c;
        if (!row_ok || col0 >= p.cout) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const int col = col0 + j;
          if (col >= p.cout) break;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
          const bool full8 = (col + 8 <= p.cout);
          if (full8) {
            {
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + c + j);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + c + j + 4);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (p.rowbias) {
              const float* rb = p.rowbias + static_cast<long long>(n) * p.rowbias_ld + col;
              if (p.rowbias_vec) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += __ldg(rb + e);
              }
            }
            if (p.act) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = apply_act(f[e], p.act);
            }
            if (p.residual) {
              const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + row * p.res_ld + col);
              const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 t = unpack_act2(ru[e]);
                f[2 * e] += t.x;
                f[2 * e + 1] += t.y;
              }
            }
            if (p.out_scale != 1.f) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] *= p.out_scale;
            }
            if (p.mode == CB_EPI_HEADS) {
              // column -> (which, head, j); row -> (batch, token); 8-column groups never straddle a head (d % 8 == 0)
              const int inner = p.hheads * p.hd;
              const int which = col / inner;
              const int cc = col - which * inner;
              const int head = cc / p.hd;
              const int jj = cc - head * p.hd;
              const long long b = row / p.htokens;
              const long long tok = row - b * p.htokens;
              act_t* optr = reinterpret_cast<act_t*>(p.out) + which * p.hwhich_stride +
                                    ((b * p.hheads + head) * p.htokens + tok) * p.hdpad + jj;
              *reinterpret_cast<uint4*>(optr) = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]),
                                                           pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
            } else if (p.out_f32) {
              float* optr = reinterpret_cast<float*>(p.out) + row * p.out_ld + col;
              *reinterpret_cast<float4*>(optr) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(optr + 4) = make_float4(f[4], f[5], f[6], f[7]);
            } else {
              act_t* optr = reinterpret_cast<act_t*>(p.out) + row * p.out_ld + col;
              *reinterpret_cast<uint4*>(optr) = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]),
                                                           pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
            }
          } else {
            // ragged tail (cout not a multiple of 8, e.g. the 4- and 3-channel output convs): scalar path
            for (int e = 0; e < 8 && col + e < p.cout; ++e) {
              float x = f[e];
              x += sbias[c + j + e];
              if (p.rowbias) x += __ldg(p.rowbias + static_cast<long long>(n) * p.rowbias_ld + col + e);
              x = apply_act(x, p.act);
              if (p.residual) x += from_act(p.residual[row * p.res_ld + col + e]);
              x *= p.out_scale;
              if (p.out_f32) reinterpret_cast<float*>(p.out)[row * p.out_ld + col + e] = x;
              else reinterpret_cast<act_t*>(p.out)[row * p.out_ld + col + e] = to_act(x);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, p.tmem_cols);
  }
}

static int pow2_cols(int bn) {
  int c = 32;
  while (c < bn) c <<= 1;
  return c;
}

}  // namespace cb

using namespace cb;

extern "C" int cb_igemm(const cb_igemm_desc* d, cudaStream_t stream) {
  CB_REQUIRE(d != nullptr, "cb_igemm: null descriptor");
  CB_REQUIRE(d->a0 && d->wgt && d->out, "cb_igemm: null a0/wgt/out pointer");
  CB_REQUIRE(d->c0 > 0 && d->c0 % 8 == 0 && d->c1 >= 0 && d->c1 % 8 == 0, "cb_igemm: channel counts must be multiples of 8 (c0=%lld c1=%lld)",
             (long long)d->c0, (long long)d->c1);
  CB_REQUIRE(d->taps >= 1 && d->taps <= 9, "cb_igemm: taps must be in [1,9]");
  CB_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->cout > 0, "cb_igemm: empty problem");
  CB_REQUIRE(d->bn >= 32 && d->bn <= 256 && d->bn % 32 == 0, "cb_igemm: bn must be a multiple of 32 in [32,256] (got %d)", d->bn);
  CB_REQUIRE(d->tw > 0 && d->th > 0 && d->tn > 0 && d->tw * d->th * d->tn == BM, "cb_igemm: tile %dx%dx%d is not 128 rows", d->tw, d->th, d->tn);
  CB_REQUIRE(d->tw <= 256 && d->th <= 256 && d->tn <= 256, "cb_igemm: tile extent > 256");
  if (d->mode == CB_EPI_GEGLU) CB_REQUIRE(d->bn % 64 == 0 && d->bias && !d->out_f32, "cb_igemm: GEGLU needs bn %% 64 == 0, a bias and bf16 output");
  if (d->mode == CB_EPI_HEADS) CB_REQUIRE(d->heads_d % 8 == 0 && d->heads_dpad >= d->heads_d && d->heads_h > 0 && d->heads_tokens > 0 && !d->out_f32, "cb_igemm: bad heads epilogue arguments");

  const int chunks0 = int((d->c0 + BK - 1) / BK);
  const int chunks1 = int((d->c1 + BK - 1) / BK);
  const int num_k = d->taps * (chunks0 + chunks1);
  const long long ktot = (long long)num_k * BK;

  const long long a0_ld = d->a0_ld > 0 ? d->a0_ld : d->c0;
  const long long a1_ld = d->a1_ld > 0 ? d->a1_ld : d->c1;
  CB_REQUIRE(a0_ld % 8 == 0 && a1_ld % 8 == 0, "cb_igemm: pixel strides must be multiples of 8 elements");

  CUtensorMap mapA0, mapA1, mapB;
  {
    uint64_t dims[4] = {(uint64_t)d->c0, (uint64_t)d->a_w, (uint64_t)d->a_h, (uint64_t)d->a_n};
    uint64_t str[4] = {1, (uint64_t)a0_ld, (uint64_t)a0_ld * d->a_w, (uint64_t)a0_ld * d->a_w * d->a_h};
    uint32_t box[4] = {BK, (uint32_t)d->tw, (uint32_t)d->th, (uint32_t)d->tn};
    int rc = make_tmap_act(&mapA0, d->a0, 4, dims, str, box);
    if (rc) return rc;
    if (chunks1 > 0) {
      CB_REQUIRE(d->a1, "cb_igemm: c1 > 0 but a1 is null");
      uint64_t dims1[4] = {(uint64_t)d->c1, (uint64_t)d->a_w, (uint64_t)d->a_h, (uint64_t)d->a_n};
      uint64_t str1[4] = {1, (uint64_t)a1_ld, (uint64_t)a1_ld * d->a_w, (uint64_t)a1_ld * d->a_w * d->a_h};
      rc = make_tmap_act(&mapA1, d->a1, 4, dims1, str1, box);
      if (rc) return rc;
    } else {
      mapA1 = mapA0;
    }
    uint64_t bdims[2] = {(uint64_t)ktot, (uint64_t)d->wgt_rows};
    uint64_t bstr[2] = {1, (uint64_t)ktot};
    uint32_t bbox[2] = {BK, (uint32_t)d->bn};
    CB_REQUIRE(d->wgt_rows > 0, "cb_igemm: wgt_rows must be > 0");
    rc = make_tmap_act(&mapB, d->wgt, 2, bdims, bstr, bbox);
    if (rc) return rc;
  }

  IGemmKParams p{};
  p.n_img = (int)d->n; p.H = (int)d->h; p.W = (int)d->w;
  p.TW = d->tw; p.TH = d->th; p.TN = d->tn;
  p.tiles_w = int((d->w + d->tw - 1) / d->tw);
  p.tiles_h = int((d->h + d->th - 1) / d->th);
  const int tiles_n = int((d->n + d->tn - 1) / d->tn);
  p.taps = d->taps; p.chunks0 = chunks0; p.chunks1 = chunks1; p.num_k = num_k;
  for (int i = 0; i < 9; ++i) { p.tap_dw[i] = d->tap_dw[i]; p.tap_dh[i] = d->tap_dh[i]; p.tap_dn[i] = d->tap_dn[i]; }
  p.cout = (int)d->cout; p.bn = d->bn;
  p.idesc = make_idesc_f16(BM, d->bn, 0, 0);
  p.tmem_cols = (uint32_t)pow2_cols(d->bn);
  p.mode = d->mode; p.act = d->act; p.out_f32 = d->out_f32;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.bias_len = int(d->mode == CB_EPI_GEGLU ? 2 * d->cout : d->cout);
  p.rowbias_vec = (d->rowbias != nullptr) && ((reinterpret_cast<uintptr_t>(d->rowbias) & 15u) == 0) && (d->rowbias_ld % 4 == 0);
  p.residual = reinterpret_cast<const act_t*>(d->residual); p.res_ld = d->res_ld;
  p.out = d->out; p.out_ld = d->out_ld;
  p.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  p.hd = d->heads_d; p.hdpad = d->heads_dpad; p.hheads = d->heads_h; p.htokens = d->heads_tokens;
  p.hwhich_stride = d->heads_which_stride;

  const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)d->bn * 128u;
  int stages = d->stages;
  if (stages <= 0) {
    // aim for two co-resident CTAs per SM (<= ~110 KiB each) with at least 2 and at most 6 stages
    stages = int((110u * 1024u) / stage_bytes);
    if (stages < 2) stages = 2;
    if (stages > 6) stages = 6;
  }
  if (stages > num_k && num_k >= 1) stages = num_k < 2 ? 2 : num_k;
  CB_REQUIRE(stages >= 2 && stages <= 12, "cb_igemm: stages out of range");
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 8 * (2 * stages + 1) + 32 + sizeof(float) * 256;
  CB_REQUIRE(smem <= 227 * 1024, "cb_igemm: tile needs %zu bytes of shared memory", smem);

  static thread_local size_t configured_smem = 0;
  if (smem > configured_smem) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    configured_smem = 227 * 1024;
  }
  const int n_tiles = int((d->mode == CB_EPI_GEGLU ? 2 * d->cout : d->cout) + d->bn - 1) / d->bn;
  dim3 grid((unsigned)n_tiles, (unsigned)(p.tiles_w * p.tiles_h * tiles_n), 1);
  igemm_kernel<<<grid, NUM_THREADS, smem, stream>>>(mapA0, mapA1, mapB, p);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}
