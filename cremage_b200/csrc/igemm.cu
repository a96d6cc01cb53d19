// Implicit-GEMM kernel for sm_100a: every contraction of the denoising path that is a "rows x K  *  K x Cout"
// product -- conv3x3 (stride 1, stride 2 via parity planes), conv1x1, nn.Linear -- runs through this one kernel.
//
//   D[row, co] = sum_{tap, c} A_src(c)[pixel(row) + shift(tap), c] * Wt[co, tap, c]      (fp32 accumulate in TMEM)
//
// * A is read straight from the NHWC 16-bit activation tensor(s) with 4-D TMA boxes {64 ch, TW, TH, TN}
//   (TW*TH*TN = 128 output pixels); a 3x3 tap is just a shifted box and the conv zero padding is TMA's
//   out-of-bounds zero fill.  No im2col buffer exists.  Two A sources give the UNet skip-concat for free
//   (reference: torch.cat([h, hs.pop()], 1) at ldm/modules/diffusionmodules/openaimodel.py:808).
// * B (weights, repacked [Cout][tap][Cin] 16-bit, K-major) is read with 2-D TMA boxes {64, BN}: either streamed
//   through the same smem ring as A, or -- when the whole K extent of one N tile fits (small-K linears / 1x1 convs) --
//   loaded ONCE per CTA and kept resident while the CTA walks the M tiles of that N tile (B-stationary), which
//   removes the weight re-reads that otherwise make K <= 640 shapes L2-bandwidth bound.
// * tcgen05.mma (cta_group::1, M=128, N=BN, K=16) issued by one thread; TWO fp32 accumulators in TMEM.
// * Persistent: one CTA per SM loops over tiles.  Warp roles: warp0 = TMA producer (runs ahead across tile
//   boundaries), warp1 = TMEM alloc + MMA issuer, warps 2..9 = epilogue (two warps per TMEM lane quarter, each
//   taking alternate 32-column chunks) draining accumulator i while the MMAs fill i+1.
// * Epilogue, two forms.  DIRECT: tcgen05.ld -> bias / per-image bias / SiLU / residual / per-head scatter ->
//   16-byte global stores from the thread that owns the row (large-K convs, where the epilogue hides under the
//   next tile's MMAs, and the ragged / fp32 / per-head outputs).  STAGED (small K, where the epilogue IS the
//   kernel): the residual panel of every chunk is prefetched by TMA into a 64B-swizzled shared-memory panel
//   while the tile's MMAs run, the thread adds it in place and the panel leaves by TMA store (bulk, asynchronous,
//   full-sector writes; rows / columns outside the tensor are clipped by the tensor map).
#include <type_traits>

#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 64;           // K elements per pipeline stage (= one 128-byte swizzle row)
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // 320
constexpr int PANEL_BYTES = 32 * 32 * 2;           // one staged epilogue panel: 32 rows x 32 columns, 16-bit
constexpr int SMEM_LIMIT = 227 * 1024;

// Division by a launch constant (tile scheduler, tile -> pixel-box coordinates): q = umulhi(n, mul) >> shr, exact for
// 0 <= n < 2^31 (mul = ceil(2^(31 + ceil_log2 d) / d)); d == 1 is the identity.  Replaces ~100-instruction integer
// divisions on the per-tile critical path of every role.
struct FastDiv {
  uint32_t mul, shr, d;
  __device__ __forceinline__ int div(int n) const { return d == 1u ? n : int(__umulhi(uint32_t(n), mul) >> shr); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * int(d); }
};
static FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, uint32_t(d)};
  if (d > 1) {
    int lg = 0;
    while ((1ll << lg) < d) ++lg;
    const int pw = 31 + lg;
    f.mul = uint32_t(((1ull << pw) + uint64_t(d) - 1) / uint64_t(d));
    f.shr = uint32_t(pw - 32);
  }
  return f;
}

struct IGemmKParams {
  // rows
  int n_img, H, W;       // output pixel grid
  int TW, TH, TN;        // tile decomposition, TW*TH*TN == 128
  int tiles_w, tiles_h;  // tiles along w / h
  FastDiv fd_tw, fd_th, fd_work;   // / tiles_w, / tiles_h, / (n_groups * ksplit)
  int total_work;        // (m_pairs or m_tiles) * n_groups * ksplit
  int m_tiles, n_tiles;  // tile grid
  int two, m_pairs;      // CTA-pair mode (cta_group::2): a cluster of 2 CTAs owns M tiles 2*mp, 2*mp+1 of one N tile
  int nsub, n_groups;    // N sub-tiles that share one A stage (1 or 2), groups of sub-tiles = ceil(n_tiles / nsub)
  int nacc;              // TMEM accumulator ring: 2 slots (nsub 1) or 3 slots (nsub 2: 3 x bn <= 512 columns)
  int ksplit, taps_per_split, num_k_split;   // split-K by tap groups: work item (tile, s) reduces taps [s*tps, (s+1)*tps)
  long long split_stride;                    // elements between the fp32 partial outputs of consecutive splits
  int resident_b;        // 1: B of this CTA's N tile stays in smem, CTA walks M tiles of that N tile
  int m_step;            // resident mode: stride between the M tiles of one CTA ( = gridDim.x / n_tiles )
  // K loop
  int taps, chunks0, chunks1, num_k;
  int tap_dw[9], tap_dh[9], tap_dn[9];
  // N
  int cout, bn, stages;
  uint32_t idesc, tmem_cols, acc_stride;
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
  // staged epilogue: panels per epilogue warp, and the pixel box {pbw, pbh, 32 / (pbw * pbh)} of a warp's 32 rows
  int npan, pbw, pbh;
  // fused GroupNorm statistics of the OUTPUT (16-bit rounded values): per (image, M tile) and channel pair, the sum
  // and the sum of squares -> gn_part[n][gn_bpi][2][cout/2], gn_bpi = tiles_w * tiles_h (M tiles per image)
  float* gn_part;
  int gn_bpi, gn_row0;   // rows per image of the table, first row of this launch
  // fused LayerNorm (staged epilogues).  PRODUCER: per output row and (N tile, epilogue half) the sum and the sum of
  // squares of the 16-bit rounded outputs -> ln_out[row][ln_out_slots][2].  CONSUMER: the A operand is the UN-normalised
  // row x; the weights are W'' = W diag(gamma) with every ROW CENTRED (sum_k W''[n, k] = 0), so x W''^T = (x - mean) W'^T,
  // b' = b + W beta, and the epilogue computes  LN(x) W^T + b = rstd * (x W''^T) + b'   (rstd from ln_in[row][slots][2])
  float* ln_out;
  int ln_out_slots;
  const float* ln_in;
  int ln_in_slots;
  float ln_inv_dim, ln_eps;
};

// ---- fused GroupNorm statistics -----------------------------------------------------------------------------------
// V[0..15] = this row's sums of 16 channel pairs, V[16..31] = their sums of squares.  Transposed butterfly: after the
// five exchange steps lane L holds the total over the warp's 32 rows of entry L (fixed order -> deterministic).
template <int M>
CB_DEVINL void gn_xchg_step(float (&V)[32], int lane) {
  constexpr int O = M / 2;
  const bool up = (lane & O) != 0;
#pragma unroll
  for (int i = 0; i < O; ++i) {
    const float send = up ? V[i] : V[i + O];
    const float keep = up ? V[i + O] : V[i];
    V[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
  }
}
// the warp's 32-row totals of one 32-column chunk go to its slot of the tile's shared-memory table [chunk][quarter][32]
CB_DEVINL void gn_store_chunk(float (&V)[32], int lane, float* __restrict__ slot) {
  gn_xchg_step<32>(V, lane);
  gn_xchg_step<16>(V, lane);
  gn_xchg_step<8>(V, lane);
  gn_xchg_step<4>(V, lane);
  gn_xchg_step<2>(V, lane);
  slot[lane] = V[0];
}
// One tile later (after the epilogue warps' named barrier): fold the four lane quarters of every chunk per image, in
// quarter order, and write the tile's row of the partial table.  tab: [8 chunks][4 quarters][32]; ew = 0..7.
CB_DEVINL void gn_flush_tile(const IGemmKParams& p, const float* __restrict__ tab, int ew, int lane, int tile_row, int img0, int nt) {
  const int qpi = 4 / p.TN;   // lane quarters per image (tw * th % 32 == 0: TN is 1, 2 or 4)
  for (int ci = ew; ci * 32 < p.bn; ci += EPI_WARPS) {
    const int col0 = nt * p.bn + ci * 32;
    if (col0 >= p.cout) break;
    const int pair = (col0 >> 1) + (lane & 15);
    if (2 * pair >= p.cout) continue;
    const float* s = tab + ci * 128 + lane;
    for (int im = 0; im < p.TN; ++im) {
      const int img = img0 + im;
      if (img >= p.n_img) break;
      float t = s[(im * qpi) * 32];
      for (int q = 1; q < qpi; ++q) t += s[(im * qpi + q) * 32];
      p.gn_part[((static_cast<long long>(img) * p.gn_bpi + p.gn_row0 + tile_row) * 2 + (lane >> 4)) * (p.cout >> 1) + pair] = t;
    }
  }
}
// accumulate the 8 final values of columns [jb, jb+8) of a chunk as they will be read back (16-bit rounded)
CB_DEVINL void gn_accum8(float (&V)[32], int jb, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  const uint32_t wv[4] = {w0, w1, w2, w3};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = unpack_act2(wv[e]);
    V[(jb >> 1) + e] = t.x + t.y;
    V[16 + (jb >> 1) + e] = fmaf(t.x, t.x, t.y * t.y);
  }
}

__device__ __forceinline__ float apply_act(float v, int act) { return act == CB_ACT_SILU ? silu_f(v) : v; }

// tile scheduler shared by the three roles: local tile i of this CTA -> (mt, nt), false when exhausted
struct TileSched {
  const IGemmKParams& p;
  __device__ explicit TileSched(const IGemmKParams& pp) : p(pp) {}
  __device__ __forceinline__ bool get(int i, int& mt, int& nt) const {
    // `nt` is the N GROUP index: the group's sub-tiles are n-tiles nt*nsub .. nt*nsub + nsub-1 (see subs());
    // with split-K the K-split index rides in the upper bits: nt = s * n_groups + group (see split() / group())
    if (p.two) {
      const int t = int(blockIdx.x >> 1) + i * int(gridDim.x >> 1);
      if (t >= p.total_work) return false;
      int q;
      p.fd_work.divmod(t, q, nt);
      mt = 2 * q + int(blockIdx.x & 1);   // may be == m_tiles (odd count): an all-out-of-bounds tile
      return true;
    }
    if (p.resident_b) {
      nt = int(blockIdx.x) % p.n_tiles;
      mt = int(blockIdx.x) / p.n_tiles + i * p.m_step;
      return mt < p.m_tiles;
    }
    const int t = int(blockIdx.x) + i * int(gridDim.x);
    if (t >= p.total_work) return false;
    p.fd_work.divmod(t, mt, nt);
    return true;
  }
  __device__ __forceinline__ int split(int nt) const { return p.ksplit == 1 ? 0 : nt / p.n_groups; }
  __device__ __forceinline__ int group(int nt) const { return p.ksplit == 1 ? nt : nt % p.n_groups; }
  template <int NSUB>
  __device__ __forceinline__ int subs(int ng) const {   // sub-tiles of group ng that exist
    if (NSUB == 1) return 1;
    const int left = p.n_tiles - ng * NSUB;
    return left < NSUB ? left : NSUB;
  }
};

// Epilogue specialisations (compile-time, so the inner loops carry no mode / flag branches):
enum : int {
  EPI_GENERIC = 0,   // every runtime flag honoured (fp32 out, ragged cout, activation, scale) -- small / rare launches
  EPI_PLAIN = 1,     // + bias                          -> 16-bit
  EPI_RES = 2,       // + bias + residual               -> 16-bit
  EPI_ROWBIAS = 3,   // + bias + per-image bias         -> 16-bit   (ResBlock conv1 with the timestep projection)
  EPI_GEGLU = 4,     // x * gelu(gate)                  -> 16-bit
  EPI_HEADS = 5,     // + bias, per-head padded scatter -> 16-bit
  EPI_COUNT = 6
};

// 32 accumulator columns [c, c+32) of one row -> global; sbias = the 32 bias values of this chunk
template <int EPI, bool STATS>
__device__ __forceinline__ void epilogue_chunk(const IGemmKParams& p, const uint32_t (&v)[32], const float* __restrict__ sbias,
                                               int nt, int c, int n, long long row, float (&V)[32]) {
  const int col0 = nt * p.bn + c;
  if (col0 >= p.cout) return;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int col = col0 + j;
    if (col >= p.cout) break;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
    if (EPI != EPI_GENERIC || col + 8 <= p.cout) {
      {
        const float4 b0 = *reinterpret_cast<const float4*>(sbias + j);
        const float4 b1 = *reinterpret_cast<const float4*>(sbias + j + 4);
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      if (EPI == EPI_ROWBIAS || (EPI == EPI_GENERIC && p.rowbias)) {
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
      if (EPI == EPI_GENERIC && p.act) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = apply_act(f[e], p.act);
      }
      if (EPI == EPI_RES || (EPI == EPI_GENERIC && p.residual)) {
        const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + row * p.res_ld + col);
        const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 t = unpack_act2(ru[e]);
          f[2 * e] += t.x;
          f[2 * e + 1] += t.y;
        }
      }
      if (EPI == EPI_GENERIC && p.out_scale != 1.f) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] *= p.out_scale;
      }
      if (EPI == EPI_HEADS) {
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
      } else if (EPI == EPI_GENERIC && p.out_f32) {
        float* optr = reinterpret_cast<float*>(p.out) + row * p.out_ld + col;
        *reinterpret_cast<float4*>(optr) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(optr + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        act_t* optr = reinterpret_cast<act_t*>(p.out) + row * p.out_ld + col;
        const uint4 o = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]), pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
        *reinterpret_cast<uint4*>(optr) = o;
        if (STATS) gn_accum8(V, j, o.x, o.y, o.z, o.w);
      }
    } else {
      // EPI_GENERIC ragged tail (cout not a multiple of 8, e.g. the 4- and 3-channel output convs): scalar path
      for (int e = 0; e < 8 && col + e < p.cout; ++e) {
        float x = f[e];
        x += sbias[j + e];
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

// DIRECT epilogue of this warp's share of one 128 x bn accumulator tile (thread = one row): chunks hh, hh+2, ...
template <int EPI, bool STATS>
__device__ __forceinline__ void epilogue_tile_direct(const IGemmKParams& p, uint32_t trow, const float* __restrict__ sbias,
                                                     int nt, int n, bool row_ok, long long row, int hh, int lane,
                                                     float* __restrict__ gtab) {
  int k = 0;
  for (int c = hh * 32; c < p.bn; c += 64, ++k) {
    if (nt * p.bn + c >= p.cout) break;
    uint32_t v[32];
    tmem_ld32(trow + uint32_t(c), v);
    tmem_ld_wait();
    float V[32];
    if (STATS) {
#pragma unroll
      for (int i = 0; i < 32; ++i) V[i] = 0.f;
    }
    if (row_ok) epilogue_chunk<EPI, STATS>(p, v, sbias + k * 32, nt, c, n, row, V);
    if (STATS) gn_store_chunk(V, lane, gtab + (c >> 5) * 128);
  }
}

CB_DEVINL void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  // no "memory" clobber: the panels are touched only through these volatile asms (ordered among themselves and
  // against the proxy fence), so ordinary loads (the bias slices) may be scheduled across the stores
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
CB_DEVINL uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// STAGED epilogue of one 32-column chunk: registers (+bias, +per-image bias, +residual read from the panel) -> the
// 64B-swizzled panel (row = lane, 16-byte piece j at  j ^ ((lane >> 1) & 3)), ready for the TMA store.
// LayerNorm fold of a staged launch: lnc = the A rows are un-normalised and the weights are gamma-folded AND centred
// (every weight row sums to zero, so the row mean drops out of the product) -- the epilogue only scales by this row's
// rstd; lnp = accumulate this row's sum / sum of squares of the 16-bit outputs into ls / lq.
struct LnRow { bool lnc, lnp; float rs; };
template <int EPI, bool STATS>
__device__ __forceinline__ void staged_chunk(const IGemmKParams& p, const uint32_t (&v)[32], const uint32_t (&g)[32],
                                             const float* __restrict__ sbx, const float* __restrict__ sbg,
                                             const float* __restrict__ rowb, bool rowb_ok, int col0, uint32_t panel, int lane,
                                             bool row_ok, float (&V)[32], const LnRow& ln, float& ls, float& lq) {
  const uint32_t rowaddr = panel + uint32_t(lane) * 64u;
  const uint32_t sw = uint32_t(lane >> 1) & 3u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(sbx + 8 * j);
    const float4 b1 = *reinterpret_cast<const float4*>(sbx + 8 * j + 4);
    if ((EPI == EPI_PLAIN || EPI == EPI_GEGLU) && ln.lnc) {
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fmaf(ln.rs, __uint_as_float(v[8 * j + e]), bb[e]);   // one FMA where the plain form has one add
    } else {
      f[0] = __uint_as_float(v[8 * j + 0]) + b0.x; f[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
      f[2] = __uint_as_float(v[8 * j + 2]) + b0.z; f[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
      f[4] = __uint_as_float(v[8 * j + 4]) + b1.x; f[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
      f[6] = __uint_as_float(v[8 * j + 6]) + b1.z; f[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
    }
    const uint32_t addr = rowaddr + ((uint32_t(j) ^ sw) << 4);
    if (EPI == EPI_GEGLU) {
      const float4 g0 = *reinterpret_cast<const float4*>(sbg + 8 * j);
      const float4 g1 = *reinterpret_cast<const float4*>(sbg + 8 * j + 4);
      const float gb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float gv[8];
      if (ln.lnc) {
#pragma unroll
        for (int e = 0; e < 8; ++e) gv[e] = fmaf(ln.rs, __uint_as_float(g[8 * j + e]), gb[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) gv[e] = __uint_as_float(g[8 * j + e]) + gb[e];
      }
      geglu_mul8(f, gv);
    }
    if (EPI == EPI_ROWBIAS) {
      if (rowb_ok && col0 + 8 * j + 8 <= p.cout) {
        const float* rb = rowb + col0 + 8 * j;
        if (p.rowbias_vec) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(rb));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
          f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
          f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += __ldg(rb + e);
        }
      }
    }
    if (EPI == EPI_RES) {
      const uint4 rv = ld_shared_v4(addr);
      const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_act2(ru[e]);
        f[2 * e] += t.x;
        f[2 * e + 1] += t.y;
      }
    }
    const uint32_t o0 = pack_act2(f[0], f[1]), o1 = pack_act2(f[2], f[3]), o2 = pack_act2(f[4], f[5]), o3 = pack_act2(f[6], f[7]);
    st_shared_v4(addr, o0, o1, o2, o3);
    if ((EPI == EPI_PLAIN || EPI == EPI_RES) && ln.lnp && row_ok && col0 + 8 * j < p.cout) {
      const uint32_t ov[4] = {o0, o1, o2, o3};      // the values as the consumer will read them back
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_act2(ov[e]);
        ls += t.x + t.y;
        lq = fmaf(t.x, t.x, fmaf(t.y, t.y, lq));
      }
    }
    if (STATS) {
      if (row_ok && col0 + 8 * j < p.cout) gn_accum8(V, 8 * j, o0, o1, o2, o3);
      else gn_accum8(V, 8 * j, 0u, 0u, 0u, 0u);
    }
  }
}

// DUAL (pair mode only): two N sub-tiles share every A stage, three accumulator slots in TMEM
template <int EPI, bool STAGED, bool TWO, bool DUAL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
             const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO,
             const __grid_constant__ CUtensorMap mapR, const IGemmKParams p) {
  pdl_launch_dependents();
  static_assert(!DUAL || TWO, "sub-tile groups exist in pair mode only");
  constexpr int NSUB = DUAL ? 2 : 1;
  constexpr int NACC = DUAL ? 3 : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve (1024-aligned): [resident B: num_k x bn*128] [ring: stages x (A 16K [+ B bn*128])]
  //                       [staging: EPI_WARPS x npan x 2K] [barriers] [bias x2]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_chunk_bytes = uint32_t(TWO ? (p.bn >> 1) : p.bn) * 128u;   // a pair CTA stages half of the B tile
  const uint32_t res_bytes = p.resident_b ? uint32_t(p.num_k) * b_chunk_bytes : 0u;
  const uint32_t pair_rank = TWO ? (blockIdx.x & 1u) : 0u;
  const uint32_t stage_bytes = A_STAGE_BYTES + (p.resident_b ? 0u : uint32_t(NSUB) * b_chunk_bytes);
  const uint32_t ring_base = smem_base + res_bytes;
  const uint32_t stg_base = ring_base + uint32_t(p.stages) * stage_bytes;
  const uint32_t bar_base = stg_base + (STAGED ? uint32_t(EPI_WARPS * p.npan) * PANEL_BYTES : 0u);
  auto full_bar = [&](int s) { return bar_base + 8u * uint32_t(s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * uint32_t(p.stages + s); };
  const uint32_t misc = bar_base + 16u * uint32_t(p.stages);
  auto acc_full = [&](int b) { return misc + 8u * uint32_t(b); };            // [3]
  auto acc_empty = [&](int b) { return misc + 24u + 8u * uint32_t(b); };      // [3]
  const uint32_t bres_bar = misc + 48u;
  const uint32_t tmem_slot = misc + 56u;
  auto resid_bar = [&](int ew) { return misc + 64u + 8u * uint32_t(ew); };
  const uint32_t bias_smem = (misc + 64u + 8u * EPI_WARPS + 15u) & ~15u;  // float[EPI_WARPS][128]: every epilogue warp stages the bias of its own chunks
  const uint32_t gn_smem = bias_smem + uint32_t(EPI_WARPS) * 128u * 4u;   // float[2 tiles][8 chunks][4 quarters][32] (only when gn_part)
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const TileSched sched(p);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.chunks1 > 0) tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    if (STAGED) {
      tma_prefetch_desc(&mapO);
      if (EPI == EPI_RES) tma_prefetch_desc(&mapR);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), EPI_WARPS * (TWO ? 2 : 1));   // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_init(bres_bar, 1);
    for (int e = 0; e < EPI_WARPS; ++e) mbar_init(resid_bar(e), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (TWO) { tmem_alloc_pair(tmem_slot, p.tmem_cols); tmem_relinquish_pair(); }
    else     { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (TWO) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // everything above ran under the previous kernel's tail; from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // (not `lane == 0`: under a divergent branch every UTMALDG / UTCHMMA is wrapped in an ELECT / BRA.U.ANY loop)
      if (p.resident_b) {
        int mt, nt;
        if (sched.get(0, mt, nt)) {
          mbar_expect_tx(bres_bar, res_bytes);
          for (int kt = 0; kt < p.num_k; ++kt)
            tma_load_2d(smem_base + uint32_t(kt) * b_chunk_bytes, &mapB, bres_bar, kt * BK, nt * p.bn);
        }
      }
      const int cpt = p.chunks0 + p.chunks1;
      int stage = 0;
      uint32_t phase = 0;
      int mt, nt;
      for (int i = 0; sched.get(i, mt, nt); ++i) {
        int tw, th, tn, mrow;
        p.fd_tw.divmod(mt, mrow, tw);
        p.fd_th.divmod(mrow, tn, th);
        const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
        const int ng = sched.group(nt), ks = sched.split(nt);
        const int subs = sched.subs<NSUB>(ng);
        const uint32_t tx_bytes = A_STAGE_BYTES + (p.resident_b ? 0u : uint32_t(subs) * b_chunk_bytes);
        const int brow0 = ng * NSUB * p.bn;   // first weight row of the group
        int tap = ks * p.taps_per_split, ch = 0;
        for (int kt = ks * p.num_k_split, kt_end = kt + p.num_k_split; kt < kt_end; ++kt) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = ring_base + uint32_t(stage) * stage_bytes;
          const int cw = w0 + p.tap_dw[tap], chh = h0 + p.tap_dh[tap], cn = n0 + p.tap_dn[tap];
          if (TWO) {
            // both CTAs load (own A rows, own half of B); every byte is counted on the LEADER's full barrier
            if (pair_rank == 0) mbar_expect_tx(full_bar(stage), 2u * tx_bytes);
            if (ch < p.chunks0) tma_load_4d_pair(sa, &mapA0, full_bar(stage), ch * BK, cw, chh, cn);
            else                tma_load_4d_pair(sa, &mapA1, full_bar(stage), (ch - p.chunks0) * BK, cw, chh, cn);
            for (int h = 0; h < subs; ++h)
              tma_load_2d_pair(sa + A_STAGE_BYTES + uint32_t(h) * b_chunk_bytes, &mapB, full_bar(stage), kt * BK,
                               brow0 + h * p.bn + int(pair_rank) * (p.bn >> 1));
          } else {
            mbar_expect_tx(full_bar(stage), tx_bytes);
            if (ch < p.chunks0) tma_load_4d(sa, &mapA0, full_bar(stage), ch * BK, cw, chh, cn);
            else                tma_load_4d(sa, &mapA1, full_bar(stage), (ch - p.chunks0) * BK, cw, chh, cn);
            if (!p.resident_b)
              for (int h = 0; h < subs; ++h)
                tma_load_2d(sa + A_STAGE_BYTES + uint32_t(h) * b_chunk_bytes, &mapB, full_bar(stage), kt * BK, brow0 + h * p.bn);
          }
          if (++ch == cpt) { ch = 0; ++tap; }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair mode: the leader CTA's thread drives both tensor cores) =====================
    if (pair_rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int mt, nt;
      bool b_ready = !p.resident_b;
      int acc_cnt = 0;   // accumulator ring position (sub-tiles issued so far)
      for (int i = 0; sched.get(i, mt, nt); ++i) {
        const bool dual = DUAL && sched.subs<NSUB>(sched.group(nt)) > 1;
        // (scalars, not arrays: a dynamically indexed array would live in local memory on the issue path)
        const int slot0 = acc_cnt % NACC, slot1 = (acc_cnt + 1) % NACC;
        mbar_wait(acc_empty(slot0), (uint32_t(acc_cnt / NACC) & 1u) ^ 1u);   // epilogue drained the slot (first round passes)
        if (DUAL && dual) mbar_wait(acc_empty(slot1), (uint32_t((acc_cnt + 1) / NACC) & 1u) ^ 1u);
        const uint32_t tacc0 = tmem_base + uint32_t(slot0) * p.acc_stride;
        const uint32_t tacc1 = tmem_base + uint32_t(slot1) * p.acc_stride;
        tc_fence_after();
        if (!b_ready) { mbar_wait(bres_bar, 0); b_ready = true; }
        // K loop; the sub-tile count is hoisted out of the issue loop (a branch per MMA costs ~20 % on this thread)
        auto k_loop = [&](auto both_tag) {
          constexpr bool BOTH = decltype(both_tag)::value;
          for (int kt = 0; kt < p.num_k_split; ++kt) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = ring_base + uint32_t(stage) * stage_bytes;
            const uint32_t sb = p.resident_b ? (smem_base + uint32_t(kt) * b_chunk_bytes) : (sa + A_STAGE_BYTES);
            const uint64_t adesc = make_sdesc_sw128(sa, 16, 1024);
            const uint64_t bdesc0 = make_sdesc_sw128(sb, 16, 1024);
            const uint64_t bdesc1 = make_sdesc_sw128(sb + b_chunk_bytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr >> 4) units
              const uint32_t acc = (kt | k) != 0;
              if (TWO) {
                umma_bf16_pair(tacc0, adesc + uint64_t(2 * k), bdesc0 + uint64_t(2 * k), p.idesc, acc);
                if (BOTH) umma_bf16_pair(tacc1, adesc + uint64_t(2 * k), bdesc1 + uint64_t(2 * k), p.idesc, acc);   // reuses the A stage
              } else {
                umma_bf16(tacc0, adesc + uint64_t(2 * k), bdesc0 + uint64_t(2 * k), p.idesc, acc);
              }
            }
            // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
            if (TWO) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        };
        if (DUAL && dual) k_loop(std::true_type{}); else k_loop(std::false_type{});
        // accumulators complete
        if (TWO) umma_commit_pair(acc_full(slot0)); else umma_commit(acc_full(slot0));
        if (DUAL && dual) umma_commit_pair(acc_full(slot1));   // (cta_group::2 instructions only in the pair instantiations)
        acc_cnt += (DUAL && dual) ? 2 : 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int hh = ew >> 2;             // which alternate 32-column chunks this warp takes
    const int r = q * 32 + lane;        // row of the tile
    const int rn = r / (p.TW * p.TH);
    const int rh = (r / p.TW) % p.TH;
    const int rw = r % p.TW;
    // origin of this warp's 32-row pixel box inside the tile (staged form)
    const int r0 = q * 32;
    const int bw0 = r0 % p.TW, bh0 = (r0 / p.TW) % p.TH, bn0 = r0 / (p.TW * p.TH);
    const uint32_t my_stg = stg_base + uint32_t(ew * p.npan) * PANEL_BYTES;
    uint32_t rphase = 0;
    float* sbias_all = reinterpret_cast<float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));
    constexpr bool GEGLU = (EPI == EPI_GEGLU);
    constexpr bool STATS_OK = (EPI == EPI_PLAIN || EPI == EPI_RES || EPI == EPI_ROWBIAS);
    const bool do_gn = STATS_OK && p.gn_part != nullptr;
    float* gn_tab = reinterpret_cast<float*>(smem_raw + (gn_smem - smem_u32(smem_raw)));
    int g_prev_row = -1, g_prev_img0 = 0, g_prev_nt = 0;
    const int ocols = GEGLU ? (p.bn >> 1) : p.bn;   // output columns per tile
    int mt, ng;
    int acc_cnt = 0;
    int ngs;
    for (int i = 0; sched.get(i, mt, ngs); ++i)
    for (int sub = 0, subs = sched.subs<NSUB>(ng = sched.group(ngs)); sub < subs; ++sub, ++acc_cnt) {
      const int nt = ng * NSUB + sub;
      const int ksi = sched.split(ngs);
      const int buf = acc_cnt % NACC;
      const uint32_t use = uint32_t(acc_cnt / NACC);
      // bias of this warp's chunks, requested first: the load latency overlaps the rest of the tile set-up
      float bx[4], bg[2] = {0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = hh * 32 + 64 * k;
        const int bc = nt * p.bn + c + lane;
        bx[k] = (c < ocols && p.bias != nullptr && bc < p.bias_len) ? __ldg(p.bias + bc) : 0.f;
        if (GEGLU && k < 2) bg[k] = (c < ocols && bc + ocols < p.bias_len) ? __ldg(p.bias + bc + ocols) : 0.f;
      }
      int tw, th, tn, mrow;
      p.fd_tw.divmod(mt, mrow, tw);
      p.fd_th.divmod(mrow, tn, th);
      const int n = tn * p.TN + rn, h = th * p.TH + rh, w = tw * p.TW + rw;
      const bool row_ok = (n < p.n_img) && (h < p.H) && (w < p.W);
      const long long row = (static_cast<long long>(n) * p.H + h) * p.W + w;
      const int bxw = tw * p.TW + bw0, bxh = th * p.TH + bh0, bxn = tn * p.TN + bn0;
      if (STAGED) {
        // the previous tile's TMA stores have finished reading this warp's panels (without a residual to load into
        // them the wait moves down to the first write of a panel, behind the accumulator wait and the TMEM load)
        if (EPI == EPI_RES) {
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
        }
        if (EPI == EPI_RES && elect_one()) {
          int cnt = 0;
          for (int c = hh * 32; c < ocols && nt * ocols + c < p.cout; c += 64) ++cnt;
          if (cnt > 0) {
            mbar_expect_tx(resid_bar(ew), uint32_t(cnt) * PANEL_BYTES);
            int k = 0;
            for (int c = hh * 32; c < ocols && nt * ocols + c < p.cout; c += 64, ++k)
              tma_load_4d(my_stg + uint32_t(k) * PANEL_BYTES, &mapR, resid_bar(ew), nt * ocols + c, bxw, bxh, bxn);
          }
        }
      }
      // this warp's bias values: chunk k of the warp at [32 k], the GEGLU gate half at [64 + 32 k].  Private to the
      // warp, so no barrier between the epilogue warps: they drift apart and hide each other's TMEM-load, fence and
      // store latencies instead of meeting them in lockstep.  (Loaded at the top of the tile, see bx / bg.)
      float* sb = sbias_all + ew * 128;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) sb[k * 32 + lane] = bx[k];
      if (GEGLU) { sb[64 + lane] = bg[0]; sb[96 + lane] = bg[1]; }
      __syncwarp();
      // fused GroupNorm statistics: every warp has stored the previous tile's chunk totals -> write that tile's row
      float* gtab = gn_tab + (acc_cnt & 1) * 1024 + q * 32;
      if (STATS_OK && do_gn) {
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
        if (g_prev_row >= 0) gn_flush_tile(p, gn_tab + ((acc_cnt - 1) & 1) * 1024, ew, lane, g_prev_row, g_prev_img0, g_prev_nt);
        g_prev_row = th * p.tiles_w + tw; g_prev_img0 = tn * p.TN; g_prev_nt = nt;
      }

      mbar_wait(acc_full(buf), use & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + uint32_t(buf) * p.acc_stride + (uint32_t(q * 32) << 16);
      if (STAGED) {
        const bool have = (hh * 32 < ocols) && (nt * ocols + hh * 32 < p.cout);
        if (EPI == EPI_RES && have) { mbar_wait(resid_bar(ew), rphase); rphase ^= 1u; }
        const float* rowb = (EPI == EPI_ROWBIAS) ? (p.rowbias + static_cast<long long>(n < p.n_img ? n : 0) * p.rowbias_ld) : nullptr;
        LnRow ln{p.ln_in != nullptr, p.ln_out != nullptr, 1.f};
        float ls = 0.f, lq = 0.f;
        if ((EPI == EPI_PLAIN || EPI == EPI_GEGLU) && ln.lnc && row_ok) {
          // this row's mean / rstd from the producer's partial sums (fixed slot order: deterministic)
          const float2* ps = reinterpret_cast<const float2*>(p.ln_in) + row * p.ln_in_slots;
          float sx = 0.f, sq = 0.f;
          for (int sl = 0; sl < p.ln_in_slots; ++sl) { const float2 t = __ldg(ps + sl); sx += t.x; sq += t.y; }
          const float mean = sx * p.ln_inv_dim;
          const float var = fmaxf(fmaf(-mean, mean, sq * p.ln_inv_dim), 0.f);
          ln.rs = rsqrtf(var + p.ln_eps);
        }
        int k = 0;
        for (int c = hh * 32; c < ocols; c += 64, ++k) {
          const int col0 = nt * ocols + c;
          if (col0 >= p.cout) break;
          uint32_t v[32], g[32];
          tmem_ld32(trow + uint32_t(c), v);
          if (GEGLU) tmem_ld32(trow + uint32_t(ocols + c), g);
          tmem_ld_wait();
          if (EPI != EPI_RES && k == 0) {
            if (elect_one()) tma_store_wait_read<0>();
            __syncwarp();
          }
          const uint32_t panel = my_stg + uint32_t(k) * PANEL_BYTES;
          float V[32];
          if (STATS_OK && do_gn) {
            staged_chunk<EPI, STATS_OK>(p, v, GEGLU ? g : v, sb + k * 32, sb + 64 + k * 32, rowb, n < p.n_img, col0, panel, lane, row_ok, V, ln, ls, lq);
            gn_store_chunk(V, lane, gtab + (c >> 5) * 128);
          } else {
            staged_chunk<EPI, false>(p, v, GEGLU ? g : v, sb + k * 32, sb + 64 + k * 32, rowb, n < p.n_img, col0, panel, lane, row_ok, V, ln, ls, lq);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {   // elect.sync names the same lane for the same mask every time: it owns the bulk groups
            tma_store_4d(&mapO, panel, col0, bxw, bxh, bxn);
            tma_store_commit();
          }
        }
        if ((EPI == EPI_PLAIN || EPI == EPI_RES) && ln.lnp && row_ok) {
          // every (N tile, epilogue half) slot of the row is written, also by a half without a chunk (zeros)
          reinterpret_cast<float2*>(p.ln_out)[row * p.ln_out_slots + nt * 2 + hh] = make_float2(ls, lq);
        }
      } else {
        // split-K: this work item's fp32 partial goes to its own slab of the workspace
        const long long orow = row + (long long)ksi * (p.split_stride / p.out_ld);
        if (STATS_OK && do_gn) epilogue_tile_direct<EPI, STATS_OK>(p, trow, sb, nt, n, row_ok, orow, hh, lane, gtab);
        else epilogue_tile_direct<EPI, false>(p, trow, sb, nt, n, row_ok, orow, hh, lane, gtab);
      }
      tc_fence_before();
      __syncwarp();                  // every lane of this warp has drained its TMEM lanes
      if (lane == 0) {
        if (TWO) mbar_arrive_leader(acc_empty(buf)); else mbar_arrive(acc_empty(buf));
      }
    }
    if (STATS_OK && do_gn && g_prev_row >= 0) {   // the last tile's row
      asm volatile("bar.sync 1, 256;" ::: "memory");
      gn_flush_tile(p, gn_tab + ((acc_cnt - 1) & 1) * 1024, ew, lane, g_prev_row, g_prev_img0, g_prev_nt);
    }
    if (STAGED && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  if (TWO) cluster_sync_all();   // no CTA of a pair leaves while its peer may still read its smem / signal its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (TWO) tmem_dealloc_pair(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
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
  if (d->mode == CB_EPI_GEGLU) CB_REQUIRE(d->bn % 64 == 0 && d->bias && !d->out_f32, "cb_igemm: GEGLU needs bn %% 64 == 0, a bias and 16-bit output");
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
    uint32_t bbox[2] = {BK, (uint32_t)(d->cta_pair ? d->bn / 2 : d->bn)};   // a pair CTA loads half of the N tile
    CB_REQUIRE(d->wgt_rows > 0, "cb_igemm: wgt_rows must be > 0");
    rc = make_tmap_act(&mapB, d->wgt, 2, bdims, bstr, bbox);
    if (rc) return rc;
  }

  const int g_num_sms = sm_count();

  IGemmKParams p{};
  p.n_img = (int)d->n; p.H = (int)d->h; p.W = (int)d->w;
  p.TW = d->tw; p.TH = d->th; p.TN = d->tn;
  p.tiles_w = int((d->w + d->tw - 1) / d->tw);
  p.tiles_h = int((d->h + d->th - 1) / d->th);
  const int tiles_n = int((d->n + d->tn - 1) / d->tn);
  p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  p.n_tiles = int((d->mode == CB_EPI_GEGLU ? 2 * d->cout : d->cout) + d->bn - 1) / d->bn;
  p.taps = d->taps; p.chunks0 = chunks0; p.chunks1 = chunks1; p.num_k = num_k;
  // split-K by tap groups (the M <= 1024 convs: too few tiles for 148 SMs, long K): every work item writes an fp32
  // partial; cb_splitk_reduce folds the partials in split order and applies the epilogue (deterministic)
  const int ksplit = d->ksplit > 1 ? d->ksplit : 1;
  if (ksplit > 1) {
    CB_REQUIRE(d->taps % ksplit == 0, "cb_igemm: ksplit %d does not divide %d taps", ksplit, d->taps);
    CB_REQUIRE(d->mode == CB_EPI_LINEAR && d->out_f32 && !d->bias && !d->rowbias && !d->residual && d->act == CB_ACT_NONE &&
               (d->out_scale == 0.f || d->out_scale == 1.f),
               "cb_igemm: split-K launches write raw fp32 partials (no bias / activation / residual; cb_splitk_reduce applies them)");
  }
  p.ksplit = ksplit;
  p.taps_per_split = d->taps / ksplit;
  p.num_k_split = p.taps_per_split * (chunks0 + chunks1);
  p.split_stride = d->n * d->h * d->w * d->out_ld;
  for (int i = 0; i < 9; ++i) { p.tap_dw[i] = d->tap_dw[i]; p.tap_dh[i] = d->tap_dh[i]; p.tap_dn[i] = d->tap_dn[i]; }
  p.cout = (int)d->cout; p.bn = d->bn;
  const bool two = d->cta_pair != 0;
  if (two) CB_REQUIRE(d->stages <= 0 || d->stages >= 2, "cb_igemm: bad stage count");
  p.two = two ? 1 : 0;
  p.m_pairs = (p.m_tiles + 1) / 2;
  p.idesc = make_idesc_f16(two ? 2 * BM : BM, d->bn, 0, 0);
  // N sub-tile groups: in pair mode two N tiles share every A stage when three accumulators fit TMEM (bn <= 160);
  // the shared-memory fill per flop -- what bounds the K >= 1152 launches -- drops from (A + B/2) to (A/2 + B/2) per tile
  p.nsub = (two && d->nsub != 1 && p.n_tiles >= 2 && 3 * d->bn <= 512) ? 2 : 1;
  p.nacc = p.nsub == 2 ? 3 : 2;
  p.n_groups = (p.n_tiles + p.nsub - 1) / p.nsub;
  p.acc_stride = p.nsub == 2 ? (uint32_t)d->bn : (uint32_t)pow2_cols(d->bn);
  p.tmem_cols = (uint32_t)pow2_cols(int(p.nacc * p.acc_stride));   // <= 512 columns
  p.mode = d->mode; p.act = d->act; p.out_f32 = d->out_f32;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.bias_len = int(d->mode == CB_EPI_GEGLU ? 2 * d->cout : d->cout);
  p.rowbias_vec = (d->rowbias != nullptr) && ((reinterpret_cast<uintptr_t>(d->rowbias) & 15u) == 0) && (d->rowbias_ld % 4 == 0);
  p.residual = reinterpret_cast<const act_t*>(d->residual); p.res_ld = d->res_ld;
  p.out = d->out; p.out_ld = d->out_ld;
  p.out_scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  p.hd = d->heads_d; p.hdpad = d->heads_dpad; p.hheads = d->heads_h; p.htokens = d->heads_tokens;
  p.hwhich_stride = d->heads_which_stride;

  // ---- epilogue specialisation
  int epi = EPI_GENERIC;
  const bool simple16 = !d->out_f32 && d->act == CB_ACT_NONE && p.out_scale == 1.f && d->cout % 8 == 0;
  if (d->mode == CB_EPI_GEGLU) epi = EPI_GEGLU;
  else if (d->mode == CB_EPI_HEADS && simple16 && !d->rowbias && !d->residual) epi = EPI_HEADS;
  else if (d->mode == CB_EPI_LINEAR && simple16) {
    if (!d->rowbias && !d->residual) epi = EPI_PLAIN;
    else if (!d->rowbias && d->residual) epi = EPI_RES;
    else if (d->rowbias && !d->residual) epi = EPI_ROWBIAS;
  }
  CB_REQUIRE(d->mode != CB_EPI_HEADS || epi == EPI_HEADS, "cb_igemm: the heads epilogue takes bias only (no activation / residual / row bias / scale)");
  if (d->gn_partials) {
    CB_REQUIRE(epi == EPI_PLAIN || epi == EPI_RES || epi == EPI_ROWBIAS, "cb_igemm: GroupNorm partials need a plain 16-bit epilogue (bias / row bias / residual)");
    CB_REQUIRE(ksplit == 1, "cb_igemm: GroupNorm partials are not produced by split-K launches");
    CB_REQUIRE((d->tw * d->th) % 32 == 0, "cb_igemm: GroupNorm partials need tw * th %% 32 == 0 (a warp's 32 rows inside one image)");
    p.gn_part = d->gn_partials;
    p.gn_bpi = p.tiles_w * p.tiles_h;
    if (d->gn_rows_per_image > 0) {
      CB_REQUIRE(d->gn_row_offset >= 0 && d->gn_row_offset + p.gn_bpi <= d->gn_rows_per_image, "cb_igemm: GroupNorm partial rows out of range");
      p.gn_row0 = (int)d->gn_row_offset;
      p.gn_bpi = (int)d->gn_rows_per_image;
    }
  }

  // staged (TMA in / TMA out) epilogue: the small-K launches whose run time IS the epilogue; large-K convs keep the
  // direct form (their epilogue hides under the next tile's MMAs and the ring keeps its depth)
  const bool tma_ok = (d->out_ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(d->out) & 15u) == 0) &&
                      (!d->residual || (d->res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15u) == 0));
  bool staged = false;
  if (epi == EPI_GEGLU) {
    CB_REQUIRE(tma_ok, "cb_igemm: the GEGLU epilogue needs a 16-byte aligned output with out_ld %% 8 == 0");
    staged = true;
  } else if (epi == EPI_PLAIN || epi == EPI_RES || epi == EPI_ROWBIAS) {
    staged = tma_ok && (d->epilogue == CB_EPILOGUE_STAGED || (d->epilogue == CB_EPILOGUE_AUTO && num_k <= 48));
  }
  if (d->ln_partials_out) {
    CB_REQUIRE(staged && (epi == EPI_PLAIN || epi == EPI_RES) && ksplit == 1,
               "cb_igemm: LayerNorm partials come from the staged 16-bit epilogue (bias / residual) of an unsplit launch");
    p.ln_out = d->ln_partials_out;
    p.ln_out_slots = 2 * p.n_tiles;
  }
  if (d->ln_partials_in) {
    CB_REQUIRE(staged && (epi == EPI_PLAIN || epi == EPI_GEGLU), "cb_igemm: the LayerNorm fold needs a staged plain or GEGLU epilogue");
    CB_REQUIRE(d->ln_in_slots > 0 && d->ln_dim > 0 && d->bias, "cb_igemm: the LayerNorm fold needs a bias, the slot count and the row width");
    p.ln_in = d->ln_partials_in; p.ln_in_slots = d->ln_in_slots;
    p.ln_inv_dim = 1.f / float(d->ln_dim); p.ln_eps = d->ln_eps;
  }
  const int ocols = (epi == EPI_GEGLU) ? d->bn / 2 : d->bn;
  p.npan = staged ? (ocols + 63) / 64 : 0;
  p.pbw = d->tw < 32 ? d->tw : 32;
  p.pbh = d->th < 32 / p.pbw ? d->th : 32 / p.pbw;
  const size_t staging = (size_t)EPI_WARPS * p.npan * PANEL_BYTES;

  const bool strided_out = d->out_w_stride > 0 || d->out_h_stride > 0 || d->out_n_stride > 0;
  if (strided_out) {
    CB_REQUIRE(staged && epi != EPI_GEGLU, "cb_igemm: a strided output pixel grid needs the staged epilogue of a plain launch");
    CB_REQUIRE(!d->residual && ksplit == 1, "cb_igemm: a strided output takes no residual / split-K");
    CB_REQUIRE(d->out_w_stride >= d->cout && d->out_w_stride % 8 == 0 && d->out_h_stride % 8 == 0 && d->out_n_stride % 8 == 0 &&
               d->out_h_stride >= d->out_w_stride * d->w && d->out_n_stride >= d->out_h_stride * d->h,
               "cb_igemm: bad output pixel strides");
  }
  CUtensorMap mapO = mapA0, mapR = mapA0;
  if (staged) {
    uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t str[4] = {1, (uint64_t)d->out_ld, (uint64_t)d->out_ld * d->w, (uint64_t)d->out_ld * d->w * d->h};
    if (strided_out) { str[1] = (uint64_t)d->out_w_stride; str[2] = (uint64_t)d->out_h_stride; str[3] = (uint64_t)d->out_n_stride; }
    uint32_t box[4] = {32, (uint32_t)p.pbw, (uint32_t)p.pbh, (uint32_t)(32 / (p.pbw * p.pbh))};
    int rc = make_tmap_act(&mapO, d->out, 4, dims, str, box, 64);
    if (rc) return rc;
    if (epi == EPI_RES) {
      uint64_t rstr[4] = {1, (uint64_t)d->res_ld, (uint64_t)d->res_ld * d->w, (uint64_t)d->res_ld * d->w * d->h};
      rc = make_tmap_act(&mapR, d->residual, 4, dims, rstr, box, 64);
      if (rc) return rc;
    }
  }

  // ---- schedule: B-stationary when one N tile's whole K extent fits beside a >= 4-stage A ring and every CTA of
  //      that N tile gets several M tiles; otherwise stream A and B through the ring
  const size_t fixed = 1024 + 16 * 12 + 160 + 16 + sizeof(float) * 128 * EPI_WARPS + 64 + (d->gn_partials ? sizeof(float) * 2048 : 0);
  const size_t b_chunk = (size_t)(two ? d->bn / 2 : d->bn) * 128;
  const size_t res_bytes = (size_t)num_k * b_chunk;
  int resident = 0;
  unsigned grid = 0;
  if (!two && ksplit == 1 && d->stages <= 0 && p.n_tiles <= g_num_sms &&
      res_bytes + 4 * (size_t)A_STAGE_BYTES + staging + fixed <= (size_t)SMEM_LIMIT) {
    const int ctas_per_nt = g_num_sms / p.n_tiles;
    if (ctas_per_nt >= 1 && p.m_tiles >= 3 * ctas_per_nt) {
      resident = 1;
      grid = (unsigned)(ctas_per_nt * p.n_tiles);
      p.m_step = ctas_per_nt;
    }
  }
  p.resident_b = resident;
  const size_t stage_bytes = A_STAGE_BYTES + (resident ? 0 : (size_t)p.nsub * b_chunk);
  int stages = d->stages;
  if (stages <= 0) {
    stages = int(((size_t)SMEM_LIMIT - fixed - staging - (resident ? res_bytes : 0)) / stage_bytes);
    if (stages > 8) stages = 8;
  }
  CB_REQUIRE(stages >= 2 && stages <= 12, "cb_igemm: %d pipeline stages do not fit / out of range", stages);
  p.stages = stages;
  const size_t smem = (resident ? res_bytes : 0) + (size_t)stages * stage_bytes + staging + fixed;
  CB_REQUIRE(smem <= (size_t)SMEM_LIMIT, "cb_igemm: tile needs %zu bytes of shared memory", smem);
  {
    const long long work = (long long)(two ? p.m_pairs : p.m_tiles) * p.n_groups * ksplit;
    CB_REQUIRE(work < (1ll << 30), "cb_igemm: %lld work items exceed the tile scheduler's range", work);
    p.total_work = int(work);
    p.fd_tw = make_fastdiv(p.tiles_w);
    p.fd_th = make_fastdiv(p.tiles_h);
    p.fd_work = make_fastdiv(p.n_groups * ksplit);
  }
  if (two) {
    const long long total = (long long)p.m_pairs * p.n_groups * ksplit;
    const long long clusters = total < g_num_sms / 2 ? total : g_num_sms / 2;
    grid = (unsigned)(2 * clusters);
  } else if (!resident) {
    const long long total = (long long)p.m_tiles * p.n_groups * ksplit;
    grid = (unsigned)(total < g_num_sms ? total : g_num_sms);
  }

  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const IGemmKParams);
  // [single | pair | pair with N sub-tile groups][staged][epilogue]
  static const KernelFn kernels[3][2][EPI_COUNT] = {
      {{igemm_kernel<EPI_GENERIC, false, false, false>, igemm_kernel<EPI_PLAIN, false, false, false>,
        igemm_kernel<EPI_RES, false, false, false>, igemm_kernel<EPI_ROWBIAS, false, false, false>, nullptr,
        igemm_kernel<EPI_HEADS, false, false, false>},
       {nullptr, igemm_kernel<EPI_PLAIN, true, false, false>, igemm_kernel<EPI_RES, true, false, false>,
        igemm_kernel<EPI_ROWBIAS, true, false, false>, igemm_kernel<EPI_GEGLU, true, false, false>, nullptr}},
      {{igemm_kernel<EPI_GENERIC, false, true, false>, igemm_kernel<EPI_PLAIN, false, true, false>,
        igemm_kernel<EPI_RES, false, true, false>, igemm_kernel<EPI_ROWBIAS, false, true, false>, nullptr,
        igemm_kernel<EPI_HEADS, false, true, false>},
       {nullptr, igemm_kernel<EPI_PLAIN, true, true, false>, igemm_kernel<EPI_RES, true, true, false>,
        igemm_kernel<EPI_ROWBIAS, true, true, false>, igemm_kernel<EPI_GEGLU, true, true, false>, nullptr}},
      {{igemm_kernel<EPI_GENERIC, false, true, true>, igemm_kernel<EPI_PLAIN, false, true, true>,
        igemm_kernel<EPI_RES, false, true, true>, igemm_kernel<EPI_ROWBIAS, false, true, true>, nullptr,
        igemm_kernel<EPI_HEADS, false, true, true>},
       {nullptr, igemm_kernel<EPI_PLAIN, true, true, true>, igemm_kernel<EPI_RES, true, true, true>,
        igemm_kernel<EPI_ROWBIAS, true, true, true>, igemm_kernel<EPI_GEGLU, true, true, true>, nullptr}}};
  static DeviceOnce configured{};
  if (device_once_needed(configured)) {
    for (int tw = 0; tw < 3; ++tw)
      for (int st = 0; st < 2; ++st)
        for (int i = 0; i < EPI_COUNT; ++i)
          if (kernels[tw][st][i])
            CB_CHECK_CUDA(cudaFuncSetAttribute(kernels[tw][st][i], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    device_once_done(configured);
  }
  const KernelFn kfn = kernels[two ? (p.nsub == 2 ? 2 : 1) : 0][staged ? 1 : 0][epi];
  CB_REQUIRE(kfn != nullptr, "cb_igemm: internal: no kernel for epilogue %d staged %d", epi, (int)staged);
  if (two) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(NUM_THREADS);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl_enabled() ? 2 : 1;
    CB_CHECK_CUDA(cudaLaunchKernelEx(&lc, kfn, mapA0, mapA1, mapB, mapO, mapR, p));
  } else {
    (void)cb::launch_k(kfn, dim3(grid), dim3(NUM_THREADS), (size_t)(smem), stream, mapA0, mapA1, mapB, mapO, mapR, p);
  }
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}


// ----------------------------------------------------------------------------------------------------------------------
// split-K reduction: out[r][c] = act?(sum_s part[s][r][c] + bias[c] + rowbias[n(r)][c]) + residual[r][c]  -> 16-bit
// (partials summed in split order -> deterministic; 8 columns per thread, 16-byte accesses)
// ----------------------------------------------------------------------------------------------------------------------
namespace cb {
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long long split_stride, long long rows,
                                     int cout, long long part_ld, const float* __restrict__ bias,
                                     const float* __restrict__ rowbias, long long rowbias_ld, long long rows_per_image,
                                     const act_t* __restrict__ residual, long long res_ld, act_t* __restrict__ out,
                                     long long out_ld) {
  pdl_prologue();
  const int cv = cout >> 3;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cv;
    const int c = int(i - r * cv) << 3;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(part + s * split_stride + r * part_ld + c);
      const float4 b = *reinterpret_cast<const float4*>(part + s * split_stride + r * part_ld + c + 4);
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += __ldg(bias + c + e);
    }
    if (rowbias) {
      const float* rb = rowbias + (r / rows_per_image) * rowbias_ld + c;
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += __ldg(rb + e);
    }
    if (residual) {
      const uint4 rv = *reinterpret_cast<const uint4*>(residual + r * res_ld + c);
      const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_act2(ru[e]);
        f[2 * e] += t.x;
        f[2 * e + 1] += t.y;
      }
    }
    *reinterpret_cast<uint4*>(out + r * out_ld + c) =
        make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]), pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
  }
}
}  // namespace cb

extern "C" int cb_splitk_reduce(const float* part, int splits, int64_t rows, int64_t cout, int64_t part_ld, const float* bias,
                                const float* rowbias, int64_t rowbias_ld, int64_t rows_per_image, const void* residual,
                                int64_t res_ld, void* out, int64_t out_ld, cudaStream_t stream) {
  CB_REQUIRE(part && out && splits >= 1 && rows > 0 && cout > 0 && cout % 8 == 0, "cb_splitk_reduce: bad arguments");
  CB_REQUIRE(part_ld % 4 == 0 && out_ld % 8 == 0 && (!residual || res_ld % 8 == 0), "cb_splitk_reduce: unaligned leading dimensions");
  CB_REQUIRE(!rowbias || rows_per_image > 0, "cb_splitk_reduce: rowbias needs rows_per_image");
  const long long total = rows * (cout / 8);
  long long g = (total + 255) / 256;
  if (g > 148LL * 8) g = 148LL * 8;
  (void)cb::launch_k(splitk_reduce_kernel, dim3((unsigned)g), dim3(256), (size_t)(0), stream, part, splits, rows * part_ld, rows, (int)cout, part_ld, bias, rowbias,
                                                        rowbias_ld, rows_per_image, (const act_t*)residual, res_ld,
                                                        (act_t*)out, out_ld);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}
