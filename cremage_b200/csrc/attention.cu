// Fused flash-style attention for sm_100a:  O = softmax(scale * Q K^T) V, never materialising the score matrix.
//   S = Q K^T      tcgen05.mma  (A = Q tile [128 x dpad], B = K tile [128 x dpad], both K-major)  -> TMEM cols [0,128)
//   online softmax 128 threads, one query row each, tcgen05.ld of S, exp2 with running max / sum in fp32
//   O += P V       tcgen05.mma  (A = P [128 x 128] bf16 written to swizzled smem, B = V tile, MN-major) -> TMEM cols [128, 128+dpad)
// Q/K/V arrive through 3-D TMA boxes from the per-head padded layout [bh][tokens][dpad] written by the QKV
// projection epilogue (CB_EPI_HEADS).  One CTA = 128 query rows of one (batch, head); for dpad = 64 two CTAs are
// co-resident per SM so one CTA's softmax overlaps the other's MMAs.
// Replaces the attention cores at ldm/modules/attention.py:418-423 (Doggettx), :646-657 (Original), :811 (xformers).
#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BN = 128;   // kv rows per iteration
constexpr int PANEL_BYTES = 128 * 128;  // [128 rows][64 bf16]

struct AttnParams {
  int nq, nk, d, dpad, np, heads, stages;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv, tmem_cols;
  act_t* out;
};

__global__ void __launch_bounds__(128, 1)
attention_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();  // swizzled tiles need a 1024-byte aligned base
  const uint32_t tile_bytes = uint32_t(p.np) * PANEL_BYTES;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + tile_bytes;                       // [stages]
  const uint32_t sV = sK + uint32_t(p.stages) * tile_bytes;  // [stages]
  const uint32_t sP = sV + uint32_t(p.stages) * tile_bytes;  // 2 panels
  const uint32_t bars = sP + 2u * PANEL_BYTES;
  const uint32_t q_bar = bars, s_bar = bars + 8;
  auto k_bar = [&](int s) { return bars + 16u + 8u * uint32_t(s); };
  auto v_bar = [&](int s) { return bars + 32u + 8u * uint32_t(s); };
  const uint32_t tmem_slot = bars + 48u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qblk = blockIdx.x, bh = blockIdx.y;
  const int nblk = (p.nk + ATT_BN - 1) / ATT_BN;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_bar, 1);
    mbar_init(s_bar, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(k_bar(s), 1); mbar_init(v_bar(s), 1); }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base;          // 128 columns
  const uint32_t tO = tmem_base + 128u;   // dpad columns

  auto load_tile = [&](const CUtensorMap* m, uint32_t dst, uint32_t bar, int row0) {
    mbar_expect_tx(bar, tile_bytes);
    for (int pn = 0; pn < p.np; ++pn) tma_load_3d(dst + uint32_t(pn) * PANEL_BYTES, m, bar, pn * 64, row0, bh);
  };
  auto issue_qk = [&](int kstage) {
    const uint32_t kb = sK + uint32_t(kstage) * tile_bytes;
    const int ksteps = p.dpad / 16;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t off = uint32_t(ks >> 2) * PANEL_BYTES + uint32_t(ks & 3) * 32u;
      umma_bf16(tS, make_sdesc_sw128(sQ + off, 16, 1024), make_sdesc_sw128(kb + off, 16, 1024), p.idesc_qk, ks != 0);
    }
  };
  auto issue_pv = [&](int vstage, bool accumulate) {
    const uint32_t vb = sV + uint32_t(vstage) * tile_bytes;
    for (int ks = 0; ks < ATT_BN / 16; ++ks) {
      const uint32_t aoff = uint32_t(ks >> 2) * PANEL_BYTES + uint32_t(ks & 3) * 32u;   // P: K-major
      const uint32_t boff = uint32_t(ks) * 2048u;                                        // V: 16 kv rows = 2 atoms
      umma_bf16(tO, make_sdesc_sw128(sP + aoff, 16, 1024), make_sdesc_sw128(vb + boff, PANEL_BYTES, 1024), p.idesc_pv,
                (accumulate || ks != 0) ? 1u : 0u);
    }
  };

  if (tid == 0) {
    load_tile(&mapQ, sQ, q_bar, qblk * ATT_BM);
    load_tile(&mapK, sK, k_bar(0), 0);
    load_tile(&mapV, sV, v_bar(0), 0);
    mbar_wait(q_bar, 0);
    mbar_wait(k_bar(0), 0);
    tc_fence_after();
    issue_qk(0);
    umma_commit(s_bar);
  }

  const int r = tid;                                   // query row within the tile == TMEM lane
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  float m_run = -INFINITY, l_run = 0.f;
  const uint32_t p_row = sP + uint32_t(r) * 128u;
  const uint32_t sw = uint32_t(r & 7);

  for (int j = 0; j < nblk; ++j) {
    mbar_wait(s_bar, uint32_t(j & 1));   // S_j ready, every earlier MMA (PV_{j-1}) retired
    tc_fence_after();
    if (tid == 0) {
      if (p.stages == 1 && j > 0) load_tile(&mapV, sV, v_bar(0), j * ATT_BN);   // V buffer freed by PV_{j-1}
      if (j + 1 < nblk) {
        const int st = (j + 1) % p.stages;
        load_tile(&mapK, sK + uint32_t(st) * tile_bytes, k_bar(st), (j + 1) * ATT_BN);
        if (p.stages == 2) load_tile(&mapV, sV + uint32_t(st) * tile_bytes, v_bar(st), (j + 1) * ATT_BN);
      }
    }
    const int kv0 = j * ATT_BN;
    const int nvalid = min(ATT_BN, p.nk - kv0);   // columns >= nvalid are padding (K rows zero filled)

    // ---- pass 1: row maximum of the raw scores
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < ATT_BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tS + lane_off + uint32_t(c), v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (c + e < nvalid) mx = fmaxf(mx, __uint_as_float(v[e]));
    }
    const float m_new = fmaxf(m_run, mx * p.scale_log2);
    const float alpha = exp2f(m_run - m_new);   // 0 on the first block (m_run = -inf)
    m_run = m_new;

    // ---- rescale the running output if any row of this warp moved its maximum
    if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
      for (int c = 0; c < p.dpad; c += 32) {
        uint32_t o[32];
        tmem_ld32(tO + lane_off + uint32_t(c), o);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
        tmem_st32(tO + lane_off + uint32_t(c), o);
      }
      tmem_st_wait();
    }

    // ---- pass 2: P = exp2(S*scale - m), row sum, bf16 -> swizzled smem (K-major A operand of the PV MMA)
    float rs = 0.f;
#pragma unroll 1
    for (int c = 0; c < ATT_BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tS + lane_off + uint32_t(c), v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float p0 = (c + e < nvalid) ? exp2f(fmaf(__uint_as_float(v[e]), p.scale_log2, -m_new)) : 0.f;
        float p1 = (c + e + 1 < nvalid) ? exp2f(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, -m_new)) : 0.f;
        // accumulate the row sum from the bf16-rounded values so numerator and denominator agree
        const uint32_t u = pack_act2(p0, p1);
        const float2 back = unpack_act2(u);
        rs += back.x + back.y;
        pk[e >> 1] = u;
      }
      const uint32_t panel = uint32_t(c >> 6) * PANEL_BYTES;
      const uint32_t chunk0 = uint32_t((c & 63) >> 3);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t addr = p_row + panel + (((chunk0 + uint32_t(g)) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * g]), "r"(pk[4 * g + 1]),
                     "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                     : "memory");
      }
    }
    l_run = l_run * alpha + rs;

    fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const int vs = j % p.stages;
      mbar_wait(v_bar(vs), uint32_t((j / p.stages) & 1));
      tc_fence_after();
      issue_pv(vs, j > 0);
      if (j + 1 < nblk) {
        const int ks = (j + 1) % p.stages;
        mbar_wait(k_bar(ks), uint32_t(((j + 1) / p.stages) & 1));
        tc_fence_after();
        issue_qk(ks);
      }
      umma_commit(s_bar);
    }
  }

  // ---- epilogue: O / l -> out[b][q][head*d + :]
  mbar_wait(s_bar, uint32_t(nblk & 1));
  tc_fence_after();
  const int q = qblk * ATT_BM + r;
  const int b = bh / p.heads, head = bh - b * p.heads;
  const float inv_l = 1.f / l_run;
  act_t* orow = p.out + (static_cast<long long>(b) * p.nq + q) * (static_cast<long long>(p.heads) * p.d) +
                        static_cast<long long>(head) * p.d;
#pragma unroll 1
  for (int c = 0; c < p.dpad; c += 32) {
    uint32_t o[32];
    tmem_ld32(tO + lane_off + uint32_t(c), o);
    tmem_ld_wait();
    if (q < p.nq) {
#pragma unroll
      for (int g = 0; g < 32; g += 8) {
        if (c + g < p.d) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[g + e]) * inv_l;
          *reinterpret_cast<uint4*>(orow + c + g) = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]),
                                                               pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_attention(const void* q, const void* k, const void* v, void* out, int64_t batch, int64_t heads,
                            int64_t nq, int64_t nk, int d, int dpad, float scale, cudaStream_t stream) {
  CB_REQUIRE(q && k && v && out, "cb_attention: null pointer");
  CB_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0, "cb_attention: empty problem");
  CB_REQUIRE(d > 0 && d % 8 == 0 && dpad % 64 == 0 && dpad >= d && dpad <= 192,
             "cb_attention: head dim %d (padded %d) unsupported: d %% 8 == 0, dpad in {64,128,192}", d, dpad);
  const int64_t bh = batch * heads;
  CB_REQUIRE(bh <= 65535, "cb_attention: batch*heads = %lld exceeds the grid limit", (long long)bh);
  CUtensorMap mq, mk, mv;
  uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {(uint64_t)dpad, (uint64_t)nq, (uint64_t)bh};
    uint64_t str[3] = {1, (uint64_t)dpad, (uint64_t)dpad * nq};
    int rc = make_tmap_act(&mq, q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)dpad, (uint64_t)nk, (uint64_t)bh};
    uint64_t str[3] = {1, (uint64_t)dpad, (uint64_t)dpad * nk};
    int rc = make_tmap_act(&mk, k, 3, dims, str, box);
    if (rc) return rc;
    rc = make_tmap_act(&mv, v, 3, dims, str, box);
    if (rc) return rc;
  }
  AttnParams p{};
  p.nq = (int)nq; p.nk = (int)nk; p.d = d; p.dpad = dpad; p.np = dpad / 64; p.heads = (int)heads;
  p.stages = dpad <= 128 ? 2 : 1;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.idesc_qk = make_idesc_f16(128, 128, 0, 0);
  p.idesc_pv = make_idesc_f16(128, dpad, 0, 1);   // B = V is MN-major
  p.tmem_cols = (128 + dpad) <= 256 ? 256u : 512u;
  p.out = (act_t*)out;
  const size_t smem = (size_t)(1 + 2 * p.stages) * p.np * PANEL_BYTES + 2 * PANEL_BYTES + 64;
  CB_REQUIRE(smem <= 227 * 1024, "cb_attention: needs %zu bytes of shared memory", smem);
  static thread_local bool configured = false;
  if (!configured) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)((nq + ATT_BM - 1) / ATT_BM), (unsigned)bh);
  attention_kernel<<<grid, 128, smem, stream>>>(mq, mk, mv, p);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}
