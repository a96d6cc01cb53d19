// Long-sequence fused attention for head dims <= 64: ONE 128-row query tile per CTA, S and P DOUBLE-BUFFERED in TMEM,
// eight softmax warps of 16 rows each (sm_100a).
//
// Same arithmetic as attention.cu (S = Q K^T in TMEM, online softmax, P packed to 16 bits and fed back to the tensor
// core from TMEM, lazy 2^8 accumulator rescale, row sum from a ones column of V, kv blocks of 128), different plan.
// What the profiles of attention.cu show (profiles/r2_attention64.md): its two query tiles per CTA are each one
// barrier-locked unit -- all four warps of a tile wait for the same Q*K^T, exponentiate at the same time and hand P over
// at the same time -- so the MUFU pipe (the bound at d = 40: 16 exponentials per clock and SM) idles whenever both tiles
// are in their MUFU-free phase, and neither more warps per tile (16 softmax warps: 0.72 ms against 0.68 ms) nor a longer
// chain per warp (software pipelining: 0.78-0.94 ms) changes that.  Here the 512 TMEM columns buy slack instead of a
// second tile: S[2] (2 x 128 columns), P[2] (2 x 64), O (64).  Q*K^T runs TWO blocks ahead of the softmax and P*V of
// block j only needs P(j), so a warp that finishes block j starts block j+1 at once -- its scores have been ready for a
// whole block and the other P buffer is free -- and the eight warps drift up to one block apart instead of marching
// in phase: at any time some are in their exponentials while others load, reduce and store.
//   * TMEM is read with the `.16x256b` shape: a warp covers 16 lanes, a row's columns are spread over the four threads
//     of a quad (thread t: rows t/4 and t/4 + 8, columns 8i + 2(t%4) + {0,1} -- tools/micro/tmem_layout.cu), so a
//     thread holds 2 rows x 32 scores (64 registers), a row maximum is a quad reduction (two shuffles) and no warp
//     needs another warp's partial result; a pair of adjacent scores packs into one 32-bit P column, which is exactly
//     the thread's element of the `.16x128b` store shape.
//   * K/V tiles are used by one query tile instead of two (twice the L2 -> shared-memory traffic of attention.cu:
//     32 KB per ~1100 clocks and SM, far below the fill rate measured in tools/micro/smem_fill_rate.cu).
//
// Warps: 0 TMA producer | 1 TMEM allocator + MMA issuer | 2 ones column of V | 3 idle |
//        4..11 softmax: row half (warp-4)/4 of TMEM lane quarter warp%4.
// Replaces the attention cores at ldm/modules/attention.py:418-423 (Doggettx), :646-657 (Original), :811 (xformers) and
// sgm/modules/attention.py:507-511 (SDPA) for the self-attentions of the top UNet levels.
#include <cstdlib>

#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

constexpr int DB_BM = 128;               // query rows per CTA tile
constexpr int DB_BN = 128;               // kv rows per block
constexpr int DB_STAGES = 4;             // K/V ring depth
constexpr int DB_TILE_BYTES = 128 * 128; // [128 rows][64 x 16-bit], SWIZZLE_128B
constexpr int DB_THREADS = 384;          // 4 control warps + 8 softmax warps
constexpr int DB_BAR_BYTES = 512;
constexpr int DB_XCH_BYTES = 128 * 4;    // [row] fp32 row sums (head dim 64: no spare column of V)
constexpr float DB_TAU = 8.0f;           // rescale O only when the row max grew by more than 2^8 (P <= 256)

struct AttnDbParams {
  int nq, nk, d, heads, bh, ksteps, npv;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv;
  act_t* out;
};

CB_DEVINL uint32_t exp2_pack_sp(float s0, float s1, float scale, float m) {
  const float a0 = fmaf(s0, scale, -m), a1 = fmaf(s1, scale, -m);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  return pack_act2(e0, e1);
}

// 16 lanes x 64 columns of fp32 (8 atoms of 8 columns): thread t holds, for atom i, v[4i], v[4i+1] = row t/4, columns
// 8i + 2(t%4) + {0, 1} and v[4i+2], v[4i+3] = row t/4 + 8, the same columns
CB_DEVINL void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
CB_DEVINL void tmem_st_16x256b_x8(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// 16 lanes x 32 columns (8 atoms of 4 columns): thread t holds, for atom i, v[2i] = row t/4, column 4i + t%4 and
// v[2i+1] = row t/4 + 8, the same column
CB_DEVINL void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
CB_DEVINL float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
CB_DEVINL float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template <bool USE_ONES>
__global__ void __launch_bounds__(DB_THREADS, 1)
attention_db_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapV, const AttnDbParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();  // swizzled tiles need a 1024-byte aligned base
  const uint32_t sQ = base;
  const uint32_t sK = sQ + DB_TILE_BYTES;                       // [stages]
  const uint32_t sV = sK + uint32_t(DB_STAGES) * DB_TILE_BYTES; // [stages]
  const uint32_t bars = sV + uint32_t(DB_STAGES) * DB_TILE_BYTES;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u + 8u * uint32_t(s); };
  auto k_empty = [&](int s) { return bars + 40u + 8u * uint32_t(s); };
  auto v_full = [&](int s) { return bars + 72u + 8u * uint32_t(s); };
  auto v_empty = [&](int s) { return bars + 104u + 8u * uint32_t(s); };
  auto s_full = [&](int b) { return bars + 136u + 8u * uint32_t(b); };    // S[b] holds Q*K^T of a block
  auto p_full = [&](int b) { return bars + 152u + 8u * uint32_t(b); };    // every softmax thread has stored its part of P[b]
  auto s_free = [&](int b) { return bars + 168u + 8u * uint32_t(b); };    // every softmax thread has S[b] in registers
  auto pv_done = [&](int b) { return bars + 184u + 8u * uint32_t(b); };   // P*V that read P[b] has completed
  auto v_ready = [&](int s) { return bars + 200u + 8u * uint32_t(s); };   // V tile landed AND its ones column written
  const uint32_t q_free = bars + 232u;                                     // every Q*K^T of the item has completed
  const uint32_t o_free = bars + 240u;                                     // the item's O has been read out of TMEM
  const uint32_t tmem_slot = bars + 256u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  float* lsum = reinterpret_cast<float*>(smem_raw + (bars + DB_BAR_BYTES - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = (p.nk + DB_BN - 1) / DB_BN;
  // persistent CTA over work items (batch*head, query tile); barrier phases and the K/V ring position run on global
  // counters across items, so the producer and the issuer run ahead into the next item while the softmax warps finish
  const int qtiles = (p.nq + DB_BM - 1) / DB_BM;
  const int total_items = qtiles * p.bh;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    mbar_init(o_free, 2 * DB_BM);
    for (int s = 0; s < DB_STAGES; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
      mbar_init(v_ready(s), 32);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1); mbar_init(p_full(b), 2 * DB_BM);
      mbar_init(s_free(b), 2 * DB_BM); mbar_init(pv_done(b), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // the set-up above overlaps the previous kernel; global memory is only touched from here on
  auto tS = [&](int b) { return tmem_base + uint32_t(b) * 128u; };
  const uint32_t tO = tmem_base + 256u;
  auto tP = [&](int b) { return tmem_base + 320u + uint32_t(b) * 64u; };   // 128 rows x 128 16-bit values = 64 columns
  struct Ring {
    int st; uint32_t ph;
    __device__ __forceinline__ void next() { if (++st == DB_STAGES) { st = 0; ph ^= 1u; } }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      Ring r{0, 0u};
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int bh = item / qtiles, q_first = (item - bh * qtiles) * DB_BM;
        const int b_idx = bh / p.heads, h_idx = bh - b_idx * p.heads;
        mbar_wait(q_free, uint32_t(it & 1) ^ 1u);     // previous item's Q*K^T are complete (first item passes)
        mbar_expect_tx(q_full, DB_TILE_BYTES);
        tma_load_4d(sQ, &mapQ, q_full, 0, q_first, h_idx, b_idx);
        for (int j = 0; j < nblk; ++j, r.next()) {
          mbar_wait(k_empty(r.st), r.ph ^ 1u);
          mbar_expect_tx(k_full(r.st), DB_TILE_BYTES);
          tma_load_4d(sK + uint32_t(r.st) * DB_TILE_BYTES, &mapK, k_full(r.st), 0, j * DB_BN, h_idx, b_idx);
          mbar_wait(v_empty(r.st), r.ph ^ 1u);
          mbar_expect_tx(v_full(r.st), DB_TILE_BYTES);
          tma_load_4d(sV + uint32_t(r.st) * DB_TILE_BYTES, &mapV, v_full(r.st), 0, j * DB_BN, h_idx, b_idx);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      const uint64_t qd0 = make_sdesc_sw128(sQ, 16, 1024), kd0 = make_sdesc_sw128(sK, 16, 1024);
      const uint64_t vd0 = make_sdesc_sw128(sV, DB_TILE_BYTES, 1024);
      int it = 0;
      Ring rq{0, 0u}, rv{0, 0u};
      int cw = 0;      // global block count at the start of the item
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        mbar_wait(q_full, uint32_t(it & 1));
        auto do_qk = [&](int j) {     // Q*K^T of block j into S[g & 1]: once the scores of block g - 2 are in registers
          const int g = cw + j, b = g & 1;
          if (g >= 2) mbar_wait(s_free(b), uint32_t(((g - 2) >> 1) & 1));
          mbar_wait(k_full(rq.st), rq.ph);
          tc_fence_after();
          const uint64_t kd = kd0 + uint64_t(rq.st) * (DB_TILE_BYTES >> 4);
          for (int ks = 0; ks < p.ksteps; ++ks)
            umma_bf16(tS(b), qd0 + uint64_t(ks) * 2u, kd + uint64_t(ks) * 2u, p.idesc_qk, ks != 0);
          umma_commit(s_full(b));
          umma_commit(k_empty(rq.st));
          if (j == nblk - 1) umma_commit(q_free);      // the item's last use of Q
          rq.next();
        };
        auto do_pv = [&](int j) {
          const int g = cw + j, b = g & 1;
          mbar_wait(p_full(b), uint32_t((g >> 1) & 1));
          mbar_wait(v_ready(rv.st), rv.ph);
          if (j == 0 && it > 0) mbar_wait(o_free, uint32_t((it - 1) & 1));   // previous item's O has been read out
          tc_fence_after();
          const uint64_t vd = vd0 + uint64_t(rv.st) * (DB_TILE_BYTES >> 4);
#pragma unroll
          for (int ks = 0; ks < DB_BN / 16; ++ks)   // V: 16 kv rows = 2048 bytes; A = P from TMEM: 16 K-values = 8 columns
            umma_ts(tO, tP(b) + uint32_t(ks) * 8u, vd + uint64_t(ks) * (2048u >> 4), p.idesc_pv, (j > 0 || ks != 0) ? 1u : 0u);
          umma_commit(pv_done(b));
          umma_commit(v_empty(rv.st));
          rv.next();
        };
        do_qk(0);
        if (nblk > 1) do_qk(1);
        for (int j = 0; j < nblk; ++j) {
          if (j + 2 < nblk) do_qk(j + 2);    // enabled at the top of block j (scores of block j in registers) ...
          do_pv(j);                          // ... P(j) follows at its end
        }
        cw += nblk;
      }
    }
  } else if (warp == 2) {
    // ===================== ones column of V =====================
    // TMA zero-fills the pad columns of a V tile; column d becomes 1.0 so that P*V also yields the softmax row sum
    const int cw = p.d & 63;
    Ring rg{0, 0u};
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      for (int j = 0; j < nblk; ++j, rg.next()) {
        mbar_wait(v_full(rg.st), rg.ph);
        if (USE_ONES) {
          const uint32_t tile = sV + uint32_t(rg.st) * DB_TILE_BYTES;
#ifdef CB_FP16
          const unsigned short one_bits = 0x3C00;   // fp16 1.0
#else
          const unsigned short one_bits = 0x3F80;   // bf16 1.0
#endif
          for (int r = lane; r < DB_BN; r += 32) {
            const uint32_t addr = tile + uint32_t(r) * 128u + (((uint32_t(cw) >> 3) ^ (uint32_t(r) & 7u)) << 4) + uint32_t(cw & 7) * 2u;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(one_bits) : "memory");
          }
          fence_proxy_async_smem();
        }
        mbar_arrive(v_ready(rg.st));
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax warps: 16 rows each, a row's scores spread over a quad =====================
    const int hh = (warp - 4) >> 2;                    // which 16 lanes of the quarter
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int qc = lane & 3;                           // column position inside the quad
    const int rowA = quarter * 32 + hh * 16 + (lane >> 2), rowB = rowA + 8;   // this thread's two query rows of the tile
    const uint32_t lane16 = uint32_t(quarter * 32 + hh * 16) << 16;
    // epilogue roles (all 32 lanes of the quarter, `.32x32b`): this warp writes the columns [32 hh, 32 hh + 32)
    const int r = quarter * 32 + lane;
    const uint32_t tOe = tO + (uint32_t(quarter * 32) << 16);
    const int pair_bar = 1 + quarter;                  // named barrier of the two warps of a quarter (head dim 64 only)
    int cw = 0;                                        // global block count at the start of the item
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int bh = item / qtiles, q_first = (item - bh * qtiles) * DB_BM;
      float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int c = cw + j, b = c & 1;
        const uint32_t ph = uint32_t((c >> 1) & 1);
        mbar_wait(s_full(b), ph);
        tc_fence_after();
        uint32_t s[64];
        tmem_ld_16x256b_x8(tS(b) + lane16 + 0u, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld_16x256b_x8(tS(b) + lane16 + 64u, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free(b));      // this warp's rows of the scores are in registers
        const int nvalid = p.nk - j * DB_BN;
        if (nvalid < DB_BN) {   // ragged last block: K rows beyond nk were zero filled -> mask
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (8 * (e >> 2) + 2 * qc + (e & 1) >= nvalid) s[e] = 0xff800000u;  // -inf
        }
        float a0 = -INFINITY, a1 = -INFINITY, b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          a0 = fmaxf(a0, __uint_as_float(s[e]));
          a1 = fmaxf(a1, __uint_as_float(s[e + 1]));
          b0 = fmaxf(b0, __uint_as_float(s[e + 2]));
          b1 = fmaxf(b1, __uint_as_float(s[e + 3]));
        }
        const float mbA = quad_max(fmaxf(a0, a1)) * p.scale_log2, mbB = quad_max(fmaxf(b0, b1)) * p.scale_log2;
        if (j == 0) {
          mA = mbA;
          mB = mbB;
        } else {
          const bool needA = (mbA - mA) > DB_TAU, needB = (mbB - mB) > DB_TAU;
          if (__any_sync(0xffffffffu, needA || needB)) {
            mbar_wait(pv_done((c - 1) & 1), uint32_t(((c - 1) >> 1) & 1));   // O holds every P*V up to block j-1
            tc_fence_after();
            const float alA = needA ? exp2f(mA - mbA) : 1.f, alB = needB ? exp2f(mB - mbB) : 1.f;
            if (needA) mA = mbA;
            if (needB) mB = mbB;
            if (!USE_ONES) { lA *= alA; lB *= alB; }
            uint32_t o[32];                                // this warp's 16 rows of the accumulator, 64 columns
            tmem_ld_16x256b_x8(tO + lane16, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * ((e & 2) ? alB : alA));
            tmem_st_16x256b_x8(tO + lane16, o);
            tmem_st_wait();
          }
        }
        if (c >= 2) mbar_wait(pv_done(b), uint32_t(((c - 2) >> 1) & 1));   // the P*V of two blocks ago has read P[b]
        float rsA = 0.f, rsB = 0.f;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {              // 8 atoms = 64 score columns = 32 P columns per store
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            pk[e >> 1] = exp2_pack_sp(__uint_as_float(s[cc + e]), __uint_as_float(s[cc + e + 1]), p.scale_log2, mA);
            pk[(e >> 1) + 1] = exp2_pack_sp(__uint_as_float(s[cc + e + 2]), __uint_as_float(s[cc + e + 3]), p.scale_log2, mB);
            if (!USE_ONES) {
              const float2 fa = unpack_act2(pk[e >> 1]), fb = unpack_act2(pk[(e >> 1) + 1]);
              rsA += fa.x + fa.y;
              rsB += fb.x + fb.y;
            }
          }
          tmem_st_16x128b_x8(tP(b) + lane16 + uint32_t(cc), pk);
        }
        tmem_st_wait();
        if (!USE_ONES) { lA += rsA; lB += rsB; }
        tc_fence_before();
        mbar_arrive(p_full(b));
      }
      cw += nblk;

      // ---- epilogue: O / l -> out[b][q][head*d + :]; row = TMEM lane (`.32x32b`), this warp's columns [32 hh, 32 hh + 32)
      if (!USE_ONES) {                                   // the row sums live in the quads that own the rows
        lA = quad_sum(lA);
        lB = quad_sum(lB);
        if (qc == 0) { lsum[rowA] = lA; lsum[rowB] = lB; }
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      }
      mbar_wait(pv_done((cw - 1) & 1), uint32_t(((cw - 1) >> 1) & 1));
      tc_fence_after();
      const int q = q_first + r;
      const int b = bh / p.heads, head = bh - b * p.heads;
      float l;
      uint32_t o[32];
      if (USE_ONES) {
        tmem_ld32(tOe + uint32_t(p.d & 32), o);          // the 32-column chunk that holds column d (the row sum)
        tmem_ld_wait();
        l = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e == (p.d & 31)) l = __uint_as_float(o[e]);
        if ((p.d & 32) != hh * 32) {
          tmem_ld32(tOe + uint32_t(hh) * 32u, o);
          tmem_ld_wait();
        }
      } else {
        l = lsum[r];
        tmem_ld32(tOe + uint32_t(hh) * 32u, o);
        tmem_ld_wait();
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // both warps have read lsum: the next item may overwrite it
      }
      tc_fence_before();
      mbar_arrive(o_free);      // the next item's first P*V may overwrite O
      const float inv_l = 1.f / l;
      if (q < p.nq) {
        act_t* orow = p.out + (static_cast<long long>(b) * p.nq + q) * (static_cast<long long>(p.heads) * p.d) +
                      static_cast<long long>(head) * p.d + hh * 32;
#pragma unroll
        for (int gq = 0; gq < 32; gq += 8) {
          if (hh * 32 + gq < p.d) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[gq + e]) * inv_l;
            *reinterpret_cast<uint4*>(orow + gq) = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]),
                                                              pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// d <= 64 (one 64-column panel), any nq / nk; called by cb_attention for long key sequences
int launch_attention_db(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld, void* out,
                        int64_t batch, int64_t heads, int64_t nq, int64_t nk, int d, float scale, int num_sms,
                        cudaStream_t stream) {
  CUtensorMap mq, mk, mv;
  uint32_t box[4] = {64, 128, 1, 1};
  {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)nq, (uint64_t)heads, (uint64_t)batch};
    uint64_t str[4] = {1, (uint64_t)q_ld, (uint64_t)d, (uint64_t)(nq * q_ld)};
    int rc = make_tmap_act(&mq, q, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)nk, (uint64_t)heads, (uint64_t)batch};
    uint64_t strk[4] = {1, (uint64_t)k_ld, (uint64_t)d, (uint64_t)(nk * k_ld)};
    uint64_t strv[4] = {1, (uint64_t)v_ld, (uint64_t)d, (uint64_t)(nk * v_ld)};
    int rc = make_tmap_act(&mk, k, 4, dims, strk, box);
    if (rc) return rc;
    rc = make_tmap_act(&mv, v, 4, dims, strv, box);
    if (rc) return rc;
  }
  const bool use_ones = d < 64;            // spare column d of V carries 1.0 -> O[:, d] = softmax row sum
  AttnDbParams p{};
  p.nq = (int)nq; p.nk = (int)nk; p.d = d; p.heads = (int)heads; p.bh = (int)(batch * heads);
  p.ksteps = (d + 15) / 16;                // the pad columns are zeros: no K-step beyond the next multiple of 16
  p.npv = use_ones ? ((d + 1 + 15) / 16) * 16 : 64;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.idesc_qk = make_idesc_f16(128, 128, 0, 0);
  p.idesc_pv = make_idesc_f16(128, p.npv, 0, 1);   // B = V is MN-major
  p.out = (act_t*)out;
  const size_t smem = (size_t)(1 + 2 * DB_STAGES) * DB_TILE_BYTES + DB_BAR_BYTES + DB_XCH_BYTES;
  static DeviceOnce configured{};
  if (device_once_needed(configured)) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_db_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_db_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    device_once_done(configured);
  }
  const long long items = ((nq + DB_BM - 1) / DB_BM) * batch * heads;
  dim3 grid((unsigned)(items < num_sms ? items : num_sms));   // persistent: one CTA per SM walks the work items
  if (use_ones) (void)cb::launch_k(attention_db_kernel<true>, dim3(grid), dim3(DB_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
  else (void)cb::launch_k(attention_db_kernel<false>, dim3(grid), dim3(DB_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

}  // namespace cb
