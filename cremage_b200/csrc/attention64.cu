// Fused attention for head dims <= 64 (SD1.5 top level d = 40, SDXL d = 64): the exponential-bound regime.
//
// Same arithmetic as attention.cu (S = Q K^T in TMEM, online softmax one thread per query row, P packed to 16 bits and
// fed back to the tensor core from TMEM, lazy accumulator rescale, row sum from a ones column of V), re-tiled so that
// THREE softmax warpgroups share an SM instead of two and every phase is half as long:
//   * kv blocks of 64 rows: a row's block of scores is 64 registers, three warpgroups + four control warps are 16
//     warps at 128 registers (registers are granted per four warps; a fourth warpgroup would cap every thread at 96
//     and force P to alias S, which puts Q K^T of the next block behind the softmax of this one: measured 0.84 ms
//     against 0.80 ms for the r1 kernel, profiles/r2_attention64.md);
//   * up to three 128-row query tiles of one (batch, head) per CTA share every K/V tile; per tile TMEM holds S (64
//     columns), P (32) and O (64): 480 of 512 columns, so S, P and O never alias and Q K^T of block j+1 is issued the
//     moment the scores of block j are in registers -- a warpgroup never waits for the tensor core in steady state;
//   * Q K^T issues ceil(d / 16) K-steps and P V uses N = ceil16(d + 1) (48 for d = 40): no MMA multiplies the padding
//     beyond the next multiple of 16.
// Optimistic softmax: the exponentials of block j use the running maximum known BEFORE the block, so the first
// exponential issues as soon as the scores are in registers and the 64-wide max reduction runs beside the MUFU pipe
// instead of in front of it; only if the block's own maximum exceeds the stale one by more than 2^8 (any lane of the
// warp) the accumulator is rescaled and the block's exponentials are redone from the registers.
// Two issuer THREADS (elect_one, so the compiler emits plain uniform-datapath code) drive the tiles as event loops over
// `mbarrier.test_wait` (non-blocking; `try_wait` suspends the thread for a time slice when the phase is incomplete).
//
// Warps: 0 TMA producer | 1 TMEM allocator + MMA issuer (tiles 0, 1) | 2 ones column of V | 3 MMA issuer (tile 2) |
//        4..15 softmax warpgroups 0..2.
// Replaces the attention cores at ldm/modules/attention.py:418-423 (Doggettx), :646-657 (Original), :811 (xformers) and
// sgm/modules/attention.py:507-511 (SDPA).
#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

constexpr int A64_BM = 128;              // query rows per tile (one softmax warpgroup)
constexpr int A64_BN = 64;               // kv rows per block
constexpr int A64_TILES = 3;             // query tiles per CTA (TMEM: 3 x (64 S + 32 P + 64 O) columns)
constexpr int A64_STAGES = 8;            // K/V ring depth (8 KB + 8 KB per stage)
constexpr int A64_QBYTES = 128 * 128;    // [128 rows][64 x 16-bit], SWIZZLE_128B
constexpr int A64_KVBYTES = 64 * 128;    // [64 rows][64 x 16-bit]
constexpr int A64_THREADS = 32 * (4 + 4 * A64_TILES);   // 512
constexpr float A64_TAU = 8.0f;          // rescale O only when the row max grew by more than 2^8 (P <= 256)

struct Attn64Params {
  int nq, nk, d, heads, bh, nt, ksteps, npv;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv;
  act_t* out;
};

CB_DEVINL uint32_t mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}

CB_DEVINL void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

template <bool SUM>
CB_DEVINL uint32_t exp2_pack64(float s0, float s1, float scale, float m, float& rs) {
  const float a0 = fmaf(s0, scale, -m), a1 = fmaf(s1, scale, -m);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  if (SUM) rs += e0 + e1;   // explicit row sum (head dim 64: no spare column of V) from the fp32 exponentials
  return pack_act2(e0, e1);
}

template <bool USE_ONES>
__global__ void __launch_bounds__(A64_THREADS, 1)
attention64_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const Attn64Params p) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sQ = base;                                          // [3] x 16 KB
  const uint32_t sK = sQ + A64_TILES * A64_QBYTES;                   // [8] x 8 KB
  const uint32_t sV = sK + A64_STAGES * A64_KVBYTES;                 // [8] x 8 KB
  const uint32_t bars = sV + A64_STAGES * A64_KVBYTES;
  const uint32_t q_full = bars, q_free = bars + 8u;
  auto k_full = [&](int s) { return bars + 16u + 8u * uint32_t(s); };
  auto k_empty = [&](int s) { return bars + 80u + 8u * uint32_t(s); };
  auto v_full = [&](int s) { return bars + 144u + 8u * uint32_t(s); };
  auto v_empty = [&](int s) { return bars + 208u + 8u * uint32_t(s); };
  auto v_ready = [&](int s) { return bars + 272u + 8u * uint32_t(s); };    // V tile landed AND its ones column written
  auto s_full = [&](int w) { return bars + 336u + 8u * uint32_t(w); };     // Q*K^T of a block complete
  auto s_free = [&](int w) { return bars + 368u + 8u * uint32_t(w); };     // its scores are in registers
  auto p_full = [&](int w) { return bars + 400u + 8u * uint32_t(w); };     // P of a block written
  auto pv_done = [&](int w) { return bars + 432u + 8u * uint32_t(w); };    // P*V of a block complete
  auto o_free = [&](int w) { return bars + 464u + 8u * uint32_t(w); };     // the item's O has been read out of TMEM
  const uint32_t tmem_slot = bars + 496u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = (p.nk + A64_BN - 1) / A64_BN;
  const int rows_per_item = p.nt * A64_BM;
  const int qgroups = (p.nq + rows_per_item - 1) / rows_per_item;
  const int total_items = qgroups * p.bh;
  auto item_q_first = [&](int item) { return (item % qgroups) * rows_per_item; };
  auto item_bh = [&](int item) { return item / qgroups; };
  auto item_nact = [&](int item) {
    const int left = p.nq - item_q_first(item);
    const int t = (left + A64_BM - 1) / A64_BM;
    return t < p.nt ? t : p.nt;
  };

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    mbar_init(q_free, 2);                   // both issuers
    for (int s = 0; s < A64_STAGES; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 2);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 2);
      mbar_init(v_ready(s), 32);
    }
    for (int w = 0; w < A64_TILES; ++w) {
      mbar_init(s_full(w), 1); mbar_init(s_free(w), A64_BM); mbar_init(p_full(w), A64_BM);
      mbar_init(pv_done(w), 1); mbar_init(o_free(w), A64_BM);
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
  auto tS = [&](int w) { return tmem_base + uint32_t(w) * 64u; };
  auto tP = [&](int w) { return tmem_base + 192u + uint32_t(w) * 32u; };       // 64 16-bit values = 32 columns
  auto tO = [&](int w) { return tmem_base + 288u + uint32_t(w) * 64u; };
  auto stage_of = [&](int g) { return g & (A64_STAGES - 1); };
  auto phase_of = [&](int g) { return uint32_t((g / A64_STAGES) & 1); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0, g = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int bh = item_bh(item), q_first = item_q_first(item), nact = item_nact(item);
        const int b_idx = bh / p.heads, h_idx = bh - b_idx * p.heads;
        mbar_wait(q_free, uint32_t(it & 1) ^ 1u);      // every Q*K^T of the previous item has completed
        mbar_expect_tx(q_full, uint32_t(nact) * A64_QBYTES);
        for (int w = 0; w < nact; ++w)
          tma_load_4d(sQ + uint32_t(w) * A64_QBYTES, &mapQ, q_full, 0, q_first + w * A64_BM, h_idx, b_idx);
        for (int j = 0; j < nblk; ++j, ++g) {
          const int st = stage_of(g);
          const uint32_t ph = phase_of(g);
          mbar_wait(k_empty(st), ph ^ 1u);
          mbar_expect_tx(k_full(st), A64_KVBYTES);
          tma_load_4d(sK + uint32_t(st) * A64_KVBYTES, &mapK, k_full(st), 0, j * A64_BN, h_idx, b_idx);
          mbar_wait(v_empty(st), ph ^ 1u);
          mbar_expect_tx(v_full(st), A64_KVBYTES);
          tma_load_4d(sV + uint32_t(st) * A64_KVBYTES, &mapV, v_full(st), 0, j * A64_BN, h_idx, b_idx);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: two threads, event loops over (up to) two query tiles each =====================
    if (elect_one()) {
      const int w0 = (warp == 1) ? 0 : 2, wmax = (warp == 1) ? 2 : 1;
      const uint64_t qd0 = make_sdesc_sw128(sQ, 16, 1024), kd0 = make_sdesc_sw128(sK, 16, 1024);
      const uint64_t vd0 = make_sdesc_sw128(sV, A64_KVBYTES, 1024);
      auto issue_qk = [&](int w, int kstage) {
        const uint64_t qd = qd0 + uint64_t(w) * (A64_QBYTES >> 4), kd = kd0 + uint64_t(kstage) * (A64_KVBYTES >> 4);
        for (int ks = 0; ks < p.ksteps; ++ks)
          umma_bf16(tS(w), qd + uint64_t(ks) * 2u, kd + uint64_t(ks) * 2u, p.idesc_qk, ks != 0);
      };
      auto issue_pv = [&](int w, int vstage, bool accumulate) {
        const uint64_t vd = vd0 + uint64_t(vstage) * (A64_KVBYTES >> 4);
#pragma unroll
        for (int ks = 0; ks < A64_BN / 16; ++ks)     // A = P from TMEM: 16 kv values of a row = 8 columns
          umma_ts(tO(w), tP(w) + uint32_t(ks) * 8u, vd + uint64_t(ks) * (2048u >> 4), p.idesc_pv,
                  (accumulate || ks != 0) ? 1u : 0u);
      };
      int it = 0, g0 = 0;
      int cw[2] = {0, 0};     // blocks each tile has been active for (phases of s_full / s_free / p_full / pv_done)
      int aw[2] = {0, 0};     // items each tile has been active for (phase of o_free)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it, g0 += nblk) {
        const int left = item_nact(item) - w0;
        const int nact = left < 0 ? 0 : (left > wmax ? wmax : left);       // my active tiles
        mbar_wait(q_full, uint32_t(it & 1));
        if (nact == 0) {
          // idle issuer (ragged last query group): stay in lockstep with the ring, release every stage it is handed
          mbar_arrive(q_free);
          for (int j = 0; j < nblk; ++j) {
            mbar_wait(k_full(stage_of(g0 + j)), phase_of(g0 + j));
            mbar_arrive(k_empty(stage_of(g0 + j)));
            mbar_wait(v_ready(stage_of(g0 + j)), phase_of(g0 + j));
            mbar_arrive(v_empty(stage_of(g0 + j)));
          }
          continue;
        }
        int jq[2] = {0, 0}, jp[2] = {0, 0};
        int remaining = 2 * nact * nblk;
        long long t_idle = 0;               // bounded spin: a protocol bug must trap, not hang the GPU box
        while (remaining > 0) {
          bool progressed = false;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i >= nact) continue;
            const int w = w0 + i;
            // ---- Q*K^T of block jq: the scores of this tile's previous block are in registers, K tile landed
            if (jq[i] < nblk) {
              const int j = jq[i], c = cw[i] + j, g = g0 + j, st = stage_of(g);
              if ((c == 0 || mbar_test(s_free(w), uint32_t((c - 1) & 1))) && mbar_test(k_full(st), phase_of(g))) {
                tc_fence_after();
                issue_qk(w, st);
                umma_commit(s_full(w));
                // my tiles move through the ring in order: the stage is done when the LAST of them has used it
                if (nact == 1 || jq[i ^ 1] > j) {
                  umma_commit(k_empty(st));
                  if (j == nblk - 1) umma_commit(q_free);      // my last use of the item's Q
                }
                jq[i] = j + 1;
                --remaining;
                progressed = true;
              }
            }
            // ---- P*V of block jp: P written, V tile (with its ones column) landed, previous item's O read out
            if (jp[i] < nblk) {
              const int j = jp[i], g = g0 + j, st = stage_of(g);
              if (mbar_test(p_full(w), uint32_t((cw[i] + j) & 1)) && mbar_test(v_ready(st), phase_of(g)) &&
                  (j > 0 || aw[i] == 0 || mbar_test(o_free(w), uint32_t((aw[i] - 1) & 1)))) {
                tc_fence_after();
                issue_pv(w, st, j > 0);
                umma_commit(pv_done(w));
                if (nact == 1 || jp[i ^ 1] > j) umma_commit(v_empty(st));
                jp[i] = j + 1;
                --remaining;
                progressed = true;
              }
            }
          }
          if (progressed) {
            t_idle = 0;
          } else {
            if (t_idle == 0) t_idle = clock64();
            else if (clock64() - t_idle > 4000000000LL) __trap();
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (i < nact) { cw[i] += nblk; ++aw[i]; }
      }
    }
  } else if (warp == 2) {
    // ===================== V ones-column warp =====================
    // TMA zero-fills the pad columns of a V tile; column d becomes 1.0 so that P*V also yields the softmax row sum.
    const uint32_t onec = uint32_t(p.d) & 63u;
#ifdef CB_FP16
    const unsigned short one_bits = 0x3C00;   // fp16 1.0
#else
    const unsigned short one_bits = 0x3F80;   // bf16 1.0
#endif
    int g = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      for (int j = 0; j < nblk; ++j, ++g) {
        const int st = stage_of(g);
        mbar_wait(v_full(st), phase_of(g));
        if (USE_ONES) {
          const uint32_t tile = sV + uint32_t(st) * A64_KVBYTES;
#pragma unroll
          for (int r = lane; r < A64_BN; r += 32) {
            const uint32_t addr = tile + uint32_t(r) * 128u + (((onec >> 3) ^ (uint32_t(r) & 7u)) << 4) + (onec & 7u) * 2u;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(one_bits) : "memory");
          }
          fence_proxy_async_smem();
        }
        mbar_arrive(v_ready(st));
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int w = (warp - 4) >> 2;
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t tSw = tS(w) + lane_off, tOw = tO(w) + lane_off, tPw = tP(w) + lane_off;
    int cw = 0, aw = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      if (w >= item_nact(item)) continue;
      const int bh = item_bh(item), q_first = item_q_first(item);
      float m_used = -INFINITY, l_run = 0.f;

      for (int j = 0; j < nblk; ++j) {
        const int c = cw + j;
        mbar_wait(s_full(w), uint32_t(c & 1));
        tc_fence_after();
        uint32_t sa[32], sb[32];                   // the row's 64 scores of this block
        tmem_ld32(tSw, sa);
        tmem_ld32(tSw + 32u, sb);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free(w));                    // the scores are in registers: Q*K^T of the next block may overwrite S
        const int nvalid = p.nk - j * A64_BN;
        if (nvalid < A64_BN) {   // ragged last block: K rows beyond nk were zero filled -> mask
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e >= nvalid) sa[e] = 0xff800000u;  // -inf
            if (e + 32 >= nvalid) sb[e] = 0xff800000u;
          }
        }
        float rs = 0.f;
        // exponentials of one 32-score half -> 16 packed registers -> straight to the P region (the store may be
        // repeated by the redo path below; P*V is not issued before this thread arrives on p_full)
        auto exps = [&](float m) {
          rs = 0.f;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float x0 = __uint_as_float(half ? sb[e] : sa[e]), x1 = __uint_as_float(half ? sb[e + 1] : sa[e + 1]);
              pk[e >> 1] = exp2_pack64<!USE_ONES>(x0, x1, p.scale_log2, m, rs);
            }
            tmem_st16(tPw + uint32_t(half * 16), pk);
          }
        };
        auto block_max = [&]() {
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[e]), __uint_as_float(sb[e])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sa[e + 1]), __uint_as_float(sb[e + 1])));
            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sa[e + 2]), __uint_as_float(sb[e + 2])));
            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sa[e + 3]), __uint_as_float(sb[e + 3])));
          }
          return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
        };
        if (c > 0) mbar_wait(pv_done(w), uint32_t((c - 1) & 1));   // the previous P*V has finished reading P (and O is current)
        if (j == 0) {
          m_used = block_max();      // first block of the row: the maximum has to be known first
          exps(m_used);
        } else {
          exps(m_used);              // optimistic: against the running maximum known before this block
          const float m_blk = block_max();
          const bool need = (m_blk - m_used) > A64_TAU;
          if (__any_sync(0xffffffffu, need)) {
            tc_fence_after();
            const float alpha = need ? exp2f(m_used - m_blk) : 1.f;
            if (need) m_used = m_blk;
            if (!USE_ONES) l_run *= alpha;
#pragma unroll 1
            for (int cc = 0; cc < p.npv; cc += 32) {
              uint32_t o[32];
              tmem_ld32(tOw + uint32_t(cc), o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tOw + uint32_t(cc), o);
            }
            exps(m_used);            // redo against the new maximum (the scores are still in registers)
          }
        }
        tmem_st_wait();
        if (!USE_ONES) l_run += rs;
        tc_fence_before();
        mbar_arrive(p_full(w));
      }
      cw += nblk;

      // ---- epilogue: O / l -> out[b][q][head*d + :]
      mbar_wait(pv_done(w), uint32_t((cw - 1) & 1));
      ++aw;
      tc_fence_after();
      const int q = q_first + w * A64_BM + r;
      const int b = bh / p.heads, head = bh - b * p.heads;
      uint32_t o0[32], o1[32];
      tmem_ld32(tOw, o0);
      if (p.npv > 32) tmem_ld32(tOw + 32u, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free(w));      // O is in registers: the next item's first P*V may overwrite it
      float inv_l;
      if (USE_ONES) {
        float l = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e == p.d) l = __uint_as_float(o0[e]);
          if (e + 32 == p.d) l = __uint_as_float(o1[e]);
        }
        inv_l = 1.f / l;
      } else {
        inv_l = 1.f / l_run;
      }
      if (q < p.nq) {
        act_t* orow = p.out + (static_cast<long long>(b) * p.nq + q) * (static_cast<long long>(p.heads) * p.d) +
                      static_cast<long long>(head) * p.d;
#pragma unroll
        for (int gq = 0; gq < 64; gq += 8) {
          if (gq < p.d) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(gq < 32 ? o0[gq + e] : o1[gq - 32 + e]) * inv_l;
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

// Host side: called by cb_attention for d <= 64 when there are enough query tiles to give every SM at least two.
int launch_attention64(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld, void* out,
                       int64_t batch, int64_t heads, int64_t nq, int64_t nk, int d, float scale, int nt, int num_sms,
                       cudaStream_t stream) {
  CUtensorMap mq, mk, mv;
  {
    uint32_t box[4] = {64, 128, 1, 1};
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)nq, (uint64_t)heads, (uint64_t)batch};
    uint64_t str[4] = {1, (uint64_t)q_ld, (uint64_t)d, (uint64_t)(nq * q_ld)};
    int rc = make_tmap_act(&mq, q, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint32_t box[4] = {64, 64, 1, 1};
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)nk, (uint64_t)heads, (uint64_t)batch};
    uint64_t strk[4] = {1, (uint64_t)k_ld, (uint64_t)d, (uint64_t)(nk * k_ld)};
    uint64_t strv[4] = {1, (uint64_t)v_ld, (uint64_t)d, (uint64_t)(nk * v_ld)};
    int rc = make_tmap_act(&mk, k, 4, dims, strk, box);
    if (rc) return rc;
    rc = make_tmap_act(&mv, v, 4, dims, strv, box);
    if (rc) return rc;
  }
  if (nt > A64_TILES) nt = A64_TILES;
  Attn64Params p{};
  const bool use_ones = d < 64;          // a spare column of the 64-wide V tile carries 1.0 -> O[:, d] = softmax row sum
  p.nq = (int)nq; p.nk = (int)nk; p.d = d; p.heads = (int)heads; p.bh = (int)(batch * heads); p.nt = nt;
  p.ksteps = (d + 15) / 16;
  p.npv = use_ones ? ((d + 1 + 15) / 16) * 16 : 64;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.idesc_qk = make_idesc_f16(128, A64_BN, 0, 0);
  p.idesc_pv = make_idesc_f16(128, p.npv, 0, 1);   // B = V is MN-major
  p.out = (act_t*)out;
  const size_t smem = (size_t)A64_TILES * A64_QBYTES + 2 * (size_t)A64_STAGES * A64_KVBYTES + 544;
  static DeviceOnce configured{};
  if (device_once_needed(configured)) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    device_once_done(configured);
  }
  const int rows_per_item = nt * A64_BM;
  const long long items = ((nq + rows_per_item - 1) / rows_per_item) * batch * heads;
  dim3 grid((unsigned)(items < num_sms ? items : num_sms));
  if (use_ones) (void)cb::launch_k(attention64_kernel<true>, dim3(grid), dim3(A64_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
  else (void)cb::launch_k(attention64_kernel<false>, dim3(grid), dim3(A64_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}

}  // namespace cb
