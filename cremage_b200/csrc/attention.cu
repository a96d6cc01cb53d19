// Fused flash-style attention for sm_100a:  O = softmax(scale * Q K^T) V, never materialising the score matrix.
//
//   S = Q K^T      tcgen05.mma  (A = Q tile [128 x dpad], B = K tile [128 x dpad], both K-major)  -> TMEM
//   online softmax one thread per query row: tcgen05.ld of S, exp2 (fp32 MUFU, one packing convert per pair), lazy
//                  running-max rescale of the TMEM accumulator, P packed to 16-bit pairs -> TMEM (tcgen05.st)
//   O += P V       tcgen05.mma  (A = P straight from TMEM, B = V tile in smem, MN-major)           -> TMEM
// P never touches shared memory: with d = 40 the P round trip (32 KB written + 32 KB read per tile and block) would
// make the kernel co-bound by shared-memory bandwidth next to the MUFU pipe, and the freed space deepens the K/V ring.
//
// Warp-specialised, one CTA = up to two 128-row query tiles of one (batch, head) sharing every K/V tile:
//   warp 0        TMA producer (Q once, K/V ring)
//   warp 1        TMEM allocator + single-thread MMA issuer of query tile 0      warp 2    MMA issuer of query tile 1
//   warp 3        ones column of V
//   warps 4..7    softmax warpgroup 0 (query tile 0)      warps 8..11   softmax warpgroup 1 (query tile 1)
// The four control warps are one ALIGNED warpgroup: they hand registers back (`setmaxnreg.dec`) and the softmax
// warpgroups take them (`setmaxnreg.inc`), so a row's 128 scores + 32 packed P + the temporaries of the exponential
// section fit without spills and ptxas has room to interleave.
// Decoupled issue: a warpgroup releases its S tile (s_free) the moment the scores are in registers, so its issuer
// thread (one per query tile, each with its own blocking wait sequence; K/V stages are released by both) issues
// Q*K^T of block j+1 while the warpgroup is still exponentiating block j; P*V of block j follows when P is written
// (p_full) and signals pv_done, which the warpgroup only consults before it overwrites P or rescales O.  In steady
// state a warpgroup never waits for the tensor core; the MUFU pipe (the bound for d = 40) is 75 % busy.
// The softmax row sum is a by-product of the tensor core when the padded head dim has a spare column (dpad > d): a
// helper warp writes 1.0 into column d of every V tile once its TMA load has landed (the pad columns arrive as
// zeros), so P*V accumulates sum_j P_ij into O[:, d] at no extra MMA.
// Q/K/V are read IN PLACE from the projection GEMM's row-major output ([tokens, (q|k|v) x heads x d], any row stride)
// through 4-D TMA boxes {d, tokens, heads, batch}: the box is 64 columns wide and the tensor map's inner extent is d,
// so TMA zero-fills the pad columns (d = 40 -> 64) -- no per-head padded copy of Q/K/V exists.
// Replaces the attention cores at ldm/modules/attention.py:418-423 (Doggettx), :646-657 (Original), :811 (xformers).
#include <cstdlib>

#include "common.cuh"
#include "cremage_b200.h"

namespace cb {

constexpr int ATT_BM = 128;             // query rows per warpgroup tile
constexpr int ATT_BN = 128;             // kv rows per iteration
constexpr int PANEL_BYTES = 128 * 128;  // [128 rows][64 x 16-bit]
constexpr int ATT_THREADS = 384;         // control warpgroup (TMA, two issuers, V ones column) + 8 softmax warps
// registers per softmax thread after the hand-over (template parameter REGS; the control warps keep (64512 - 256 REGS) / 128):
// measured per head-dim class, because ptxas schedules the exponential section differently with every budget --
// d 40: 168 (no hand-over) 0.692 ms, 176 0.682, 184 0.693, 192 0.688, 200 0.760, 224 0.808; d 80: 0.082, 0.081, 0.074, 0.074, 0.075, 0.075
constexpr int ATT_REGS_D64 = 176, ATT_REGS_WIDE = 192;
constexpr float RESCALE_TAU = 8.0f;     // rescale O only when the row max grew by more than 2^8 (P <= 256)

struct AttnParams {
  int nq, nk, d, dpad, np, heads, bh, stages, nwg, use_ones, p_alias, pingpong, ksteps, split_from, total_items;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv, tmem_cols;
  act_t* out;
};

// p = 2^(s * scale - m) for two scores -> packed 16-bit pair.  fp32 MUFU.EX2 then one packing convert: the packed
// `ex2.approx.f16x2` form lowers to two MUFU.EX2.F16 plus a PRMT (same MUFU count, one more ALU instruction per pair).
CB_DEVINL uint32_t exp2_pack(float s0, float s1, float scale, float m) {
  const float a0 = fmaf(s0, scale, -m), a1 = fmaf(s1, scale, -m);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  return pack_act2(e0, e1);
}

template <bool USE_ONES, int REGS>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();  // swizzled tiles need a 1024-byte aligned base
  const uint32_t tile_bytes = uint32_t(p.np) * PANEL_BYTES;
  const uint32_t sQ = base;                                    // [nwg]
  const uint32_t sK = sQ + uint32_t(p.nwg) * tile_bytes;       // [stages]
  const uint32_t sV = sK + uint32_t(p.stages) * tile_bytes;    // [stages]
  const uint32_t bars = sV + uint32_t(p.stages) * tile_bytes;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u + 8u * uint32_t(s); };
  auto k_empty = [&](int s) { return bars + 40u + 8u * uint32_t(s); };
  auto v_full = [&](int s) { return bars + 72u + 8u * uint32_t(s); };
  auto v_empty = [&](int s) { return bars + 104u + 8u * uint32_t(s); };
  auto s_full = [&](int w) { return bars + 136u + 8u * uint32_t(w); };
  auto p_full = [&](int w) { return bars + 152u + 8u * uint32_t(w); };
  auto s_free = [&](int w) { return bars + 168u + 8u * uint32_t(w); };
  auto pv_done = [&](int w) { return bars + 184u + 8u * uint32_t(w); };
  auto v_ready = [&](int s) { return bars + 200u + 8u * uint32_t(s); };   // V tile landed AND its ones column written
  const uint32_t q_free = bars + 232u;                                     // every Q*K^T of the item has been issued and completed
  auto o_free = [&](int w) { return bars + 240u + 8u * uint32_t(w); };     // the item's O has been read out of TMEM
  const uint32_t tmem_slot = bars + 256u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = (p.nk + ATT_BN - 1) / ATT_BN;
  // persistent CTA: work items (batch*head, query-tile pair) blockIdx.x, + gridDim.x, ...  Barrier phases and the
  // K/V ring position run on GLOBAL counters across items, so the producer prefetches the next item's Q / K / V
  // while the softmax groups still finish the current one, and the per-CTA set-up is paid once.
  // Items below `split_from` are pairs of query tiles; the pairs that would form a last, mostly empty round over the
  // CTAs are handed out as single tiles instead (item >= split_from: pair split_from + (item - split_from) / 2, tile
  // (item - split_from) % 2), so that round costs one tile's time on twice as many SMs (cb_attention decides).
  const int rows_per_item = p.nwg * ATT_BM;
  const int qpairs = (p.nq + rows_per_item - 1) / rows_per_item;
  const int total_items = p.total_items;
  auto item_pair = [&](int item) { return item < p.split_from ? item : p.split_from + ((item - p.split_from) >> 1); };
  auto item_q_first = [&](int item) {
    return (item_pair(item) % qpairs) * rows_per_item + (item < p.split_from ? 0 : ((item - p.split_from) & 1) * ATT_BM);
  };
  auto item_bh = [&](int item) { return item_pair(item) / qpairs; };
  auto item_nact = [&](int item) {
    return (p.nwg == 2 && item < p.split_from && item_q_first(item) + ATT_BM < p.nq) ? 2 : 1;
  };

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    mbar_init(q_free, uint32_t(p.nwg));
    for (int s = 0; s < 4; ++s) {   // a stage is free once BOTH issuers are done with it (an idle one just arrives)
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), uint32_t(p.nwg));
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), uint32_t(p.nwg));
    }
    for (int s = 0; s < 4; ++s) mbar_init(v_ready(s), 32);
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full(s), 1); mbar_init(p_full(s), ATT_BM);
      mbar_init(s_free(s), ATT_BM); mbar_init(pv_done(s), 1);
      mbar_init(o_free(s), ATT_BM);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // the set-up above overlaps the previous kernel; global memory is only touched from here on
  auto tS = [&](int w) { return tmem_base + uint32_t(w) * 128u; };
  auto tO = [&](int w) { return tmem_base + uint32_t(p.nwg) * 128u + uint32_t(w) * uint32_t(p.dpad); };
  // P: 128 rows x 128 16-bit values = 64 columns; its own region when TMEM has room, else the first half of S
  auto tP = [&](int w) {
    return p.p_alias ? tS(w) : tmem_base + uint32_t(p.nwg) * (128u + uint32_t(p.dpad)) + uint32_t(w) * 64u;
  };
  // position in the K/V ring, advanced block by block (no `% stages` on the issue paths)
  struct Ring {
    int st; uint32_t ph;
    __device__ __forceinline__ void next(int stages) { if (++st == stages) { st = 0; ph ^= 1u; } }
  };

  if (warp < 4) {
  // (the register hand-over sits INSIDE the role branches: behind a merge of the two paths ptxas has to assume the
  // smaller budget for everything that follows)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"((64512 - 256 * REGS) / 128));
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // (not `lane == 0`: a divergent branch wraps every UTMALDG / UTCHMMA in an ELECT / BRA.U.ANY loop)
      int it = 0;
      Ring r{0, 0u};
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int bh = item_bh(item), q_first = item_q_first(item), nact = item_nact(item);
        const int b_idx = bh / p.heads, h_idx = bh - b_idx * p.heads;
        mbar_wait(q_free, uint32_t(it & 1) ^ 1u);     // previous item's Q*K^T are complete (first item passes)
        mbar_expect_tx(q_full, uint32_t(nact) * tile_bytes);
        for (int w = 0; w < nact; ++w)
          for (int pn = 0; pn < p.np; ++pn)
            tma_load_4d(sQ + uint32_t(w) * tile_bytes + uint32_t(pn) * PANEL_BYTES, &mapQ, q_full, pn * 64,
                        q_first + w * ATT_BM, h_idx, b_idx);
        for (int j = 0; j < nblk; ++j, r.next(p.stages)) {
          const int st = r.st;
          const uint32_t ph = r.ph;
          mbar_wait(k_empty(st), ph ^ 1u);
          mbar_expect_tx(k_full(st), tile_bytes);
          for (int pn = 0; pn < p.np; ++pn)
            tma_load_4d(sK + uint32_t(st) * tile_bytes + uint32_t(pn) * PANEL_BYTES, &mapK, k_full(st), pn * 64, j * ATT_BN,
                        h_idx, b_idx);
          mbar_wait(v_empty(st), ph ^ 1u);
          mbar_expect_tx(v_full(st), tile_bytes);
          for (int pn = 0; pn < p.np; ++pn)
            tma_load_4d(sV + uint32_t(st) * tile_bytes + uint32_t(pn) * PANEL_BYTES, &mapV, v_full(st), pn * 64, j * ATT_BN,
                        h_idx, b_idx);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (one thread per query tile) =====================
    const int w = warp - 1;
    if (w < p.nwg && elect_one()) {
      // descriptors are 64-bit adds on precomputed bases (start address field = bytes >> 4)
      const uint64_t qd0 = make_sdesc_sw128(sQ + uint32_t(w) * tile_bytes, 16, 1024), kd0 = make_sdesc_sw128(sK, 16, 1024);
      const uint64_t vd0 = make_sdesc_sw128(sV, PANEL_BYTES, 1024);
      const uint32_t ts = tS(w), tp = tP(w), to = tO(w);
      const int ksteps = p.ksteps;   // ceil(d / 16): the pad columns are zeros, no K-step beyond the next multiple of 16
      auto issue_qk = [&](int kstage) {
        const uint64_t kd = kd0 + uint64_t(kstage) * (tile_bytes >> 4);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t off = uint64_t(ks >> 2) * (PANEL_BYTES >> 4) + uint64_t(ks & 3) * 2u;
          umma_bf16(ts, qd0 + off, kd + off, p.idesc_qk, ks != 0);
        }
      };
      auto issue_pv = [&](int vstage, bool accumulate) {
        const uint64_t vd = vd0 + uint64_t(vstage) * (tile_bytes >> 4);
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks) {
          // V: 16 kv rows = 2 atoms = 2048 bytes; A = P from TMEM: 16 K-values of a row = 8 columns
          umma_ts(to, tp + uint32_t(ks) * 8u, vd + uint64_t(ks) * (2048u >> 4), p.idesc_pv, (accumulate || ks != 0) ? 1u : 0u);
        }
      };
      int it = 0;              // item count of this CTA
      Ring rq{0, 0u}, rv{0, 0u};   // ring positions of the next Q*K^T / the next P*V (both advance nblk per item)
      int cw = 0, aw = 0;      // blocks / items this query tile has been ACTIVE for (phases of its private barriers)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const bool active = w < item_nact(item);
        mbar_wait(q_full, uint32_t(it & 1));
        if (!active) {
          // idle tile (ragged last query pair): stay in lockstep with the ring, release every stage it is handed
          mbar_arrive(q_free);
          for (int j = 0; j < nblk; ++j, rq.next(p.stages), rv.next(p.stages)) {
            mbar_wait(k_full(rq.st), rq.ph);
            mbar_arrive(k_empty(rq.st));
            mbar_wait(v_ready(rv.st), rv.ph);
            mbar_arrive(v_empty(rv.st));
          }
          continue;
        }
        auto do_qk = [&](int j) {     // Q*K^T of block j: once the scores of this tile's previous block are in registers
          if (cw + j > 0) mbar_wait(s_free(w), uint32_t((cw + j - 1) & 1));
          mbar_wait(k_full(rq.st), rq.ph);
          tc_fence_after();
          issue_qk(rq.st);
          umma_commit(s_full(w));
          umma_commit(k_empty(rq.st));
          if (j == nblk - 1) umma_commit(q_free);      // the item's last use of Q
          rq.next(p.stages);
        };
        auto do_pv = [&](int j) {
          mbar_wait(p_full(w), uint32_t((cw + j) & 1));
          mbar_wait(v_ready(rv.st), rv.ph);
          if (j == 0 && aw > 0) mbar_wait(o_free(w), uint32_t((aw - 1) & 1));   // previous item's O has been read out
          tc_fence_after();
          issue_pv(rv.st, j > 0);
          umma_commit(pv_done(w));
          umma_commit(v_empty(rv.st));
          rv.next(p.stages);
        };
        do_qk(0);
        for (int j = 0; j < nblk; ++j) {
          if (p.p_alias) {            // P overwrites S: the next Q*K^T may only follow this block's P*V
            do_pv(j);
            if (j + 1 < nblk) do_qk(j + 1);
          } else {
            if (j + 1 < nblk) do_qk(j + 1);
            do_pv(j);
          }
        }
        cw += nblk;
        ++aw;
      }
    }
  } else if (warp == 3) {
    // ===================== V ones-column warp =====================
    // TMA zero-fills the pad columns of a V tile; column d becomes 1.0 so that P*V also yields the softmax row sum.
    const int pn = p.d >> 6, cw = p.d & 63;
    Ring rg{0, 0u};
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      for (int j = 0; j < nblk; ++j, rg.next(p.stages)) {
        const int st = rg.st;
        mbar_wait(v_full(st), rg.ph);
        if (USE_ONES) {
          const uint32_t tile = sV + uint32_t(st) * tile_bytes + uint32_t(pn) * PANEL_BYTES;
#ifdef CB_FP16
          const unsigned short one_bits = 0x3C00;   // fp16 1.0
#else
          const unsigned short one_bits = 0x3F80;   // bf16 1.0
#endif
          for (int r = lane; r < ATT_BN; r += 32) {
            const uint32_t addr = tile + uint32_t(r) * 128u + (((uint32_t(cw) >> 3) ^ (uint32_t(r) & 7u)) << 4) + uint32_t(cw & 7) * 2u;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(one_bits) : "memory");
          }
          fence_proxy_async_smem();
        }
        mbar_arrive(v_ready(st));
      }
    }
  }
  } else {
    // ===================== softmax warpgroups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
    const int w = (warp - 4) >> 2;
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t tSw = tS(w) + lane_off, tOw = tO(w) + lane_off, tPw = tP(w) + lane_off;
    int cw = 0;                                        // blocks this tile has been active for (barrier phases)
    // Ping-pong of the MUFU sections.  The two warpgroups share each scheduler's MUFU pipe, one warp alone drives 90 %
    // of it (tools/micro/mufu_rate.cu), and a block's MUFU-free part (S load, row maximum, P store, barriers) is about
    // as long as its 128 exponentials -- but left alone the two warpgroups drift into the same phase and then both
    // wait (profiles/r2_attention64.md).  Two named barriers hand the pipe back and forth: a warpgroup enters its
    // exponentials only when the other one has left them.  An idle tile (ragged last item) still passes the token.
    const bool pingpong = p.nwg == 2 && p.pingpong;
    const int bar_mine = 2 + w, bar_other = 2 + (w ^ 1);
    auto pp_wait = [&]() { if (pingpong) asm volatile("bar.sync %0, 256;" ::"r"(bar_mine) : "memory"); };
    auto pp_pass = [&]() { if (pingpong) asm volatile("bar.arrive %0, 256;" ::"r"(bar_other) : "memory"); };
    if (w == 1) pp_pass();                             // warpgroup 0 goes first
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      if (w >= item_nact(item)) {
        for (int j = 0; j < nblk; ++j) { pp_wait(); pp_pass(); }
        continue;
      }
      const int bh = item_bh(item), q_first = item_q_first(item);
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int c = cw + j;
        mbar_wait(s_full(w), uint32_t(c & 1));
        tc_fence_after();
        uint32_t s[128];
        tmem_ld32(tSw + 0u, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld32(tSw + 32u, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld32(tSw + 64u, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
        tmem_ld32(tSw + 96u, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free(w));      // the scores are in registers: Q*K^T of the next block may overwrite S
        const int nvalid = p.nk - j * ATT_BN;
        if (nvalid < ATT_BN) {   // ragged last block: K rows beyond nk were zero filled -> mask
#pragma unroll
          for (int e = 0; e < 128; ++e)
            if (e >= nvalid) s[e] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 128; e += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[e]));
          mx1 = fmaxf(mx1, __uint_as_float(s[e + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[e + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[e + 3]));
        }
        const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
        if (j == 0) {
          m_used = m_blk;
        } else {
          const bool need = (m_blk - m_used) > RESCALE_TAU;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(pv_done(w), uint32_t((c - 1) & 1));   // O holds every P*V up to block j-1
            tc_fence_after();
            const float alpha = need ? exp2f(m_used - m_blk) : 1.f;
            if (need) m_used = m_blk;
            if (!USE_ONES) l_run *= alpha;
#pragma unroll 1
            for (int cc = 0; cc < p.dpad; cc += 32) {
              uint32_t o[32];
              tmem_ld32(tOw + uint32_t(cc), o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tOw + uint32_t(cc), o);
            }
            tmem_st_wait();
          }
        }
        float rs = 0.f;
        // exponent arguments in place (FMA pipe), and the wait for the previous P*V (it read P) BEFORE the MUFU section:
        // the section below is then straight-line code -- a spin loop in its middle is a basic-block boundary that
        // kept the second 64 exponentials behind the pack / store of the first 64, with the MUFU pipe idle in between
#pragma unroll
        for (int e = 0; e < ATT_BN; ++e) s[e] = __float_as_uint(fmaf(__uint_as_float(s[e]), p.scale_log2, -m_used));
        if (c > 0) mbar_wait(pv_done(w), uint32_t((c - 1) & 1));
        pp_wait();
#pragma unroll
        for (int cc = 0; cc < ATT_BN; cc += 64) {
          uint32_t pk[32];
#pragma unroll
          for (int e = 0; e < 64; e += 2) {
            float e0, e1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(__uint_as_float(s[cc + e])));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(__uint_as_float(s[cc + e + 1])));
            pk[e >> 1] = pack_act2(e0, e1);
            if (!USE_ONES) rs += e0 + e1;   // the fp32 exponentials (no unpack of the rounded pair: two instructions less per score pair)
          }
          tmem_st32(tPw + uint32_t(cc >> 1), pk);
        }
        pp_pass();
        tmem_st_wait();
        if (!USE_ONES) l_run += rs;
        tc_fence_before();
        mbar_arrive(p_full(w));
      }
      cw += nblk;

      // ---- epilogue: O / l -> out[b][q][head*d + :]
      mbar_wait(pv_done(w), uint32_t((cw - 1) & 1));
      tc_fence_after();
      const int q = q_first + w * ATT_BM + r;
      const int b = bh / p.heads, head = bh - b * p.heads;
      float inv_l;
      if (USE_ONES) {
        uint32_t o[32];
        tmem_ld32(tOw + uint32_t(p.d & ~31), o);
        tmem_ld_wait();
        float l = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e == (p.d & 31)) l = __uint_as_float(o[e]);
        inv_l = 1.f / l;
      } else {
        inv_l = 1.f / l_run;
      }
      act_t* orow = p.out + (static_cast<long long>(b) * p.nq + q) * (static_cast<long long>(p.heads) * p.d) +
                    static_cast<long long>(head) * p.d;
#pragma unroll 1
      for (int cc = 0; cc < p.d; cc += 32) {
        uint32_t o[32];
        tmem_ld32(tOw + uint32_t(cc), o);
        tmem_ld_wait();
        if (cc + 32 >= p.d) {        // last TMEM read of this item's O: the next item's first P*V may overwrite it
          tc_fence_before();
          mbar_arrive(o_free(w));
        }
        if (q < p.nq) {
#pragma unroll
          for (int gq = 0; gq < 32; gq += 8) {
            if (cc + gq < p.d) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[gq + e]) * inv_l;
              *reinterpret_cast<uint4*>(orow + cc + gq) = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]),
                                                                     pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
            }
          }
        }
      }
    }
    if (w == 0) pp_wait();   // absorbs warpgroup 1's last token so that no barrier is left half-arrived
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace cb

using namespace cb;

namespace cb {
int launch_attention64(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld, void* out,
                       int64_t batch, int64_t heads, int64_t nq, int64_t nk, int d, float scale, int nt, int num_sms,
                       cudaStream_t stream);   // attention64.cu
}

extern "C" int cb_attention(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld,
                            void* out, int64_t batch, int64_t heads, int64_t nq, int64_t nk, int d, float scale,
                            cudaStream_t stream) {
  CB_REQUIRE(q && k && v && out, "cb_attention: null pointer");
  CB_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0, "cb_attention: empty problem");
  CB_REQUIRE(d > 0 && d % 8 == 0 && d <= 192, "cb_attention: head dim %d unsupported (multiple of 8, <= 192)", d);
  CB_REQUIRE(q_ld >= heads * d && k_ld >= heads * d && v_ld >= heads * d && q_ld % 8 == 0 && k_ld % 8 == 0 && v_ld % 8 == 0,
             "cb_attention: row strides must be >= heads * d and multiples of 8 elements");
  const int dpad = (d + 63) / 64 * 64;
  const int64_t bh = batch * heads;
  CB_REQUIRE(bh * ((nq + 127) / 128) < (1LL << 30), "cb_attention: problem too large");
  const int num_sms = sm_count();
  if (dpad == 64) {
    // head dims <= 64 with SHORT key sequences (cross-attention, nk = 77 * k): three query tiles per CTA over 64-row kv
    // blocks (attention64.cu) -- 0.058 ms against 0.089 ms for bh 128, 4096 x 77, d 40.  For long key sequences the
    // two-warpgroup kernel below (128-row kv blocks, half as many barrier round trips per exponential) is still the
    // faster one (0.80 ms against 0.88 ms at 4096 x 4096; profiles/r2_attention64.md).  CB_ATTN64=0 / =2: never / always.
    static int use64 = -1;
    if (use64 < 0) {
      const char* e = getenv("CB_ATTN64");
      use64 = e ? atoi(e) : 1;
    }
    const long long qtiles = (nq + 127) / 128, tiles = bh * qtiles;
    // two or three query tiles per CTA: rounds over the SMs x (per-item chain ~6 + one unit of MUFU work per tile) --
    // SDXL's 1024 x 77 level (40 x 8 tiles) is 160 items = two rounds with two tiles, 120 items = one round with three
    int nt = (int)(tiles / num_sms < 3 ? tiles / num_sms : 3);
    if (nt == 2) {
      auto cost = [&](int t) { const long long items = bh * ((qtiles + t - 1) / t); return ((items + num_sms - 1) / num_sms) * (6 + t); };
      if (cost(3) < cost(2)) nt = 3;
    }
    if (nt >= 2 && (use64 == 2 || (use64 == 1 && nk <= 256)))
      return launch_attention64(q, q_ld, k, k_ld, v, v_ld, out, batch, heads, nq, nk, d, scale, nt, num_sms, stream);
  }
  CUtensorMap mq, mk, mv;
  uint32_t box[4] = {64, 128, 1, 1};
  {
    // element (j, token, head, b) = base[(b * tokens + token) * ld + head * d + j]; inner extent d < box 64 -> zero fill
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
  AttnParams p{};
  p.nq = (int)nq; p.nk = (int)nk; p.d = d; p.dpad = dpad; p.np = dpad / 64; p.heads = (int)heads; p.bh = (int)bh;
  p.nwg = dpad <= 128 ? 2 : 1;           // TMEM: nwg * (128 + dpad) columns <= 512
  p.stages = dpad <= 64 ? 4 : (dpad <= 128 ? 2 : 1);   // smem: (nwg + 2*stages) * np panels of 16 KB
  p.p_alias = (p.nwg * (128 + dpad + 64) > 512) ? 1 : 0;   // no room for a separate P region: P overwrites S
  p.use_ones = dpad > d;                 // spare column d of V carries 1.0 -> O[:, d] = softmax row sum
  p.scale_log2 = scale * 1.4426950408889634f;
  p.idesc_qk = make_idesc_f16(128, 128, 0, 0);
  p.ksteps = (d + 15) / 16;
  // B = V is MN-major; N stops at the next multiple of 16 behind the ones column (d 40: 48 instead of 64 columns)
  p.idesc_pv = make_idesc_f16(128, p.use_ones ? ((d + 1 + 15) / 16) * 16 : dpad, 0, 1);
  p.tmem_cols = 512u;
  p.out = (act_t*)out;
  {
    // measured (profiles/r2_attention64.md): helps the 128-column head dims (d 80: 0.090 -> 0.086 ms), costs at d <= 64
    // (d 40: 0.684 -> 0.718 ms), where the hand-over latency outweighs the overlap it enforces
    static int pp = -2;
    if (pp == -2) { const char* e = getenv("CB_ATTN_PINGPONG"); pp = e ? atoi(e) : -1; }
    p.pingpong = pp >= 0 ? pp : (dpad > 64 ? 1 : 0);
  }
  const size_t smem = (size_t)(p.nwg + 2 * p.stages) * p.np * PANEL_BYTES + 384;
  CB_REQUIRE(smem <= 227 * 1024, "cb_attention: needs %zu bytes of shared memory", smem);
  static DeviceOnce configured{};
  if (device_once_needed(configured)) {
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<true, ATT_REGS_D64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<false, ATT_REGS_D64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<true, ATT_REGS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<false, ATT_REGS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // the register hand-over is balanced for a launch at 168 registers per thread (384 x 168 = 128 x control + 256 x
    // REGS): with any other count the increase would wait for registers nobody releases -- fail loudly instead
    cudaFuncAttributes fa{};
    CB_CHECK_CUDA(cudaFuncGetAttributes(&fa, attention_kernel<true, ATT_REGS_D64>));
    CB_REQUIRE(fa.numRegs == 168, "cb_attention: attention_kernel was compiled to %d registers per thread, the setmaxnreg split assumes 168", fa.numRegs);
    CB_CHECK_CUDA(cudaFuncGetAttributes(&fa, attention_kernel<true, ATT_REGS_WIDE>));
    CB_REQUIRE(fa.numRegs == 168, "cb_attention: attention_kernel was compiled to %d registers per thread, the setmaxnreg split assumes 168", fa.numRegs);
    CB_CHECK_CUDA(cudaFuncGetAttributes(&fa, attention_kernel<false, ATT_REGS_D64>));
    CB_REQUIRE(fa.numRegs == 168, "cb_attention: attention_kernel was compiled to %d registers per thread, the setmaxnreg split assumes 168", fa.numRegs);
    CB_CHECK_CUDA(cudaFuncGetAttributes(&fa, attention_kernel<false, ATT_REGS_WIDE>));
    CB_REQUIRE(fa.numRegs == 168, "cb_attention: attention_kernel was compiled to %d registers per thread, the setmaxnreg split assumes 168", fa.numRegs);
    device_once_done(configured);
  }
  const int rows_per_item = p.nwg * ATT_BM;
  const long long items = ((nq + rows_per_item - 1) / rows_per_item) * bh;
  dim3 grid((unsigned)(items < num_sms ? items : num_sms));   // persistent: one CTA per SM walks the work items
  // The last round over the CTAs: `left` pairs on `grid` CTAs.  Handed out as 2 * left single tiles it costs one
  // tile's time (0.73 of a pair's: 2 376 against 3 250 clocks per kv block, profiles/r2_attention64.md) when they
  // fit one round -- SDXL's 1024-token level is 160 pairs on 148 SMs: 1 + 0.73 instead of 2 rounds.
  p.split_from = (int)items;
  p.total_items = (int)items;
  {
    static int tail = -1;
    if (tail < 0) { const char* e = getenv("CB_ATTN_TAIL_SPLIT"); tail = e ? atoi(e) : 1; }
    const long long left = items % (long long)grid.x;
    if (tail && p.nwg == 2 && nq % rows_per_item == 0 && left > 0 && 2 * left <= (long long)grid.x) {
      p.split_from = (int)(items - left);
      p.total_items = (int)(items + left);
    }
  }
  if (dpad == 64) {
    if (p.use_ones) (void)cb::launch_k(attention_kernel<true, ATT_REGS_D64>, dim3(grid), dim3(ATT_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
    else (void)cb::launch_k(attention_kernel<false, ATT_REGS_D64>, dim3(grid), dim3(ATT_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
  } else {
    if (p.use_ones) (void)cb::launch_k(attention_kernel<true, ATT_REGS_WIDE>, dim3(grid), dim3(ATT_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
    else (void)cb::launch_k(attention_kernel<false, ATT_REGS_WIDE>, dim3(grid), dim3(ATT_THREADS), (size_t)(smem), stream, mq, mk, mv, p);
  }
  CB_CHECK_CUDA(cudaGetLastError());
  CB_LAUNCHED(1);
  return CB_OK;
}
