// Tiling policy of the implicit GEMM, inside the library: a host that is not the Python mirror (or a reference-side
// binding) fills a cb_igemm_desc with the PROBLEM only and lets cb_igemm_plan / cb_igemm_auto choose the
// 128-row pixel tile, the N tile, CTA pairs, the dual-N schedule and split-K.  The choices are the ones measured in
// round 1 (profiles/r1_per_shape_timings_v3.txt, r1_igemm_pair_ncu_summary.md); the knobs are environment variables.
#include <cstdlib>

#include "common.cuh"
#include "cremage_b200.h"

namespace cb {
namespace {

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

int num_sms() { return sm_count(); }

long long cdiv(long long a, long long b) { return (a + b - 1) / b; }
int pow2_ceil(long long x) { int p = 1; while (p < x) p <<= 1; return p; }

// 128-row tile {tw, th, tn} over an (n, h, w) pixel grid; overhang is masked by the kernel
void choose_tile(long long n, long long h, long long w, int* tw, int* th, int* tn) {
  if (h == 1 && n == 1) { *tw = 128; *th = 1; *tn = 1; return; }
  long long low = w & -w;                         // largest power of two dividing w
  *tw = int(low < 128 ? low : 128);
  const int hp = pow2_ceil(h);
  *th = (128 / *tw) < hp ? (128 / *tw) : hp;
  *tn = 128 / (*tw * *th);
}

// N tile of the persistent single-CTA kernel: minimise rounds x per-tile cost (about bn MMA columns plus a fixed
// prologue / epilogue share); ties go to the wider tile (fewer A re-reads)
int choose_bn(long long cols, long long m_tiles, int multiple, int sms) {
  const int cand[4] = {160, 128, 64, 32};
  long long best = -1; int best_bn = 128;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (bn % multiple) continue;
    const long long tiles = m_tiles * cdiv(cols, bn);
    const long long cost = cdiv(tiles, sms) * (bn + 24);
    if (best < 0 || cost < best) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

// (bn, nsub, ksplit) in CTA-pair mode (sms / 2 clusters, each a 256-row tile).  nsub = 2: two N tiles share every A
// stage (3 * bn <= 512 TMEM columns).  ksplit = 3: K split by tap groups over otherwise idle clusters.  A round costs
// about the MMA columns of the tile group; tiles narrower than 256 columns are shared-memory-fill bound (x 1.25), below
// 128 columns the single issuing thread cannot keep the tensor core fed; a split adds a per-item and a reduce cost.
void choose_bn_pair(long long cols, long long m_tiles, int multiple, bool allow_split, int sms, int* bn_out, int* nsub_out,
                    int* ks_out) {
  const int cand[5] = {256, 160, 128, 64, 32};
  const long long m_pairs = (m_tiles + 1) / 2;
  double best = -1.0;
  for (int i = 0; i < 5; ++i) {
    const int bn = cand[i];
    if (bn % multiple) continue;
    const long long n_tiles = cdiv(cols, bn);
    for (int nsub = 2; nsub >= 1; --nsub) {
      if (nsub == 2 && !(3 * bn <= 512 && n_tiles >= 2)) continue;
      const long long groups = cdiv(n_tiles, nsub);
      const int width = nsub * bn;
      const double per = (nsub * (bn > 128 ? bn : 128) + 16) * (width >= 256 ? 1.0 : 1.25);
      for (int ks = 1; ks <= (allow_split ? 3 : 1); ks += 2) {
        const long long items = m_pairs * groups * ks;
        const double cost = double(cdiv(items, sms / 2)) * (per / ks + (ks > 1 ? 24 : 0)) + (ks > 1 ? 30 : 0);
        if (best < 0 || cost < best) { best = cost; *bn_out = bn; *nsub_out = nsub; *ks_out = ks; }
      }
    }
  }
}

// split-K factor (1, 3 or 9 tap groups) of a single-CTA 3x3 conv launch with fewer output tiles than SMs
int choose_ksplit_single(long long tiles, int num_k, int sms) {
  if (tiles >= sms) return 1;
  int best_ks = 1; double best = -1.0;
  const int cand[3] = {1, 3, 9};
  for (int i = 0; i < 3; ++i) {
    const int ks = cand[i];
    const double cost = double(cdiv(tiles * ks, sms)) * (double(num_k) / ks) + (ks > 1 ? 8 * ks : 0);
    if (best < 0 || cost < best) { best = cost; best_ks = ks; }
  }
  return best_ks;
}

// (bn, ksplit) of a single-CTA launch with fewer output tiles than SMs (small batch, the 8x8 / 16x16 levels).  Such a
// launch is bound by how fast each CTA can fill its shared memory -- about 100 GB/s per SM with the ring this kernel
// has room for (tools/micro/smem_fill_rate.cu, profiles/r2_small_batch.md) -- so the model is bytes per work item
// (K blocks x (16 KB of A + 128 B x bn of B)) times rounds over the SMs, plus the reduce pass of a split.  Times in us.
void choose_single_small(long long m_tiles, long long ncols, int num_k, long long rows, int multiple, bool can_split, int sms,
                         int* bn_out, int* ks_out) {
  const int cand[5] = {256, 160, 128, 64, 32};
  const int splits[3] = {1, 3, 9};
  double best = -1.0;
  for (int i = 0; i < 5; ++i) {
    const int bn = cand[i];
    if (bn % multiple) continue;
    for (int j = 0; j < (can_split ? 3 : 1); ++j) {
      const int ks = splits[j];
      const long long items = m_tiles * cdiv(ncols, bn) * ks;
      const double fill = double(cdiv(items, sms)) * (double(num_k) / ks) * (16384.0 + 128.0 * bn) / 100e3;
      const double reduce = ks > 1 ? 3.0 + double(ks) * double(rows) * double(ncols) * 8.0 / 3e6 : 0.0;
      const double cost = fill + reduce + 0.05 * cdiv(ncols, bn);   // ties: the wider tile
      if (best < 0 || cost < best) { best = cost; *bn_out = bn; *ks_out = ks; }
    }
  }
}

}  // namespace
}  // namespace cb

using namespace cb;

extern "C" int cb_igemm_plan(const cb_igemm_desc* d, cb_igemm_plan_t* plan) {
  CB_REQUIRE(d != nullptr && plan != nullptr, "cb_igemm_plan: null argument");
  CB_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->cout > 0 && d->taps >= 1 && d->taps <= 9 && d->c0 > 0,
             "cb_igemm_plan: empty or malformed problem");
  const int sms = num_sms();
  static const int pair_min_k = env_int("CB_PAIR_MIN_K_CHUNKS", 18);   // CTA pairs for K >= 1152 (the MMA-bound launches)
  static const int geglu_pair = env_int("CB_GEGLU_PAIR", 1);
  static const int gn_fuse = env_int("CB_GN_FUSE", 1);
  static const int gn_min_k = env_int("CB_GN_FUSE_MIN_K_CHUNKS", 1);
  static const long long gn_min_bytes = env_int("CB_GN_FUSE_MIN_BYTES", 0);
  static const int pair_default = env_int("CB_PAIR", 1), split_default = env_int("CB_SPLITK", 1);
  const int pair_min_m_tiles = 8;

  int tw = d->tw, th = d->th, tn = d->tn;
  if (tw <= 0 || th <= 0 || tn <= 0) choose_tile(d->n, d->h, d->w, &tw, &th, &tn);
  const long long m_tiles = cdiv(d->w, tw) * cdiv(d->h, th) * cdiv(d->n, tn);
  const long long rows = d->n * d->h * d->w;
  const long long ncols = d->mode == CB_EPI_GEGLU ? 2 * d->cout : d->cout;
  const int num_k = d->taps * int(cdiv(d->c0, 64) + cdiv(d->c1, 64));
  const float oscale = d->out_scale == 0.f ? 1.f : d->out_scale;
  const bool strided = d->out_w_stride > 0 || d->out_h_stride > 0 || d->out_n_stride > 0;
  const long long out_ld = d->out_ld > 0 ? d->out_ld : d->cout;

  bool pair;
  if (d->cta_pair > 0) pair = true;
  else if (d->cta_pair < 0) pair = false;
  else if (d->mode == CB_EPI_GEGLU) pair = geglu_pair && m_tiles >= pair_min_m_tiles;
  else pair = pair_default && num_k >= pair_min_k && m_tiles >= pair_min_m_tiles;
  // small-K linears with a wide output (q|k|v projections, 640-wide to_out / proj): the launch is bound by the epilogue
  // and by re-reading A once per N tile, both of which a 256-row x 256-column pair tile halves (tools/sweep_plan.py:
  // 320 -> 960 at 65 536 rows 56 -> 46 us, 640 -> 1920 at 16 384 rows 46 -> 36 us; 320 -> 320 stays single)
  static const int wide_pair = env_int("CB_PAIR_WIDE_LINEAR", 1);
  bool wide_linear = false;
  if (!pair && d->cta_pair == 0 && wide_pair && pair_default && d->taps == 1 && d->mode != CB_EPI_GEGLU && ncols >= 640 &&
      m_tiles >= pair_min_m_tiles && m_tiles * cdiv(ncols, 256) >= sms) {
    pair = true;
    wide_linear = true;
  }

  // split-K (by tap groups) is available to plain 16-bit 3x3 convs; cb_splitk_reduce applies bias / row bias / residual
  bool can_split = d->mode == CB_EPI_LINEAR && d->taps == 9 && !d->out_f32 && d->act == CB_ACT_NONE && oscale == 1.f &&
                   d->cout % 8 == 0 && out_ld % 8 == 0;
  int bn = d->bn, nsub = d->nsub, ksplit = d->ksplit;
  const int multiple = d->mode == CB_EPI_GEGLU ? 64 : 32;
  if (bn <= 0) {
    if (pair && wide_linear) {
      // one 256-column tile per pair unless that pads N by more than a tenth (then two 160-column sub-tiles per A stage)
      if ((cdiv(ncols, 256) * 256 - ncols) * 10 <= ncols) { bn = 256; if (nsub <= 0) nsub = 1; }
      else { bn = 160; if (nsub <= 0) nsub = 2; }
      if (ksplit <= 0) ksplit = 1;
    } else if (pair) {
      int a_bn = 128, a_nsub = 1, a_ks = 1;
      choose_bn_pair(ncols, m_tiles, multiple, can_split && ksplit <= 0 && split_default, sms, &a_bn, &a_nsub, &a_ks);
      bn = a_bn;
      if (nsub <= 0) nsub = a_nsub;
      if (ksplit <= 0) ksplit = a_ks;
    } else if (m_tiles * cdiv(ncols, 64) < sms) {
      int a_bn = 64, a_ks = 1;
      choose_single_small(m_tiles, ncols, num_k, rows, multiple, can_split && ksplit <= 0 && split_default, sms, &a_bn, &a_ks);
      bn = a_bn;
      if (ksplit <= 0) ksplit = a_ks;
    } else {
      bn = choose_bn(ncols, m_tiles, multiple, sms);
      if (can_split && ksplit <= 0 && split_default) ksplit = choose_ksplit_single(m_tiles * cdiv(ncols, bn), num_k, sms);
    }
  }
  if (strided) can_split = false;
  if (!(ksplit > 1 && can_split)) ksplit = 1;

  plan->tw = tw; plan->th = th; plan->tn = tn;
  plan->bn = bn; plan->cta_pair = pair ? 1 : 0; plan->nsub = nsub > 0 ? nsub : 0; plan->ksplit = ksplit;
  plan->m_tiles = m_tiles;
  plan->workspace_bytes = ksplit > 1 ? (long long)ksplit * rows * d->cout * 4 : 0;
  plan->gn_rows_per_image = cdiv(d->w, tw) * cdiv(d->h, th);
  plan->gn_fusable = (gn_fuse && ksplit == 1 && d->mode == CB_EPI_LINEAR && !d->out_f32 && d->act == CB_ACT_NONE && oscale == 1.f &&
                      d->cout % 8 == 0 && num_k >= gn_min_k && 2 * rows * d->cout >= gn_min_bytes && (tw * th) % 32 == 0 &&
                      !strided && out_ld == d->cout) ? 1 : 0;
  // the staged (TMA in / TMA out) epilogue is what carries the LayerNorm fold: small-K 16-bit launches (cb_igemm: AUTO
  // stages num_k <= 48) and every GEGLU launch
  const bool staged16 = !d->out_f32 && d->act == CB_ACT_NONE && oscale == 1.f && d->cout % 8 == 0 && out_ld % 8 == 0 && ksplit == 1 &&
                        (d->epilogue == CB_EPILOGUE_STAGED || (d->epilogue == CB_EPILOGUE_AUTO && num_k <= 48));
  plan->ln_out_slots = (d->mode == CB_EPI_LINEAR && staged16 && !d->rowbias && !strided) ? int(2 * cdiv(ncols, bn)) : 0;
  plan->ln_foldable = ((d->mode == CB_EPI_LINEAR && staged16 && !d->rowbias && !d->residual) ||
                       (d->mode == CB_EPI_GEGLU && !d->out_f32 && out_ld % 8 == 0)) ? 1 : 0;
  return CB_OK;
}

extern "C" int cb_igemm_auto(const cb_igemm_desc* d, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  cb_igemm_plan_t plan;
  int rc = cb_igemm_plan(d, &plan);
  if (rc) return rc;
  cb_igemm_desc dd = *d;
  dd.tw = plan.tw; dd.th = plan.th; dd.tn = plan.tn;
  dd.bn = plan.bn; dd.cta_pair = plan.cta_pair; dd.nsub = plan.nsub;
  if (dd.out_ld <= 0) dd.out_ld = d->cout;
  if (d->gn_partials) CB_REQUIRE(plan.gn_fusable || d->gn_rows_per_image > 0, "cb_igemm_auto: this launch cannot produce GroupNorm partials");
  if (plan.ksplit <= 1) {
    dd.ksplit = 1;
    return cb_igemm(&dd, stream);
  }
  CB_REQUIRE(workspace != nullptr && workspace_bytes >= plan.workspace_bytes,
             "cb_igemm_auto: split-K needs a workspace of %lld bytes (cb_igemm_plan reports it)", (long long)plan.workspace_bytes);
  CB_REQUIRE(!d->gn_partials, "cb_igemm_auto: GroupNorm partials are not produced by split-K launches");
  dd.ksplit = plan.ksplit;
  dd.out = workspace; dd.out_f32 = 1; dd.out_ld = d->cout;
  dd.bias = nullptr; dd.rowbias = nullptr; dd.residual = nullptr; dd.rowbias_ld = 0; dd.res_ld = 0;
  rc = cb_igemm(&dd, stream);
  if (rc) return rc;
  return cb_splitk_reduce(static_cast<const float*>(workspace), plan.ksplit, d->n * d->h * d->w, d->cout, d->cout, d->bias,
                          d->rowbias, d->rowbias_ld, d->h * d->w, d->residual, d->res_ld, d->out,
                          d->out_ld > 0 ? d->out_ld : d->cout, stream);
}
