/* cremage_b200 -- C ABI of the B200-native (sm_100a) Stable Diffusion denoising hot path.
 *
 * The reference (HowToSD/cremage) has no native boundary: its hot path is a tree of torch.nn modules selected by
 * `target:` strings (ldm/util.py:81-96).  This header is the boundary the B200 implementation creates underneath
 * those modules; each entry point names the reference operator(s) it replaces (paths relative to the reference
 * root, `modules/` prefix omitted).  The Python mirror of the reference operator API lives in cremage_b200/ and
 * binds these symbols with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *  - every call enqueues work on `stream` and returns immediately (no synchronisation, no allocation ->
 *    CUDA-graph capturable);
 *  - return value: 0 = ok, negative = error (cb_last_error() gives a thread-local message);
 *  - activations are NHWC in the build's 16-bit type ("pixel rows x channels": fp16 in libcremage_b200_fp16.so -- the
 *    default, the reference's own GPU precision -- bf16 in libcremage_b200_bf16.so; cb_act_dtype() tells which; "act16"
 *    in the comments below stands for that type), statistics / latents / schedules are fp32.
 */
#ifndef CREMAGE_B200_H_
#define CREMAGE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ---------------------------------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------------------------------- */
const char* cb_last_error(void);
int cb_version(void);
/* 16-bit storage type this build of the library computes in: 1 = fp16, 2 = bf16 ("act16" in the comments below
 * means this type).  fp16 is the reference's own GPU precision (model.half() + autocast) and the default. */
int cb_act_dtype(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
int64_t cb_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * implicit GEMM on tcgen05/TMEM, operands by TMA
 *   replaces nn.Conv2d 3x3 / 1x1 and nn.Linear on the path:
 *     ResBlock in_layers/out_layers/skip_connection   ldm/modules/diffusionmodules/openaimodel.py:205-245
 *     Downsample.op / Upsample.conv                   openaimodel.py:111,155
 *     SpatialTransformer.proj_in/proj_out             ldm/modules/attention.py:1017-1021,1048
 *     CrossAttention to_q/to_k/to_v/to_out            ldm/modules/attention.py:571-578
 *     GEGLU proj + FeedForward net[2]                 ldm/modules/attention.py:78,96,145
 *     time_embed / emb_layers                         openaimodel.py:538-543,222-228
 *     VAE Decoder convs, AttnBlock q/k/v/proj_out     ldm/modules/diffusionmodules/model.py:89-209,469-575
 * ------------------------------------------------------------------------------------------------------------- */
enum { CB_EPI_LINEAR = 0, CB_EPI_GEGLU = 1, CB_EPI_HEADS = 2 };
enum { CB_ACT_NONE = 0, CB_ACT_SILU = 1 };
/* epilogue form: AUTO picks STAGED (residual panel in by TMA, result out by TMA store) for the small-K launches */
enum { CB_EPILOGUE_AUTO = 0, CB_EPILOGUE_DIRECT = 1, CB_EPILOGUE_STAGED = 2 };

typedef struct cb_igemm_desc {
  /* A operand: one or two NHWC act16 tensors [a_n][a_h][a_w][c] sharing the pixel grid; channels of source 1
   * follow those of source 0 in the K order (the UNet skip concat).  A plain [M,K] matrix is n=h=1, w=M. */
  const void* a0; int64_t c0; int64_t a0_ld; /* a0_ld: elements between pixels (0 = c0) */
  const void* a1; int64_t c1; int64_t a1_ld;
  int64_t a_n, a_h, a_w;
  /* output pixel grid (rows = n*h*w) and its 128-row tile {tw, th, tn} */
  int64_t n, h, w;
  int tw, th, tn;
  /* taps: A box of tap t is read at pixel offset (tap_dw, tap_dh, tap_dn)[t] from the output pixel */
  int taps;
  int tap_dw[9], tap_dh[9], tap_dn[9];
  /* weights: act16 [wgt_rows][taps * (ceil64(c0) + ceil64(c1))], K-major, zero padded */
  const void* wgt; int64_t wgt_rows;
  int64_t cout;          /* valid output columns (GEGLU: columns of the gated output = wgt_rows / 2) */
  /* epilogue */
  int mode;              /* CB_EPI_* */
  int act;               /* CB_ACT_* applied after bias + rowbias, before the residual */
  const float* bias;     /* [cout] fp32 or NULL (GEGLU: [2*cout], permuted like the weight rows) */
  const float* rowbias;  /* [n][rowbias_ld] fp32 per-image bias (timestep embedding) or NULL */
  int64_t rowbias_ld;
  const void* residual;  /* act16 [rows][res_ld] added last, or NULL */
  int64_t res_ld;
  void* out; int64_t out_ld; int out_f32; /* act16 (0) or fp32 (1) output, row stride out_ld */
  float out_scale;       /* multiplies the final value (0 = 1.0) */
  /* CB_EPI_HEADS: column -> (which, head, j), row -> (batch, token); out[which][batch*heads+head][token][dpad] */
  int heads_d, heads_dpad, heads_h, heads_tokens;
  int64_t heads_which_stride;
  /* tiling */
  int bn;                /* N tile, multiple of 32, <= 256 */
  int stages;            /* smem pipeline depth, 0 = auto */
  int epilogue;          /* CB_EPILOGUE_* (tuning / test knob; results are identical) */
  int cta_pair;          /* 1: 256-row tiles on CTA pairs (tcgen05 cta_group::2, cluster of 2); bn % 32 == 0 */
  int nsub;              /* pair mode: 0 = auto (two N tiles share each A stage when 3 * bn <= 512), 1 = never */
  int ksplit;            /* > 1: split K by tap groups; `out` must be an fp32 workspace [ksplit][rows][out_ld], no
                          * bias / rowbias / residual / activation here -- cb_splitk_reduce applies them */
  int64_t out_w_stride, out_h_stride, out_n_stride; /* optional (0 = dense rows of out_ld): element strides of the
                          * output pixel grid -- row (n, h, w) is written at out + n*out_n_stride + h*out_h_stride +
                          * w*out_w_stride.  Lets one launch fill every other pixel of a larger NHWC tensor (the four parity
                          * classes of a nearest-2x-upsample + conv3x3 folded into 2x2 convs).  Staged epilogue only
                          * (epilogue = CB_EPILOGUE_STAGED), no residual, 16-bit output, strides multiples of 8 elements. */
  float* gn_partials;    /* optional: fused GroupNorm statistics of the (16-bit rounded) output for a following
                          * cb_groupnorm_from_partials: fp32 [n][cb_gn_partial_blocks(h, w, tw, th)][2][cout/2], per M tile of
                          * the image and channel pair the sum and the sum of squares; plain 16-bit epilogues only
                          * (bias / rowbias / residual), ksplit <= 1, tw * th % 32 == 0.  Every entry is written. */
  int64_t gn_rows_per_image, gn_row_offset; /* optional: this launch fills rows [gn_row_offset, + its own M tiles per image)
                          * of a table with gn_rows_per_image rows per image (0 = a table of its own): the four parity
                          * launches of a folded upsample conv share one table */
  /* fused LayerNorm (BasicTransformerBlock norm1/2/3, ldm/modules/attention.py:900-912): the LayerNorm launch and its
   * read + write of the token matrix disappear.
   *   producer (the GEMM that writes the residual stream x): ln_partials_out = fp32 [rows][cb_igemm_plan().ln_out_slots][2],
   *     per row and slot the sum and the sum of squares of the 16-bit outputs (staged epilogue, bias / residual only).
   *   consumer (the GEMM that reads LayerNorm(x)): A = x itself; weights W'' = W diag(gamma) with every row centred
   *     (W''[n][k] -= mean_k W'[n][k], so x W''^T = (x - mean(x)) W'^T: the row mean needs no separate term), bias
   *     b' = b + W beta; ln_partials_in / ln_in_slots = the producer's table, ln_dim = row width, ln_eps; the epilogue
   *     computes rstd * (x W''^T) + b'. */
  float* ln_partials_out;
  const float* ln_partials_in;
  int ln_in_slots;
  int64_t ln_dim;
  float ln_eps;
} cb_igemm_desc;

int cb_igemm(const cb_igemm_desc* d, cudaStream_t stream);

/* Tiling policy inside the library (SURVEY 8b "plan" entries): a host fills the PROBLEM fields of a cb_igemm_desc and
 * leaves the tiling fields at 0 -- tw/th/tn, bn, nsub, ksplit: 0 = choose; cta_pair: 0 = choose, 1 = CTA pairs,
 * -1 = single CTAs; ksplit: 1 = never split -- cb_igemm_plan reports what the library would run on the current device:
 * the 128-row pixel tile, the N tile, CTA pairs (tcgen05 cta_group::2), the dual-N schedule, split-K by tap groups
 * and the fp32 workspace a split needs; whether the launch can write fused GroupNorm partials and how many rows per
 * image its partial table has.  Environment knobs: CB_PAIR, CB_PAIR_MIN_K_CHUNKS, CB_GEGLU_PAIR, CB_SPLITK, CB_GN_FUSE,
 * CB_GN_FUSE_MIN_K_CHUNKS, CB_GN_FUSE_MIN_BYTES. */
typedef struct cb_igemm_plan_t {
  int tw, th, tn;
  int bn, cta_pair, nsub, ksplit;
  int64_t m_tiles;
  int64_t workspace_bytes;      /* fp32 [ksplit][rows][cout] partials of a split-K launch, 0 otherwise */
  int gn_fusable;               /* cb_igemm_desc.gn_partials may be set for this launch */
  int64_t gn_rows_per_image;    /* rows per image of its partial table: fp32 [n][rows][2][cout/2] */
  int ln_out_slots;             /* > 0: cb_igemm_desc.ln_partials_out may be set, fp32 [rows][ln_out_slots][2] */
  int ln_foldable;              /* cb_igemm_desc.ln_partials_in may be set (staged plain / GEGLU epilogue) */
} cb_igemm_plan_t;
int cb_igemm_plan(const cb_igemm_desc* d, cb_igemm_plan_t* plan);
/* plan + run: cb_igemm with the library's tiling and, for a split-K plan, cb_splitk_reduce applying the descriptor's
 * bias / row bias / residual.  `workspace` (device, >= cb_igemm_plan().workspace_bytes; may be NULL when that is 0). */
int cb_igemm_auto(const cb_igemm_desc* d, void* workspace, int64_t workspace_bytes, cudaStream_t stream);

/* weight repacking on the device: fp32 OIHW conv weight (or [O][I] linear weight, kh = kw = 1), input channels split
 * into two sources c0 | c1 (c1 = 0: one source) -> 16-bit [cout][kh*kw][ceil64(c0) | ceil64(c1)], K-major, zero padded:
 * the `wgt` layout of cb_igemm_desc.  Done once per checkpoint load. */
int cb_pack_weight(const float* w, int64_t cout, int64_t c0, int64_t c1, int taps, void* out, cudaStream_t stream);

/* fold the fp32 partials of a split-K cb_igemm in split order (deterministic) and apply the epilogue:
 * out[r][c] = sum_s part[s][r][c] + bias[c] + rowbias[r / rows_per_image][c] + residual[r][c]  -> 16-bit */
int cb_splitk_reduce(const float* part, int splits, int64_t rows, int64_t cout, int64_t part_ld, const float* bias,
                     const float* rowbias, int64_t rowbias_ld, int64_t rows_per_image, const void* residual,
                     int64_t res_ld, void* out, int64_t out_ld, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * fused flash-style attention on tcgen05 (QK^T -> online softmax -> PV), replaces
 *   CrossAttention / CrossAttentionOriginal / MemoryEfficientCrossAttention cores  ldm/modules/attention.py:418-423,646-657,811
 * q, k, v: 16-bit, read in place from the projection GEMMs' row-major outputs: element (b, token, head, j) lives at
 * base[(b * tokens + token) * ld + head * d + j] (ld = elements between rows, >= heads * d, multiple of 8; q/k/v may be
 * column slices of one [tokens, 3 * heads * d] tensor).  The head dim is padded to a multiple of 64 inside the kernel by
 * TMA's out-of-bounds zero fill.  out: 16-bit [b][nq][heads*d]
 * ------------------------------------------------------------------------------------------------------------- */
int cb_attention(const void* q, int64_t q_ld, const void* k, int64_t k_ld, const void* v, int64_t v_ld, void* out,
                 int64_t batch, int64_t heads, int64_t nq, int64_t nk, int d, float scale, cudaStream_t stream);

/* row softmax: dst[r][:] = softmax(scale * src[r][:]); src fp32 (src_f32 = 1) or act16, dst act16 (may alias a act16
 * src); VAE AttnBlock, ldm/modules/diffusionmodules/model.py:196-198 */
int cb_softmax_rows(const void* src, int src_f32, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                    int64_t cols, float scale, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) over NHWC act16, fp32 statistics; replaces GroupNorm32+SiLU (ldm/modules/diffusionmodules/
 * util.py:214-216, openaimodel.py:205-207,229-231), Normalize (attention.py:189, model.py:45) + nonlinearity
 * (model.py:40-42).  Two sources = normalise the channel concat without materialising it.
 *   stats: workspace of cb_groupnorm_workspace_bytes(c0 + c1, n, hw, groups) bytes; its first n*groups*2 floats
 *   receive (sum, sum of squares).  Reductions use no floating-point atomics: results are run-to-run identical.
 * ------------------------------------------------------------------------------------------------------------- */
int64_t cb_groupnorm_workspace_bytes(int64_t c, int64_t n, int64_t hw, int groups);
int cb_groupnorm_nhwc(const void* x0, int64_t c0, const void* x1, int64_t c1, int64_t n, int64_t hw, int groups,
                      float eps, const float* gamma, const float* beta, int silu, void* out, float* stats,
                      cudaStream_t stream);

/* GroupNorm (+SiLU) whose statistics were accumulated by the producing cb_igemm launches (gn_partials): a fold of
 * the partials in fixed order (deterministic, fp64 accumulation) + ONE streaming pass (one read, one write per
 * element -- the algorithmic minimum).  part0 / part1: the partial buffers of the two sources with bpi0 / bpi1 blocks
 * per image (cb_gn_partial_blocks of the producing launch); stats: n * 32 * (c0 + c1) floats of workspace (the
 * partial tables reduced to 32 rows per image; the final fold runs in the apply kernel's prologue). */
int64_t cb_gn_partial_blocks(int64_t h, int64_t w, int tw, int th);   /* 0: tile not supported */
int cb_groupnorm_from_partials(const void* x0, int64_t c0, const float* part0, int64_t bpi0, const void* x1, int64_t c1,
                               const float* part1, int64_t bpi1, int64_t n, int64_t hw, int groups, float eps,
                               const float* gamma, const float* beta, int silu, void* out, float* stats,
                               cudaStream_t stream);

/* LayerNorm over the last dim of act16 [rows][c] (nn.LayerNorm, ldm/modules/attention.py:900-902) */
int cb_layernorm(const void* x, int64_t rows, int64_t c, float eps, const float* gamma, const float* beta, void* out,
                 cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * layout / small elementwise kernels
 * ------------------------------------------------------------------------------------------------------------- */
/* ControlNet residual injection, replaces `h += control.pop()` / `hs.pop() + control.pop()` of
 * ControlledUnetModel.forward (cldm/cldm.py:59-66): dst[n][p][ch] = base[n][p][ch] + ctrl[n][ch][p]; base / dst NHWC
 * 16-bit (dst may alias base), ctrl NCHW fp32 / fp16 / bf16 (ctrl_dtype 0 / 1 / 2) as the ControlNet returns it */
int cb_add_nchw_to_nhwc(const void* base, const void* ctrl, int ctrl_dtype, int64_t n, int64_t c, int64_t hw, void* dst,
                        cudaStream_t stream);
/* NCHW (fp32, or fp16/bf16 when src_dtype = 1/2) -> NHWC act16 with channel padding to c_pad (zeros), times scale */
int cb_nchw_to_nhwc(const void* src, int src_dtype, int64_t n, int64_t c, int64_t hw, int64_t c_pad, float scale,
                    void* dst, cudaStream_t stream);
/* 1x1 channel mix fused with the layout change: dst[n][p][co] = sum_ci w[co][ci] * src[n][ci][p] * scale + b[co]
 * (AutoencoderKL.post_quant_conv + the 1/scale_factor of decode_first_stage, ldm/models/autoencoder.py:303,336,
 * ldm/models/diffusion/ddpm.py:794-798); src NCHW fp32, w fp32 [cout][c], dst NHWC act16 [n][hw][c_pad] */
int cb_pointwise_nchw_to_nhwc(const float* src, int64_t n, int64_t c, int64_t hw, const float* w, const float* b,
                              int64_t cout, int64_t c_pad, float scale, void* dst, cudaStream_t stream);
/* NHWC (act16, or fp32 when src_f32) [n][hw][c_ld] -> NCHW fp32 [n][c][hw], first c channels */
int cb_nhwc_to_nchw_f32(const void* src, int src_f32, int64_t n, int64_t c, int64_t hw, int64_t c_ld, float* dst,
                        cudaStream_t stream);
/* nearest-neighbour 2x upsample NHWC act16 (F.interpolate(scale_factor=2, 'nearest'), openaimodel.py:120, model.py:61) */
int cb_upsample2x_nhwc(const void* src, int64_t n, int64_t h, int64_t w, int64_t c, void* dst, cudaStream_t stream);
/* parity split for stride-2 convs: src [n][h][w][c] -> dst [2*ph+pw][n][h/2][w/2][c] */
int cb_parity_split_nhwc(const void* src, int64_t n, int64_t h, int64_t w, int64_t c, void* dst, cudaStream_t stream);
/* sinusoidal timestep embedding (ldm/modules/diffusionmodules/util.py:151-171): t fp32 [n], freqs fp32 [dim/2]
 * (host-built with the reference's expression) -> act16 [n][dim] = [cos(t*f) | sin(t*f)] */
int cb_timestep_embedding(const float* t, int64_t n, int dim, const float* freqs, void* out, cudaStream_t stream);
/* direct 3x3 conv for tiny channel counts (UNet conv_in 4->320, VAE conv_in 4->512): src NHWC act16 [n][h][w][cin_ld],
 * wgt fp32 [3][3][cin][cout], out NHWC act16 */
int cb_conv3x3_small_cin(const void* src, int64_t n, int64_t h, int64_t w, int cin, int64_t cin_ld, const float* wgt,
                         const float* bias, int64_t cout, void* out, cudaStream_t stream);
/* DiagonalGaussianDistribution of AutoencoderKL.encode (ldm/modules/distributions/distributions.py:24-37;
 * ldm/models/autoencoder.py:324-331; LatentDiffusion.get_first_stage_encoding ddpm.py:585-594): moments fp32 NCHW
 * [n][2c][hw] -> mean, std = exp(0.5 clamp(logvar, -30, 20)), sample = scale * (mean + std * noise); any output may
 * be NULL, noise NULL = the mode; scale 0 = 1 */
int cb_diag_gaussian(const float* moments, const float* noise, int64_t n, int64_t c, int64_t hw, float scale,
                     float* mean_out, float* std_out, float* sample_out, cudaStream_t stream);
/* y = silu(x) or y = silu(x + add) over act16 vectors (SDXL label_emb path) */
int cb_silu_add(const void* x, const void* add, int64_t count, void* out, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * sampler latent updates, fp32 latents (any layout, `count` elements), one launch per step.
 * eps_u / eps_c are the unconditional / conditional halves of the CFG-doubled UNet output (uncond first:
 * ldm/models/diffusion/ldm_wrapper_for_k_diffusion.py:67-99, ddim.py:538-561).  Without guidance pass the same
 * pointer twice and cfg_scale 0.  With is_denoised = 1, eps_u already holds the denoised prediction (generic
 * `model(x, sigma)` callers) and eps_c / cfg_scale / the CompVis step are skipped.
 * ------------------------------------------------------------------------------------------------------------- */
/* x_in = x * c_in duplicated for both CFG halves: out [2][count] <- x [count]   (k_diffusion/external.py:111-114) */
int cb_cfg_scale_input(const float* x, int64_t per_batch, int64_t b, float c_in, float* out, cudaStream_t stream);
/* out = a*x + b*y (y may be NULL): CompVisDenoiser's input*c_in and input + eps*c_out (external.py:111-114) */
int cb_axpby_f32(const float* x, float a, const float* y, float b, int64_t count, float* out, cudaStream_t stream);
/* out = uncond + scale * (cond - uncond)   (ldm_wrapper_for_k_diffusion.py:99) */
int cb_cfg_mix_f32(const float* uncond, const float* cond, float scale, int64_t count, float* out, cudaStream_t stream);
/* out = keep * mask + (1 - mask) * fresh over fp32 NCHW latents [n][c][hw]; mask [n][mask_c][hw], mask_c = 1 (broadcast
 * over the channels) or c.  DDIM inpainting blend, ldm/models/diffusion/ddim.py:171-174. */
int cb_blend_mask_f32(const float* keep, const float* fresh, const float* mask, int64_t n, int64_t c, int64_t hw, int mask_c,
                      float* out, cudaStream_t stream);
/* Euler-ancestral (k_diffusion/sampling.py:147-163): denoised = x - sigma*(eu + s*(ec-eu)) [per half, then mixed];
 *   x' = x + (x-denoised)/sigma*(sigma_down - sigma) + noise*sigma_up ; noise may be NULL (last step) */
int cb_step_euler_ancestral(const float* x, const float* eps_u, const float* eps_c, int is_denoised, const float* noise,
                            int64_t count, float cfg_scale, float sigma, float sigma_down, float sigma_up, float* x_out,
                            float* denoised_out, cudaStream_t stream);
/* DPM++ 2M (k_diffusion/sampling.py:593-615): x' = ratio*x - em1*(c_new*denoised - c_old*old_denoised);
 *   old_denoised NULL = first / last step form */
int cb_step_dpmpp_2m(const float* x, const float* eps_u, const float* eps_c, int is_denoised, const float* old_denoised,
                     int64_t count, float cfg_scale, float sigma, float ratio, float em1, float c_new, float c_old,
                     float* x_out, float* denoised_out, cudaStream_t stream);
/* DDIM (ldm/models/diffusion/ddim.py:590-611): e = eu + s*(ec-eu); pred_x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);
 *   x' = sqrt(a_prev) pred_x0 + dir_coef * e + sigma_t*noise, dir_coef = sqrt(1-a_prev-sigma_t^2) */
int cb_step_ddim(const float* x, const float* eps_u, const float* eps_c, const float* noise, int64_t count,
                 float cfg_scale, float sqrt_at, float sqrt_one_minus_at, float sqrt_aprev, float dir_coef, float sigma_t,
                 float* x_out, float* pred_x0_out, cudaStream_t stream);
/* hires-fix latent upscale (sd/image_generator.py:975: F.interpolate(samples, scale_factor, 'bilinear',
 * align_corners=False)): fp32 [planes][h][w] -> [planes][h*factor][w*factor] */
int cb_bilinear_upsample_f32(const float* src, int64_t planes, int64_t h, int64_t w, int factor, float* dst,
                             cudaStream_t stream);
/* image post-process (sd/image_generator.py:1017-1018,1151-1152): NHWC fp32 [n][hw][c_ld] -> uint8 HWC [n][hw][3],
 * clamp((x+1)/2,0,1)*255 truncated */
int cb_image_to_u8(const void* src, int64_t n, int64_t hw, int64_t c_ld, uint8_t* dst, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CREMAGE_B200_H_ */
