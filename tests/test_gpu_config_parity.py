"""Parity at the sizes BASELINE.json's configs name, against goldens produced by the UNMODIFIED reference in fp32 on
the CPU (oracle/make_golden_configs.py; inputs and weights are re-derived here from the same seeds):

  cfg1  configs[0] verbatim -- SD1.5, batch 1, the complete 20-step Euler-ancestral run, every per-step latent + image
  cfg2  configs[1] at batch 2 -- the complete DDIM-50 run, final latent + pred_x0 at steps 0 / 25 / 49
  cfg3  configs[2] at batch 2 -- AutoencoderKL decode of a 64x64 latent
  hires configs[3]'s shape   -- one SD1.5 UNet forward at 128x128 (16 384-token self-attention inside the network)
  sdxl  configs[4]'s network -- one full-shape sd_xl_base UNet forward (CFG pair, 128x128, ctx 77x2048, y 2816)
  sdxlvae a25 -- first-stage decode with activations beyond fp16's range, and a 128x128-latent (1024^2) decode

Stated tolerances.  Latents: max|dx| <= TOL * max(1, |x|max) per step with TOL = 2e-2 for the fp16 build (the
reference's own GPU precision) -- the bound is relative because these random-weight trajectories run at |x| ~ 100-300
under CFG 7.5; the absolute error and the error in units of the step's noise level sigma_i are printed beside it and
collected in gpurun_out/parity_<dtype>.json.  bf16 (8 mantissa bits instead of 11: unit round-off 8x larger) is held to
BF16_FACTOR x the same bounds.  Images: PSNR >= 35 dB (peak-to-peak 2.0) in either build.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import build_dtype as _dtype
from tests._models import build_ldm, build_unet, build_vae, gold, psnr, randn, tol

pytestmark = pytest.mark.gpu


_TABLE = {}


def _record(name, **kv):
    _TABLE[name] = {k: (float(v) if not isinstance(v, (list, str)) else v) for k, v in kv.items()}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, f"parity_{_dtype()}.json")
        prev = {}
        if os.path.exists(path):
            with open(path) as f:
                prev = json.load(f)
        prev.update(_TABLE)
        with open(path, "w") as f:
            json.dump(prev, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _latent_check(name, got, want, sigma=None, bound=2e-2):
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    rms = ((got - want) ** 2).mean().sqrt().item() / max((want ** 2).mean().sqrt().item(), 1e-12)
    rel = err / max(ref, 1.0)
    msg = f"[parity] {name}: max|dx|={err:.4e} |x|max={ref:.2f} rel={rel:.3e} rel_rms={rms:.3e}"
    if sigma is not None and sigma > 0:
        msg += f" max|dx|/sigma={err / sigma:.3e} (sigma={sigma:.4f})"
    print(msg)
    assert rel <= tol(bound), msg
    return err, ref, rms


@pytest.fixture(scope="module")
def sd15():
    usd = O.make_weights(O.unet_param_shapes(O.SD15_UNET), seed=0)
    vsd = O.make_weights(O.decoder_param_shapes(O.SD15_VAE), seed=1)
    chk = O.weights_checksum(usd)
    ldm = build_ldm(O.SD15_UNET, usd, O.SD15_VAE, vsd)
    del usd, vsd
    yield ldm, chk
    del ldm
    torch.cuda.empty_cache()


def test_cfg1_euler20_full_trajectory_and_image(sd15):
    """BASELINE configs[0]: k_diffusion/sampling.py:147 through CompVisDenoiser + LDMWrapperForKDiffusion, then
    decode_first_stage (ddpm.py:794-798)."""
    from cremage_b200.k_diffusion.external import CompVisDenoiser
    from cremage_b200.k_diffusion.sampling import sample_euler_ancestral
    from cremage_b200.ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    ldm, chk = sd15
    g = gold("cfg1_euler20.npz")
    assert chk == pytest.approx(float(g["weights_checksum"]), rel=1e-12)
    den = CompVisDenoiser(ldm, False).cuda()
    sig = torch.from_numpy(g["sigmas"]).cuda()
    cond, uncond = randn((1, 77, 768), 21).cuda(), randn((1, 77, 768), 22).cuda()
    x_T, noise = randn((1, 4, 64, 64), 23).cuda(), randn((20, 1, 4, 64, 64), 24).cuda()
    wrapper = LDMWrapperForKDiffusion(den, cond, uncond, 7.5)
    trace, dens = [], []
    it = iter(range(20))
    xf = sample_euler_ancestral(wrapper, x_T * sig[0], sig, disable=True, noise_sampler=lambda s, sn: noise[next(it)],
                                callback=lambda d: (trace.append(d["x"].clone()), dens.append(d["denoised"].clone())))
    got = torch.stack(trace[1:] + [xf]).cpu()
    want = torch.from_numpy(g["trace"])
    sigmas = g["sigmas"]
    rows = []
    for i in range(20):
        err, ref, _ = _latent_check(f"cfg1 euler_a step {i:2d}", got[i], want[i], sigma=float(sigmas[i]))
        rows.append([i, float(sigmas[i]), err, ref])
    derr = (torch.stack(dens).cpu() - torch.from_numpy(g["denoised"])).abs().amax(dim=(1, 2, 3, 4))
    print("[parity] cfg1 denoised max|d| per step:", " ".join(f"{v:.3f}" for v in derr.tolist()))
    # the fused (callback-free) path the benchmark runs: ONE kernel per step for CFG mix + c_out + update + noise, a
    # different rounding sequence from the callback path above, held to the same bound against the same golden
    it = iter(range(20))
    x_fused = sample_euler_ancestral(wrapper, x_T * sig[0], sig, disable=True, noise_sampler=lambda s, sn: noise[next(it)])
    ferr, _, _ = _latent_check("cfg1 euler_a final latent, fused path", x_fused.cpu(), want[-1])
    img = ldm.decode_first_stage(x_fused)
    want_img = torch.from_numpy(g["image"].astype(np.float32))
    p = psnr(img.float().cpu(), want_img)
    print(f"[parity] cfg1 decoded 512x512 image: PSNR={p:.1f} dB")
    _record("cfg1_euler20", steps=rows, final_abs=rows[-1][2], final_rel=rows[-1][2] / max(rows[-1][3], 1.0),
            worst_rel=max(r[2] / max(r[3], 1.0) for r in rows), image_psnr=p, fused_final_abs=ferr)
    assert p >= 35.0


def test_cfg2_ddim50_full_run(sd15):
    """BASELINE configs[1] (the benchmarked workload) at batch 2: ldm/models/diffusion/ddim.py:78-190."""
    from cremage_b200.ldm.models.diffusion.ddim import DDIMSampler
    ldm, _ = sd15
    g = gold("cfg2_ddim50.npz")
    cond, uncond = randn((2, 77, 768), 31).cuda(), randn((2, 77, 768), 32).cuda()
    x_T = randn((2, 4, 64, 64), 33).cuda()
    smp = DDIMSampler(ldm)
    inter = {}
    x, _ = smp.sample(S=50, batch_size=2, shape=[4, 64, 64], conditioning=cond, eta=0.0, x_T=x_T,
                      unconditional_guidance_scale=7.5, unconditional_conditioning=uncond, verbose=False,
                      img_callback=lambda pred_x0, i: inter.__setitem__(i, pred_x0.clone()))
    assert smp.ddim_timesteps[0] == 1 and smp.ddim_timesteps[-1] == 981 and len(smp.ddim_timesteps) == 50
    rec = {}
    for j, step in enumerate(g["pred_x0_steps"].tolist()):
        err, ref, _ = _latent_check(f"cfg2 ddim pred_x0 step {step}", inter[step].cpu(), torch.from_numpy(g["pred_x0"][j]))
        rec[f"pred_x0_{step}_abs"], rec[f"pred_x0_{step}_absmax"] = err, ref
    err, ref, rms = _latent_check("cfg2 ddim final latent", x.cpu(), torch.from_numpy(g["final"]))
    _record("cfg2_ddim50", final_abs=err, final_absmax=ref, final_rel_rms=rms, **rec)
    x2, _ = smp.sample(S=50, batch_size=2, shape=[4, 64, 64], conditioning=cond, eta=0.0, x_T=x_T,
                       unconditional_guidance_scale=7.5, unconditional_conditioning=uncond, verbose=False)
    _latent_check("cfg2 ddim final latent, callback-free path", x2.cpu(), torch.from_numpy(g["final"]))


def test_cfg3_vae_decode_64(sd15):
    ldm, _ = sd15
    g = gold("cfg3_vae64.npz")
    z = randn((2, 4, 64, 64), 41).cuda()
    img = ldm.first_stage_model.decode(z)
    want = torch.from_numpy(g["out"].astype(np.float32))
    p = psnr(img.float().cpu(), want)
    err = (img.float().cpu() - want).abs().max().item()
    print(f"[parity] cfg3 vae decode 64x64 latent x2: PSNR={p:.1f} dB max_abs_err={err:.3e}")
    _record("cfg3_vae64", psnr=p, max_abs=err)
    assert p >= 35.0 and tuple(img.shape) == (2, 3, 512, 512)


def test_full_size_decode_and_unet_forward_are_bit_reproducible(sd15):
    """The same input through the same kernels must give the same bytes (the property the multi-GPU parity line of
    bench.py rests on): 60 AutoencoderKL decodes of one 8 x 64 x 64 latent batch -- the shape whose GroupNorm ring exposed
    the cross-proxy race of round 2 -- and 8 UNet forwards of one CFG-doubled batch."""
    ldm, _ = sd15
    z = randn((8, 4, 64, 64), 43).cuda()
    first = ldm.first_stage_model.decode(z).clone()
    bad = sum(0 if torch.equal(ldm.first_stage_model.decode(z), first) else 1 for _ in range(60))
    assert bad == 0, f"{bad} of 60 decodes differ"
    x, t, ctx = randn((4, 4, 64, 64), 44).cuda(), torch.tensor([801., 801., 12., 12.]).cuda(), randn((4, 77, 768), 45).cuda()
    u0 = ldm.model.diffusion_model(x, t, context=ctx).clone()
    bad = sum(0 if torch.equal(ldm.model.diffusion_model(x, t, context=ctx), u0) else 1 for _ in range(8))
    assert bad == 0, f"{bad} of 8 UNet forwards differ"


def test_hires_unet_forward_128(sd15):
    """configs[3]: the second pass runs the same UNet at 128x128 -- 16 384-token self-attention at the top level."""
    ldm, _ = sd15
    g = gold("hires_unet128.npz")
    x, ctx = randn((1, 4, 128, 128), 51).cuda(), randn((1, 77, 768), 52).cuda()
    y = ldm.model.diffusion_model(x, torch.from_numpy(g["t"]).cuda(), context=ctx)
    err, ref, rms = _latent_check("hires UNet 128x128 eps", y.float().cpu(), torch.from_numpy(g["out"]), bound=1e-2)
    _record("hires_unet128", max_abs=err, absmax=ref, rel_rms=rms)
    assert rms <= tol(5e-3)


def test_sdxl_unet_full_shape_forward():
    """configs[4]'s network: sd_xl_base.yaml:17-33 through sgm/modules/diffusionmodules/openaimodel.py:828-874."""
    from oracle import sgm_oracle as S
    from tests.test_sgm import _build_sgm_unet
    g = gold("sdxl_unet128.npz")
    cfg = S.SDXL_UNET
    sd = O.make_weights(S.sgm_unet_param_shapes(cfg), seed=300)
    assert O.weights_checksum(sd) == pytest.approx(float(g["weights_checksum"]), rel=1e-12)
    m = _build_sgm_unet(cfg, sd)
    del sd
    x = randn((2, 4, 128, 128), 61).cuda()
    ctx, y = randn((2, 77, cfg.context_dim), 62).cuda(), randn((2, cfg.adm_in_channels), 63).cuda()
    out = m(x, torch.from_numpy(g["t"]).cuda(), context=ctx, y=y)
    err, ref, rms = _latent_check("sdxl UNet 128x128 eps (CFG pair)", out.float().cpu(), torch.from_numpy(g["out"]), bound=1e-2)
    _record("sdxl_unet128", max_abs=err, absmax=ref, rel_rms=rms)
    assert rms <= tol(5e-3)
    del m
    torch.cuda.empty_cache()


def _sdxl_engine(vsd):
    """DiffusionEngine (sgm/models/diffusion.py) with only its first stage populated, sd_xl_base.yaml's flags."""
    from cremage_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    from cremage_b200.sgm.models.diffusion import DiffusionEngine
    from tests._models import full_vae_weights, vae_kwargs
    kw = vae_kwargs(O.SD15_VAE)
    with torch.device("meta"):
        vae = AutoencoderKLInferenceWrapper(embed_dim=kw["embed_dim"], ddconfig=kw["ddconfig"], lossconfig=None)
    vae = vae.to_empty(device="cpu")
    vae.load_state_dict(full_vae_weights(O.SD15_VAE, vsd), strict=True)
    eng = DiffusionEngine.__new__(DiffusionEngine)
    torch.nn.Module.__init__(eng)
    eng.first_stage_model = vae.cuda().eval()
    eng.scale_factor = 0.13025
    eng.disable_first_stage_autocast = True
    return eng


def test_sdxl_first_stage_survives_activations_beyond_fp16_range():
    """a25: the reference decodes the SDXL first stage in fp32 (disable_first_stage_autocast, sgm/models/diffusion.py:125)
    because its activations overflow fp16.  Fixture: conv_in scaled by 32768 -> the mid block carries |h| ~ 2e5 (> 65504)
    while every GroupNorm downstream is scale invariant, so the image is unchanged in exact arithmetic."""
    from oracle.make_golden_configs import scale_decoder_for_overflow
    g = gold("sdxl_vae_overflow.npz")
    assert float(g["mid_absmax"]) > 65504.0
    vsd = O.make_weights(O.decoder_param_shapes(O.SD15_VAE), seed=2)
    eng = _sdxl_engine(scale_decoder_for_overflow(vsd, float(g["gain"])))
    z = randn((1, 4, 32, 32), 71).cuda()
    img = eng.decode_first_stage(z * eng.scale_factor)           # decode_first_stage divides by scale_factor (:120)
    want = torch.from_numpy(g["out"].astype(np.float32))
    assert torch.isfinite(img).all()
    p = psnr(img.float().cpu(), want)
    print(f"[parity] sdxl first stage, |h| up to {float(g['mid_absmax']):.0f}: PSNR={p:.1f} dB")
    _record("sdxl_vae_overflow", psnr=p, mid_absmax=float(g["mid_absmax"]))
    assert p >= 35.0


def test_sdxl_first_stage_decode_1024():
    """a25 at BASELINE configs[4]'s size: 128x128 latent -> 1024x1024 (16 384-token single-head d=512 mid attention)."""
    g = gold("sdxl_vae128.npz")
    vsd = O.make_weights(O.decoder_param_shapes(O.SD15_VAE), seed=2)
    eng = _sdxl_engine(vsd)
    z = randn((1, 4, 128, 128), 72).cuda()
    img = eng.decode_first_stage(z * eng.scale_factor).float().cpu()
    assert tuple(img.shape) == (1, 3, 1024, 1024)
    crop = torch.from_numpy(g["crop"].astype(np.float32))
    p_crop = psnr(img[0, :, 384:640, 384:640], crop)
    pooled = torch.nn.functional.avg_pool2d(img, 8)[0]
    p_pool = psnr(pooled, torch.from_numpy(g["pooled"]))
    print(f"[parity] sdxl first stage 1024^2: centre-crop PSNR={p_crop:.1f} dB, 8x8-pooled whole image PSNR={p_pool:.1f} dB")
    _record("sdxl_vae128", psnr_crop=p_crop, psnr_pooled=p_pool)
    assert p_crop >= 35.0 and p_pool >= 35.0
