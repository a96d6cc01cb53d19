"""Hires-fix second pass (latent upscaler): sd/image_generator.py:969-999 + img2img_sampling (DDIM branch).
CPU: the oracle restatement against the golden produced by the reference's own DDIMSampler.stochastic_encode /
decode.  GPU: cremage_b200.hires.hires_fix_latent against the same golden."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import tol  # noqa: E402
from tests._models import build_ldm, gold


def _setup():
    g = gold("tiny_sampling.npz")
    _, ac, _ = O.alphas_cumprod_from_betas(O.make_beta_schedule_linear())
    return g, ac


def test_oracle_hires_matches_reference_golden():
    g, ac = _setup()
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    eps_fn = lambda x, t, c: O.unet_forward(sd, O.TINY_UNET, x, t, c)
    with torch.no_grad():
        x = O.hires_fix_latent_ddim(eps_fn, ac, torch.from_numpy(g["ddim_final"]), torch.from_numpy(g["cond"]),
                                    torch.from_numpy(g["uncond"]), float(g["cfg_scale"]), S=5,
                                    strength=float(g["hires_strength"]), noise=torch.from_numpy(g["hires_noise"]))
    assert tuple(x.shape) == (2, 4, 32, 32)
    assert np.abs(x.numpy() - g["hires_final"]).max() < 5e-4


@pytest.mark.gpu
def test_bilinear_upsample_matches_reference_interpolate():
    from cremage_b200 import ops
    g, _ = _setup()
    up = ops.bilinear_upsample(torch.from_numpy(g["ddim_final"]).cuda(), 2)
    assert np.abs(up.cpu().numpy() - g["hires_upsampled"]).max() < 1e-5
    x = torch.randn(1, 3, 5, 7, generator=torch.Generator().manual_seed(1))
    want = torch.nn.functional.interpolate(x, scale_factor=3, mode="bilinear", align_corners=False)
    assert (ops.bilinear_upsample(x.cuda(), 3).cpu() - want).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_hires_fix_second_pass_vs_reference_golden():
    from cremage_b200.hires import hires_fix_latent
    from cremage_b200.ldm.models.diffusion.ddim import DDIMSampler
    g, _ = _setup()
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    ldm = build_ldm(O.TINY_UNET, sd)
    smp = DDIMSampler(ldm)
    x = hires_fix_latent(smp, torch.from_numpy(g["ddim_final"]).cuda(), torch.from_numpy(g["cond"]).cuda(),
                         torch.from_numpy(g["uncond"]).cuda(), float(g["cfg_scale"]), sampling_steps=5,
                         strength=float(g["hires_strength"]), noise=torch.from_numpy(g["hires_noise"]).cuda())
    want = torch.from_numpy(g["hires_final"])
    err = (x.cpu() - want).abs().max().item()
    print(f"[parity] hires second pass: max_abs_err={err:.4e} latent_absmax={want.abs().max():.2f}")
    assert tuple(x.shape) == (2, 4, 32, 32)
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)
