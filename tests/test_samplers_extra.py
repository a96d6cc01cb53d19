"""Remaining k-diffusion samplers of the reference's front ends: Heun, DPM-2, DPM-2 ancestral, LMS,
DPM++ 2S ancestral.  CPU: oracle restatements against goldens produced by the reference's own functions
(oracle/make_golden_samplers.py).  GPU: cremage_b200.k_diffusion.sampling through the real wrapper chain."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import tol  # noqa: E402
from tests._models import build_ldm, gold

STEPS = 5
NAMES = ["heun", "lms", "dpm_2", "dpm_2_ancestral", "dpmpp_2s_ancestral"]


def _setup():
    g, e = gold("tiny_sampling.npz"), gold("tiny_samplers_extra.npz")
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    return g, e, sd


def _sigmas(e, name):
    return torch.from_numpy(e["sigmas_discrete" if name in ("heun", "lms") else "sigmas_karras"])


def test_schedules_of_the_front_ends_bit_exact():
    """HeunSampler / LmsSampler use get_sigmas(n); the DPM-2 family get_sigmas_karras(n, sigma_min, sigma_max)
    (k_diffusion_samplers.py:324,335,345,355,366)."""
    _, e, _ = _setup()
    _, ac, _ = O.alphas_cumprod_from_betas(O.make_beta_schedule_linear())
    sched = O.DiscreteSchedule(ac)
    assert np.array_equal(sched.get_sigmas(STEPS).numpy(), e["sigmas_discrete"])
    from cremage_b200.k_diffusion.sampling import get_sigmas_karras
    assert np.array_equal(get_sigmas_karras(STEPS, float(sched.sigma_min), float(sched.sigma_max)).numpy(), e["sigmas_karras"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    g, e, sd = _setup()
    _, ac, _ = O.alphas_cumprod_from_betas(O.make_beta_schedule_linear())
    den = O.OracleDenoiser(lambda x, t, c: O.unet_forward(sd, O.TINY_UNET, x, t, c), ac, torch.from_numpy(g["cond"]),
                           torch.from_numpy(g["uncond"]), float(g["cfg_scale"]))
    sig = _sigmas(e, name)
    x0 = torch.from_numpy(g["x_T"]) * sig[0]
    noise = torch.from_numpy(g["noise"])
    with torch.no_grad():
        if name in ("dpm_2_ancestral", "dpmpp_2s_ancestral"):
            x = getattr(O, "sample_" + name)(den, x0, sig, noise)
        else:
            x = getattr(O, "sample_" + name)(den, x0, sig)
    assert np.abs(x.numpy() - e[name]).max() < 1e-3 * max(1.0, np.abs(e[name]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_samplers_vs_reference_golden(name):
    from cremage_b200.k_diffusion import sampling
    from cremage_b200.k_diffusion.external import CompVisDenoiser
    from cremage_b200.ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    g, e, sd = _setup()
    ldm = build_ldm(O.TINY_UNET, sd)
    den = CompVisDenoiser(ldm, False).cuda()
    wrapper = LDMWrapperForKDiffusion(den, torch.from_numpy(g["cond"]).cuda(), torch.from_numpy(g["uncond"]).cuda(),
                                      float(g["cfg_scale"]))
    sig = _sigmas(e, name).cuda()
    x0 = (torch.from_numpy(g["x_T"]) * _sigmas(e, name)[0]).cuda()
    keep = x0.clone()
    noise = torch.from_numpy(g["noise"]).cuda()
    it = iter(range(STEPS))
    kw = {"noise_sampler": lambda s, sn: noise[next(it)]} if "ancestral" in name else {}
    calls = []
    x = getattr(sampling, "sample_" + name)(wrapper, x0, sig, disable=True, callback=lambda d: calls.append(d["i"]), **kw)
    assert torch.equal(x0, keep)                      # the caller's latent is not mutated
    assert calls == list(range(STEPS))                # the reference's callback contract: once per step, in order
    want = torch.from_numpy(e[name])
    err = (x.cpu() - want).abs().max().item()
    print(f"[parity] sample_{name}: max_abs_err={err:.4e} latent_absmax={want.abs().max():.2f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)


@pytest.mark.gpu
def test_front_end_classes_exist_and_sample():
    from cremage_b200.ldm.models.diffusion import k_diffusion_samplers as K
    g, _, sd = _setup()
    ldm = build_ldm(O.TINY_UNET, sd)
    for cls in (K.HeunSampler, K.Dpm2Sampler, K.Dpm2AncestralSampler, K.LmsSampler, K.Dpmpp2sAncestralSampler):
        smp = cls(ldm)
        x, _ = smp.sample(S=3, batch_size=2, shape=[4, 16, 16], conditioning=torch.from_numpy(g["cond"]).cuda(),
                          unconditional_guidance_scale=float(g["cfg_scale"]),
                          unconditional_conditioning=torch.from_numpy(g["uncond"]).cuda(),
                          x_T=torch.from_numpy(g["x_T"]).cuda())
        assert tuple(x.shape) == (2, 4, 16, 16) and torch.isfinite(x).all()
