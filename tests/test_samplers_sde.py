"""SDE family, stochastic churn and the sampler front ends.  Goldens: the UNMODIFIED reference
functions with injected noise (oracle/make_golden_samplers2.py).  CPU: the oracle restatements against those goldens.
GPU: cremage_b200.k_diffusion.sampling / ldm.models.diffusion.k_diffusion_samplers through the C ABI."""
import contextlib

import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import tol  # noqa: E402
from tests._models import build_ldm, gold

STEPS = 5
# name -> (oracle function, sigma schedule, kwargs)
CASES = {
    "dpmpp_sde": ("sample_dpmpp_sde", "sigmas_karras", {}),
    "dpmpp_2m_sde": ("sample_dpmpp_2m_sde", "sigmas_karras", {}),
    "dpmpp_2m_sde_heun": ("sample_dpmpp_2m_sde", "sigmas_karras", {"solver_type": "heun"}),
    "dpmpp_3m_sde": ("sample_dpmpp_3m_sde", "sigmas_karras", {}),
    "dpmpp_2m_sde_eta0": ("sample_dpmpp_2m_sde", "sigmas_karras", {"eta": 0.0}),
}
CHURN = {
    "euler_churn": ("sample_euler", "sigmas_discrete", {}),
    "heun_churn": ("sample_heun", "sigmas_discrete", {"s_noise": 1.003}),
    "dpm_2_churn": ("sample_dpm_2", "sigmas_karras", {"s_tmin": 0.05, "s_tmax": 10.0}),
}


@contextlib.contextmanager
def injected_randn_like(seq):
    real = torch.randn_like
    it = iter(seq)
    torch.randn_like = lambda x, *a, **k: next(it).to(x.device)
    try:
        yield
    finally:
        torch.randn_like = real


def _setup():
    g, e = gold("tiny_sampling.npz"), gold("tiny_samplers_sde.npz")
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    return g, e, sd


def _oracle_den(g, sd):
    _, ac, _ = O.alphas_cumprod_from_betas(O.make_beta_schedule_linear())
    return O.OracleDenoiser(lambda x, t, c: O.unet_forward(sd, O.TINY_UNET, x, t, c), ac, torch.from_numpy(g["cond"]),
                            torch.from_numpy(g["uncond"]), float(g["cfg_scale"])), ac


@pytest.mark.parametrize("name", list(CASES) + list(CHURN))
def test_oracle_matches_reference_golden(name):
    g, e, sd = _setup()
    den, _ = _oracle_den(g, sd)
    fn, sched, kw = (CASES.get(name) or CHURN[name])
    sig = torch.from_numpy(e[sched])
    # the churn goldens all start from x_T * sigmas_discrete[0] (oracle/make_golden_samplers2.py)
    x0 = torch.from_numpy(g["x_T"]) * (torch.from_numpy(e["sigmas_discrete"])[0] if name in CHURN else sig[0])
    noise = list(torch.from_numpy(e["noise"]))
    with torch.no_grad():
        if name in CHURN:
            x = getattr(O, fn + "_churn")(den, x0, sig, noise, float(e["s_churn"]), **kw)
        else:
            x = getattr(O, fn)(den, x0, sig, noise, **kw)
    assert np.abs(x.numpy() - e[name]).max() < 1e-3 * max(1.0, np.abs(e[name]).max())


def test_oracle_stochastic_encode_matches_reference_golden():
    g, e, _ = _setup()
    _, ac, _ = O.alphas_cumprod_from_betas(O.make_beta_schedule_linear())
    noise = torch.from_numpy(e["noise"])
    x_T = torch.from_numpy(g["x_T"])
    a = O.kdiff_stochastic_encode(ac, x_T, torch.tensor([2, 2]), 5, noise[0])
    b = O.kdiff_stochastic_encode(ac, x_T, torch.tensor([1, 3]), 5, noise[1])
    assert np.abs(a.numpy() - e["front_stochastic_encode"]).max() < 1e-6
    assert np.abs(b.numpy() - e["front_stochastic_encode_ragged"]).max() < 1e-6


def test_brownian_path_is_one_consistent_process():
    """The torchsde-free default noise source: increments over nested / adjacent intervals add up, unit variance after
    the sampler's normalisation, deterministic for a seed."""
    from cremage_b200.k_diffusion.sampling import BrownianPath, BrownianTreeNoiseSampler
    x = torch.zeros(4, 64, 64)
    p = BrownianPath(x, 0.03, 14.6, seed=7)
    full = p(0.03, 14.6)
    assert torch.allclose(p(0.03, 2.0) + p(2.0, 14.6), full, atol=1e-5)
    assert torch.allclose(p(0.5, 1.0) + p(1.0, 2.0), p(0.5, 2.0), atol=1e-5)
    q = BrownianPath(x, 0.03, 14.6, seed=7)
    assert torch.equal(q(0.03, 14.6), full)
    ns = BrownianTreeNoiseSampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=3)
    n = ns(torch.tensor(10.0), torch.tensor(7.0))
    assert abs(float(n.std()) - 1.0) < 0.05 and abs(float(n.mean())) < 0.05


def _gpu_wrapper(g, sd):
    from cremage_b200.k_diffusion.external import CompVisDenoiser
    from cremage_b200.ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    ldm = build_ldm(O.TINY_UNET, sd)
    den = CompVisDenoiser(ldm, False).cuda()
    return ldm, LDMWrapperForKDiffusion(den, torch.from_numpy(g["cond"]).cuda(), torch.from_numpy(g["uncond"]).cuda(),
                                        float(g["cfg_scale"]))


def _check(name, x, want):
    err = (x.float().cpu() - want).abs().max().item()
    print(f"[parity] {name}: max_abs_err={err:.4e} latent_absmax={want.abs().max():.2f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_sde_samplers_vs_reference_golden(name):
    from cremage_b200.k_diffusion import sampling
    g, e, sd = _setup()
    _, wrapper = _gpu_wrapper(g, sd)
    fn, sched, kw = CASES[name]
    sig = torch.from_numpy(e[sched])
    x0 = (torch.from_numpy(g["x_T"]) * sig[0]).cuda()
    keep = x0.clone()
    noise = torch.from_numpy(e["noise"]).cuda()
    it = iter(range(noise.shape[0]))
    calls = []
    x = getattr(sampling, fn)(wrapper, x0, sig.cuda(), disable=True, noise_sampler=lambda s, sn: noise[next(it)],
                              callback=lambda d: calls.append(d["i"]), **kw)
    assert torch.equal(x0, keep) and calls == list(range(STEPS))
    _check(name, x, torch.from_numpy(e[name]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CHURN))
def test_cuda_churn_vs_reference_golden(name):
    from cremage_b200.k_diffusion import sampling
    g, e, sd = _setup()
    _, wrapper = _gpu_wrapper(g, sd)
    fn, sched, kw = CHURN[name]
    sig = torch.from_numpy(e[sched])
    x0 = (torch.from_numpy(g["x_T"]) * torch.from_numpy(e["sigmas_discrete"])[0]).cuda()
    with injected_randn_like(list(torch.from_numpy(e["noise"]).cuda())):
        x = getattr(sampling, fn)(wrapper, x0, sig.cuda(), disable=True, s_churn=float(e["s_churn"]), **kw)
    _check(name, x, torch.from_numpy(e[name]))


@pytest.mark.gpu
def test_cuda_default_sde_noise_runs_without_torchsde():
    """No noise_sampler given: the Brownian path default (torchsde is not installed) -- finite, seed-deterministic."""
    from cremage_b200.k_diffusion import sampling
    g, e, sd = _setup()
    _, wrapper = _gpu_wrapper(g, sd)
    sig = torch.from_numpy(e["sigmas_karras"])
    x0 = (torch.from_numpy(g["x_T"]) * sig[0]).cuda()
    outs = []
    for _ in range(2):
        torch.manual_seed(5)
        outs.append(sampling.sample_dpmpp_2m_sde(wrapper, x0, sig.cuda(), disable=True))
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])


@pytest.mark.gpu
def test_cuda_front_ends_vs_reference_golden():
    """KDiffusionSamplerBase.sample / stochastic_encode (k_diffusion_samplers.py:197-297) through the mirrors."""
    from cremage_b200.ldm.models.diffusion import k_diffusion_samplers as K
    g, e, sd = _setup()
    ldm, _ = _gpu_wrapper(g, sd)
    cond, uncond = torch.from_numpy(g["cond"]).cuda(), torch.from_numpy(g["uncond"]).cuda()
    x_T = torch.from_numpy(g["x_T"]).cuda()
    noise = torch.from_numpy(e["noise"]).cuda()
    common = dict(batch_size=2, shape=[4, 16, 16], conditioning=cond, unconditional_guidance_scale=float(g["cfg_scale"]),
                  unconditional_conditioning=uncond)
    sig_d = torch.from_numpy(e["sigmas_discrete"])
    with injected_randn_like(list(noise)):
        xa, aux = K.EulerAncestralSampler(ldm).sample(S=STEPS, x0=x_T * float(sig_d[0]), **common)
    assert aux is None
    _check("EulerAncestralSampler.sample", xa, torch.from_numpy(e["front_euler_a"]))
    smp = K.Dpmpp2mSampler(ldm)
    xm, _ = smp.sample(S=6, x0=x_T * 2.0, denoising_steps=3, **common)
    assert np.array_equal(smp.sigmas.cpu().numpy(), e["front_dpmpp2m_img2img_sigmas"])       # last t+1 sigmas, bit-exact
    _check("Dpmpp2mSampler.sample(denoising_steps=3)", xm, torch.from_numpy(e["front_dpmpp2m_img2img"]))
    enc = smp.stochastic_encode(x_T, torch.tensor([2, 2]).cuda(), 5, noise=noise[0])
    assert (enc.cpu() - torch.from_numpy(e["front_stochastic_encode"])).abs().max().item() < 1e-5
    enc2 = smp.stochastic_encode(x_T, torch.tensor([1, 3]).cuda(), 5, noise=noise[1])
    assert (enc2.cpu() - torch.from_numpy(e["front_stochastic_encode_ragged"])).abs().max().item() < 1e-5
    for cls in (K.DpmppSdeSampler, K.Dpmpp2mSdeSampler, K.Dpmpp3mSdeSampler):       # :373-411, default Brownian noise
        x, _ = cls(ldm).sample(S=3, x0=x_T * 14.0, **common)
        assert tuple(x.shape) == (2, 4, 16, 16) and torch.isfinite(x).all()


def test_oracle_ddim_inpainting_matches_reference_golden():
    g, _, sd = _setup()
    m = gold("tiny_ddim_mask.npz")
    _, ac, _ = O.alphas_cumprod_from_betas(O.make_beta_schedule_linear())
    eps = lambda x, t, c: O.unet_forward(sd, O.TINY_UNET, x, t, c)
    with torch.no_grad():
        x = O.ddim_sample(eps, ac, torch.from_numpy(g["x_T"]), torch.from_numpy(g["cond"]), torch.from_numpy(g["uncond"]),
                          float(g["cfg_scale"]), 5, mask=torch.from_numpy(m["mask"]), x0=torch.from_numpy(m["x0"]),
                          mask_noise=list(torch.from_numpy(m["noise"])))
    assert np.abs(x.numpy() - m["final"]).max() < 1e-3 * max(1.0, np.abs(m["final"]).max())


@pytest.mark.gpu
def test_cuda_ddim_inpainting_vs_reference_golden():
    """DDIMSampler.sample(mask=, x0=) (ldm/models/diffusion/ddim.py:171-174): q_sample + masked blend before every step."""
    from cremage_b200.ldm.models.diffusion.ddim import DDIMSampler
    g, _, sd = _setup()
    m = gold("tiny_ddim_mask.npz")
    ldm, _ = _gpu_wrapper(g, sd)
    with injected_randn_like(list(torch.from_numpy(m["noise"]).cuda())):
        x, _ = DDIMSampler(ldm).sample(S=5, batch_size=2, shape=[4, 16, 16], conditioning=torch.from_numpy(g["cond"]).cuda(),
                                       eta=0.0, x_T=torch.from_numpy(g["x_T"]).cuda(), mask=torch.from_numpy(m["mask"]).cuda(),
                                       x0=torch.from_numpy(m["x0"]).cuda(), unconditional_guidance_scale=float(g["cfg_scale"]),
                                       unconditional_conditioning=torch.from_numpy(g["uncond"]).cuda(), verbose=False)
    _check("DDIMSampler.sample(mask, x0)", x, torch.from_numpy(m["final"]))
