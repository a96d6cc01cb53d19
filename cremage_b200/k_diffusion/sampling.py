"""Drop-in mirror of the reference's `k_diffusion.sampling` functions on the SD path (HowToSD/cremage
modules/k_diffusion/sampling.py): schedules (:13-44), to_d (:46), get_ancestral_step (:51), sample_euler (:118),
sample_euler_ancestral (:147), sample_dpmpp_2m (:593) -- same signatures, same callback dictionary.

The per-step latent arithmetic runs in ONE fused CUDA kernel per step (classifier-free-guidance mix + CompVis
c_out step + sampler update [+ noise]) when `model` is cremage_b200's LDMWrapperForKDiffusion and no callback needs
the intermediate `denoised`; otherwise `model(x, sigma)` is treated as opaque and only the update is fused.
Step scalars (sigma_down, sigma_up, expm1(-h), ...) are computed with the reference's own fp32 torch expressions.
"""
from __future__ import annotations

import math

import torch
from tqdm.auto import trange

from .. import ops
from . import utils


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7., device='cpu'):
    """Constructs the noise schedule of Karras et al. (2022)."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return append_zero(sigmas).to(device)


def get_sigmas_exponential(n, sigma_min, sigma_max, device='cpu'):
    """Constructs an exponential noise schedule."""
    sigmas = torch.linspace(math.log(sigma_max), math.log(sigma_min), n, device=device).exp()
    return append_zero(sigmas)


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1., device='cpu'):
    """Constructs an polynomial in log sigma noise schedule."""
    ramp = torch.linspace(1, 0, n, device=device) ** rho
    sigmas = torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min))
    return append_zero(sigmas)


def get_sigmas_vp(n, beta_d=19.9, beta_min=0.1, eps_s=1e-3, device='cpu'):
    """Constructs a continuous VP noise schedule."""
    t = torch.linspace(1, eps_s, n, device=device)
    sigmas = torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1)
    return append_zero(sigmas)


def to_d(x, sigma, denoised):
    """Converts a denoiser output to a Karras ODE derivative."""
    return (x - denoised) / utils.append_dims(sigma, x.ndim)


def get_ancestral_step(sigma_from, sigma_to, eta=1.):
    """Calculates the noise level (sigma_down) to step down to and the amount
    of noise to add (sigma_up) when doing an ancestral sampling step."""
    if not eta:
        return sigma_to, 0.
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def default_noise_sampler(x):
    return lambda sigma, sigma_next: torch.randn_like(x)


def _prep(model, x, sigmas, extra_args, callback):
    if not x.is_cuda:
        raise RuntimeError("cremage_b200 has no CPU path: move the latents to a CUDA device")
    fused = getattr(model, "cb_fused_eps", None)
    if extra_args or callback is not None:
        fused = None
    sig_cpu = sigmas.detach().to(device="cpu", dtype=torch.float32)  # one sync per sampling run, not per step
    plan = fused and model.cb_plan(sig_cpu[:-1], x)  # per-step (c_in, device timestep rows), from the schedule only
    return fused, sig_cpu, plan


@torch.no_grad()
def sample_euler(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                 s_tmax=float('inf'), s_noise=1.):
    """Implements Algorithm 2 (Euler steps) from Karras et al. (2022)."""
    extra_args = {} if extra_args is None else extra_args
    if s_churn != 0.:
        raise NotImplementedError("cremage_b200: sample_euler with s_churn > 0 is not implemented")
    fused, sig, plan = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    for i in trange(len(sigmas) - 1, disable=disable):
        sigma = float(sig[i])
        if fused:
            eps2, cfg = model.cb_fused_eps(x, plan, i)
            x, _ = ops.step_euler_ancestral(x, eps2, None, cfg, sigma, float(sig[i + 1]), 0.0)
        else:
            denoised = model(x.to(dt_x), sigmas[i] * s_in, **extra_args)
            if callback is not None:
                callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
            x, _ = ops.step_euler_ancestral(x, None, None, 0.0, sigma, float(sig[i + 1]), 0.0,
                                            denoised=denoised.float())
    return x.to(dt_x)


@torch.no_grad()
def sample_euler_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                           noise_sampler=None):
    """Ancestral sampling with Euler method steps."""
    extra_args = {} if extra_args is None else extra_args
    noise_sampler = default_noise_sampler(x) if noise_sampler is None else noise_sampler
    fused, sig, plan = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    for i in trange(len(sigmas) - 1, disable=disable):
        sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1], eta=eta)
        noise = None
        if sig[i + 1] > 0:
            noise = noise_sampler(sigmas[i], sigmas[i + 1]).float().contiguous()
        su = float(sigma_up) * s_noise if s_noise != 1. else float(sigma_up)
        if fused:
            eps2, cfg = model.cb_fused_eps(x, plan, i)
            x, _ = ops.step_euler_ancestral(x, eps2, noise, cfg, float(sig[i]), float(sigma_down), su)
        else:
            denoised = model(x.to(dt_x), sigmas[i] * s_in, **extra_args)
            if callback is not None:
                callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
            x, _ = ops.step_euler_ancestral(x, None, noise, 0.0, float(sig[i]), float(sigma_down), su,
                                            denoised=denoised.float())
    return x.to(dt_x)


@torch.no_grad()
def sample_dpmpp_2m(model, x, sigmas, extra_args=None, callback=None, disable=None):
    """DPM-Solver++(2M)."""
    extra_args = {} if extra_args is None else extra_args
    fused, sig, plan = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    old_denoised = None
    for i in trange(len(sigmas) - 1, disable=disable):
        t, t_next = t_fn(sig[i]), t_fn(sig[i + 1])
        h = t_next - t
        ratio = float(sigma_fn(t_next) / sigma_fn(t))
        em1 = float((-h).expm1())
        if old_denoised is None or sig[i + 1] == 0:
            old, c_new, c_old = None, 1.0, 0.0
        else:
            h_last = t - t_fn(sig[i - 1])
            r = h_last / h
            old, c_new, c_old = old_denoised, float(1 + 1 / (2 * r)), float(1 / (2 * r))
        if fused:
            eps2, cfg = model.cb_fused_eps(x, plan, i)
            x, old_denoised = ops.step_dpmpp_2m(x, eps2, old, cfg, float(sig[i]), ratio, em1, c_new, c_old)
        else:
            denoised = model(x.to(dt_x), sigmas[i] * s_in, **extra_args)
            if callback is not None:
                callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
            x, old_denoised = ops.step_dpmpp_2m(x, None, old, 0.0, float(sig[i]), ratio, em1, c_new, c_old,
                                                denoised=denoised.float())
    return x.to(dt_x)
