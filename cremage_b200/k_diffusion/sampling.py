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


# ----------------------------------------------------------------------------------------------------------------------
# The remaining samplers the reference's front ends reach (ldm/models/diffusion/k_diffusion_samplers.py:321-372):
# Heun (:167), DPM-2 (:196), DPM-2 ancestral (:227), LMS (:269), DPM++ 2S ancestral (:517).  Host-side control flow
# over the same kernels: `model(x, sigma)` is one CFG-doubled UNet call, every latent update is an axpby / Euler-step
# launch on fp32 latents.  Step scalars come from the reference's own fp32 torch expressions on the CPU schedule copy.
# ----------------------------------------------------------------------------------------------------------------------
def _no_churn(s_churn, name):
    if s_churn != 0.:
        raise NotImplementedError(f"cremage_b200: {name} with s_churn > 0 is not implemented")


def _to_d(x, sigma: float, denoised):
    """(x - denoised) / sigma as one launch (reference to_d, :46-48)."""
    inv = 1.0 / sigma
    return ops.axpby(x, inv, denoised, -inv)


def _call(model, x, dt_x, sigma, s_in, extra_args, callback, i, sigmas):
    denoised = model(x.to(dt_x), sigma * s_in, **extra_args).float().contiguous()
    if callback is not None:
        callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
    return denoised


@torch.no_grad()
def sample_heun(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                s_tmax=float('inf'), s_noise=1.):
    """Implements Algorithm 2 (Heun steps) from Karras et al. (2022)."""
    extra_args = {} if extra_args is None else extra_args
    _no_churn(s_churn, "sample_heun")
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        d = _to_d(x, float(sig[i]), denoised)
        dt = float(sig[i + 1] - sig[i])
        if sig[i + 1] == 0:
            x = ops.axpby(x, 1.0, d, dt)                       # Euler method
        else:
            x_2 = ops.axpby(x, 1.0, d, dt)                     # Heun's method
            denoised_2 = model(x_2.to(dt_x), sigmas[i + 1] * s_in, **extra_args).float().contiguous()
            d_2 = _to_d(x_2, float(sig[i + 1]), denoised_2)
            d_prime = ops.axpby(d, 0.5, d_2, 0.5)
            x = ops.axpby(x, 1.0, d_prime, dt)
    return x.to(dt_x)


@torch.no_grad()
def sample_dpm_2(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                 s_tmax=float('inf'), s_noise=1.):
    """A sampler inspired by DPM-Solver-2 and Algorithm 2 from Karras et al. (2022)."""
    extra_args = {} if extra_args is None else extra_args
    _no_churn(s_churn, "sample_dpm_2")
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        d = _to_d(x, float(sig[i]), denoised)
        if sig[i + 1] == 0:
            x = ops.axpby(x, 1.0, d, float(sig[i + 1] - sig[i]))
        else:
            sigma_mid = sig[i].log().lerp(sig[i + 1].log(), 0.5).exp()
            dt_1, dt_2 = float(sigma_mid - sig[i]), float(sig[i + 1] - sig[i])
            x_2 = ops.axpby(x, 1.0, d, dt_1)
            denoised_2 = model(x_2.to(dt_x), sigma_mid.to(x.device) * s_in, **extra_args).float().contiguous()
            d_2 = _to_d(x_2, float(sigma_mid), denoised_2)
            x = ops.axpby(x, 1.0, d_2, dt_2)
    return x.to(dt_x)


@torch.no_grad()
def sample_dpm_2_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                           noise_sampler=None):
    """Ancestral sampling with DPM-Solver second-order steps."""
    extra_args = {} if extra_args is None else extra_args
    noise_sampler = default_noise_sampler(x) if noise_sampler is None else noise_sampler
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1], eta=eta)
        d = _to_d(x, float(sig[i]), denoised)
        if sigma_down == 0:
            x = ops.axpby(x, 1.0, d, float(sigma_down - sig[i]))
        else:
            sigma_mid = sig[i].log().lerp(sigma_down.log(), 0.5).exp()
            dt_1, dt_2 = float(sigma_mid - sig[i]), float(sigma_down - sig[i])
            x_2 = ops.axpby(x, 1.0, d, dt_1)
            denoised_2 = model(x_2.to(dt_x), sigma_mid.to(x.device) * s_in, **extra_args).float().contiguous()
            d_2 = _to_d(x_2, float(sigma_mid), denoised_2)
            x = ops.axpby(x, 1.0, d_2, dt_2)
            noise = noise_sampler(sigmas[i], sigmas[i + 1]).float().contiguous()
            x = ops.axpby(x, 1.0, noise, float(s_noise * sigma_up))
    return x.to(dt_x)


def linear_multistep_coeff(order, t, i, j):
    from scipy import integrate
    if order - 1 > i:
        raise ValueError(f'Order {order} too high for step {i}')

    def fn(tau):
        prod = 1.
        for k in range(order):
            if j == k:
                continue
            prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod
    return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]


@torch.no_grad()
def sample_lms(model, x, sigmas, extra_args=None, callback=None, disable=None, order=4):
    extra_args = {} if extra_args is None else extra_args
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    sigmas_cpu = sigmas.detach().cpu().numpy()
    ds = []
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        ds.append(_to_d(x, float(sig[i]), denoised))
        if len(ds) > order:
            ds.pop(0)
        cur_order = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur_order, sigmas_cpu, i, j) for j in range(cur_order)]
        # x + sum(coeff * d): the reference's sum() starts from 0 and adds left to right
        acc = None
        for coeff, d in zip(coeffs, reversed(ds)):
            acc = ops.axpby(d, float(coeff)) if acc is None else ops.axpby(acc, 1.0, d, float(coeff))
        x = ops.axpby(x, 1.0, acc, 1.0)
    return x.to(dt_x)


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                              noise_sampler=None):
    """Ancestral sampling with DPM-Solver++(2S) second-order steps."""
    extra_args = {} if extra_args is None else extra_args
    noise_sampler = default_noise_sampler(x) if noise_sampler is None else noise_sampler
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1], eta=eta)
        if sigma_down == 0:
            d = _to_d(x, float(sig[i]), denoised)
            x = ops.axpby(x, 1.0, d, float(sigma_down - sig[i]))
        else:
            t, t_next = t_fn(sig[i]), t_fn(sigma_down)
            r = 1 / 2
            h = t_next - t
            s = t + r * h
            x_2 = ops.axpby(x, float(sigma_fn(s) / sigma_fn(t)), denoised, -float((-h * r).expm1()))
            denoised_2 = model(x_2.to(dt_x), sigma_fn(s).to(x.device) * s_in, **extra_args).float().contiguous()
            x = ops.axpby(x, float(sigma_fn(t_next) / sigma_fn(t)), denoised_2, -float((-h).expm1()))
        if sig[i + 1] > 0:
            noise = noise_sampler(sigmas[i], sigmas[i + 1]).float().contiguous()
            x = ops.axpby(x, 1.0, noise, float(s_noise * sigma_up))
    return x.to(dt_x)
