"""Drop-in mirror of the reference's `k_diffusion.sampling` functions on the SD path (HowToSD/cremage
modules/k_diffusion/sampling.py): schedules (:13-44), to_d (:46), get_ancestral_step (:51), sample_euler (:118),
sample_euler_ancestral (:147), sample_dpmpp_2m (:593), the remaining deterministic / ancestral samplers and the SDE
family (sample_dpmpp_sde :551, sample_dpmpp_2m_sde :619, sample_dpmpp_3m_sde :664) -- same signatures, same callback
dictionary, same consumption of the global torch RNG (a draw the reference makes and discards is made and discarded).

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


def _churn(x, sig, i, n_steps, s_churn, s_tmin, s_tmax, s_noise):
    """Karras' stochastic churn in front of a step (sampling.py:124-133, :171-175, :200-204): returns (x, sigma_hat).
    The reference draws `randn_like(x)` on every step whether or not gamma > 0; so does this (RNG-stream parity)."""
    gamma = min(s_churn / n_steps, 2 ** 0.5 - 1) if s_tmin <= sig[i] <= s_tmax else 0.
    eps = torch.randn_like(x)
    sigma_hat = sig[i] * (gamma + 1)
    if gamma > 0:
        x = ops.axpby(x, 1.0, eps.float().contiguous(), float(s_noise * (sigma_hat ** 2 - sig[i] ** 2) ** 0.5))
    return x, sigma_hat


@torch.no_grad()
def sample_euler(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                 s_tmax=float('inf'), s_noise=1.):
    """Implements Algorithm 2 (Euler steps) from Karras et al. (2022)."""
    extra_args = {} if extra_args is None else extra_args
    fused, sig, plan = _prep(model, x, sigmas, extra_args, callback)
    if s_churn != 0.:
        fused = None        # sigma_hat != sigma_i: the fused step's per-step plan is built from the schedule alone
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    n = len(sigmas) - 1
    for i in trange(n, disable=disable):
        x, sigma_hat = _churn(x, sig, i, n, s_churn, s_tmin, s_tmax, s_noise)
        if fused:
            eps2, cfg = model.cb_fused_eps(x, plan, i)
            x, _ = ops.step_euler_ancestral(x, eps2, None, cfg, float(sig[i]), float(sig[i + 1]), 0.0)
        else:
            sh = sigmas[i] if s_churn == 0. else sigma_hat.to(x.device)
            denoised = model(x.to(dt_x), sh * s_in, **extra_args)
            if callback is not None:
                callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sh, 'denoised': denoised})
            x, _ = ops.step_euler_ancestral(x, None, None, 0.0, float(sigma_hat), float(sig[i + 1]), 0.0,
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
def _to_d(x, sigma: float, denoised):
    """(x - denoised) / sigma as one launch (reference to_d, :46-48)."""
    inv = 1.0 / sigma
    return ops.axpby(x, inv, denoised, -inv)


def _call(model, x, dt_x, sigma, s_in, extra_args, callback, i, sigmas):
    denoised = model(x.to(dt_x), sigma * s_in, **extra_args).float().contiguous()
    if callback is not None:
        callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigma, 'denoised': denoised})
    return denoised


def _lin(*terms):
    """sum(coef * tensor) over fp32 latents, left to right, two terms per launch."""
    (c0, t0), rest = terms[0], terms[1:]
    if not rest:
        return ops.axpby(t0, float(c0))
    acc = ops.axpby(t0, float(c0), rest[0][1], float(rest[0][0]))
    for c, t in rest[1:]:
        acc = ops.axpby(acc, 1.0, t, float(c))
    return acc


@torch.no_grad()
def sample_heun(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                s_tmax=float('inf'), s_noise=1.):
    """Implements Algorithm 2 (Heun steps) from Karras et al. (2022)."""
    extra_args = {} if extra_args is None else extra_args
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    n = len(sigmas) - 1
    for i in trange(n, disable=disable):
        x, sigma_hat = _churn(x, sig, i, n, s_churn, s_tmin, s_tmax, s_noise)
        denoised = _call(model, x, dt_x, sigmas[i] if s_churn == 0. else sigma_hat.to(x.device), s_in, extra_args,
                         callback, i, sigmas)
        d = _to_d(x, float(sigma_hat), denoised)
        dt = float(sig[i + 1] - sigma_hat)
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
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    n = len(sigmas) - 1
    for i in trange(n, disable=disable):
        x, sigma_hat = _churn(x, sig, i, n, s_churn, s_tmin, s_tmax, s_noise)
        denoised = _call(model, x, dt_x, sigmas[i] if s_churn == 0. else sigma_hat.to(x.device), s_in, extra_args,
                         callback, i, sigmas)
        d = _to_d(x, float(sigma_hat), denoised)
        if sig[i + 1] == 0:
            x = ops.axpby(x, 1.0, d, float(sig[i + 1] - sigma_hat))
        else:
            sigma_mid = sigma_hat.log().lerp(sig[i + 1].log(), 0.5).exp()
            dt_1, dt_2 = float(sigma_mid - sigma_hat), float(sig[i + 1] - sigma_hat)
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



# ----------------------------------------------------------------------------------------------------------------------
# SDE family (sampling.py:551-617, :619-662, :664-717).  `noise_sampler(sigma, sigma_next)` is the reference's hook; the
# default is a Brownian-tree sampler so that the noise of nested / successive sigma intervals is one consistent path.
# ----------------------------------------------------------------------------------------------------------------------
class BrownianPath:
    """W(t) of one Brownian motion per batch item on [t0, t1], sampled lazily: a query time between two known times is
    drawn from the Brownian bridge between them, so increments over nested or adjacent intervals are consistent.
    Used when torchsde (the reference's un-vendored dependency, requirements.txt torchsde==0.2.6) is not importable:
    statistically the same process as torchsde.BrownianTree, not the same bits."""

    def __init__(self, x, t0, t1, seed=None):
        self.gen = torch.Generator(device=x.device)
        self.gen.manual_seed(int(torch.randint(0, 2 ** 63 - 1, []).item()) if seed is None else int(seed))
        self.shape, self.device, self.dtype = tuple(x.shape), x.device, torch.float32
        self.times = [float(t0), float(t1)]
        w1 = self._randn() * math.sqrt(float(t1) - float(t0))
        self.values = [torch.zeros(self.shape, device=self.device, dtype=self.dtype), w1]

    def _randn(self):
        return torch.randn(self.shape, generator=self.gen, device=self.device, dtype=self.dtype)

    def w(self, t: float):
        import bisect
        t = float(t)
        k = bisect.bisect_left(self.times, t)
        if k < len(self.times) and self.times[k] == t:
            return self.values[k]
        if k == 0:                                   # before the first known time: independent increment backwards
            v = self.values[0] - self._randn() * math.sqrt(self.times[0] - t)
        elif k == len(self.times):                   # beyond the last known time
            v = self.values[-1] + self._randn() * math.sqrt(t - self.times[-1])
        else:                                        # Brownian bridge between the neighbours
            ta, tb = self.times[k - 1], self.times[k]
            wa, wb = self.values[k - 1], self.values[k]
            lam = (t - ta) / (tb - ta)
            v = wa + lam * (wb - wa) + self._randn() * math.sqrt((t - ta) * (tb - t) / (tb - ta))
        self.times.insert(k, t)
        self.values.insert(k, v)
        return v

    def __call__(self, ta, tb):
        return self.w(tb) - self.w(ta)


class BatchedBrownianTree:
    """sampling.py:65-89: torchsde.BrownianTree per seed when torchsde is importable, else `BrownianPath`."""

    def __init__(self, x, t0, t1, seed=None, **kwargs):
        t0, t1, self.sign = self.sort(t0, t1)
        try:
            import torchsde
            if not hasattr(torchsde, "BrownianTree"):
                raise ImportError("stub")
        except ImportError:
            torchsde = None
        self.batched = False
        if torchsde is None:
            self.trees = [BrownianPath(x, t0, t1, seed if not isinstance(seed, (list, tuple)) else seed[0])]
            return
        w0 = kwargs.get('w0', torch.zeros_like(x))
        if seed is None:
            seed = torch.randint(0, 2 ** 63 - 1, []).item()
        self.batched = True
        try:
            assert len(seed) == x.shape[0]
            w0 = w0[0]
        except TypeError:
            seed = [seed]
            self.batched = False
        self.trees = [torchsde.BrownianTree(t0, w0, t1, entropy=s, **kwargs) for s in seed]

    @staticmethod
    def sort(a, b):
        return (a, b, 1) if a < b else (b, a, -1)

    def __call__(self, t0, t1):
        t0, t1, sign = self.sort(t0, t1)
        w = torch.stack([tree(t0, t1) for tree in self.trees]) * (self.sign * sign)
        return w if self.batched else w[0]


class BrownianTreeNoiseSampler:
    """sampling.py:92-115: unit-variance noise for the interval (sigma, sigma_next) from one Brownian path."""

    def __init__(self, x, sigma_min, sigma_max, seed=None, transform=lambda x: x):
        self.transform = transform
        t0, t1 = self.transform(torch.as_tensor(sigma_min)), self.transform(torch.as_tensor(sigma_max))
        self.tree = BatchedBrownianTree(x, t0, t1, seed)

    def __call__(self, sigma, sigma_next):
        t0, t1 = self.transform(torch.as_tensor(sigma)), self.transform(torch.as_tensor(sigma_next))
        return self.tree(t0, t1) / (t1 - t0).abs().sqrt()


def _sde_prep(x, sigmas, noise_sampler):
    if noise_sampler is not None:
        return noise_sampler
    pos = sigmas[sigmas > 0]
    return BrownianTreeNoiseSampler(x, pos.min().cpu(), sigmas.max().cpu())


def _noise(noise_sampler, a, b, x):
    return noise_sampler(a.to(x.device), b.to(x.device)).float().contiguous()


@torch.no_grad()
def sample_dpmpp_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                     noise_sampler=None, r=1 / 2):
    """DPM-Solver++ (stochastic)."""
    noise_sampler = _sde_prep(x, sigmas, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        if sig[i + 1] == 0:
            d = _to_d(x, float(sig[i]), denoised)                          # Euler method
            x = ops.axpby(x, 1.0, d, float(sig[i + 1] - sig[i]))
        else:
            t, t_next = t_fn(sig[i]), t_fn(sig[i + 1])
            h = t_next - t
            s = t + h * r
            fac = 1 / (2 * r)
            # Step 1
            sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(s), eta)
            s_ = t_fn(sd)
            x_2 = _lin((sigma_fn(s_) / sigma_fn(t), x), (-(t - s_).expm1(), denoised),
                       (s_noise * su, _noise(noise_sampler, sigma_fn(t), sigma_fn(s), x)))
            denoised_2 = model(x_2.to(dt_x), sigma_fn(s).to(x.device) * s_in, **extra_args).float().contiguous()
            # Step 2
            sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(t_next), eta)
            t_next_ = t_fn(sd)
            denoised_d = ops.axpby(denoised, float(1 - fac), denoised_2, float(fac))
            x = _lin((sigma_fn(t_next_) / sigma_fn(t), x), (-(t - t_next_).expm1(), denoised_d),
                     (s_noise * su, _noise(noise_sampler, sigma_fn(t), sigma_fn(t_next), x)))
    return x.to(dt_x)


@torch.no_grad()
def sample_dpmpp_2m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                        noise_sampler=None, solver_type='midpoint'):
    """DPM-Solver++(2M) SDE."""
    if solver_type not in {'heun', 'midpoint'}:
        raise ValueError('solver_type must be \'heun\' or \'midpoint\'')
    noise_sampler = _sde_prep(x, sigmas, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    old_denoised = None
    h_last = None
    h = None
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        if sig[i + 1] == 0:
            x = denoised                                                    # Denoising step
        else:
            t, s = -sig[i].log(), -sig[i + 1].log()
            h = s - t
            eta_h = eta * h
            c_den = (-h - eta_h).expm1().neg()
            terms = [(sig[i + 1] / sig[i] * (-eta_h).exp(), x)]
            if old_denoised is not None:
                r = h_last / h
                if solver_type == 'heun':
                    c = ((-h - eta_h).expm1().neg() / (-h - eta_h) + 1) * (1 / r)
                else:
                    c = 0.5 * (-h - eta_h).expm1().neg() * (1 / r)
                terms += [(c_den + c, denoised), (-c, old_denoised)]        # c_den * den + c * (den - old)
            else:
                terms += [(c_den, denoised)]
            if eta:
                terms += [(sig[i + 1] * (-2 * eta_h).expm1().neg().sqrt() * s_noise,
                           _noise(noise_sampler, sig[i], sig[i + 1], x))]
            x = _lin(*terms)
        old_denoised = denoised
        h_last = h
    return x.to(dt_x)


@torch.no_grad()
def sample_dpmpp_3m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                        noise_sampler=None):
    """DPM-Solver++(3M) SDE."""
    noise_sampler = _sde_prep(x, sigmas, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    _, sig, _ = _prep(model, x, sigmas, extra_args, callback)
    dt_x = x.dtype
    x = x.float().contiguous()
    s_in = x.new_ones([x.shape[0]])
    denoised_1, denoised_2 = None, None
    h, h_1, h_2 = None, None, None
    for i in trange(len(sigmas) - 1, disable=disable):
        denoised = _call(model, x, dt_x, sigmas[i], s_in, extra_args, callback, i, sigmas)
        if sig[i + 1] == 0:
            x = denoised                                                    # Denoising step
        else:
            t, s = -sig[i].log(), -sig[i + 1].log()
            h = s - t
            h_eta = h * (eta + 1)
            # x = exp(-h_eta) x + (1 - exp(-h_eta)) den + phi_2 d1 - phi_3 d2, with d1 / d2 linear in the three
            # denoised tensors: fold everything into one coefficient per tensor (the reference's own scalars, :689-704)
            cx, c0, c1, c2 = torch.exp(-h_eta), (-h_eta).expm1().neg(), 0.0, 0.0
            if h_2 is not None:
                r0 = h_1 / h
                r1 = h_2 / h
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                phi_3 = phi_2 / h_eta - 0.5
                # d1_0 = (den - den1)/r0 ; d1_1 = (den1 - den2)/r1 ; d1 = d1_0 + (d1_0 - d1_1) r0/(r0+r1) ; d2 = (d1_0 - d1_1)/(r0+r1)
                a = phi_2 * (1 + r0 / (r0 + r1)) - phi_3 / (r0 + r1)        # coefficient of d1_0
                b = -phi_2 * r0 / (r0 + r1) + phi_3 / (r0 + r1)              # coefficient of d1_1
                c0, c1, c2 = c0 + a / r0, -a / r0 + b / r1, -b / r1
            elif h_1 is not None:
                r = h_1 / h
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                c0, c1 = c0 + phi_2 / r, -phi_2 / r
            terms = [(cx, x), (c0, denoised)]
            if denoised_1 is not None and float(c1) != 0.0:
                terms.append((c1, denoised_1))
            if denoised_2 is not None and float(c2) != 0.0:
                terms.append((c2, denoised_2))
            if eta:
                terms.append((sig[i + 1] * (-2 * h * eta).expm1().neg().sqrt() * s_noise,
                              _noise(noise_sampler, sig[i], sig[i + 1], x)))
            x = _lin(*terms)
        denoised_1, denoised_2 = denoised, denoised_1
        h_1, h_2 = h, h_1
    return x.to(dt_x)
