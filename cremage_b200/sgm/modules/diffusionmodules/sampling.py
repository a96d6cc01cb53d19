"""Mirror of the sgm samplers on the SDXL path (modules/sdxl/sgm/modules/diffusionmodules/sampling.py):
BaseDiffusionSampler (:28-122: prepare_sampling_loop with x *= sqrt(1 + sigma_0^2), denoise through the guider),
EulerEDMSampler / HeunEDMSampler (:147-220,309-358, incl. s_churn > 0), LinearMultistepSampler (:271-306),
EulerAncestralSampler (:361-385), DPMPP2SAncestralSampler (:384-457), DPMPP2MSampler (:459-573).
`sampler(denoiser, x, cond, uc, num_steps)` as in the reference; `denoiser(input, sigma, c)` is opaque, the latent
update of each step is one fused kernel.  Step multipliers use the reference's fp32 torch expressions."""
from typing import Dict, Union

import torch
from tqdm import tqdm

from .... import ops
from ...util import default, instantiate_from_config
from .sampling_utils import get_ancestral_step, linear_multistep_coeff, to_neg_log_sigma, to_sigma

DEFAULT_GUIDER = {"target": "sgm.modules.diffusionmodules.guiders.IdentityGuider"}


class BaseDiffusionSampler:
    def __init__(self, discretization_config: Union[Dict, object], num_steps: Union[int, None] = None,
                 guider_config: Union[Dict, object, None] = None, verbose: bool = False, device: str = "cuda"):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(default(guider_config, DEFAULT_GUIDER))
        self.verbose = verbose
        self.device = device

    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device="cpu")
        uc = default(uc, cond)
        # reference: x *= torch.sqrt(1.0 + sigmas[0] ** 2.0)   (:83)
        x = ops.axpby(x.float().contiguous(), float(torch.sqrt(1.0 + sigmas[0] ** 2.0)))
        num_sigmas = len(sigmas)
        s_in = x.new_ones([x.shape[0]])
        return x, s_in, sigmas, num_sigmas, cond, uc

    def denoise(self, x, denoiser, sigma, cond, uc):
        denoised = denoiser(*self.guider.prepare_inputs(x, sigma, cond, uc))
        return self.guider(denoised, sigma)

    def get_sigma_gen(self, num_sigmas):
        gen = range(num_sigmas - 1)
        if self.verbose:
            gen = tqdm(gen, total=num_sigmas - 1, desc=f"Sampling with {self.__class__.__name__}")
        return gen


class EDMSampler(BaseDiffusionSampler):
    """:147-220.  gamma > 0 (s_churn) adds `randn_like(x) * s_noise * sqrt(sigma_hat^2 - sigma^2)` in front of the step
    and evaluates the denoiser at sigma_hat; like the reference the draw only happens when gamma > 0."""

    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise

    def _churn(self, x, sigmas, i, num_sigmas):
        gamma = min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1) if self.s_tmin <= sigmas[i] <= self.s_tmax else 0.0
        sigma_hat = sigmas[i] * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x)
            x = ops.axpby(x, 1.0, eps.float().contiguous(), float(self.s_noise * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5))
        return x, sigma_hat


class EulerEDMSampler(EDMSampler):
    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        for i in self.get_sigma_gen(num_sigmas):
            x, sigma_hat = self._churn(x, sigmas, i, num_sigmas)
            denoised = self.denoise(x, denoiser, s_in * sigma_hat.to(x.device), cond, uc)
            # d = (x - denoised) / sigma_hat ; x + d * (next_sigma - sigma_hat)      (:165-193)
            x, _ = ops.step_euler_ancestral(x, None, None, 0.0, float(sigma_hat), float(sigmas[i + 1]), 0.0,
                                            denoised=denoised.float())
        return x


class HeunEDMSampler(EDMSampler):
    """EDMSampler.sampler_step + HeunEDMSampler.possible_correction_step (:165-193,321-358): Euler predictor, one more
    denoiser call at next_sigma, x + (d + d_new) / 2 * dt where next_sigma > 0.  Two UNet evaluations per step."""

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        for i in self.get_sigma_gen(num_sigmas):
            x, sigma_hat = self._churn(x, sigmas, i, num_sigmas)
            sigma, nxt = float(sigma_hat), float(sigmas[i + 1])
            denoised = self.denoise(x, denoiser, s_in * sigma_hat.to(x.device), cond, uc).float().contiguous()
            d = ops.axpby(x, 1.0 / sigma, denoised, -1.0 / sigma)             # to_d: (x - denoised) / sigma_hat
            dt = nxt - sigma
            euler = ops.axpby(x, 1.0, d, dt)
            if float(torch.sum(s_in.cpu() * sigmas[i + 1])) < 1e-14:           # all noise levels 0: no correction (:335-337)
                x = euler
                continue
            denoised2 = self.denoise(euler, denoiser, s_in * dev_sig[i + 1], cond, uc).float().contiguous()
            d_new = ops.axpby(euler, 1.0 / nxt, denoised2, -1.0 / nxt)
            d_prime = ops.axpby(d, 0.5, d_new, 0.5)
            x = ops.axpby(x, 1.0, d_prime, dt) if nxt > 0.0 else euler        # torch.where(next_sigma > 0, ...)  (:354-356)
        return x


class LinearMultistepSampler(BaseDiffusionSampler):
    """:271-306: x + sum_j coeff_j * d_{i-j}, coefficients by host quadrature (linear_multistep_coeff)."""

    def __init__(self, order=4, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.order = order

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, **kwargs):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        ds = []
        sigmas_cpu = sigmas.detach().cpu().numpy()
        for i in self.get_sigma_gen(num_sigmas):
            sigma = float(sigmas[i])
            denoised = denoiser(*self.guider.prepare_inputs(x, s_in * dev_sig[i], cond, uc), **kwargs)
            denoised = self.guider(denoised, s_in * dev_sig[i]).float().contiguous()
            ds.append(ops.axpby(x, 1.0 / sigma, denoised, -1.0 / sigma))
            if len(ds) > self.order:
                ds.pop(0)
            cur_order = min(i + 1, self.order)
            coeffs = [linear_multistep_coeff(cur_order, sigmas_cpu, i, j) for j in range(cur_order)]
            acc = None   # the reference's sum() adds left to right starting from 0
            for coeff, d in zip(coeffs, reversed(ds)):
                acc = ops.axpby(d, float(coeff)) if acc is None else ops.axpby(acc, 1.0, d, float(coeff))
            x = ops.axpby(x, 1.0, acc, 1.0)
        return x


class AncestralSampler(BaseDiffusionSampler):
    """:222-268.  `ancestral_step` calls `noise_sampler(x)` on EVERY step -- also the last one, where torch.where
    discards it -- so the global RNG advances exactly as in the reference."""

    def __init__(self, eta=1.0, s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.eta, self.s_noise = eta, s_noise
        self.noise_sampler = lambda x: torch.randn_like(x)

    def _noise(self, x, next_sigma):
        noise = self.noise_sampler(x)                       # drawn unconditionally (:246-250)
        return noise.float().contiguous() if float(next_sigma) > 0.0 else None


class EulerAncestralSampler(AncestralSampler):
    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        for i in self.get_sigma_gen(num_sigmas):
            sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=self.eta)
            denoised = self.denoise(x, denoiser, s_in * dev_sig[i], cond, uc)
            # euler step to sigma_down, then x + noise * s_noise * sigma_up where next_sigma > 0   (:338-358,374-383)
            noise = self._noise(x, sigmas[i + 1])
            x, _ = ops.step_euler_ancestral(x, None, noise, 0.0, float(sigmas[i]), float(sigma_down),
                                            float(sigma_up) * self.s_noise, denoised=denoised.float())
        return x


class DPMPP2SAncestralSampler(AncestralSampler):
    """:384-457: Euler-ancestral skeleton with a DPM-Solver++(2S) midpoint correction (a second denoiser call at
    sigma(s), s = t + h/2 in -log sigma) wherever sigma_down > 0."""

    def get_variables(self, sigma, sigma_down):
        t, t_next = [to_neg_log_sigma(s) for s in (sigma, sigma_down)]
        h = t_next - t
        s = t + 0.5 * h
        return h, s, t, t_next

    def get_mult(self, h, s, t, t_next):
        mult1 = to_sigma(s) / to_sigma(t)
        mult2 = (-0.5 * h).expm1()
        mult3 = to_sigma(t_next) / to_sigma(t)
        mult4 = (-h).expm1()
        return mult1, mult2, mult3, mult4

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        for i in self.get_sigma_gen(num_sigmas):
            sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=self.eta)
            denoised = self.denoise(x, denoiser, s_in * dev_sig[i], cond, uc).float().contiguous()
            if float(torch.sum(s_in.cpu() * sigma_down)) < 1e-14:            # all noise levels 0: plain Euler step (:424-426)
                x, _ = ops.step_euler_ancestral(x, None, None, 0.0, float(sigmas[i]), float(sigma_down), 0.0,
                                                denoised=denoised)
            else:
                h, s, t, t_next = self.get_variables(sigmas[i], sigma_down)
                m1, m2, m3, m4 = self.get_mult(h, s, t, t_next)
                x2 = ops.axpby(x, float(m1), denoised, -float(m2))
                denoised2 = self.denoise(x2, denoiser, s_in * to_sigma(s).to(x.device), cond, uc).float().contiguous()
                x = ops.axpby(x, float(m3), denoised2, -float(m4))
            noise = self._noise(x, sigmas[i + 1])
            if noise is not None:
                x = ops.axpby(x, 1.0, noise, float(self.s_noise * sigma_up))
        return x


class DPMPP2MSampler(BaseDiffusionSampler):
    def get_variables(self, sigma, next_sigma, previous_sigma=None):
        t, t_next = [to_neg_log_sigma(s) for s in (sigma, next_sigma)]
        h = t_next - t
        if previous_sigma is not None:
            h_last = t - to_neg_log_sigma(previous_sigma)
            r = h_last / h
            return h, r, t, t_next
        return h, None, t, t_next

    def get_mult(self, h, r, t, t_next, previous_sigma):
        mult1 = to_sigma(t_next) / to_sigma(t)
        mult2 = (-h).expm1()
        if previous_sigma is not None:
            mult3 = 1 + 1 / (2 * r)
            mult4 = 1 / (2 * r)
            return mult1, mult2, mult3, mult4
        return mult1, mult2

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, **kwargs):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        old_denoised = None
        for i in self.get_sigma_gen(num_sigmas):
            prev = None if i == 0 else sigmas[i - 1]
            denoised = self.denoise(x, denoiser, s_in * dev_sig[i], cond, uc).float().contiguous()
            h, r, t, t_next = self.get_variables(sigmas[i], sigmas[i + 1], prev)
            mult = self.get_mult(h, r, t, t_next, prev)
            # x_standard = m0 x - m1 denoised ; x_advanced = m0 x - m1 (m2 denoised - m3 old) where next_sigma > 0  (:527-546)
            advanced = old_denoised is not None and float(sigmas[i + 1]) > 0.0 and float(torch.sum(sigmas[i + 1])) >= 1e-14
            if advanced:
                x, _ = ops.step_dpmpp_2m(x, None, old_denoised, 0.0, float(sigmas[i]), float(mult[0]), float(mult[1]),
                                         float(mult[2]), float(mult[3]), denoised=denoised, want_denoised=False)
            else:
                x, _ = ops.step_dpmpp_2m(x, None, None, 0.0, float(sigmas[i]), float(mult[0]), float(mult[1]), 1.0, 0.0,
                                         denoised=denoised, want_denoised=False)
            old_denoised = denoised
        return x
