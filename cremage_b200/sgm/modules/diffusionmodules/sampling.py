"""Mirror of the sgm samplers on the SDXL path (modules/sdxl/sgm/modules/diffusionmodules/sampling.py):
BaseDiffusionSampler (:28-122: prepare_sampling_loop with x *= sqrt(1 + sigma_0^2), denoise through the guider),
EulerEDMSampler / HeunEDMSampler (:147-220,309-358, s_churn = 0), LinearMultistepSampler (:271-306),
EulerAncestralSampler (:361-385), DPMPP2MSampler (:459-573).
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


class EulerEDMSampler(BaseDiffusionSampler):
    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if s_churn != 0.0:
            raise NotImplementedError("cremage_b200: EulerEDMSampler with s_churn > 0 is not implemented")

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        for i in self.get_sigma_gen(num_sigmas):
            denoised = self.denoise(x, denoiser, s_in * dev_sig[i], cond, uc)
            # d = (x - denoised) / sigma ; x + d * (next_sigma - sigma)      (:189-200)
            x, _ = ops.step_euler_ancestral(x, None, None, 0.0, float(sigmas[i]), float(sigmas[i + 1]), 0.0,
                                            denoised=denoised.float())
        return x


class HeunEDMSampler(BaseDiffusionSampler):
    """EDMSampler.sampler_step + HeunEDMSampler.possible_correction_step (:165-193,321-358): Euler predictor, one more
    denoiser call at next_sigma, x + (d + d_new) / 2 * dt where next_sigma > 0.  Two UNet evaluations per step."""

    def __init__(self, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if s_churn != 0.0:
            raise NotImplementedError("cremage_b200: HeunEDMSampler with s_churn > 0 is not implemented")

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        for i in self.get_sigma_gen(num_sigmas):
            sigma, nxt = float(sigmas[i]), float(sigmas[i + 1])
            denoised = self.denoise(x, denoiser, s_in * dev_sig[i], cond, uc).float().contiguous()
            d = ops.axpby(x, 1.0 / sigma, denoised, -1.0 / sigma)             # to_d: (x - denoised) / sigma
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


class EulerAncestralSampler(BaseDiffusionSampler):
    def __init__(self, eta=1.0, s_noise=1.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.eta, self.s_noise = eta, s_noise
        self.noise_sampler = lambda x: torch.randn_like(x)

    @torch.no_grad()
    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        dev_sig = sigmas.to(x.device)
        for i in self.get_sigma_gen(num_sigmas):
            sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=self.eta)
            denoised = self.denoise(x, denoiser, s_in * dev_sig[i], cond, uc)
            # euler step to sigma_down, then x + noise * s_noise * sigma_up where next_sigma > 0   (:338-358,374-383)
            noise = self.noise_sampler(x).float().contiguous() if float(sigmas[i + 1]) > 0.0 else None
            x, _ = ops.step_euler_ancestral(x, None, noise, 0.0, float(sigmas[i]), float(sigma_down),
                                            float(sigma_up) * self.s_noise, denoised=denoised.float())
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
