"""Drop-in mirror of the reference's k-diffusion sampler front ends (HowToSD/cremage
modules/ldm/models/diffusion/k_diffusion_samplers.py): KDiffusionSamplerBase (:63-297) and the subclasses that map
UI names to sampling functions -- EulerSampler (:299), EulerAncestralSampler (:310), Dpmpp2mSampler (:383).

`sample(S, batch_size, shape, conditioning, ..., x0=, unconditional_guidance_scale=, unconditional_conditioning=)`
returns `(x, None)` like the reference.  Like the reference (:166-171) the initial latent is `randn(size)` or the
caller's `x0` as given -- the caller scales by sigma_max where it wants to.
"""
from __future__ import annotations

import numpy as np
import torch

from ....k_diffusion.external import CompVisDenoiser
from ....k_diffusion.sampling import (get_sigmas_karras, sample_dpm_2, sample_dpm_2_ancestral, sample_dpmpp_2m,
                                         sample_dpmpp_2m_sde, sample_dpmpp_2s_ancestral, sample_dpmpp_3m_sde,
                                         sample_dpmpp_sde, sample_euler, sample_euler_ancestral, sample_heun,
                                         sample_lms)
from .ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion


class KDiffusionSamplerBase(object):
    def __init__(self, model, sigma_min=0.0316386, sigma_max=14.5521805, beta_d=19.9, beta_min=0.1, eps_s=1e-3):
        self.ldm_model = model
        self.ddpm_num_timesteps = model.num_timesteps
        alphas_cumprod = self.ldm_model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.beta_d = beta_d
        self.beta_min = beta_min
        self.eps_s = eps_s
        self.device = self.ldm_model.device
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.ldm_model.device)
        self.register_buffer('betas', to_torch(self.ldm_model.betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(self.ldm_model.alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(alphas_cumprod.cpu())))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod.cpu())))

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    @torch.no_grad()
    def compute_sigmas(self, n: int):
        return None

    @torch.no_grad()
    def _sample_common_prep(self, S, batch_size, shape, conditioning=None, x0=None, unconditional_guidance_scale=1.,
                            unconditional_conditioning=None, **kwargs):
        C, H, W = shape
        size = (batch_size, C, H, W)
        self.x = torch.randn(size, device=self.device) if x0 is None else x0
        self.compviz_wrapper_model = CompVisDenoiser(self.ldm_model, False).to(self.device)
        self.ldm_wrapper_model = LDMWrapperForKDiffusion(self.compviz_wrapper_model, conditioning,
                                                         unconditional_conditioning, unconditional_guidance_scale)
        self.sigmas = self.compute_sigmas(S)
        if "denoising_steps" in kwargs:  # partial denoising (img2img): the last t+1 sigmas (:188-194)
            t = kwargs["denoising_steps"]
            self.sigmas = self.sigmas[-(t + 1):]
            assert self.sigmas.shape[0] == t + 1

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        self._sample_common_prep(S=S, batch_size=batch_size, shape=shape, conditioning=conditioning, x0=x0,
                                 unconditional_guidance_scale=unconditional_guidance_scale,
                                 unconditional_conditioning=unconditional_conditioning, **kwargs)
        return self.do_sample()

    @torch.no_grad()
    def do_sample(self):
        return self.x, None

    @torch.no_grad()
    def stochastic_encode(self, x0, t, sampling_steps, noise=None):
        """VP forward noising used by img2img (:260-297): index t*1000/steps into the DDPM tables."""
        if noise is None:
            noise = torch.randn_like(x0)
        from .... import ops
        t = torch.as_tensor(t, device=self.sqrt_alphas_cumprod.device).reshape(-1)
        idx = (t * 1000.0 / sampling_steps).long().tolist()  # (:292) t is a per-sample index tensor
        if len(idx) == 1:
            idx = idx * x0.shape[0]
        x0f, nf = x0.float().contiguous(), noise.float().contiguous()
        if all(i == idx[0] for i in idx):
            out = ops.axpby(x0f, float(self.sqrt_alphas_cumprod[idx[0]]), nf,
                            float(self.sqrt_one_minus_alphas_cumprod[idx[0]]))
        else:
            out = torch.cat([ops.axpby(x0f[j:j + 1].contiguous(), float(self.sqrt_alphas_cumprod[i]),
                                       nf[j:j + 1].contiguous(), float(self.sqrt_one_minus_alphas_cumprod[i]))
                             for j, i in enumerate(idx)])
        return out.to(x0.dtype)


class EulerSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return self.compviz_wrapper_model.get_sigmas(n).to(self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_euler(self.ldm_wrapper_model, self.x, self.sigmas), None


class EulerAncestralSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return self.compviz_wrapper_model.get_sigmas(n).to(self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_euler_ancestral(self.ldm_wrapper_model, self.x, self.sigmas), None


class Dpmpp2mSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpmpp_2m(self.ldm_wrapper_model, self.x, self.sigmas), None


class HeunSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return self.compviz_wrapper_model.get_sigmas(n).to(self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_heun(self.ldm_wrapper_model, self.x, self.sigmas), None


class Dpm2Sampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpm_2(self.ldm_wrapper_model, self.x, self.sigmas), None


class Dpm2AncestralSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpm_2_ancestral(self.ldm_wrapper_model, self.x, self.sigmas), None


class LmsSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return self.compviz_wrapper_model.get_sigmas(n).to(self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_lms(self.ldm_wrapper_model, self.x, self.sigmas), None


class Dpmpp2sAncestralSampler(KDiffusionSamplerBase):
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpmpp_2s_ancestral(self.ldm_wrapper_model, self.x, self.sigmas), None


class DpmppSdeSampler(KDiffusionSamplerBase):                    # :373-381
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpmpp_sde(self.ldm_wrapper_model, self.x, self.sigmas), None


class Dpmpp2mSdeSampler(KDiffusionSamplerBase):                  # :393-401
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpmpp_2m_sde(self.ldm_wrapper_model, self.x, self.sigmas), None


class Dpmpp3mSdeSampler(KDiffusionSamplerBase):                  # :403-411
    @torch.no_grad()
    def compute_sigmas(self, n):
        return get_sigmas_karras(n, self.sigma_min, self.sigma_max, device=self.device)

    @torch.no_grad()
    def do_sample(self):
        return sample_dpmpp_3m_sde(self.ldm_wrapper_model, self.x, self.sigmas), None
