"""Inference-only mirror of the parts of the reference's `ldm.models.diffusion.ddpm` that the denoising path touches
(HowToSD/cremage modules/ldm/models/diffusion/ddpm.py): LatentDiffusion (:446) -- register_schedule (:134-186),
apply_model (:926-1039, live branch :1031-1034), decode_first_stage (:741-798) -- and DiffusionWrapper (:1501-1531,
'crossattn' conditioning).  No Lightning, no training code, no text encoder: `context` tensors are inputs.

State-dict layout matches SD checkpoints: `model.diffusion_model.*`, `first_stage_model.*`, schedule buffers.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..autoencoder import AutoencoderKL
from ...modules.diffusionmodules.openaimodel import UNetModel
from ...modules.diffusionmodules.util import make_beta_schedule


class DiffusionWrapper(nn.Module):
    def __init__(self, diffusion_model, conditioning_key="crossattn"):
        super().__init__()
        self.diffusion_model = diffusion_model
        self.conditioning_key = conditioning_key
        if conditioning_key != "crossattn":
            raise NotImplementedError("cremage_b200: only 'crossattn' conditioning is on the SD1.5 path")

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None):
        # (a one-element list is passed through as is: torch.cat would hand the UNet a fresh copy every step and defeat
        # its per-context K/V cache; the values are the same)
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, t, context=cc)


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config, first_stage_config=None, timesteps=1000, linear_start=0.00085, linear_end=0.012,
                 beta_schedule="linear", scale_factor=0.18215, conditioning_key="crossattn", parameterization="eps",
                 **ignored):
        super().__init__()
        unet = unet_config if isinstance(unet_config, nn.Module) else UNetModel(**unet_config)
        self.model = DiffusionWrapper(unet, conditioning_key)
        if first_stage_config is not None:
            self.first_stage_model = first_stage_config if isinstance(first_stage_config, nn.Module) \
                else AutoencoderKL(**first_stage_config)
        else:
            self.first_stage_model = None
        self.parameterization = parameterization
        self.scale_factor = scale_factor
        self.register_schedule(beta_schedule, timesteps, linear_start, linear_end)

    def register_schedule(self, beta_schedule="linear", timesteps=1000, linear_start=1e-4, linear_end=2e-2,
                          cosine_s=8e-3):
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end,
                                   cosine_s=cosine_s)
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        to_torch = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer('betas', to_torch(betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod)))

    @property
    def device(self):
        return self.betas.device

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """ddpm.py:296-299: sqrt(abar_t) x_0 + sqrt(1 - abar_t) noise, t a per-sample index tensor."""
        from .... import ops
        if noise is None:
            noise = torch.randn_like(x_start)
        idx = torch.as_tensor(t).reshape(-1).tolist()
        if len(idx) == 1:
            idx = idx * x_start.shape[0]
        sa, so = self.sqrt_alphas_cumprod.cpu(), self.sqrt_one_minus_alphas_cumprod.cpu()
        x0f, nf = x_start.float().contiguous(), noise.float().contiguous()
        if all(i == idx[0] for i in idx):
            out = ops.axpby(x0f, float(sa[idx[0]]), nf, float(so[idx[0]]))
        else:
            out = torch.cat([ops.axpby(x0f[j:j + 1].contiguous(), float(sa[i]), nf[j:j + 1].contiguous(), float(so[i]))
                             for j, i in enumerate(idx)])
        return out.to(x_start.dtype)

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if isinstance(cond, dict):
            pass
        else:
            if not isinstance(cond, list):
                cond = [cond]
            cond = {'c_crossattn': cond}
        return self.model(x_noisy, t, **cond)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False, to_uint8=False):
        if self.first_stage_model is None:
            raise RuntimeError("LatentDiffusion was built without a first stage model")
        return self.first_stage_model.decode_first_stage(z, self.scale_factor, to_uint8=to_uint8)

    @torch.no_grad()
    def encode_first_stage(self, x):
        """ddpm.py:861-899 (live branch :899: self.first_stage_model.encode(x)): the first stage's posterior."""
        if self.first_stage_model is None:
            raise RuntimeError("LatentDiffusion was built without a first stage model")
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior, noise=None):
        """ddpm.py:575-582: scale_factor * posterior.sample() (a tensor passes through scaled)."""
        if isinstance(encoder_posterior, torch.Tensor):
            return self.scale_factor * encoder_posterior
        return encoder_posterior.sample(noise=noise, scale=self.scale_factor)

    def get_learned_conditioning(self, c):
        raise NotImplementedError("cremage_b200: the text encoder is outside the hot-path scope; pass context tensors")
