"""Hires-fix glue of the reference's SD1.5 generator (latent upscaler mode), mirrored on the CUDA kernels:
`modules/sd/image_generator.py:969-999` (bilinear x2 of the latents, `t_enc = int(strength * steps)`) followed by
`img2img_sampling` (`image_generator.py:147-248`, non-ControlNet branches).  The second pass runs the UNet at twice
the latent resolution (128x128 -> 16 384-token self-attention at the top level for 512 -> 1024 images)."""
from __future__ import annotations

import torch

from . import ops
from .ldm.models.diffusion.ddim import DDIMSampler


@torch.no_grad()
def hires_fix_latent(sampler, samples: torch.Tensor, c, uc, scale: float, sampling_steps: int, strength: float,
                     upscale_factor: int = 2, eta: float = 0.0, noise: torch.Tensor = None) -> torch.Tensor:
    """samples: first-pass latents [b, 4, h, w] -> second-pass latents [b, 4, h*f, w*f]."""
    b = samples.shape[0]
    up = ops.bilinear_upsample(samples, upscale_factor)                       # image_generator.py:975
    t_enc = int(strength * sampling_steps)                                    # :978
    shape = (up.shape[1], up.shape[2], up.shape[3])
    if isinstance(sampler, DDIMSampler):
        sampler.make_schedule(ddim_num_steps=sampling_steps, ddim_eta=eta, verbose=False)   # :971-972
        z_enc = sampler.stochastic_encode(up, torch.tensor([t_enc] * b, device=up.device), noise=noise)
        return sampler.decode(z_enc, c, t_enc, unconditional_guidance_scale=scale, unconditional_conditioning=uc)
    z_enc = sampler.stochastic_encode(up, torch.tensor([t_enc] * b, device=up.device), sampling_steps=sampling_steps,
                                      noise=noise)
    out, _ = sampler.sample(S=sampling_steps, conditioning=c, batch_size=b, shape=shape, verbose=False,
                            unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=eta, x0=z_enc,
                            denoising_steps=t_enc)
    return out
