"""Drop-in mirror of the reference's `LDMWrapperForKDiffusion` (HowToSD/cremage
modules/ldm/models/diffusion/ldm_wrapper_for_k_diffusion.py:30-106): classifier-free guidance around a
CompVisDenoiser -- batch doubling with the UNCONDITIONAL half first (:67-92) and `uncond + s * (cond - uncond)` on the
denoised predictions (:99).

Besides the reference's `apply_model` / `forward`, it offers the samplers in cremage_b200.k_diffusion.sampling a fused
protocol (`cb_plan`, `cb_fused_eps`): the sampler gets the raw CFG-doubled eps and folds the CompVis c_out step, the
guidance mix and its own update into one kernel.
"""
from __future__ import annotations

import torch

from .... import ops


class LDMWrapperForKDiffusion(torch.nn.Module):
    def __init__(self, compviz_wrapper_model, c, unconditional_conditioning, unconditional_guidance_scale):
        super().__init__()
        self.compviz_model = compviz_wrapper_model
        self.alphas_cumprod = compviz_wrapper_model.inner_model.alphas_cumprod
        self.ddpm_num_timesteps = compviz_wrapper_model.inner_model.num_timesteps
        self.c = c
        self.unconditional_conditioning = unconditional_conditioning
        self.unconditional_guidance_scale = unconditional_guidance_scale
        self._cc = None

    # -- reference API ---------------------------------------------------------------------------------------------
    def apply_model(self, x, t, **kwargs):
        """x: noisy latent, t: per-sample sigma. Returns the guided denoised prediction."""
        c = self.c
        uc = self.unconditional_conditioning
        scale = self.unconditional_guidance_scale
        if uc is None or scale == 1.:
            return self.compviz_model(x, t, c)
        x_in = torch.cat([x] * 2)
        t_in = torch.cat([t] * 2)
        if isinstance(c, dict):
            assert isinstance(uc, dict)
            c_in = dict()
            for k in c:
                if isinstance(c[k], list):
                    c_in[k] = [torch.cat([uc[k][i], c[k][i]]) for i in range(len(c[k]))]
                else:
                    c_in[k] = torch.cat([uc[k], c[k]])
        else:
            c_in = torch.cat([uc, c])
        c_in = {"cond": {"c_crossattn": [c_in]}}
        d_uncond, d_cond = self.compviz_model(x_in, t_in, **c_in).chunk(2)
        return ops.cfg_mix(d_uncond.float().contiguous(), d_cond.float().contiguous(), float(scale)).to(x.dtype)

    def forward(self, *args, **kwargs):
        return self.apply_model(*args, **kwargs)

    # -- fused protocol --------------------------------------------------------------------------------------------
    @property
    def cb_fused_eps(self):
        uc, scale = self.unconditional_conditioning, self.unconditional_guidance_scale
        if uc is None or scale == 1. or isinstance(self.c, dict) or getattr(self.compviz_model, "quantize", False):
            return None
        return self._fused_eps

    def cb_plan(self, sigmas_cpu, x):
        """Per-step scalars that depend on the schedule only: c_in (external.py:97-100) and the (fractional) timestep
        rows sigma_to_t(sigma) (external.py:66-78) for the doubled batch, computed with the denoiser's own methods."""
        den = self.compviz_model
        _, c_in = den.get_scalings(sigmas_cpu)
        t = den.sigma_to_t(sigmas_cpu.to(den.log_sigmas.device)).to(device=x.device, dtype=torch.float32)
        t_rows = t[:, None].expand(t.shape[0], 2 * x.shape[0]).contiguous()
        if self._cc is None or self._cc.device != x.device:
            self._cc = torch.cat([self.unconditional_conditioning, self.c]).to(x.device)
        return {"c_in": c_in.tolist(), "t_rows": t_rows}

    def _fused_eps(self, x, plan, i):
        x_in = ops.cfg_scale_input(x, plan["c_in"][i])
        eps2 = self.compviz_model.inner_model.apply_model(x_in, plan["t_rows"][i], cond={"c_crossattn": [self._cc]})
        return eps2.float().contiguous(), float(self.unconditional_guidance_scale)
