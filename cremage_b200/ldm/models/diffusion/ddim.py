"""Drop-in mirror of the reference's `DDIMSampler` (HowToSD/cremage modules/ldm/models/diffusion/ddim.py:15-676):
make_schedule (:38-75), sample (:78-135), ddim_sampling (:138-190), p_sample_ddim (:193, 530-612),
stochastic_encode (:615), decode (:657) -- same signatures and return values.

Schedule tensors come from the reference's own numpy / torch expressions (bit-exact).  Each step is one UNet call on
the CFG-doubled batch (unconditional half first, :538-561) plus ONE fused kernel doing the guidance mix, pred_x0 and the
x_{t-1} update (:561, 590-611).
"""
from __future__ import annotations

import numpy as np
import torch
from tqdm import tqdm

from .... import ops
from ...modules.diffusionmodules.util import make_ddim_sampling_parameters, make_ddim_timesteps


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def register_buffer(self, name, attr):
        if type(attr) == torch.Tensor and attr.device != self.model.device:
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        self.register_buffer('betas', to_torch(self.model.betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(self.model.alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(alphas_cumprod.cpu())))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod.cpu())))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch(np.log(1. - alphas_cumprod.cpu())))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod.cpu())))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod.cpu() - 1)))
        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=alphas_cumprod.cpu(), ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        # kept on the host: they are per-step scalars (the reference indexes them with a python int, :575-578)
        self.ddim_sigmas = ddim_sigmas
        self.ddim_alphas = ddim_alphas
        self.ddim_alphas_prev = ddim_alphas_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - ddim_alphas)
        sigmas_for_original_sampling_steps = ddim_eta * torch.sqrt(
            (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * (
                    1 - self.alphas_cumprod / self.alphas_cumprod_prev))
        self.register_buffer('ddim_sigmas_for_original_num_steps', sigmas_for_original_sampling_steps)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        if quantize_x0 or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("cremage_b200: DDIM quantize_x0 / score_corrector / noise_dropout are not "
                                      "on the Stable Diffusion path and are not implemented")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback, mask=mask, x0=x0,
                                  ddim_use_original_steps=False, temperature=temperature, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None):
        if ddim_use_original_steps:
            raise NotImplementedError("cremage_b200: ddim_use_original_steps is not implemented")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        cc = self._cfg_cond(cond, unconditional_conditioning, unconditional_guidance_scale)
        # all timestep rows in one host->device copy (the reference builds torch.full(...) every step, :169)
        ts_all = torch.as_tensor(np.ascontiguousarray(time_range), device=device, dtype=torch.long)
        iterator = tqdm(time_range, desc='DDIM Sampler', total=total_steps, disable=not hasattr(tqdm, "_instances") and False)
        out_dtype = img.dtype
        img = img.float().contiguous()
        for i, step in enumerate(iterator):
            index = total_steps - i - 1
            ts = ts_all[i].expand(b)
            if mask is not None:       # inpainting (:171-174): the known region is re-noised to this step's level
                assert x0 is not None
                img_orig = self.model.q_sample(x0, [int(step)])     # host-known step: no device sync
                img = ops.blend_mask(img_orig, img, mask)
            img, pred_x0 = self._p_sample(img, cond, cc, ts, index, temperature, unconditional_guidance_scale,
                                          want_x0=(img_callback is not None) or index % log_every_t == 0
                                          or index == total_steps - 1)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img.to(out_dtype), intermediates

    @staticmethod
    def _cfg_cond(c, uc, scale):
        if uc is None or scale == 1.:
            return None
        if isinstance(c, dict):
            assert isinstance(uc, dict)
            c_in = dict()
            for k in c:
                if isinstance(c[k], list):
                    c_in[k] = [torch.cat([uc[k][i], c[k][i]]) for i in range(len(c[k]))]
                else:
                    c_in[k] = torch.cat([uc[k], c[k]])
            return c_in
        return torch.cat([uc, c])

    def _p_sample(self, x, c, cc, t, index, temperature, scale, want_x0=True, noise=None):
        a_t = torch.tensor(self.ddim_alphas[index], dtype=torch.float32)
        a_prev = torch.tensor(self.ddim_alphas_prev[index], dtype=torch.float32)
        sigma_t = torch.tensor(self.ddim_sigmas[index], dtype=torch.float32)
        sqrt_one_minus_at = torch.tensor(self.ddim_sqrt_one_minus_alphas[index], dtype=torch.float32)
        # the reference evaluates these as fp32 tensor ops (:575-603)
        sqrt_at = float(a_t.sqrt())
        sqrt_aprev = float(a_prev.sqrt())
        dir_coef = float((1. - a_prev - sigma_t ** 2).sqrt())
        sig = float(sigma_t)
        if noise is None:
            # the reference draws noise_like(x.shape) on EVERY step, also when sigma_t == 0 (:597): same RNG consumption
            noise = torch.randn(x.shape, device=x.device)
            if temperature != 1.:
                noise = noise * temperature
        if sig == 0.0:
            noise = None
        if cc is None:
            e_t = self.model.apply_model(x, t, c).float().contiguous()
            return ops.step_ddim(x, None, noise, 0.0, sqrt_at, float(sqrt_one_minus_at), sqrt_aprev, dir_coef, sig,
                                 want_x0=want_x0, eps=e_t)
        x_in = ops.cfg_scale_input(x, 1.0)
        t_in = torch.cat([t] * 2)
        eps2 = self.model.apply_model(x_in, t_in, cc).float().contiguous()
        return ops.step_ddim(x, eps2, noise, float(scale), sqrt_at, float(sqrt_one_minus_at), sqrt_aprev, dir_coef, sig,
                             want_x0=want_x0)

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None):
        """One DDIM step (:193, 530-612). Returns (x_prev, pred_x0)."""
        if use_original_steps or quantize_denoised or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("cremage_b200: unsupported p_sample_ddim option")
        cc = self._cfg_cond(c, unconditional_conditioning, unconditional_guidance_scale)
        xo, x0 = self._p_sample(x.float().contiguous(), c, cc, t, index, temperature, unconditional_guidance_scale)
        return xo.to(x.dtype), x0.to(x.dtype)

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """x_t = sqrt(a_t) x_0 + sqrt(1 - a_t) noise (:615-655); t indexes the DDIM (or DDPM) alpha table."""
        if use_original_steps:
            sa, s1 = self.sqrt_alphas_cumprod.cpu(), self.sqrt_one_minus_alphas_cumprod.cpu()
        else:
            sa = torch.sqrt(torch.as_tensor(self.ddim_alphas, dtype=torch.float32))
            s1 = torch.as_tensor(self.ddim_sqrt_one_minus_alphas, dtype=torch.float32)
        if noise is None:
            noise = torch.randn_like(x0)
        idx = torch.as_tensor(t).reshape(-1).tolist()
        if len(idx) == 1:
            idx = idx * x0.shape[0]
        x0f, nf = x0.float().contiguous(), noise.float().contiguous()
        if all(i == idx[0] for i in idx):
            out = ops.axpby(x0f, float(sa[idx[0]]), nf, float(s1[idx[0]]))
        else:
            out = torch.cat([ops.axpby(x0f[j:j + 1].contiguous(), float(sa[i]), nf[j:j + 1].contiguous(), float(s1[i]))
                             for j, i in enumerate(idx)])
        return out.to(x0.dtype)

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
               use_original_steps=False, callback=None):
        """img2img denoise from step t_start (:657-676)."""
        if use_original_steps:
            raise NotImplementedError("cremage_b200: use_original_steps is not implemented")
        timesteps = self.ddim_timesteps[:t_start]
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        cc = self._cfg_cond(cond, unconditional_conditioning, unconditional_guidance_scale)
        ts_all = torch.as_tensor(np.ascontiguousarray(time_range), device=x_latent.device, dtype=torch.long)
        x_dec = x_latent.float().contiguous()
        for i in range(total_steps):
            index = total_steps - i - 1
            x_dec, _ = self._p_sample(x_dec, cond, cc, ts_all[i].expand(x_latent.shape[0]), index, 1.,
                                      unconditional_guidance_scale, want_x0=False)
            if callback:
                callback(i)
        return x_dec.to(x_latent.dtype)
