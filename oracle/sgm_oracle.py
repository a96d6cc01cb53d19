"""ORACLE (test infrastructure) -- fp32 PyTorch restatement of the SDXL side of the path (the reference's vendored
`sgm`, modules/sdxl/sgm): UNetModel with vector conditioning / per-level transformer depth / linear projections,
DiscreteDenoiser + EpsScaling + VanillaCFG, EDM / legacy-DDPM discretizations and the DPM++ 2M sampler.
Pinned like oracle/sd_oracle.py: goldens generated from the unmodified reference by oracle/make_golden_sgm.py,
checked in tests/test_sgm.py.  Citations relative to modules/sdxl/."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import sd_oracle as O

Tensor = torch.Tensor


@dataclass
class SgmUNetConfig:  # configs/inference/sd_xl_base.yaml:17-33
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Tuple[int, ...] = (4, 2)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4)
    num_head_channels: int = 64
    transformer_depth: Tuple[int, ...] = (1, 2, 10)
    context_dim: int = 2048
    adm_in_channels: int = 2816


SDXL_UNET = SgmUNetConfig()
TINY_SGM_UNET = SgmUNetConfig(model_channels=64, attention_resolutions=(2,), num_res_blocks=1, channel_mult=(1, 2),
                              num_head_channels=32, transformer_depth=(1, 2), context_dim=64, adm_in_channels=96)


def sgm_unet_layout(cfg: SgmUNetConfig):
    """sgm/modules/diffusionmodules/openaimodel.py:632-826: blocks of ('conv'|'res'|'st'|'down'|'up', cin, cout|level)."""
    mc = cfg.model_channels
    inp = [[("conv", cfg.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    nlev = len(cfg.channel_mult)
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(("st", ch, level))
            inp.append(layers)
            chans.append(ch)
        if level != nlev - 1:
            inp.append([("down", ch, ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("st", ch, nlev - 1), ("res", ch, ch)]
    out = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(("st", ch, level))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", ch, ch))
                ds //= 2
            out.append(layers)
    return inp, mid, out


def sgm_unet_param_shapes(cfg: SgmUNetConfig) -> Dict[str, Tuple[int, ...]]:
    mc, ted = cfg.model_channels, cfg.model_channels * 4
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, i, o, bias=True):
        s[p + ".weight"] = (o, i)
        if bias:
            s[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".in_layers.0", cin)
        conv(p + ".in_layers.2", cin, cout, 3)
        lin(p + ".emb_layers.1", ted, cout)
        norm(p + ".out_layers.0", cout)
        conv(p + ".out_layers.3", cout, cout, 3)
        if cin != cout:
            conv(p + ".skip_connection", cin, cout, 1)

    def st(p, c, level):
        inner = c  # heads * d_head with heads = c // num_head_channels
        norm(p + ".norm", c)
        lin(p + ".proj_in", c, inner)     # use_linear_in_transformer: True
        for d in range(cfg.transformer_depth[level]):
            b = f"{p}.transformer_blocks.{d}"
            for a, cd in (("attn1", inner), ("attn2", cfg.context_dim)):
                lin(f"{b}.{a}.to_q", inner, inner, bias=False)
                lin(f"{b}.{a}.to_k", cd, inner, bias=False)
                lin(f"{b}.{a}.to_v", cd, inner, bias=False)
                lin(f"{b}.{a}.to_out.0", inner, inner)
            lin(f"{b}.ff.net.0.proj", inner, inner * 8)
            lin(f"{b}.ff.net.2", inner * 4, inner)
            for nrm in ("norm1", "norm2", "norm3"):
                norm(f"{b}.{nrm}", inner)
        lin(p + ".proj_out", inner, c)

    def block(p, layers):
        for j, (kind, cin, cout) in enumerate(layers):
            q = f"{p}.{j}"
            if kind == "conv":
                conv(q, cin, cout, 3)
            elif kind == "res":
                res(q, cin, cout)
            elif kind == "st":
                st(q, cin, cout)
            elif kind == "down":
                conv(q + ".op", cin, cout, 3)
            elif kind == "up":
                conv(q + ".conv", cin, cout, 3)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    lin("label_emb.0.0", cfg.adm_in_channels, ted)
    lin("label_emb.0.2", ted, ted)
    inp, mid, out = sgm_unet_layout(cfg)
    for i, layers in enumerate(inp):
        block(f"input_blocks.{i}", layers)
    block("middle_block", mid)
    for i, layers in enumerate(out):
        block(f"output_blocks.{i}", layers)
    norm("out.0", mc)
    conv("out.2", mc, cfg.out_channels, 3)
    return s


def _sgm_spatial_transformer(sd, p: str, x: Tensor, context: Tensor, heads: int, depth: int) -> Tensor:
    """sgm/modules/attention.py:1068-1133 with use_linear=True: norm -> rearrange -> Linear proj_in -> blocks ->
    Linear proj_out -> rearrange -> + x_in."""
    b, c, h, w = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = F.linear(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    for d in range(depth):
        x = O._transformer_block(sd, f"{p}.transformer_blocks.{d}", x, context, heads)
    x = F.linear(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x + x_in


def sgm_unet_forward(sd, cfg: SgmUNetConfig, x: Tensor, timesteps: Tensor, context: Tensor, y: Tensor) -> Tensor:
    """sgm UNetModel.forward, openaimodel.py:828-874."""
    inp, mid, out = sgm_unet_layout(cfg)
    t_emb = O.timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    lab = F.linear(y, sd["label_emb.0.0.weight"], sd["label_emb.0.0.bias"])
    lab = F.linear(F.silu(lab), sd["label_emb.0.2.weight"], sd["label_emb.0.2.bias"])
    emb = emb + lab

    def run(prefix, layers, h):
        for j, (kind, cin, cout) in enumerate(layers):
            q = f"{prefix}.{j}"
            if kind == "conv":
                h = F.conv2d(h, sd[q + ".weight"], sd[q + ".bias"], padding=1)
            elif kind == "res":
                h = O._resblock(sd, q, h, emb)
            elif kind == "st":
                h = _sgm_spatial_transformer(sd, q, h, context, cin // cfg.num_head_channels, cfg.transformer_depth[cout])
            elif kind == "down":
                h = F.conv2d(h, sd[q + ".op.weight"], sd[q + ".op.bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[q + ".conv.weight"], sd[q + ".conv.bias"], padding=1)
        return h

    hs: List[Tensor] = []
    h = x
    for i, layers in enumerate(inp):
        h = run(f"input_blocks.{i}", layers, h)
        hs.append(h)
    h = run("middle_block", mid, h)
    for i, layers in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = run(f"output_blocks.{i}", layers, h)
    h = F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5)
    return F.conv2d(F.silu(h), sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ----------------------------------------------------------------------------------------------------------------------
# discretizations, denoiser, guider, sampler
# ----------------------------------------------------------------------------------------------------------------------
def edm_sigmas(n: int, sigma_min=0.002, sigma_max=80.0, rho=7.0) -> Tensor:
    """EDMDiscretization.get_sigmas + append_zero, sgm/modules/diffusionmodules/discretizer.py:28-48,18-22."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def legacy_ddpm_sigma_table(num_idx: int = 1000) -> Tensor:
    """DiscreteDenoiser's table: LegacyDDPMDiscretization(num_idx, do_append_zero=False, flip=True),
    discretizer.py:51-78 + denoiser.py:55-57 -> ascending sigmas."""
    betas = O.make_beta_schedule_linear(1000, 0.00085, 0.0120)
    ac = np.cumprod(1.0 - betas, axis=0)
    sigmas = torch.tensor((1 - ac) / ac, dtype=torch.float32) ** 0.5
    sigmas = torch.flip(sigmas, (0,))      # get_sigmas returns descending
    return torch.flip(sigmas, (0,))        # flip=True in DiscreteDenoiser


def discrete_denoise(net_fn, table: Tensor, x: Tensor, sigma: Tensor, cond: dict) -> Tensor:
    """DiscreteDenoiser.forward with EpsScaling, denoiser.py:23-39,61-75, denoiser_scaling.py:29-37."""
    idx = (sigma - table[:, None]).abs().argmin(dim=0).view(sigma.shape)
    sigma_q = table[idx]
    s = sigma_q[(...,) + (None,) * (x.ndim - sigma_q.ndim)]
    c_skip, c_out, c_in, c_noise = torch.ones_like(s), -s, 1 / (s ** 2 + 1.0) ** 0.5, s.clone()
    c_noise = (c_noise.reshape(sigma.shape) - table[:, None]).abs().argmin(dim=0).view(sigma.shape)
    return net_fn(x * c_in, c_noise, cond) * c_out + x * c_skip


def sample_dpmpp_2m_sgm(denoise_fn, x: Tensor, sigmas: Tensor, cond: dict, uc: dict, scale: float,
                        trace: Optional[List[Tensor]] = None) -> Tensor:
    """DPMPP2MSampler.__call__ with VanillaCFG, sampling.py:49-122,459-573, guiders.py:24-65.
    denoise_fn(x, sigma, cond) -> denoised (the `denoiser` lambda of sdxl_image_generator_utils.py:697-700)."""
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    old = None
    for i in range(len(sigmas) - 1):
        sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
        prev = None if i == 0 else s_in * sigmas[i - 1]
        c_in = {k: torch.cat((uc[k], cond[k]), 0) for k in cond}
        d_u, d_c = denoise_fn(torch.cat([x] * 2), torch.cat([sigma] * 2), c_in).chunk(2)
        den = d_u + scale * (d_c - d_u)
        t, t_next = sigma.log().neg(), nxt.log().neg()
        h = t_next - t
        ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
        m0, m1 = ap(t_next.neg().exp() / t.neg().exp()), ap((-h).expm1())
        x_std = m0 * x - m1 * den
        if old is None or torch.sum(nxt) < 1e-14:
            x = x_std
        else:
            r = (t - prev.log().neg()) / h
            dd = ap(1 + 1 / (2 * r)) * den - ap(1 / (2 * r)) * old
            x = torch.where(ap(nxt) > 0.0, m0 * x - m1 * dd, x_std)
        old = den
        if trace is not None:
            trace.append(x.clone())
    return x


def _cfg_denoise(denoise_fn, x: Tensor, sigma: Tensor, cond: dict, uc: dict, scale: float) -> Tensor:
    """BaseDiffusionSampler.denoise through VanillaCFG: sampling.py:97-122, guiders.py:24-65."""
    c_in = {k: torch.cat((uc[k], cond[k]), 0) for k in cond}
    d_u, d_c = denoise_fn(torch.cat([x] * 2), torch.cat([sigma] * 2), c_in).chunk(2)
    return d_u + scale * (d_c - d_u)


def sample_heun_sgm(denoise_fn, x: Tensor, sigmas: Tensor, cond: dict, uc: dict, scale: float) -> Tensor:
    """HeunEDMSampler with s_churn = 0: EDMSampler.sampler_step / __call__ (sampling.py:165-220) +
    possible_correction_step (:332-358)."""
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
    for i in range(len(sigmas) - 1):
        sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
        den = _cfg_denoise(denoise_fn, x, sigma, cond, uc, scale)
        d = (x - den) / ap(sigma)
        dt = ap(nxt - sigma)
        euler = x + dt * d
        if torch.sum(nxt) < 1e-14:
            x = euler
        else:
            den2 = _cfg_denoise(denoise_fn, euler, nxt, cond, uc, scale)
            d_new = (euler - den2) / ap(nxt)
            d_prime = (d + d_new) / 2.0
            x = torch.where(ap(nxt) > 0.0, x + d_prime * dt, euler)
    return x


def sample_lms_sgm(denoise_fn, x: Tensor, sigmas: Tensor, cond: dict, uc: dict, scale: float, order: int = 4) -> Tensor:
    """LinearMultistepSampler.__call__, sampling.py:282-306; linear_multistep_coeff, sampling_utils.py:7-19."""
    from scipy import integrate

    def coeff(order_, t, i, j):
        def fn(tau):
            prod = 1.0
            for k in range(order_):
                if j == k:
                    continue
                prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
            return prod
        return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]

    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
    ds: List[Tensor] = []
    sig_np = sigmas.detach().cpu().numpy()
    for i in range(len(sigmas) - 1):
        sigma = s_in * sigmas[i]
        den = _cfg_denoise(denoise_fn, x, sigma, cond, uc, scale)
        ds.append((x - den) / ap(sigma))
        if len(ds) > order:
            ds.pop(0)
        cur = min(i + 1, order)
        cs = [coeff(cur, sig_np, i, j) for j in range(cur)]
        x = x + sum(c * d for c, d in zip(cs, reversed(ds)))
    return x


def _sgm_ancestral_step(sigma_from: Tensor, sigma_to: Tensor, eta: float = 1.0):
    """sampling_utils.py:23-39."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = torch.minimum(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    return (sigma_to ** 2 - sigma_up ** 2) ** 0.5, sigma_up


def sample_euler_ancestral_sgm(denoise_fn, x: Tensor, sigmas: Tensor, cond: dict, uc: dict, scale: float,
                               noise, eta: float = 1.0, s_noise: float = 1.0) -> Tensor:
    """EulerAncestralSampler, sampling.py:222-268,361-382; noise[i] = the (always drawn) noise_sampler output of step i."""
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
    for i in range(len(sigmas) - 1):
        sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
        sd, su = _sgm_ancestral_step(sigma, nxt, eta)
        den = _cfg_denoise(denoise_fn, x, sigma, cond, uc, scale)
        x = x + ap(sd - sigma) * ((x - den) / ap(sigma))
        x = torch.where(ap(nxt) > 0.0, x + noise[i] * s_noise * ap(su), x)
    return x


def sample_dpmpp_2s_ancestral_sgm(denoise_fn, x: Tensor, sigmas: Tensor, cond: dict, uc: dict, scale: float,
                                  noise, eta: float = 1.0, s_noise: float = 1.0) -> Tensor:
    """DPMPP2SAncestralSampler, sampling.py:384-457."""
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
    for i in range(len(sigmas) - 1):
        sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
        sd, su = _sgm_ancestral_step(sigma, nxt, eta)
        den = _cfg_denoise(denoise_fn, x, sigma, cond, uc, scale)
        x_euler = x + ap(sd - sigma) * ((x - den) / ap(sigma))
        if torch.sum(sd) < 1e-14:
            x = x_euler
        else:
            t, t_next = sigma.log().neg(), sd.log().neg()
            h = t_next - t
            s = t + 0.5 * h
            x2 = ap(s.neg().exp() / t.neg().exp()) * x - ap((-0.5 * h).expm1()) * den
            den2 = _cfg_denoise(denoise_fn, x2, s.neg().exp(), cond, uc, scale)
            x_d = ap(t_next.neg().exp() / t.neg().exp()) * x - ap((-h).expm1()) * den2
            x = torch.where(ap(sd) > 0.0, x_d, x_euler)
        x = torch.where(ap(nxt) > 0.0, x + noise[i] * s_noise * ap(su), x)
    return x


def sample_euler_edm_sgm(denoise_fn, x: Tensor, sigmas: Tensor, cond: dict, uc: dict, scale: float, noise=None,
                         s_churn: float = 0.0, s_tmin: float = 0.0, s_tmax: float = float("inf"),
                         s_noise: float = 1.0) -> Tensor:
    """EulerEDMSampler incl. s_churn, sampling.py:147-220 (the draw happens only when gamma > 0)."""
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
    it = iter(noise) if noise is not None else None
    n = len(sigmas)
    for i in range(n - 1):
        gamma = min(s_churn / (n - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        sigma, nxt = s_in * sigmas[i], s_in * sigmas[i + 1]
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            x = x + next(it) * s_noise * ap(sigma_hat ** 2 - sigma ** 2) ** 0.5
        den = _cfg_denoise(denoise_fn, x, sigma_hat, cond, uc, scale)
        x = x + ap(nxt - sigma_hat) * ((x - den) / ap(sigma_hat))
    return x
