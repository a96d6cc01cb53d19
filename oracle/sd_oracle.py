"""ORACLE -- test infrastructure, not product code.

Plain-PyTorch fp32 restatement of the reference's (HowToSD/cremage v4.0.1) Stable Diffusion denoising path:
UNetModel.forward, the k-diffusion / DDIM sampler arithmetic and schedules, and the AutoencoderKL decoder.  It is a
functional re-expression over a state dict that uses the reference's own parameter names, so the same weights load
into the reference modules, this oracle and the CUDA implementation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (cremage_b200/) never does.

Pinning: tests/test_oracle_pin.py checks every function here against the reference's own modules imported from
/root/reference (when present, i.e. in the authoring container) and against golden vectors generated from the
reference by oracle/make_golden.py and committed under tests/golden/.  The reference's own tests hold no numeric
fixture for this path, so those reference-generated goldens are the pin.

All citations are relative to the reference root, `modules/` prefix omitted.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ======================================================================================================================
# configs (configs/ldm/configs/stable-diffusion/v1-inference.yaml:29-67)
# ======================================================================================================================
@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_heads: int = 8
    transformer_depth: int = 1
    context_dim: int = 768


@dataclass
class DecoderConfig:
    ch: int = 128
    out_ch: int = 3
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    resolution: int = 256
    embed_dim: int = 4


SD15_UNET = UNetConfig()
SD15_VAE = DecoderConfig()
TINY_UNET = UNetConfig(model_channels=64, attention_resolutions=(2, 1), num_res_blocks=1, channel_mult=(1, 2),
                       num_heads=2, context_dim=64)
TINY_VAE = DecoderConfig(ch=64, ch_mult=(1, 2), num_res_blocks=1, resolution=32)


# ======================================================================================================================
# block structure (ldm/modules/diffusionmodules/openaimodel.py:548-756)
# ======================================================================================================================
def unet_layout(cfg: UNetConfig):
    """Returns (input_blocks, middle, output_blocks): lists of blocks, each block a list of
    ('conv'|'res'|'st'|'down'|'up', cin, cout) in module order -- the reference's nn.ModuleList indexing."""
    mc = cfg.model_channels
    inp = [[("conv", cfg.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(("st", ch, ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inp.append([("down", ch, ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("st", ch, ch), ("res", ch, ch)]
    out = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(("st", ch, ch))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", ch, ch))
                ds //= 2
            out.append(layers)
    return inp, mid, out


def unet_param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """Every parameter of the reference UNetModel for `cfg`, by reference state-dict key."""
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

    def st(p, c):
        d_head = c // cfg.num_heads
        inner = cfg.num_heads * d_head
        norm(p + ".norm", c)
        conv(p + ".proj_in", c, inner, 1)
        for d in range(cfg.transformer_depth):
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
        conv(p + ".proj_out", inner, c, 1)

    def block(p, layers):
        for j, (kind, cin, cout) in enumerate(layers):
            q = f"{p}.{j}"
            if kind == "conv":
                conv(q, cin, cout, 3)
            elif kind == "res":
                res(q, cin, cout)
            elif kind == "st":
                st(q, cin)
            elif kind == "down":
                conv(q + ".op", cin, cout, 3)
            elif kind == "up":
                conv(q + ".conv", cin, cout, 3)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    inp, mid, out = unet_layout(cfg)
    for i, layers in enumerate(inp):
        block(f"input_blocks.{i}", layers)
    block("middle_block", mid)
    for i, layers in enumerate(out):
        block(f"output_blocks.{i}", layers)
    norm("out.0", mc)
    conv("out.2", mc, cfg.out_channels, 3)
    return s


def decoder_param_shapes(cfg: DecoderConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameters of the reference VAE Decoder (ldm/modules/diffusionmodules/model.py:470-540) + post_quant_conv
    (ldm/models/autoencoder.py:303), keys as under `first_stage_model.` in a checkpoint."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, i, o, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[-1]
    conv("post_quant_conv", cfg.embed_dim, cfg.z_channels, 1)
    conv("decoder.conv_in", cfg.z_channels, block_in, 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for i_level in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[i_level]
        for i_block in range(cfg.num_res_blocks + 1):
            res(f"decoder.up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            conv(f"decoder.up.{i_level}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, cfg.out_ch, 3)
    return s


def make_weights(shapes: Dict[str, Tuple[int, ...]], seed: int) -> SD:
    """Deterministic synthetic weights ("random-init weights of that architecture"), independent of torch's module
    initialisers: every tensor is drawn from N(0, std) with a fan-in std (biases / norm offsets small, norm gains near
    one), including the tensors the reference zero-initialises (zero_module: openaimodel.py:233,755, attention.py:1002)
    -- otherwise the UNet output is identically zero."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: SD = {}
    for key in sorted(shapes):
        shp = shapes[key]
        if key.endswith(".bias"):
            t = torch.randn(shp, generator=g) * 0.05
        elif len(shp) == 1:  # norm gain
            t = 1.0 + torch.randn(shp, generator=g) * 0.1
        else:
            fan_in = int(np.prod(shp[1:]))
            t = torch.randn(shp, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[key] = t
    return sd


def weights_checksum(sd: SD) -> float:
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


# ======================================================================================================================
# schedules (bit-exact restatements; same torch / numpy ops in the same order as the reference)
# ======================================================================================================================
def make_beta_schedule_linear(n_timestep: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012) -> np.ndarray:
    """ldm/modules/diffusionmodules/util.py:21-25 (schedule == 'linear')."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return betas.numpy()


def alphas_cumprod_from_betas(betas: np.ndarray):
    """ldm/models/diffusion/ddpm.py:134-160 register_schedule: float64 numpy cumprod, stored as fp32 tensors."""
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return f32(betas), f32(ac), f32(ac_prev)


class DiscreteSchedule:
    """k_diffusion/external.py:41-84 (DiscreteSchedule) with sigmas from DiscreteEpsDDPMDenoiser.__init__ :93."""

    def __init__(self, alphas_cumprod: Tensor):
        self.sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
        self.log_sigmas = self.sigmas.log()

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def get_sigmas(self, n: Optional[int] = None) -> Tensor:  # external.py:59-64
        if n is None:
            return torch.cat([self.sigmas.flip(0), self.sigmas.new_zeros([1])])
        t_max = len(self.sigmas) - 1
        t = torch.linspace(t_max, 0, n, device=self.sigmas.device)
        s = self.t_to_sigma(t)
        return torch.cat([s, s.new_zeros([1])])

    def sigma_to_t(self, sigma: Tensor) -> Tensor:  # external.py:66-78 (quantize=False)
        log_sigma = sigma.log()
        dists = log_sigma - self.log_sigmas[:, None]
        low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=self.log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = self.log_sigmas[low_idx], self.log_sigmas[high_idx]
        w = (low - log_sigma) / (low - high)
        w = w.clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape)

    def t_to_sigma(self, t: Tensor) -> Tensor:  # external.py:80-84
        t = t.float()
        low_idx, high_idx, w = t.floor().long(), t.ceil().long(), t.frac()
        log_sigma = (1 - w) * self.log_sigmas[low_idx] + w * self.log_sigmas[high_idx]
        return log_sigma.exp()


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> Tensor:
    """k_diffusion/sampling.py:17-23."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def get_ancestral_step(sigma_from, sigma_to, eta: float = 1.0):
    """k_diffusion/sampling.py:51-58."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def make_ddim_timesteps(num_ddim_timesteps: int, num_ddpm_timesteps: int = 1000) -> np.ndarray:
    """ldm/modules/diffusionmodules/util.py:46-60 ('uniform')."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    return ddim_timesteps + 1


def make_ddim_sampling_parameters(alphacums: Tensor, ddim_timesteps: np.ndarray, eta: float):
    """ldm/modules/diffusionmodules/util.py:63-74 (alphacums is the fp32 cpu tensor, as passed at ddim.py:63)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def timestep_embedding(timesteps: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """ldm/modules/diffusionmodules/util.py:151-171."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        embedding = torch.cat([embedding, torch.zeros_like(embedding[:, :1])], dim=-1)
    return embedding


# ======================================================================================================================
# UNet (ldm/modules/diffusionmodules/openaimodel.py, ldm/modules/attention.py)
# ======================================================================================================================
def _resblock(sd: SD, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """ResBlock._forward, openaimodel.py:259-279 (no up/down, no scale-shift norm); GroupNorm32 eps 1e-5."""
    h = F.group_norm(x, 32, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + emb_out[:, :, None, None]
    h = F.group_norm(h, 32, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


# "einsum": CrossAttentionOriginal (attention.py:611-693), the class the fp32 CPU goldens come from.  "sdpa": the same
# maths through torch's fused kernel, as the reference's GPU classes do (MemoryEfficientCrossAttention -> xformers
# attention.py:769-861, sgm CrossAttention -> F.scaled_dot_product_attention sgm/modules/attention.py:507-511); only
# bench.py's torch-on-GPU baseline switches this.
ATTENTION_IMPL = "einsum"


def _cross_attention(sd: SD, p: str, x: Tensor, context: Optional[Tensor], heads: int,
                     ipa: Optional[Tuple[float, int]] = None) -> Tensor:
    """CrossAttentionOriginal.forward, attention.py:611-693 with no LoRA / mask.  ipa = (ipa_scale, ipa_num_tokens): the
    IP-Adapter path (:623-627,660-681) -- the last ipa_num_tokens context tokens go through to_k_ipa / to_v_ipa, a second
    softmax attention with the same queries, and `out + ipa_scale * out_ipa` before to_out."""
    ctx = x if context is None else context
    ipa_ctx = None
    if ipa is not None and ipa[1] > 0:
        end = ctx.shape[1] - ipa[1]
        ctx, ipa_ctx = ctx[:, :end, :], ctx[:, end:, :]
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads
    if ATTENTION_IMPL == "sdpa" and ipa_ctx is None:
        h4 = lambda t: t.view(b, -1, heads, d).transpose(1, 2)
        out = F.scaled_dot_product_attention(h4(q), h4(k), h4(v)).transpose(1, 2).reshape(b, n, inner)
        return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    split = lambda t: t.view(b, -1, heads, d).permute(0, 2, 1, 3).reshape(b * heads, -1, d)
    q, k, v = split(q), split(k), split(v)
    merge = lambda t: t.view(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    sim = torch.einsum("b i d, b j d -> b i j", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = merge(torch.einsum("b i j, b j d -> b i d", attn, v))
    if ipa_ctx is not None:
        k2 = split(F.linear(ipa_ctx, sd[p + ".to_k_ipa.weight"]))
        v2 = split(F.linear(ipa_ctx, sd[p + ".to_v_ipa.weight"]))
        attn2 = (torch.einsum("b i d, b j d -> b i j", q, k2) * (d ** -0.5)).softmax(dim=-1)
        out = out + ipa[0] * merge(torch.einsum("b i j, b j d -> b i d", attn2, v2))
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _transformer_block(sd: SD, p: str, x: Tensor, context: Tensor, heads: int,
                       ipa: Optional[Tuple[float, int]] = None) -> Tensor:
    """BasicTransformerBlock._forward, attention.py:908-912; FeedForward/GEGLU :88-96,157-168 (exact erf GELU)."""
    ln = lambda t, nm: F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{nm}.weight"], sd[f"{p}.{nm}.bias"], 1e-5)
    x = _cross_attention(sd, p + ".attn1", ln(x, "norm1"), None, heads) + x
    x = _cross_attention(sd, p + ".attn2", ln(x, "norm2"), context, heads, ipa) + x   # only attn2 takes IPA (:895-901)
    y = F.linear(ln(x, "norm3"), sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
    y, gate = y.chunk(2, dim=-1)
    y = y * F.gelu(gate)
    y = F.linear(y, sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"])
    return y + x


def _spatial_transformer(sd: SD, p: str, x: Tensor, context: Tensor, heads: int, depth: int,
                         ipa: Optional[Tuple[float, int]] = None) -> Tensor:
    """SpatialTransformer.forward, attention.py:1031-1057; Normalize eps 1e-6 (:189); conv proj_in/out."""
    b, c, h, w = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    x = F.conv2d(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    for d in range(depth):
        x = _transformer_block(sd, f"{p}.transformer_blocks.{d}", x, context, heads, ipa)
    x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + x_in


def unet_forward(sd: SD, cfg: UNetConfig, x: Tensor, timesteps: Tensor, context: Tensor,
                 control: Optional[List[Tensor]] = None, only_mid_control: bool = False,
                 ipa: Optional[Tuple[float, int]] = None) -> Tensor:
    """UNetModel.forward, openaimodel.py:780-816 (fp32, y=None).  ipa = (ipa_scale, ipa_num_tokens) as given to the
    constructor (:479-480), see _cross_attention.  With `control`: ControlledUnetModel.forward,
    cldm/cldm.py:44-70 -- the residuals are popped from the END of the list (a copy here): one after the middle block
    (:59-60), one per output block added to the skip before the concat (:62-66) unless only_mid_control."""
    inp, mid, out = unet_layout(cfg)
    t_emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    def run(prefix: str, layers, h: Tensor) -> Tensor:
        for j, (kind, cin, cout) in enumerate(layers):
            q = f"{prefix}.{j}"
            if kind == "conv":
                h = F.conv2d(h, sd[q + ".weight"], sd[q + ".bias"], padding=1)
            elif kind == "res":
                h = _resblock(sd, q, h, emb)
            elif kind == "st":
                h = _spatial_transformer(sd, q, h, context, cfg.num_heads, cfg.transformer_depth, ipa)
            elif kind == "down":  # Downsample.forward openaimodel.py:162
                h = F.conv2d(h, sd[q + ".op.weight"], sd[q + ".op.bias"], stride=2, padding=1)
            elif kind == "up":  # Upsample.forward openaimodel.py:113-123
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[q + ".conv.weight"], sd[q + ".conv.bias"], padding=1)
        return h

    hs: List[Tensor] = []
    h = x.float()
    for i, layers in enumerate(inp):
        h = run(f"input_blocks.{i}", layers, h)
        hs.append(h)
    h = run("middle_block", mid, h)
    control = list(control) if control is not None else None
    if control is not None:
        h = h + control.pop()
    for i, layers in enumerate(out):
        if only_mid_control or control is None:
            h = torch.cat([h, hs.pop()], dim=1)
        else:
            h = torch.cat([h, hs.pop() + control.pop()], dim=1)
        h = run(f"output_blocks.{i}", layers, h)
    h = F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5)
    return F.conv2d(F.silu(h), sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ======================================================================================================================
# VAE decoder (ldm/modules/diffusionmodules/model.py, ldm/models/autoencoder.py)
# ======================================================================================================================
def _vae_resnet(sd: SD, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward, model.py:128-148 (temb None); Normalize eps 1e-6 (:45), swish (:40-42)."""
    h = F.group_norm(x, 32, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6)
    h = F.conv2d(h * torch.sigmoid(h), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.group_norm(h, 32, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6)
    h = F.conv2d(h * torch.sigmoid(h), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def _vae_attn(sd: SD, p: str, x: Tensor) -> Tensor:
    """AttnBlock.forward, model.py:185-209."""
    h_ = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    q = F.conv2d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, h, w = q.shape
    if ATTENTION_IMPL == "sdpa":   # sgm AttnBlock.attention, sgm/modules/diffusionmodules/model.py:180-195
        t4 = lambda t: t.reshape(b, 1, c, h * w).transpose(2, 3)
        h_ = F.scaled_dot_product_attention(t4(q), t4(k), t4(v)).transpose(2, 3).reshape(b, c, h, w)
        return x + F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + h_


def decoder_forward(sd: SD, cfg: DecoderConfig, z: Tensor) -> Tensor:
    """Decoder.forward, model.py:542-575 (attn_resolutions = [], give_pre_end / tanh_out False)."""
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _vae_resnet(sd, "decoder.mid.block_1", h)
    h = _vae_attn(sd, "decoder.mid.attn_1", h)
    h = _vae_resnet(sd, "decoder.mid.block_2", h)
    for i_level in reversed(range(len(cfg.ch_mult))):
        for i_block in range(cfg.num_res_blocks + 1):
            h = _vae_resnet(sd, f"decoder.up.{i_level}.block.{i_block}", h)
        if i_level != 0:  # Upsample.forward model.py:60-64
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up.{i_level}.upsample.conv.weight"], sd[f"decoder.up.{i_level}.upsample.conv.bias"], padding=1)
    h = F.group_norm(h, 32, sd["decoder.norm_out.weight"], sd["decoder.norm_out.bias"], 1e-6)
    h = h * torch.sigmoid(h)
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def vae_decode(sd: SD, cfg: DecoderConfig, z: Tensor) -> Tensor:
    """AutoencoderKL.decode, autoencoder.py:333-338: post_quant_conv then Decoder."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(sd, cfg, z)


def decode_first_stage(sd: SD, cfg: DecoderConfig, z: Tensor, scale_factor: float = 0.18215) -> Tensor:
    """LatentDiffusion.decode_first_stage live branch, ddpm.py:794-798."""
    return vae_decode(sd, cfg, 1.0 / scale_factor * z)


def images_to_uint8(x: Tensor) -> Tensor:
    """sd/image_generator.py:1017-1018,1151-1152: clamp((x+1)/2,0,1); 255*x -> uint8 HWC (truncating astype)."""
    x = torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0)
    x = (255.0 * x.permute(0, 2, 3, 1)).cpu().numpy().astype(np.uint8)
    return torch.from_numpy(x)


# ======================================================================================================================
# denoiser wrappers + samplers
# ======================================================================================================================
class OracleDenoiser:
    """CompVisDenoiser (k_diffusion/external.py:87-147) around LDMWrapperForKDiffusion's CFG
    (ldm/models/diffusion/ldm_wrapper_for_k_diffusion.py:48-101): returns the guided *denoised* prediction."""

    def __init__(self, eps_fn: Callable[[Tensor, Tensor, Tensor], Tensor], alphas_cumprod: Tensor, cond: Tensor,
                 uncond: Tensor, cfg_scale: float):
        self.eps_fn = eps_fn
        self.schedule = DiscreteSchedule(alphas_cumprod)
        self.cond, self.uncond, self.cfg_scale = cond, uncond, cfg_scale

    def _compvis(self, x: Tensor, sigma: Tensor, ctx: Tensor) -> Tensor:  # external.py:111-114
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + 1.0 ** 2) ** 0.5
        ap = lambda v: v[(...,) + (None,) * (x.ndim - v.ndim)]
        eps = self.eps_fn(x * ap(c_in), self.schedule.sigma_to_t(sigma), ctx)
        return x + eps * ap(c_out)

    def __call__(self, x: Tensor, sigma: Tensor) -> Tensor:
        x_in = torch.cat([x] * 2)
        t_in = torch.cat([sigma] * 2)
        c_in = torch.cat([self.uncond, self.cond])
        d_uncond, d_cond = self._compvis(x_in, t_in, c_in).chunk(2)
        return d_uncond + self.cfg_scale * (d_cond - d_uncond)


def sample_euler_ancestral(model, x: Tensor, sigmas: Tensor, noise: Optional[Sequence[Tensor]] = None, eta: float = 1.0,
                           s_noise: float = 1.0, trace: Optional[List[Tensor]] = None) -> Tensor:
    """k_diffusion/sampling.py:147-163; `noise[i]` is the injected noise_sampler output of step i."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        d = (x - denoised) / sigmas[i]
        dt = sigma_down - sigmas[i]
        x = x + d * dt
        if sigmas[i + 1] > 0:
            x = x + noise[i] * s_noise * sigma_up
        if trace is not None:
            trace.append(x.clone())
    return x


def sample_dpmpp_2m(model, x: Tensor, sigmas: Tensor, trace: Optional[List[Tensor]] = None) -> Tensor:
    """k_diffusion/sampling.py:593-615."""
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    old_denoised = None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        if old_denoised is None or sigmas[i + 1] == 0:
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            denoised_d = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old_denoised
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised_d
        old_denoised = denoised
        if trace is not None:
            trace.append(x.clone())
    return x


def ddim_sample(eps_fn: Callable[[Tensor, Tensor, Tensor], Tensor], alphas_cumprod: Tensor, x_T: Tensor, cond: Tensor,
                uncond: Tensor, cfg_scale: float, S: int, eta: float = 0.0,
                trace: Optional[List[Tensor]] = None, mask: Optional[Tensor] = None, x0: Optional[Tensor] = None,
                mask_noise: Optional[Sequence[Tensor]] = None) -> Tensor:
    """DDIMSampler.sample -> ddim_sampling -> p_sample_ddim (ldm/models/diffusion/ddim.py:78-190,530-612), eta = 0
    path with classifier-free guidance; integer timesteps (torch.long)."""
    ddim_timesteps = make_ddim_timesteps(S, alphas_cumprod.shape[0])
    sig, alphas, alphas_prev = make_ddim_sampling_parameters(alphas_cumprod.cpu(), ddim_timesteps, eta)
    sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)  # ddim.py:71 (a tensor op on the fp32 tensor)
    img = x_T
    b = x_T.shape[0]
    time_range = np.flip(ddim_timesteps)
    total = ddim_timesteps.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), device=x_T.device, dtype=torch.long)
        if mask is not None:  # inpainting, ddim.py:171-174 + LatentDiffusion.q_sample ddpm.py:296-299
            ac_t = alphas_cumprod[int(step)]
            img_orig = torch.sqrt(ac_t) * x0 + torch.sqrt(1.0 - ac_t) * mask_noise[i]
            img = img_orig * mask + (1.0 - mask) * img
        x_in = torch.cat([img] * 2)
        t_in = torch.cat([ts] * 2)
        c_in = torch.cat([uncond, cond])
        e_t_uncond, e_t = eps_fn(x_in, t_in, c_in).chunk(2)
        e_t = e_t_uncond + cfg_scale * (e_t - e_t_uncond)
        dev = x_T.device
        a_t = torch.full((b, 1, 1, 1), alphas[index], device=dev)
        a_prev = torch.full((b, 1, 1, 1), alphas_prev[index], device=dev)
        sigma_t = torch.full((b, 1, 1, 1), sig[index], device=dev)
        sqrt_one_minus_at = torch.full((b, 1, 1, 1), sqrt_one_minus_alphas[index], device=dev)
        pred_x0 = (img - sqrt_one_minus_at * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        img = a_prev.sqrt() * pred_x0 + dir_xt  # + sigma_t * noise, identically zero for eta = 0
        if trace is not None:
            trace.append(img.clone())
    return img


def ddim_stochastic_encode(alphas_cumprod: Tensor, x0: Tensor, t_index: int, S: int, noise: Tensor) -> Tensor:
    """DDIMSampler.stochastic_encode, ddim.py:615-655 (use_original_steps=False): index into the S-step DDIM tables."""
    ts = make_ddim_timesteps(S, alphas_cumprod.shape[0])
    _, alphas, _ = make_ddim_sampling_parameters(alphas_cumprod.cpu(), ts, 0.0)
    sqrt_a = torch.sqrt(alphas)
    sqrt_1ma = np.sqrt(1.0 - alphas)
    return float(sqrt_a[t_index]) * x0 + float(sqrt_1ma[t_index]) * noise


def ddim_decode(eps_fn, alphas_cumprod: Tensor, x_latent: Tensor, cond: Tensor, uncond: Tensor, cfg_scale: float,
                S: int, t_start: int) -> Tensor:
    """DDIMSampler.decode, ddim.py:657-676: the last `t_start` DDIM steps (eta = 0) with classifier-free guidance."""
    ts_all = make_ddim_timesteps(S, alphas_cumprod.shape[0])
    sig, alphas, alphas_prev = make_ddim_sampling_parameters(alphas_cumprod.cpu(), ts_all, 0.0)
    sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)
    timesteps = ts_all[:t_start]
    x = x_latent
    b = x.shape[0]
    total = timesteps.shape[0]
    for i, step in enumerate(np.flip(timesteps)):
        index = total - i - 1
        ts = torch.full((b,), int(step), device=x.device, dtype=torch.long)
        e_u, e_c = eps_fn(torch.cat([x] * 2), torch.cat([ts] * 2), torch.cat([uncond, cond])).chunk(2)
        e_t = e_u + cfg_scale * (e_c - e_u)
        a_t = torch.full((b, 1, 1, 1), alphas[index])
        a_prev = torch.full((b, 1, 1, 1), alphas_prev[index])
        s1 = torch.full((b, 1, 1, 1), sqrt_one_minus_alphas[index])
        pred_x0 = (x - s1 * e_t) / a_t.sqrt()
        x = a_prev.sqrt() * pred_x0 + (1.0 - a_prev).sqrt() * e_t
    return x


def hires_fix_latent_ddim(eps_fn, alphas_cumprod: Tensor, samples: Tensor, cond: Tensor, uncond: Tensor, cfg_scale: float,
                          S: int, strength: float, noise: Tensor, factor: int = 2) -> Tensor:
    """sd/image_generator.py:969-999 (latent upscaler) + img2img_sampling :147-248 (DDIM branch)."""
    up = F.interpolate(samples, scale_factor=factor, mode="bilinear", align_corners=False)
    t_enc = int(strength * S)
    z_enc = ddim_stochastic_encode(alphas_cumprod, up, t_enc, S, noise)
    return ddim_decode(eps_fn, alphas_cumprod, z_enc, cond, uncond, cfg_scale, S, t_enc)


# ======================================================================================================================
# remaining k-diffusion samplers of the reference's front ends, s_churn = 0
# ======================================================================================================================
def sample_heun(model, x: Tensor, sigmas: Tensor) -> Tensor:
    """k_diffusion/sampling.py:167-193."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        d = (x - denoised) / sigmas[i]
        dt = sigmas[i + 1] - sigmas[i]
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            denoised_2 = model(x_2, sigmas[i + 1] * s_in)
            d_2 = (x_2 - denoised_2) / sigmas[i + 1]
            x = x + ((d + d_2) / 2) * dt
    return x


def sample_dpm_2(model, x: Tensor, sigmas: Tensor) -> Tensor:
    """k_diffusion/sampling.py:196-224."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        d = (x - denoised) / sigmas[i]
        if sigmas[i + 1] == 0:
            x = x + d * (sigmas[i + 1] - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigmas[i + 1].log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigmas[i])
            denoised_2 = model(x_2, sigma_mid * s_in)
            d_2 = (x_2 - denoised_2) / sigma_mid
            x = x + d_2 * (sigmas[i + 1] - sigmas[i])
    return x


def sample_dpm_2_ancestral(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], eta: float = 1.0,
                           s_noise: float = 1.0) -> Tensor:
    """k_diffusion/sampling.py:227-252; `noise[i]` is the injected noise_sampler output of step i."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        d = (x - denoised) / sigmas[i]
        if sigma_down == 0:
            x = x + d * (sigma_down - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigma_down.log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigmas[i])
            denoised_2 = model(x_2, sigma_mid * s_in)
            d_2 = (x_2 - denoised_2) / sigma_mid
            x = x + d_2 * (sigma_down - sigmas[i])
            x = x + noise[i] * s_noise * sigma_up
    return x


def linear_multistep_coeff(order: int, t, i: int, j: int) -> float:
    """k_diffusion/sampling.py:255-266 (scipy.integrate.quad, epsrel 1e-4)."""
    from scipy import integrate

    def fn(tau):
        prod = 1.0
        for k in range(order):
            if j == k:
                continue
            prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod
    return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]


def sample_lms(model, x: Tensor, sigmas: Tensor, order: int = 4) -> Tensor:
    """k_diffusion/sampling.py:269-286."""
    s_in = x.new_ones([x.shape[0]])
    sigmas_cpu = sigmas.detach().cpu().numpy()
    ds: List[Tensor] = []
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        ds.append((x - denoised) / sigmas[i])
        if len(ds) > order:
            ds.pop(0)
        cur_order = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur_order, sigmas_cpu, i, j) for j in range(cur_order)]
        x = x + sum(coeff * d for coeff, d in zip(coeffs, reversed(ds)))
    return x


def sample_dpmpp_2s_ancestral(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], eta: float = 1.0,
                              s_noise: float = 1.0) -> Tensor:
    """k_diffusion/sampling.py:517-548."""
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        if sigma_down == 0:
            x = x + ((x - denoised) / sigmas[i]) * (sigma_down - sigmas[i])
        else:
            t, t_next = t_fn(sigmas[i]), t_fn(sigma_down)
            r = 1 / 2
            h = t_next - t
            s = t + r * h
            x_2 = (sigma_fn(s) / sigma_fn(t)) * x - (-h * r).expm1() * denoised
            denoised_2 = model(x_2, sigma_fn(s) * s_in)
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised_2
        if sigmas[i + 1] > 0:
            x = x + noise[i] * s_noise * sigma_up
    return x


# ======================================================================================================================
# SDE family and stochastic churn.  `noise` is the sequence the injected noise_sampler / randn_like returns.
# ======================================================================================================================
def _churned(x: Tensor, sigmas: Tensor, i: int, noise_it, s_churn: float, s_tmin: float, s_tmax: float, s_noise: float):
    """k_diffusion/sampling.py:124-133: the reference draws on every step, the draw only enters when gamma > 0."""
    gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
    eps = next(noise_it) * s_noise
    sigma_hat = sigmas[i] * (gamma + 1)
    if gamma > 0:
        x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
    return x, sigma_hat


def sample_euler_churn(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], s_churn: float, s_tmin: float = 0.0,
                       s_tmax: float = float("inf"), s_noise: float = 1.0) -> Tensor:
    """k_diffusion/sampling.py:118-144."""
    s_in, it = x.new_ones([x.shape[0]]), iter(noise)
    for i in range(len(sigmas) - 1):
        x, sigma_hat = _churned(x, sigmas, i, it, s_churn, s_tmin, s_tmax, s_noise)
        denoised = model(x, sigma_hat * s_in)
        x = x + ((x - denoised) / sigma_hat) * (sigmas[i + 1] - sigma_hat)
    return x


def sample_heun_churn(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], s_churn: float, s_tmin: float = 0.0,
                      s_tmax: float = float("inf"), s_noise: float = 1.0) -> Tensor:
    """k_diffusion/sampling.py:167-193."""
    s_in, it = x.new_ones([x.shape[0]]), iter(noise)
    for i in range(len(sigmas) - 1):
        x, sigma_hat = _churned(x, sigmas, i, it, s_churn, s_tmin, s_tmax, s_noise)
        denoised = model(x, sigma_hat * s_in)
        d = (x - denoised) / sigma_hat
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            d_2 = (x_2 - model(x_2, sigmas[i + 1] * s_in)) / sigmas[i + 1]
            x = x + ((d + d_2) / 2) * dt
    return x


def sample_dpm_2_churn(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], s_churn: float, s_tmin: float = 0.0,
                       s_tmax: float = float("inf"), s_noise: float = 1.0) -> Tensor:
    """k_diffusion/sampling.py:196-224."""
    s_in, it = x.new_ones([x.shape[0]]), iter(noise)
    for i in range(len(sigmas) - 1):
        x, sigma_hat = _churned(x, sigmas, i, it, s_churn, s_tmin, s_tmax, s_noise)
        denoised = model(x, sigma_hat * s_in)
        d = (x - denoised) / sigma_hat
        if sigmas[i + 1] == 0:
            x = x + d * (sigmas[i + 1] - sigma_hat)
        else:
            sigma_mid = sigma_hat.log().lerp(sigmas[i + 1].log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigma_hat)
            d_2 = (x_2 - model(x_2, sigma_mid * s_in)) / sigma_mid
            x = x + d_2 * (sigmas[i + 1] - sigma_hat)
    return x


def sample_dpmpp_sde(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], eta: float = 1.0, s_noise: float = 1.0,
                     r: float = 0.5) -> Tensor:
    """k_diffusion/sampling.py:551-590: two noise draws per step (midpoint, end)."""
    s_in, it = x.new_ones([x.shape[0]]), iter(noise)
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        if sigmas[i + 1] == 0:
            x = x + ((x - denoised) / sigmas[i]) * (sigmas[i + 1] - sigmas[i])
            continue
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        s = t + h * r
        fac = 1 / (2 * r)
        sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(s), eta)
        s_ = t_fn(sd)
        x_2 = (sigma_fn(s_) / sigma_fn(t)) * x - (t - s_).expm1() * denoised + next(it) * s_noise * su
        denoised_2 = model(x_2, sigma_fn(s) * s_in)
        sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(t_next), eta)
        t_next_ = t_fn(sd)
        denoised_d = (1 - fac) * denoised + fac * denoised_2
        x = (sigma_fn(t_next_) / sigma_fn(t)) * x - (t - t_next_).expm1() * denoised_d + next(it) * s_noise * su
    return x


def sample_dpmpp_2m_sde(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], eta: float = 1.0, s_noise: float = 1.0,
                        solver_type: str = "midpoint") -> Tensor:
    """k_diffusion/sampling.py:619-662."""
    s_in, it = x.new_ones([x.shape[0]]), iter(noise)
    old_denoised, h_last, h = None, None, None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        if sigmas[i + 1] == 0:
            x = denoised
        else:
            t, s = -sigmas[i].log(), -sigmas[i + 1].log()
            h = s - t
            eta_h = eta * h
            x = sigmas[i + 1] / sigmas[i] * (-eta_h).exp() * x + (-h - eta_h).expm1().neg() * denoised
            if old_denoised is not None:
                r = h_last / h
                if solver_type == "heun":
                    x = x + ((-h - eta_h).expm1().neg() / (-h - eta_h) + 1) * (1 / r) * (denoised - old_denoised)
                else:
                    x = x + 0.5 * (-h - eta_h).expm1().neg() * (1 / r) * (denoised - old_denoised)
            if eta:
                x = x + next(it) * sigmas[i + 1] * (-2 * eta_h).expm1().neg().sqrt() * s_noise
        old_denoised, h_last = denoised, h
    return x


def sample_dpmpp_3m_sde(model, x: Tensor, sigmas: Tensor, noise: Sequence[Tensor], eta: float = 1.0,
                        s_noise: float = 1.0) -> Tensor:
    """k_diffusion/sampling.py:664-717."""
    s_in, it = x.new_ones([x.shape[0]]), iter(noise)
    den_1, den_2, h, h_1, h_2 = None, None, None, None, None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        if sigmas[i + 1] == 0:
            x = denoised
        else:
            t, s = -sigmas[i].log(), -sigmas[i + 1].log()
            h = s - t
            h_eta = h * (eta + 1)
            x = torch.exp(-h_eta) * x + (-h_eta).expm1().neg() * denoised
            if h_2 is not None:
                r0, r1 = h_1 / h, h_2 / h
                d1_0 = (denoised - den_1) / r0
                d1_1 = (den_1 - den_2) / r1
                d1 = d1_0 + (d1_0 - d1_1) * r0 / (r0 + r1)
                d2 = (d1_0 - d1_1) / (r0 + r1)
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                phi_3 = phi_2 / h_eta - 0.5
                x = x + phi_2 * d1 - phi_3 * d2
            elif h_1 is not None:
                r = h_1 / h
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                x = x + phi_2 * ((denoised - den_1) / r)
            if eta:
                x = x + next(it) * sigmas[i + 1] * (-2 * h * eta).expm1().neg().sqrt() * s_noise
        den_1, den_2 = denoised, den_1
        h_1, h_2 = h, h_1
    return x


def kdiff_stochastic_encode(alphas_cumprod: Tensor, x0: Tensor, t: Tensor, sampling_steps: int, noise: Tensor) -> Tensor:
    """KDiffusionSamplerBase.stochastic_encode, ldm/models/diffusion/k_diffusion_samplers.py:260-297: per-sample index
    `t * 1000 / steps` (truncated) into the DDPM sqrt(alpha-bar) tables (np.sqrt of the fp32 table, :112-113)."""
    sqrt_ac = torch.from_numpy(np.sqrt(alphas_cumprod.cpu().numpy())).float()
    sqrt_1m = torch.from_numpy(np.sqrt(1.0 - alphas_cumprod.cpu().numpy())).float()
    idx = (t * 1000.0 / sampling_steps).long()
    shape = (-1,) + (1,) * (x0.ndim - 1)
    return sqrt_ac[idx].reshape(shape) * x0 + sqrt_1m[idx].reshape(shape) * noise


# ======================================================================================================================
# VAE encoder: the step before the path for img2img
# ======================================================================================================================
def encoder_param_shapes(cfg: DecoderConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameters of the reference VAE Encoder (ldm/modules/diffusionmodules/model.py:375-448) + quant_conv
    (ldm/models/autoencoder.py:302), keys as under `first_stage_model.` in a checkpoint (attn_resolutions = [])."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, i, o, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    nres = len(cfg.ch_mult)
    in_ch_mult = (1,) + tuple(cfg.ch_mult)
    conv("encoder.conv_in", 3, cfg.ch, 3)
    block_in = cfg.ch
    for i_level in range(nres):
        block_in = cfg.ch * in_ch_mult[i_level]
        block_out = cfg.ch * cfg.ch_mult[i_level]
        for i_block in range(cfg.num_res_blocks):
            res(f"encoder.down.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != nres - 1:
            conv(f"encoder.down.{i_level}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    norm("encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", block_in, 2 * cfg.z_channels, 3)
    conv("quant_conv", 2 * cfg.z_channels, 2 * cfg.embed_dim, 1)
    return s


def encoder_forward(sd: SD, cfg: DecoderConfig, x: Tensor) -> Tensor:
    """Encoder.forward, model.py:440-466; Downsample = pad (0,1,0,1) then conv3x3 stride 2 padding 0 (:79-84)."""
    nres = len(cfg.ch_mult)
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i_level in range(nres):
        for i_block in range(cfg.num_res_blocks):
            h = _vae_resnet(sd, f"encoder.down.{i_level}.block.{i_block}", h)
        if i_level != nres - 1:
            p = f"encoder.down.{i_level}.downsample.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = _vae_resnet(sd, "encoder.mid.block_1", h)
    h = _vae_attn(sd, "encoder.mid.attn_1", h)
    h = _vae_resnet(sd, "encoder.mid.block_2", h)
    h = F.group_norm(h, 32, sd["encoder.norm_out.weight"], sd["encoder.norm_out.bias"], 1e-6)
    return F.conv2d(h * torch.sigmoid(h), sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)


def vae_encode_moments(sd: SD, cfg: DecoderConfig, x: Tensor) -> Tensor:
    """AutoencoderKL.encode up to the posterior's parameters, autoencoder.py:324-331: quant_conv(encoder(x))."""
    return F.conv2d(encoder_forward(sd, cfg, x), sd["quant_conv.weight"], sd["quant_conv.bias"])


def gaussian_sample(moments: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution, ldm/modules/distributions/distributions.py:24-37: mean + exp(0.5 * clamp(logvar,
    -30, 20)) * noise (`noise` = the injected torch.randn draw)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * noise


# ======================================================================================================================
# LoRA side branches of the UNet: ldm/modules/attention.py:79-96,148-168,306-376,966-1055
# ======================================================================================================================
LORA_BASE = {"q": "to_q", "k": "to_k", "v": "to_v", "out": "to_out.0", "proj": "proj", "net_2": "net.2",
             "proj_in": "proj_in", "proj_out": "proj_out"}


def make_lora_weights(shapes: Dict[str, Tuple[int, ...]], seed: int) -> SD:
    """Deterministic non-trivial LoRA tensors for the keys `*_lora_downs.i.weight`, `*_lora_ups.i.weight`,
    `*_lora_alphas.i` (the reference zero-initialises downs / ups, which would make the branches no-ops)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: SD = {}
    for key in sorted(shapes):
        shp = shapes[key]
        if "_lora_alphas." in key:
            sd[key] = torch.tensor(3.0) + torch.randn((), generator=g)
        else:
            fan_in = int(np.prod(shp[1:]))
            sd[key] = torch.randn(shp, generator=g) * (0.7 / math.sqrt(fan_in))
    return sd


def lora_merge(sd: SD, lora_sd: SD, lora_ranks: Sequence[int], lora_weights: Sequence[float]) -> SD:
    """Every branch computes  base(x) + sum_i up_i(down_i(x)) * lora_weights[i] * (alpha_i / rank_i)  on the SAME input as
    the base projection (attention.py:344-348,1036-1041 ...), i.e. the linear map  W + sum_i s_i up_i @ down_i."""
    out = dict(sd)
    for key in lora_sd:
        if "_lora_downs." not in key:
            continue
        head, rest = key.rsplit("_lora_downs.", 1)
        i = int(rest.split(".")[0])
        prefix, _, branch = head.rpartition(".")
        # net_2 / proj_in / proj_out contain an underscore themselves: the branch is everything after the last dot
        base_key = f"{prefix}.{LORA_BASE[branch]}.weight" if prefix else f"{LORA_BASE[branch]}.weight"
        down = lora_sd[key].reshape(lora_sd[key].shape[0], -1)
        up = lora_sd[f"{head}_lora_ups.{i}.weight"]
        up = up.reshape(up.shape[0], -1)
        scale = lora_weights[i] * (lora_sd[f"{head}_lora_alphas.{i}"] / lora_ranks[i])
        w = out[base_key]
        out[base_key] = w + (scale * (up @ down)).reshape(w.shape)
    return out
