"""Test helpers: build the cremage_b200 drop-in modules for an oracle config and load oracle-named weights."""
import os

import numpy as np
import torch

from oracle import sd_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator(device="cpu").manual_seed(seed))


def unet_kwargs(cfg: O.UNetConfig):
    return dict(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                model_channels=cfg.model_channels, attention_resolutions=list(cfg.attention_resolutions),
                num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult), num_heads=cfg.num_heads,
                use_spatial_transformer=True, transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                use_checkpoint=True, legacy=False)


def vae_kwargs(cfg: O.DecoderConfig):
    return dict(ddconfig=dict(double_z=True, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=3,
                              out_ch=cfg.out_ch, ch=cfg.ch, ch_mult=list(cfg.ch_mult),
                              num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0),
                lossconfig=None, embed_dim=cfg.embed_dim)


def build_unet(cfg, sd, device="cuda"):
    from cremage_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device("meta"):
        m = UNetModel(**unet_kwargs(cfg))
    m = m.to_empty(device="cpu")
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


ENCODER_SEED = 400


def full_vae_weights(cfg, sd):
    """Decoder-side weights `sd` completed with the (seeded) encoder-side tensors: the checkpoint layout of
    `first_stage_model.*` (encoder.*, quant_conv, post_quant_conv, decoder.*)."""
    out = dict(O.make_weights(O.encoder_param_shapes(cfg), seed=ENCODER_SEED))
    out.update(sd)
    return out


def build_vae(cfg, sd, device="cuda"):
    from cremage_b200.ldm.models.autoencoder import AutoencoderKL
    with torch.device("meta"):
        m = AutoencoderKL(**vae_kwargs(cfg))
    m = m.to_empty(device="cpu")
    m.load_state_dict(full_vae_weights(cfg, sd), strict=True)
    return m.to(device).eval()


def build_ldm(ucfg, usd, vcfg=None, vsd=None, device="cuda"):
    from cremage_b200.ldm.models.diffusion.ddpm import LatentDiffusion
    unet = build_unet(ucfg, usd, device)
    vae = build_vae(vcfg, vsd, device) if vcfg is not None else None
    return LatentDiffusion(unet, vae).to(device).eval()


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


BF16_FACTOR = 4.0


def build_dtype() -> str:
    """'fp16' (default build, the reference's own GPU precision) or 'bf16' (CREMAGE_B200_DTYPE=bf16)."""
    from cremage_b200 import _lib
    return _lib.DEFAULT_DTYPE


def tol(fp16_value: float) -> float:
    """Stated tolerance for the running build: bf16 keeps 8 mantissa bits instead of fp16's 11 (unit round-off 8x
    larger); its bounds are BF16_FACTOR x the fp16 ones."""
    return fp16_value * (BF16_FACTOR if build_dtype() == "bf16" else 1.0)
