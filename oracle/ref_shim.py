"""ORACLE support -- test infrastructure. Imports the UNMODIFIED reference modules from /root/reference.

Only usable where /root/reference exists (the authoring container): used by oracle/make_golden.py to generate the
committed golden vectors and by tests/test_oracle_pin.py to pin oracle/sd_oracle.py against the reference itself.
Nothing that runs on the GPU box may import this (the reference tree does not travel).

Recipe: stub the un-vendored imports the path never executes (omegaconf, torchdiffeq, torchsde),
pretend CUDA is available *before* importing so the reference's "no CUDA => Apple MPS => cast to fp16" branches stay
off (openaimodel.py:85-90,794-795), and select CrossAttentionOriginal with GPU_DEVICE=cpu (attention.py:877-881).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modules", "ldm"))


_DONE = False


def install():
    global _DONE
    if _DONE:
        return
    if not available():
        raise RuntimeError("reference tree not present at /root/reference")
    import torch

    os.environ["GPU_DEVICE"] = "cpu"
    torch.cuda.is_available = lambda: True  # keeps the reference in fp32 on CPU (see module docstring)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class ListConfig(list):
        pass

    if "omegaconf" not in sys.modules:
        oc = stub("omegaconf", ListConfig=ListConfig)
        oc.listconfig = stub("omegaconf.listconfig", ListConfig=ListConfig)
    if "torchdiffeq" not in sys.modules:
        stub("torchdiffeq", odeint=None)
    if "torchsde" not in sys.modules:
        stub("torchsde")
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "modules"))
    _DONE = True


def install_lightning_stub():
    """cldm/cldm.py imports ldm.models.diffusion.ddpm (LatentDiffusion, a pytorch_lightning.LightningModule) at module
    level; pytorch_lightning is not installed.  Nothing of it runs on the UNet path: a LightningModule = nn.Module stub
    is enough to import the reference's ControlledUnetModel unmodified."""
    install()
    if "pytorch_lightning" in sys.modules:
        return
    import torch.nn as nn

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class LightningModule(nn.Module):
        pass

    stub("pytorch_lightning", LightningModule=LightningModule)
    for sub in ("utilities", "utilities.distributed", "utilities.rank_zero"):
        stub("pytorch_lightning." + sub, rank_zero_only=lambda f: f)


def reference_unet(cfg, sd):
    """Reference UNetModel (ldm/modules/diffusionmodules/openaimodel.py:417) for an oracle UNetConfig, loaded with `sd`
    (strict: proves the oracle's key naming equals the reference's)."""
    install()
    from ldm.modules.diffusionmodules.openaimodel import UNetModel

    m = UNetModel(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                  model_channels=cfg.model_channels, attention_resolutions=list(cfg.attention_resolutions),
                  num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult), num_heads=cfg.num_heads,
                  use_spatial_transformer=True, transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                  use_checkpoint=False, legacy=False)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def reference_decoder(cfg, sd):
    """Reference VAE Decoder (ldm/modules/diffusionmodules/model.py:469) + post_quant_conv (autoencoder.py:303)."""
    install()
    import torch
    from ldm.modules.diffusionmodules.model import Decoder

    dec = Decoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=tuple(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                  attn_resolutions=[], dropout=0.0, in_channels=3, resolution=cfg.resolution,
                  z_channels=cfg.z_channels, double_z=True)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    pq = torch.nn.Conv2d(cfg.embed_dim, cfg.z_channels, 1)
    pq.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
    return dec.eval(), pq.eval()


class DuckLDM:
    """The attributes of LatentDiffusion that CompVisDenoiser (k_diffusion/external.py:139-147),
    LDMWrapperForKDiffusion (ldm_wrapper_for_k_diffusion.py:42-43) and DDIMSampler (ddim.py:25,41-49,536) touch."""

    def __init__(self, unet, betas, alphas_cumprod, alphas_cumprod_prev):
        import torch

        self.unet = unet
        self.betas = betas
        self.alphas_cumprod = alphas_cumprod
        self.alphas_cumprod_prev = alphas_cumprod_prev
        self.num_timesteps = int(alphas_cumprod.shape[0])
        self.device = torch.device("cpu")
        self.parameterization = "eps"

    def q_sample(self, x_start, t, noise=None):  # ddpm.py:296-299 (LatentDiffusion itself needs Lightning to import)
        import torch

        noise = torch.randn_like(x_start) if noise is None else noise
        sa = torch.sqrt(self.alphas_cumprod)[t].reshape(-1, 1, 1, 1)
        so = torch.sqrt(1.0 - self.alphas_cumprod)[t].reshape(-1, 1, 1, 1)
        return sa * x_start + so * noise

    def apply_model(self, x, t, cond):  # ddpm.py:926-1039 live branch + DiffusionWrapper 'crossattn' :1517-1519
        import torch

        if isinstance(cond, dict):
            cc = torch.cat(cond["c_crossattn"], 1)
        else:
            cc = cond
        return self.unet(x, t, context=cc)
