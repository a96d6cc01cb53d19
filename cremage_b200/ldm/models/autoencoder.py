"""Drop-in mirror of the reference's `ldm.models.autoencoder.AutoencoderKL` (HowToSD/cremage
modules/ldm/models/autoencoder.py:285-338) for the decode side of the path.

`decode(z)` = post_quant_conv (1x1, embed_dim -> z_channels) then Decoder.  The 1x1 conv is fused into the
NCHW->NHWC conversion kernel.  `decode_first_stage(z)` additionally applies the 1/scale_factor of
LatentDiffusion.decode_first_stage (ldm/models/diffusion/ddpm.py:794-798) inside the same kernel and can return the
uint8 HWC image of the reference's post-processing (sd/image_generator.py:1017-1018,1151-1152).
The encoder (img2img) is not part of this round.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...engine import PackedModule, f32, require_cuda
from ..modules.diffusionmodules.model import Decoder


class AutoencoderKL(PackedModule):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=[], image_key="image",
                 colorize_nlabels=None, monitor=None):
        super().__init__()
        if ckpt_path is not None:
            raise NotImplementedError("cremage_b200: load weights with load_state_dict (ckpt_path is not supported)")
        self.image_key = image_key
        ddconfig = dict(ddconfig)
        assert ddconfig.get("double_z", True)
        ddconfig.pop("double_z", None)
        self.decoder = Decoder(**ddconfig)
        self.embed_dim = embed_dim
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        if embed_dim > 8 or ddconfig["z_channels"] > 8:
            raise NotImplementedError("cremage_b200: latent channel counts above 8 are not implemented")

    def _own_params(self):
        return list(self.post_quant_conv.parameters())

    def _pack(self, device):
        zc = self.post_quant_conv.out_channels
        return {"w": f32(self.post_quant_conv.weight, device).reshape(zc, self.embed_dim).contiguous(),
                "b": f32(self.post_quant_conv.bias, device)}

    def _decode_nhwc(self, z: torch.Tensor, scale: float) -> torch.Tensor:
        require_cuda(z, "AutoencoderKL.decode")
        p = self.packed(z.device)
        zq = ops.pointwise_nchw_to_nhwc(z, p["w"], p["b"], c_pad=8, scale=scale)
        return self.decoder._run(zq)  # fp32 NHWC [n, H, W, 4]

    def decode(self, z):
        """autoencoder.py:333-338. z: [n, embed_dim, h, w] -> [n, out_ch, 8h, 8w]."""
        o = self._decode_nhwc(z, 1.0)
        return ops.nhwc_to_nchw_f32(o, self.decoder.out_ch).to(z.dtype)

    def decode_first_stage(self, z, scale_factor: float = 0.18215, to_uint8: bool = False):
        o = self._decode_nhwc(z, 1.0 / scale_factor)
        if to_uint8:
            return ops.image_to_u8(o)  # [n, H, W, 3] uint8
        return ops.nhwc_to_nchw_f32(o, self.decoder.out_ch).to(z.dtype)

    def encode(self, x):
        raise NotImplementedError("cremage_b200: the VAE encoder is outside this round's hot-path scope (SURVEY 8f N2)")

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("cremage_b200: only AutoencoderKL.decode is on the denoising path")
