"""Drop-in mirror of the reference's `ldm.models.autoencoder.AutoencoderKL` (HowToSD/cremage
modules/ldm/models/autoencoder.py:285-338) for the decode side of the path.

`decode(z)` = post_quant_conv (1x1, embed_dim -> z_channels) then Decoder.  The 1x1 conv is fused into the
NCHW->NHWC conversion kernel.  `decode_first_stage(z)` additionally applies the 1/scale_factor of
LatentDiffusion.decode_first_stage (ldm/models/diffusion/ddpm.py:794-798) inside the same kernel and can return the
uint8 HWC image of the reference's post-processing (sd/image_generator.py:1017-1018,1151-1152).
`encode(x)` (autoencoder.py:324-331) = Encoder then quant_conv (1x1, folded into conv_out's weights at pack time: two
linear maps with nothing between them) -> DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...engine import PackedModule, f32, require_cuda
from ..modules.diffusionmodules.model import Decoder, Encoder


class DiagonalGaussianDistribution(object):
    """ldm/modules/distributions/distributions.py:24-37 on the encoder's moments (fp32 NCHW [n, 2c, h, w])."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.deterministic = deterministic
        self._ms = None

    def _mean_std(self):
        if self._ms is None:
            _, mean, std = ops.diag_gaussian(self.parameters, None, want_mean_std=True)
            if self.deterministic:
                std = torch.zeros_like(mean)
            self._ms = (mean, std)
        return self._ms

    @property
    def mean(self):
        return self._mean_std()[0]

    @property
    def std(self):
        return self._mean_std()[1]

    @property
    def var(self):
        return self.std * self.std

    @property
    def logvar(self):
        return torch.clamp(torch.chunk(self.parameters, 2, dim=1)[1], -30.0, 20.0)

    def sample(self, noise=None, scale: float = 1.0):
        """mean + std * randn (:35); `noise` injects the draw (tests / bit-reproducible multi-GPU sharding)."""
        if self.deterministic:
            return ops.diag_gaussian(self.parameters, None, scale)
        if noise is None:
            n, c2, h, w = self.parameters.shape
            noise = torch.randn((n, c2 // 2, h, w), device=self.parameters.device)
        return ops.diag_gaussian(self.parameters, noise.float().contiguous(), scale)

    def mode(self):
        return ops.diag_gaussian(self.parameters, None)


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
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.embed_dim = embed_dim
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        if embed_dim > 8 or ddconfig["z_channels"] > 8:
            raise NotImplementedError("cremage_b200: latent channel counts above 8 are not implemented")

    def _own_params(self):
        return (list(self.post_quant_conv.parameters()) + list(self.quant_conv.parameters()) +
                list(self.encoder.conv_out.parameters()))

    def _pack(self, device):
        zc = self.post_quant_conv.out_channels
        # quant_conv o conv_out: W' = Wq Wout, b' = Wq bout + bq   (exact: both are linear, nothing in between)
        wq = f32(self.quant_conv.weight, device).reshape(self.quant_conv.out_channels, -1)
        wo = f32(self.encoder.conv_out.weight, device)
        wfold = torch.einsum("om,mikl->oikl", wq, wo)
        bfold = wq @ f32(self.encoder.conv_out.bias, device) + f32(self.quant_conv.bias, device)
        return {"w": f32(self.post_quant_conv.weight, device).reshape(zc, self.embed_dim).contiguous(),
                "b": f32(self.post_quant_conv.bias, device),
                "enc_w": ops.pack_weight(wfold), "enc_b": bfold.contiguous()}

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
        """autoencoder.py:324-331. x: [n, 3, H, W] in [-1, 1] -> DiagonalGaussianDistribution over [n, embed_dim, H/8, W/8]."""
        require_cuda(x, "AutoencoderKL.encode")
        p = self.packed(x.device)
        co = self.quant_conv.out_channels
        o = self.encoder._run(ops.nchw_to_nhwc(x.float(), c_pad=8), out_w=p["enc_w"], out_b=p["enc_b"], out_c=co)
        return DiagonalGaussianDistribution(ops.nhwc_to_nchw_f32(o, co))

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("cremage_b200: only AutoencoderKL.decode is on the denoising path")
