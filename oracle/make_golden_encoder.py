"""ORACLE support -- golden for the VAE encoder from the UNMODIFIED reference Encoder
(modules/ldm/modules/diffusionmodules/model.py:375-466) + a 1x1 quant_conv, tiny config, seeded weights / image.
    python oracle/make_golden_encoder.py  ->  tests/golden/tiny_vae_encoder.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import sd_oracle as O  # noqa: E402
from oracle.make_golden import randn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref_shim.install()
    from ldm.modules.diffusionmodules.model import Encoder
    from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
    cfg = O.TINY_VAE
    sd = O.make_weights(O.encoder_param_shapes(cfg), seed=400)
    enc = Encoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=tuple(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                  attn_resolutions=[], dropout=0.0, in_channels=3, resolution=cfg.resolution, z_channels=cfg.z_channels,
                  double_z=True)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    qc = torch.nn.Conv2d(2 * cfg.z_channels, 2 * cfg.embed_dim, 1)
    qc.load_state_dict({"weight": sd["quant_conv.weight"], "bias": sd["quant_conv.bias"]})
    x = randn((2, 3, 32, 48), 11).clamp(-1, 1)      # non-square on purpose
    with torch.no_grad():
        moments = qc(enc.eval()(x))
    post = DiagonalGaussianDistribution(moments)
    noise = randn(tuple(post.mean.shape), 12)
    sample = post.mean + post.std * noise            # distributions.py:35 with the randn draw injected
    print("moments", tuple(moments.shape), float(moments.abs().max()), "sample", float(sample.abs().max()))
    np.savez_compressed(os.path.join(GOLD, "tiny_vae_encoder.npz"), x=x.numpy(), moments=moments.numpy(),
                        noise=noise.numpy(), sample=sample.numpy(), mode=post.mode().numpy(),
                        weights_checksum=np.float64(O.weights_checksum(sd)))


if __name__ == "__main__":
    main()
