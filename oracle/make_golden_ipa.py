"""ORACLE support -- golden for the IP-Adapter token path from the UNMODIFIED reference UNet built with
ipa_num_tokens / ipa_scale (modules/ldm/modules/attention.py:338-341 and CrossAttentionOriginal :623-627,660-681;
openaimodel.py:479-480), tiny config; the context carries ipa_num_tokens extra tokens at its end
(modules/sd/image_generator.py:810-814).
    python oracle/make_golden_ipa.py  ->  tests/golden/tiny_unet_ipa.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import sd_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
IPA_TOKENS, IPA_SCALE = 4, 0.8


def main():
    ref_shim.install()
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    cfg = O.TINY_UNET
    unet = UNetModel(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                     model_channels=cfg.model_channels, attention_resolutions=list(cfg.attention_resolutions),
                     num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult), num_heads=cfg.num_heads,
                     use_spatial_transformer=True, transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                     use_checkpoint=False, legacy=False, ipa_scale=IPA_SCALE, ipa_num_tokens=IPA_TOKENS).eval()
    base = O.make_weights(O.unet_param_shapes(cfg), seed=100)
    ipa_shapes = {k: tuple(v.shape) for k, v in unet.state_dict().items() if k not in base}
    assert ipa_shapes and all(k.endswith(("attn2.to_k_ipa.weight", "attn2.to_v_ipa.weight")) for k in ipa_shapes), sorted(ipa_shapes)[:4]
    ipa = O.make_weights(ipa_shapes, seed=700)
    unet.load_state_dict({**base, **ipa}, strict=True)
    g = np.load(os.path.join(GOLD, "tiny_unet.npz"))
    x, t, ctx = (torch.from_numpy(g[k]) for k in ("x", "t", "context"))
    gen = torch.Generator().manual_seed(701)
    ipa_tokens = torch.randn(ctx.shape[0], IPA_TOKENS, ctx.shape[2], generator=gen)
    ctx_full = torch.cat([ctx, ipa_tokens], dim=1)
    with torch.no_grad():
        out = unet(x, t, context=ctx_full)
    print(f"{len(ipa_shapes)} IPA tensors; output moves by {float((out - torch.from_numpy(g['out'])).abs().max()):.4f} "
          f"(abs max {float(out.abs().max()):.3f})")
    keys = sorted(ipa_shapes)
    np.savez_compressed(os.path.join(GOLD, "tiny_unet_ipa.npz"), out=out.numpy(), ipa_tokens=ipa_tokens.numpy(),
                        ipa_keys=np.array(keys), ipa_shapes=np.array([str(ipa_shapes[k]) for k in keys]),
                        ipa_num_tokens=np.int64(IPA_TOKENS), ipa_scale=np.float64(IPA_SCALE),
                        ipa_checksum=np.float64(O.weights_checksum(ipa)))


if __name__ == "__main__":
    main()
