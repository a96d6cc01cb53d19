"""ORACLE support -- golden for the LoRA side branches from the UNMODIFIED reference UNet built with
lora_ranks / lora_weights (modules/ldm/modules/attention.py:79-96,148-168,306-376,966-1055), tiny config.
    python oracle/make_golden_lora.py  ->  tests/golden/tiny_unet_lora.npz"""
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
RANKS, WEIGHTS = [4, 2], [0.7, 1.3]


def main():
    ref_shim.install()
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    cfg = O.TINY_UNET
    unet = UNetModel(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                     model_channels=cfg.model_channels, attention_resolutions=list(cfg.attention_resolutions),
                     num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult), num_heads=cfg.num_heads,
                     use_spatial_transformer=True, transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                     use_checkpoint=False, legacy=False, lora_ranks=RANKS, lora_weights=WEIGHTS).eval()
    base = O.make_weights(O.unet_param_shapes(cfg), seed=100)
    lora_shapes = {k: tuple(v.shape) for k, v in unet.state_dict().items() if k not in base}
    assert all("_lora_" in k for k in lora_shapes), [k for k in lora_shapes if "_lora_" not in k][:5]
    lora = O.make_lora_weights(lora_shapes, seed=500)
    unet.load_state_dict({**base, **lora}, strict=True)
    g = np.load(os.path.join(GOLD, "tiny_unet.npz"))
    x, t, ctx = (torch.from_numpy(g[k]) for k in ("x", "t", "context"))
    with torch.no_grad():
        out = unet(x, t, context=ctx)
    delta = float((out - torch.from_numpy(g["out"])).abs().max())
    print(f"{len(lora_shapes)} LoRA tensors; output moves by {delta:.4f} (abs max {float(out.abs().max()):.3f})")
    keys = sorted(lora_shapes)
    np.savez_compressed(os.path.join(GOLD, "tiny_unet_lora.npz"), out=out.numpy(), lora_keys=np.array(keys),
                        lora_shapes=np.array([str(lora_shapes[k]) for k in keys]), ranks=np.array(RANKS),
                        weights=np.array(WEIGHTS, dtype=np.float64),
                        lora_checksum=np.float64(O.weights_checksum(lora)))


if __name__ == "__main__":
    main()
