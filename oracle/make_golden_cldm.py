"""ORACLE support -- golden for the ControlNet residual injection from the UNMODIFIED reference
`ControlledUnetModel` (modules/cldm/cldm.py:28-70), tiny UNet config, random residuals of the right shapes (the
ControlNet that would produce them is out of scope: its outputs are inputs of the hot path).
    python oracle/make_golden_cldm.py  ->  tests/golden/tiny_unet_control.npz"""
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


def main():
    ref_shim.install_lightning_stub()
    from cldm.cldm import ControlledUnetModel
    cfg = O.TINY_UNET
    unet = ControlledUnetModel(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                               model_channels=cfg.model_channels, attention_resolutions=list(cfg.attention_resolutions),
                               num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult),
                               num_heads=cfg.num_heads, use_spatial_transformer=True,
                               transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                               use_checkpoint=False, legacy=False).eval()
    unet.load_state_dict(O.make_weights(O.unet_param_shapes(cfg), seed=100), strict=True)
    g = np.load(os.path.join(GOLD, "tiny_unet.npz"))
    x, t, ctx = (torch.from_numpy(g[k]) for k in ("x", "t", "context"))
    # ControlledUnetModel.forward casts to fp16 whenever the tensors are not on a CUDA device (cldm.py:49-50,68-69: the
    # reference equates "not cuda" with Apple MPS), which cannot even run against fp32 weights on the CPU.  Same remedy
    # as ref_shim's torch.cuda.is_available patch: keep the fp32 CPU run in fp32 by making .half() a no-op here.
    torch.Tensor.half = lambda self, *a, **k: self
    # shapes of the residuals = shapes of the skip tensors (input order) + the middle block's output (last)
    shapes = []
    hooks = [m.register_forward_hook(lambda _m, _i, o: shapes.append(tuple(o.shape))) for m in unet.input_blocks]
    hooks.append(unet.middle_block.register_forward_hook(lambda _m, _i, o: shapes.append(tuple(o.shape))))
    with torch.no_grad():
        plain = unet(x, t, context=ctx)
    for h in hooks:
        h.remove()
    assert float((plain - torch.from_numpy(g["out"])).abs().max()) < 1e-5   # control=None == the plain UNet golden
    gen = torch.Generator().manual_seed(900)
    control = [torch.randn(s, generator=gen) * 0.5 for s in shapes]
    with torch.no_grad():
        lst = list(control)
        out_all = unet(x, t, context=ctx, control=lst)
        assert len(lst) == 0                       # the reference consumed the list from the end
        lst = list(control)
        out_mid = unet(x, t, context=ctx, control=lst, only_mid_control=True)
        assert len(lst) == len(control) - 1
    print(f"{len(control)} residuals; control moves the output by {float((out_all - plain).abs().max()):.4f}, "
          f"mid-only by {float((out_mid - plain).abs().max()):.4f} (abs max {float(plain.abs().max()):.3f})")
    np.savez_compressed(os.path.join(GOLD, "tiny_unet_control.npz"), out_all=out_all.numpy(), out_mid=out_mid.numpy(),
                        **{f"control_{i}": c.numpy() for i, c in enumerate(control)}, n_control=np.int64(len(control)))


if __name__ == "__main__":
    main()
