"""ControlNet residual injection: `ControlledUnetModel.forward(x, timesteps, context, control,
only_mid_control)` of modules/cldm/cldm.py:28-70.  The golden comes from the UNMODIFIED reference class
(oracle/make_golden_cldm.py).  CPU: the oracle restatement against it.  GPU: the CUDA mirror against it, including the
reference's consumption of the caller's `control` list (popped from the end) and `control=None` == plain UNet."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import gold, unet_kwargs


def _inputs():
    g, base = gold("tiny_unet_control.npz"), gold("tiny_unet.npz")
    control = [torch.from_numpy(g[f"control_{i}"]) for i in range(int(g["n_control"]))]
    x, t, ctx = (torch.from_numpy(base[k]) for k in ("x", "t", "context"))
    return g, base, x, t, ctx, control


def test_oracle_control_injection_matches_reference_golden():
    g, base, x, t, ctx, control = _inputs()
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    with torch.no_grad():
        out_all = O.unet_forward(sd, O.TINY_UNET, x, t, ctx, control=control)
        out_mid = O.unet_forward(sd, O.TINY_UNET, x, t, ctx, control=control, only_mid_control=True)
    assert np.abs(out_all.numpy() - g["out_all"]).max() < 5e-5
    assert np.abs(out_mid.numpy() - g["out_mid"]).max() < 5e-5
    assert np.abs(g["out_all"] - base["out"]).max() > 0.3 and np.abs(g["out_mid"] - g["out_all"]).max() > 0.1


def test_controlled_unet_keeps_the_plain_key_layout():
    from cremage_b200.cldm.cldm import ControlledUnetModel
    with torch.device("meta"):
        m = ControlledUnetModel(**unet_kwargs(O.TINY_UNET))
    assert set(m.state_dict()) == set(O.unet_param_shapes(O.TINY_UNET))
    # the attribute walk of cldm.py:44-66 finds the same containers
    assert len(m.input_blocks) + 1 == int(gold("tiny_unet_control.npz")["n_control"]) and len(m.output_blocks) == len(m.input_blocks)


@pytest.mark.gpu
def test_cuda_controlled_unet_vs_reference_golden():
    from cremage_b200.cldm.cldm import ControlledUnetModel
    g, base, x, t, ctx, control = _inputs()
    with torch.device("meta"):
        m = ControlledUnetModel(**unet_kwargs(O.TINY_UNET))
    m = m.to_empty(device="cpu")
    m.load_state_dict(O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100), strict=True)
    m = m.cuda().eval()
    x, t, ctx = x.cuda(), t.cuda(), ctx.cuda()
    from tests._models import tol as _tol
    tol = _tol(2e-2) * max(float(np.abs(g["out_all"]).max()), 1.0)

    lst = [c.cuda() for c in control]
    out_all = m(x, t, context=ctx, control=lst)
    assert lst == []                                             # consumed from the end, like the reference
    err = (out_all.cpu() - torch.from_numpy(g["out_all"])).abs().max().item()
    print(f"[parity] tiny UNet + 5 control residuals: max_abs_err={err:.4e}")
    assert err <= tol

    lst = [c.cuda() for c in control]
    out_mid = m(x, t, context=ctx, control=lst, only_mid_control=True)
    assert len(lst) == len(control) - 1
    assert (out_mid.cpu() - torch.from_numpy(g["out_mid"])).abs().max().item() <= tol

    out_plain = m(x, t, context=ctx)                             # control=None: the plain UNet
    assert (out_plain.cpu() - torch.from_numpy(base["out"])).abs().max().item() <= tol
    # graph replay == the first (capturing) call, and fp16 residuals are accepted
    again = m(x, t, context=ctx, control=[c.cuda() for c in control])
    assert torch.equal(again, out_all)
    half = m(x, t, context=ctx, control=[c.cuda().half() for c in control])
    assert (half.cpu() - torch.from_numpy(g["out_all"])).abs().max().item() <= tol
    with pytest.raises(IndexError):
        m(x, t, context=ctx, control=[control[0].cuda()])
