"""LoRA side branches of the UNet.  The reference adds `up(down(x)) * w * alpha / rank` next to every
attention / feed-forward / proj_in / proj_out projection; cremage_b200 keeps the reference's parameter names and merges
the branches into the packed base weights.  CPU: key layout == the reference's (stored with the golden), oracle merge
reproduces the reference's side-branch forward.  GPU: the CUDA UNet with LoRA against the same golden."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import tol  # noqa: E402
from tests._models import gold, unet_kwargs


def _setup():
    g, base_g = gold("tiny_unet_lora.npz"), gold("tiny_unet.npz")
    ranks, weights = [int(r) for r in g["ranks"]], [float(w) for w in g["weights"]]
    return g, base_g, ranks, weights


def _build(ranks, weights, device):
    from cremage_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device(device):
        return UNetModel(**unet_kwargs(O.TINY_UNET), lora_ranks=ranks, lora_weights=weights)


def _lora_weights(m, g):
    base_shapes = O.unet_param_shapes(O.TINY_UNET)
    lora_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k not in base_shapes}
    lora = O.make_lora_weights(lora_shapes, seed=500)
    assert abs(O.weights_checksum(lora) - float(g["lora_checksum"])) < 1e-6
    return lora


def test_lora_state_dict_keys_equal_the_reference():
    g, _, ranks, weights = _setup()
    m = _build(ranks, weights, "meta")
    base_shapes = O.unet_param_shapes(O.TINY_UNET)
    have = {k: str(tuple(v.shape)) for k, v in m.state_dict().items() if k not in base_shapes}
    want = dict(zip([str(k) for k in g["lora_keys"]], [str(s) for s in g["lora_shapes"]]))
    assert have == want and len(have) == 504
    # without LoRA the key set is exactly the base layout (empty ModuleLists add nothing)
    assert set(_build(None, None, "meta").state_dict()) == set(base_shapes)


def test_oracle_merge_reproduces_reference_side_branches():
    g, base_g, ranks, weights = _setup()
    m = _build(ranks, weights, "meta")
    sd = O.lora_merge(O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100), _lora_weights(m, g), ranks, weights)
    with torch.no_grad():
        out = O.unet_forward(sd, O.TINY_UNET, torch.from_numpy(base_g["x"]), torch.from_numpy(base_g["t"]),
                             torch.from_numpy(base_g["context"]))
    assert np.abs(out.numpy() - g["out"]).max() < 5e-5
    assert np.abs(g["out"] - base_g["out"]).max() > 0.5      # the branches matter in this fixture


@pytest.mark.gpu
def test_cuda_unet_with_lora_vs_reference_golden():
    g, base_g, ranks, weights = _setup()
    m = _build(ranks, weights, "cpu")
    m.load_state_dict({**O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100), **_lora_weights(m, g)}, strict=True)
    m = m.cuda().eval()
    x, t, ctx = (torch.from_numpy(base_g[k]).cuda() for k in ("x", "t", "context"))
    out = m(x, t, context=ctx)
    want = torch.from_numpy(g["out"])
    err = (out.cpu() - want).abs().max().item()
    print(f"[parity] tiny UNet + LoRA ranks {ranks}: max_abs_err={err:.4e} ref_absmax={want.abs().max():.3f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)
    # changing a LoRA tensor in place re-packs (version counters) and changes the output
    with torch.no_grad():
        m.input_blocks[1][1].transformer_blocks[0].attn1.q_lora_alphas[0].mul_(2.0)
    assert not torch.equal(m(x, t, context=ctx), out)
