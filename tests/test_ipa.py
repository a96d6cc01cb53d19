"""IP-Adapter tokens: `ipa_num_tokens` / `ipa_scale` of the reference UNet
(modules/ldm/modules/attention.py:338-341; CrossAttentionOriginal :623-627,660-681).  Golden from the unmodified reference
(oracle/make_golden_ipa.py).  CPU: key layout == the reference's, oracle restatement vs the golden.  GPU: CUDA mirror."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import tol  # noqa: E402
from tests._models import gold, unet_kwargs


def _setup():
    g, base = gold("tiny_unet_ipa.npz"), gold("tiny_unet.npz")
    return g, base, int(g["ipa_num_tokens"]), float(g["ipa_scale"])


def _build(device, t, s):
    from cremage_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device(device):
        return UNetModel(**unet_kwargs(O.TINY_UNET), ipa_num_tokens=t, ipa_scale=s)


def _ipa_weights(m, g):
    base_shapes = O.unet_param_shapes(O.TINY_UNET)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k not in base_shapes}
    w = O.make_weights(shapes, seed=700)
    assert abs(O.weights_checksum(w) - float(g["ipa_checksum"])) < 1e-6
    return w


def _context(g, base):
    return torch.cat([torch.from_numpy(base["context"]), torch.from_numpy(g["ipa_tokens"])], dim=1)


def test_ipa_state_dict_keys_equal_the_reference():
    g, _, t, s = _setup()
    m = _build("meta", t, s)
    base_shapes = O.unet_param_shapes(O.TINY_UNET)
    have = {k: str(tuple(v.shape)) for k, v in m.state_dict().items() if k not in base_shapes}
    assert have == dict(zip([str(k) for k in g["ipa_keys"]], [str(x) for x in g["ipa_shapes"]])) and len(have) == 14
    assert set(_build("meta", 0, 1.0).state_dict()) == set(base_shapes)     # no IPA: exactly the base layout


def test_oracle_ipa_matches_reference_golden():
    g, base, t, s = _setup()
    sd = {**O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100), **_ipa_weights(_build("meta", t, s), g)}
    with torch.no_grad():
        out = O.unet_forward(sd, O.TINY_UNET, torch.from_numpy(base["x"]), torch.from_numpy(base["t"]), _context(g, base),
                             ipa=(s, t))
    assert np.abs(out.numpy() - g["out"]).max() < 5e-5
    assert np.abs(g["out"] - base["out"]).max() > 0.5      # the adapter tokens matter in this fixture


@pytest.mark.gpu
def test_cuda_unet_with_ipa_vs_reference_golden():
    g, base, t, s = _setup()
    m = _build("cpu", t, s)
    m.load_state_dict({**O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100), **_ipa_weights(m, g)}, strict=True)
    m = m.cuda().eval()
    x, ts, ctx = torch.from_numpy(base["x"]).cuda(), torch.from_numpy(base["t"]).cuda(), _context(g, base).cuda()
    out = m(x, ts, context=ctx)
    want = torch.from_numpy(g["out"])
    err = (out.cpu() - want).abs().max().item()
    print(f"[parity] tiny UNet + {t} IP-Adapter tokens (scale {s}): max_abs_err={err:.4e} ref_absmax={want.abs().max():.3f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)
    assert torch.equal(m(x, ts, context=ctx), out)          # graph replay
    with torch.no_grad():                                   # ipa_scale is baked into a packed weight: re-pack on change
        m.input_blocks[1][1].transformer_blocks[0].attn2.to_k_ipa.weight.mul_(1.5)
    assert not torch.equal(m(x, ts, context=ctx), out)
