"""Parity of the fused tcgen05 attention kernel (cb_attention) against softmax(QK^T * scale) V in torch fp32."""
import pytest
import torch

from cremage_b200.ops import ACT  # fp16 (default) or bf16 build of the library

pytestmark = pytest.mark.gpu


def _mk(bh, n, d, dpad, seed, ones_col=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(bh, n, d, generator=g)
    xp = torch.zeros(bh, n, dpad)
    xp[..., :d] = x
    if ones_col and dpad > d:
        xp[..., d] = 1.0  # cb_attention contract: V's first pad column carries the row-sum ones
    return xp.to(ACT)


@pytest.mark.parametrize("batch,heads,nq,nk,d,dpad", [
    (1, 2, 128, 128, 40, 64),      # one block
    (2, 8, 1024, 1024, 40, 64),    # multi-block self attention, two CTAs per SM
    (1, 8, 4096, 4096, 40, 64),    # SD1.5 top level
    (2, 8, 1024, 77, 40, 64),      # cross attention, ragged kv
    (2, 8, 256, 256, 80, 128),     # 32x32 level
    (2, 8, 100, 333, 160, 192),    # deepest level head dim, ragged q and kv, single-stage K/V
    (1, 4, 300, 300, 64, 64),      # SDXL head dim (no pad column: explicit row sum)
    (2, 8, 4096, 4096, 40, 64),    # two query tiles per CTA, 32 kv blocks, lazy rescale
    (1, 2, 129, 130, 40, 64),      # second query tile nearly empty
    (1, 2, 640, 640, 128, 128),    # dpad 128: two warpgroups, single-stage K/V, explicit row sum
])
def test_attention_matches_torch(batch, heads, nq, nk, d, dpad):
    from cremage_b200 import ops
    bh = batch * heads
    q = _mk(bh, nq, d, dpad, 1)
    k = _mk(bh, nk, d, dpad, 2)
    v = _mk(bh, nk, d, dpad, 3, ones_col=True)
    scale = d ** -0.5
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), batch, heads, nq, nk, d, dpad, scale)
    torch.cuda.synchronize()
    qf, kf, vf = (t[..., :d].float().cuda() for t in (q, k, v))
    want = torch.softmax(qf @ kf.transpose(1, 2) * scale, dim=-1) @ vf          # [bh, nq, d]
    want = want.view(batch, heads, nq, d).permute(0, 2, 1, 3).reshape(batch * nq, heads * d)
    err = (out.float() - want).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


def test_attention_peaked_scores_rescale_path():
    """Large score range forces the running-max rescale of the TMEM accumulator in every block."""
    from cremage_b200 import ops
    batch, heads, n, d, dpad = 1, 2, 512, 40, 64
    q = _mk(heads, n, d, dpad, 4) * 4
    k = _mk(heads, n, d, dpad, 5) * 4
    v = _mk(heads, n, d, dpad, 6, ones_col=True)
    scale = d ** -0.5
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), batch, heads, n, n, d, dpad, scale)
    torch.cuda.synchronize()
    qf, kf, vf = (t[..., :d].float().cuda() for t in (q, k, v))
    want = torch.softmax(qf @ kf.transpose(1, 2) * scale, dim=-1) @ vf
    want = want.view(batch, heads, n, d).permute(0, 2, 1, 3).reshape(batch * n, heads * d)
    err = (out.float() - want).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"
