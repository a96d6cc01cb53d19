"""Parity of the fused tcgen05 attention kernel (cb_attention) against softmax(QK^T * scale) V in torch fp32."""
import pytest
import torch

from cremage_b200.ops import ACT  # fp16 (default) or bf16 build of the library

pytestmark = pytest.mark.gpu


def _mk(rows, cols, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(ACT)


def _ref(q, k, v, batch, heads, nq, nk, d, scale):
    def heads_view(t, n):
        return t.float().cuda().view(batch, n, heads, d).permute(0, 2, 1, 3)            # [b, h, n, d]
    qf, kf, vf = heads_view(q, nq), heads_view(k, nk), heads_view(v, nk)
    o = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1) @ vf                    # [b, h, nq, d]
    return o.permute(0, 2, 1, 3).reshape(batch * nq, heads * d)


@pytest.mark.parametrize("batch,heads,nq,nk,d", [
    (1, 2, 128, 128, 40),      # one block
    (2, 8, 1024, 1024, 40),    # multi-block self attention
    (1, 8, 4096, 4096, 40),    # SD1.5 top level
    (2, 8, 1024, 77, 40),      # cross attention, ragged kv
    (2, 8, 256, 256, 80),      # 32x32 level (P aliases S)
    (2, 8, 100, 333, 160),     # deepest level head dim, ragged q and kv, single-stage K/V
    (1, 4, 300, 300, 64),      # SDXL head dim (no spare columns: explicit row sum)
    (2, 8, 4096, 4096, 40),    # two query tiles per CTA, 32 kv blocks, lazy rescale
    (1, 2, 129, 130, 40),      # second query tile nearly empty
    (1, 2, 640, 640, 128),     # d 128: explicit row sum
    (3, 5, 200, 200, 48),      # d a multiple of 16 below 64: row sums in columns 48..63
    (2, 3, 150, 90, 8),        # tiny head dim
    (1, 8, 16384, 16384, 40),  # hires-fix second pass (BASELINE configs[3]): 16 384-token self-attention, 128 kv blocks
    (2, 20, 1024, 1024, 64),   # SDXL 32x32 level: 20 heads of 64; 160 pairs of query tiles: the last round runs as single tiles
    (20, 8, 512, 512, 80),     # 320 pairs at d 80 (P aliases S, MUFU hand-over between the warpgroups): tail as single tiles
    (2, 10, 4096, 77, 64),     # SDXL cross attention at the 64x64 level
])
@pytest.mark.parametrize("packed_qkv", [False, True])
def test_attention_matches_torch(batch, heads, nq, nk, d, packed_qkv):
    """q / k / v are read in place: separate [tokens, heads*d] tensors, or column slices of one [tokens, 3*heads*d]
    projection output (self-attention) -- the kernel pads the head dim with TMA's zero fill."""
    from cremage_b200 import ops
    inner = heads * d
    if packed_qkv:
        if nq != nk:
            pytest.skip("packed q|k|v is the self-attention layout")
        qkv = _mk(batch * nq, 3 * inner, 1).cuda()
        q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
    else:
        q, k, v = _mk(batch * nq, inner, 1).cuda(), _mk(batch * nk, inner, 2).cuda(), _mk(batch * nk, inner, 3).cuda()
    scale = d ** -0.5
    out = ops.attention(q, k, v, batch, heads, nq, nk, d, scale)
    torch.cuda.synchronize()
    err = (out.float() - _ref(q, k, v, batch, heads, nq, nk, d, scale)).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


def test_attention_peaked_scores_rescale_path():
    """Large score range forces the running-max rescale of the TMEM accumulator in every block."""
    from cremage_b200 import ops
    batch, heads, n, d = 1, 2, 512, 40
    q, k, v = _mk(n, heads * d, 4, 4.0).cuda(), _mk(n, heads * d, 5, 4.0).cuda(), _mk(n, heads * d, 6).cuda()
    scale = d ** -0.5
    out = ops.attention(q, k, v, batch, heads, n, n, d, scale)
    torch.cuda.synchronize()
    err = (out.float() - _ref(q, k, v, batch, heads, n, n, d, scale)).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"
