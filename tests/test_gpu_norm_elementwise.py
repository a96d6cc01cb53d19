"""Parity of the bandwidth kernels (GroupNorm+SiLU, LayerNorm, row softmax, layout / sampler kernels) vs torch fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

from cremage_b200.ops import ACT  # fp16 (default) or bf16 build of the library

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale + shift


@pytest.mark.parametrize("n,h,w,c0,c1,silu,eps", [
    (2, 64, 64, 320, 0, True, 1e-5), (2, 16, 16, 1280, 640, True, 1e-5), (3, 8, 8, 1280, 1280, True, 1e-5),
    (1, 32, 32, 640, 320, False, 1e-6), (1, 128, 128, 128, 0, True, 1e-6), (2, 5, 7, 64, 0, True, 1e-6),
])
def test_groupnorm_silu(n, h, w, c0, c1, silu, eps):
    from cremage_b200 import ops
    c = c0 + c1
    x = _rand(n, c, h, w, seed=1, scale=1.5, shift=0.3).to(ACT)
    gamma = _rand(c, seed=2, scale=0.2, shift=1.0)
    beta = _rand(c, seed=3, scale=0.2)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    x0 = xn[..., :c0].contiguous()
    x1 = xn[..., c0:].contiguous() if c1 else None
    out = ops.groupnorm(x0, gamma.cuda(), beta.cuda(), eps, silu, x1=x1)
    torch.cuda.synchronize()
    want = F.group_norm(x.float(), 32, gamma, beta, eps)
    if silu:
        want = F.silu(want)
    got = out.float().cpu().permute(0, 3, 1, 2)
    assert (got - want).abs().max().item() < 3e-2


@pytest.mark.parametrize("rows,c", [(1000, 320), (64, 640), (257, 1280), (33, 64), (5, 2048)])
def test_layernorm(rows, c):
    from cremage_b200 import ops
    x = _rand(rows, c, seed=1, scale=2.0, shift=-0.5).to(ACT)
    gamma = _rand(c, seed=2, scale=0.2, shift=1.0)
    beta = _rand(c, seed=3, scale=0.2)
    out = ops.layernorm(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    torch.cuda.synchronize()
    want = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    assert (out.float().cpu() - want).abs().max().item() < 3e-2


def test_softmax_rows_inplace():
    from cremage_b200 import ops
    s = _rand(300, 4096, seed=1, scale=3.0).to(ACT)
    got = ops.softmax_rows_(s.clone().cuda(), 0.5)
    torch.cuda.synchronize()
    want = torch.softmax(s.float() * 0.5, dim=-1)
    assert (got.float().cpu() - want).abs().max().item() < 2e-3


def test_layout_roundtrip_and_upsample_and_parity():
    from cremage_b200 import ops
    x = _rand(2, 4, 16, 16, seed=1)
    nhwc = ops.nchw_to_nhwc(x.cuda(), c_pad=8, scale=0.5)
    want = torch.zeros(2, 16, 16, 8)
    want[..., :4] = (x * 0.5).permute(0, 2, 3, 1)
    assert torch.equal(nhwc.float().cpu(), want.to(ACT).float())
    back = ops.nhwc_to_nchw_f32(nhwc, 4)
    assert torch.equal(back.cpu(), want[..., :4].to(ACT).float().permute(0, 3, 1, 2))
    y = _rand(2, 64, 6, 10, seed=2).to(ACT)
    wide = ops.nchw_to_nhwc(y.cuda())
    assert torch.equal(wide.cpu(), y.permute(0, 2, 3, 1))
    up = ops.upsample2x(wide)
    assert torch.equal(up.cpu().permute(0, 3, 1, 2).float(), F.interpolate(y.float(), scale_factor=2, mode="nearest"))
    ps = ops.parity_split(wide).cpu()
    for ph in range(2):
        for pw in range(2):
            assert torch.equal(ps[2 * ph + pw], y.permute(0, 2, 3, 1)[:, ph::2, pw::2, :])


def test_timestep_embedding_and_small_conv():
    from cremage_b200 import ops
    dim = 320
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    t = torch.tensor([999.0, 946.4210205, 0.0, 1.0])
    got = ops.timestep_embedding(t.cuda(), dim, freqs.cuda())
    args = t[:, None] * freqs[None]
    want = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    assert (got.float().cpu() - want).abs().max().item() < 1e-2
    x = _rand(2, 4, 16, 16, seed=1).to(ACT)
    w = _rand(320, 4, 3, 3, seed=2, scale=1 / 6)
    b = _rand(320, seed=3)
    xn = ops.nchw_to_nhwc(x.cuda(), c_pad=8)
    out = ops.conv3x3_small_cin(xn, 4, w.permute(2, 3, 1, 0).contiguous().cuda(), b.cuda(), 320)
    torch.cuda.synchronize()
    want = F.conv2d(x.float(), w, b, padding=1)
    assert (out.float().cpu().permute(0, 3, 1, 2) - want).abs().max().item() < 3e-2


def test_sampler_steps_match_reference_expressions():
    from cremage_b200 import ops
    b = 3
    x = _rand(b, 4, 16, 16, seed=1, scale=5.0)
    eps2 = _rand(2 * b, 4, 16, 16, seed=2)
    noise = _rand(b, 4, 16, 16, seed=3)
    s = 7.5
    sigma, sd, su = 6.2049351, 4.1, 2.3
    xo, den = ops.step_euler_ancestral(x.cuda(), eps2.cuda(), noise.cuda(), s, sigma, sd, su, want_denoised=True)
    du, dc = (x + eps2[:b] * -sigma), (x + eps2[b:] * -sigma)
    dref = du + s * (dc - du)
    xr = x + (x - dref) / sigma * (sd - sigma)
    xr = xr + noise * su
    assert (den.cpu() - dref).abs().max().item() < 1e-4
    assert (xo.cpu() - xr).abs().max().item() < 1e-4
    old = _rand(b, 4, 16, 16, seed=4)
    xo, den = ops.step_dpmpp_2m(x.cuda(), eps2.cuda(), old.cuda(), s, sigma, 0.7, -0.3, 1.4, 0.4)
    xr = 0.7 * x - (-0.3) * (1.4 * dref - 0.4 * old)
    assert (xo.cpu() - xr).abs().max().item() < 1e-4
    xo, x0 = ops.step_ddim(x.cuda(), eps2.cuda(), None, s, 0.8, 0.6, 0.85, 0.5267, 0.0)
    e = eps2[:b] + s * (eps2[b:] - eps2[:b])
    p0 = (x - 0.6 * e) / 0.8
    xr = 0.85 * p0 + 0.5267 * e
    assert (x0.cpu() - p0).abs().max().item() < 1e-4
    assert (xo.cpu() - xr).abs().max().item() < 1e-4
    xin = ops.cfg_scale_input(x.cuda(), 0.25)
    assert torch.equal(xin.cpu(), torch.cat([x * 0.25, x * 0.25]))


# ---------------------------------------------------------------------------------------------------------------
# GroupNorm statistics fused into the producing cb_igemm launch (gn_partials) + fold + streaming apply
# ---------------------------------------------------------------------------------------------------------------
def _conv_with_stats(ops, x_nhwc, cin, cout, seed, *, epilogue=0, pair=None, residual=None, rowbias=None, bn=None, taps3=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    k = 3 if taps3 else 1
    wt = torch.randn(cout, cin, k, k, generator=g) * (k * k * cin) ** -0.5
    b = torch.randn(cout, generator=g) * 0.5
    out = ops.igemm(x_nhwc, ops.pack_weight(wt).cuda(), cout, taps=ops.TAPS_3X3 if taps3 else ops.TAPS_1X1, bias=b.cuda(),
                    residual=residual, rowbias=rowbias, epilogue=epilogue, pair=pair, bn=bn, ksplit=1, gn_stats=True)
    return out


@pytest.mark.parametrize("n,h,w,cin,cout,epilogue,pair,what", [
    (2, 32, 32, 256, 320, 1, False, "plain"),     # direct epilogue, single CTA, ragged N tiles (320 = 2 x 160)
    (2, 32, 32, 256, 320, 2, False, "res"),       # staged epilogue + residual panel
    (2, 32, 32, 256, 320, 0, True, "rowb"),       # CTA pairs (auto epilogue), per-image bias
    (3, 8, 8, 256, 128, 0, False, "res"),         # two images per 128-row tile (tn = 2), odd image count
    (1, 24, 40, 128, 64, 1, False, "plain"),      # tiles overhang the pixel grid
    (2, 16, 16, 128, 256, 2, True, "plain"),      # staged + pairs
])
def test_fused_groupnorm_partials_match_tensor_sums(monkeypatch, n, h, w, cin, cout, epilogue, pair, what):
    from cremage_b200 import ops
    monkeypatch.setattr(ops, "GN_FUSE_MIN_K_CHUNKS", 1)
    monkeypatch.setattr(ops, "GN_FUSE_MIN_BYTES", 0)
    x = _rand(n, h, w, cin, seed=11).to(ACT).cuda()
    res = _rand(n * h * w, cout, seed=12).to(ACT).cuda() if what == "res" else None
    rowb = _rand(n, cout, seed=13).cuda() if what == "rowb" else None
    out = _conv_with_stats(ops, x, cin, cout, 14, epilogue=epilogue, pair=pair, residual=res, rowbias=rowb)
    torch.cuda.synchronize()
    part = getattr(out, "_gn_part", None)
    assert part is not None and part.shape[0] == n and part.shape[2] == 2 and part.shape[3] == cout // 2
    o = out.float().view(n, h * w, cout // 2, 2)
    want_sum = o.sum(dim=(1, 3))
    want_sq = (o * o).sum(dim=(1, 3))
    got = part.double().sum(dim=1)
    scale = want_sq.abs().max().item()
    assert (got[:, 0].float() - want_sum).abs().max().item() <= 1e-4 * max(1.0, want_sum.abs().max().item())
    assert (got[:, 1].float() - want_sq).abs().max().item() <= 1e-4 * max(1.0, scale)


@pytest.mark.parametrize("n,h,w,c0,c1,silu", [(2, 32, 32, 320, 0, True), (2, 16, 16, 128, 64, True), (3, 8, 8, 256, 128, False),
                                               (1, 64, 64, 128, 0, True)])
def test_groupnorm_from_fused_partials(monkeypatch, n, h, w, c0, c1, silu):
    """cb_groupnorm_from_partials (fold + one streaming pass) == the stand-alone GroupNorm on the same tensors; with
    c1 > 0 the 32 groups straddle the two sources of the concat."""
    from cremage_b200 import ops
    monkeypatch.setattr(ops, "GN_FUSE_MIN_K_CHUNKS", 1)
    monkeypatch.setattr(ops, "GN_FUSE_MIN_BYTES", 0)
    x = _rand(n, h, w, 64, seed=21).to(ACT).cuda()
    a = ops.nhwc(_conv_with_stats(ops, x, 64, c0, 22), n, h, w, c0)
    b = ops.nhwc(_conv_with_stats(ops, x, 64, c1, 23, taps3=False), n, h, w, c1) if c1 else None
    assert getattr(a, "_gn_part", None) is not None and (b is None or getattr(b, "_gn_part", None) is not None)
    c = c0 + c1
    gamma = _rand(c, seed=2, scale=0.2, shift=1.0).cuda()
    beta = _rand(c, seed=3, scale=0.2).cuda()
    prof = ops.LaunchProfile()
    with prof:
        got = ops.groupnorm(a, gamma, beta, 1e-5, silu, x1=b)
    assert any(k[0] == "cb_groupnorm_from_partials" for k in prof.by_shape()), prof.by_shape()
    plain_a = a.clone()
    plain_b = b.clone() if b is not None else None
    ref = ops.groupnorm(plain_a, gamma, beta, 1e-5, silu, x1=plain_b)
    torch.cuda.synchronize()
    xa = torch.cat([a, b], dim=-1) if b is not None else a
    want = F.group_norm(xa.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    if silu:
        want = F.silu(want)
    want = want.permute(0, 2, 3, 1)
    assert (got.float() - want).abs().max().item() < 3e-2
    assert (got.float() - ref.float()).abs().max().item() < 1e-2


@pytest.mark.parametrize("n,h,w,c0,c1", [(2, 128, 128, 128, 0), (2, 64, 64, 256, 128), (1, 96, 80, 64, 0)])
def test_streaming_groupnorm_is_bit_reproducible(monkeypatch, n, h, w, c0, c1):
    """The bulk-copy ring of the apply pass re-fills its stages while other warps still compute, and the statistics are
    folded from per-tile partials: twenty runs must agree bit for bit (a stage re-used too early or an unordered
    reduction would show up as run-to-run differences), and with the torch fp32 reference."""
    from cremage_b200 import ops
    monkeypatch.setattr(ops, "GN_FUSE_MIN_K_CHUNKS", 1)
    monkeypatch.setattr(ops, "GN_FUSE_MIN_BYTES", 0)
    x = _rand(n, h, w, 64, seed=31).to(ACT).cuda()
    a = ops.nhwc(_conv_with_stats(ops, x, 64, c0, 32), n, h, w, c0)
    b = ops.nhwc(_conv_with_stats(ops, x, 64, c1, 33), n, h, w, c1) if c1 else None
    c = c0 + c1
    gamma, beta = _rand(c, seed=2, scale=0.2, shift=1.0).cuda(), _rand(c, seed=3, scale=0.2).cuda()
    first = ops.groupnorm(a, gamma, beta, 1e-6, True, x1=b)
    for _ in range(20):
        assert torch.equal(ops.groupnorm(a, gamma, beta, 1e-6, True, x1=b), first)
    xa = torch.cat([a, b], dim=-1) if b is not None else a
    want = F.silu(F.group_norm(xa.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-6)).permute(0, 2, 3, 1)
    assert (first.float() - want).abs().max().item() < 3e-2


def test_streaming_groupnorm_ring_reuse_at_the_largest_vae_tensor(monkeypatch):
    """8 x 512 x 512 x 128 (the VAE decoder's last level at batch 8): 32 768 chunks go through the apply pass's
    bulk-copy ring per call.  The stage refill is an async-proxy write over shared memory that generic-proxy loads have
    just read; without the proxy fence in front of the CTA barrier about one chunk in a million was read after the
    next one had started to land -- one wrong pixel in ~1-3 % of the calls (tools/gn_repro.py), far too rare for the
    small shapes above.  400 calls must agree bit for bit."""
    from cremage_b200 import ops
    monkeypatch.setattr(ops, "GN_FUSE_MIN_K_CHUNKS", 1)
    monkeypatch.setattr(ops, "GN_FUSE_MIN_BYTES", 0)
    n, h, w, c = 8, 512, 512, 128
    x = _rand(n, h, w, 64, seed=41).to(ACT).cuda()
    a = ops.nhwc(_conv_with_stats(ops, x, 64, c, 42), n, h, w, c)
    assert getattr(a, "_gn_part", None) is not None and a._gn_part.shape[1] > 32   # the folded-table path
    gamma, beta = _rand(c, seed=2, scale=0.2, shift=1.0).cuda(), _rand(c, seed=3, scale=0.2).cuda()
    first = ops.groupnorm(a, gamma, beta, 1e-6, True).clone()
    bad = sum(0 if torch.equal(ops.groupnorm(a, gamma, beta, 1e-6, True), first) else 1 for _ in range(400))
    assert bad == 0, f"{bad} of 400 repeats differ"


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 128), (1, 32, 32, 128, 64), (3, 8, 8, 64, 320)])
def test_groupnorm_after_folded_upsample_conv_uses_the_shared_partial_table(monkeypatch, n, h, w, cin, cout):
    """The four parity launches of conv3x3_up2x fill disjoint row ranges of one GroupNorm partial table."""
    from cremage_b200 import ops
    monkeypatch.setattr(ops, "GN_FUSE_MIN_BYTES", 0)
    x = _rand(n, h, w, cin, seed=41).to(ACT).cuda()
    wt = _rand(cout, cin, 3, 3, seed=42, scale=(9 * cin) ** -0.5)
    b = _rand(cout, seed=43, scale=0.5)
    y = ops.conv3x3_up2x(x, ops.pack_weight_up2x(wt.cuda()), cout, b.cuda())
    part = getattr(y, "_gn_part", None)
    assert part is not None and part.shape[0] == n and part.shape[1] % 4 == 0
    o = y.float().view(n, 4 * h * w, cout // 2, 2)
    got = part.double().sum(dim=1)
    assert (got[:, 0].float() - o.sum(dim=(1, 3))).abs().max().item() <= 1e-4 * max(1.0, o.sum(dim=(1, 3)).abs().max().item())
    assert (got[:, 1].float() - (o * o).sum(dim=(1, 3))).abs().max().item() <= 1e-4 * max(1.0, (o * o).sum(dim=(1, 3)).abs().max().item())
    gamma, beta = _rand(cout, seed=2, scale=0.2, shift=1.0).cuda(), _rand(cout, seed=3, scale=0.2).cuda()
    prof = ops.LaunchProfile()
    with prof:
        out = ops.groupnorm(y, gamma, beta, 1e-5, True)
    assert any(k[0] == "cb_groupnorm_from_partials" for k in prof.by_shape())
    want = F.silu(F.group_norm(y.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert (out.float() - want).abs().max().item() < 3e-2
