"""Parity of the tcgen05 implicit-GEMM kernel (cb_igemm) against plain torch fp32 on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

from cremage_b200.ops import ACT  # fp16 (default) or bf16 build of the library

pytestmark = pytest.mark.gpu


def _ops():
    from cremage_b200 import ops
    return ops


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def _close(got, want, rtol=2e-2, atol=2e-2):
    got = got.float().cpu()
    want = want.float().cpu()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= atol + rtol * ref, f"max abs err {err} (ref max {ref})"


@pytest.mark.parametrize("m,k,n", [(128, 64, 32), (1000, 320, 320), (154, 768, 640), (4096, 1280, 1280), (256, 96, 72)])
def test_plain_gemm_bias_residual(m, k, n):
    ops = _ops()
    a = _rand(m, k, seed=1).to(ACT)
    w = _rand(n, k, scale=k ** -0.5, seed=2)
    b = _rand(n, seed=3)
    r = _rand(m, n, seed=4).to(ACT)
    wp = ops.pack_weight(w).cuda()
    out = ops.igemm(a.cuda(), wp, n, bias=b.cuda(), residual=r.cuda())
    torch.cuda.synchronize()
    want = a.float() @ w.to(ACT).float().t() + b + r.float()
    _close(out, want)


def test_gemm_f32_out_ragged_cout_and_silu():
    ops = _ops()
    m, k, n = 300, 128, 4
    a = _rand(m, k, seed=1).to(ACT)
    w = _rand(n, k, scale=k ** -0.5, seed=2)
    b = _rand(n, seed=3)
    out = ops.igemm(a.cuda(), ops.pack_weight(w).cuda(), n, bias=b.cuda(), out_f32=True, act=ops.ACT_SILU)
    torch.cuda.synchronize()
    want = F.silu(a.float() @ w.to(ACT).float().t() + b)
    assert out.dtype == torch.float32 and out.shape == (m, n)
    _close(out, want, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 128), (3, 8, 8, 128, 64), (1, 64, 64, 64, 160), (2, 32, 32, 192, 64), (1, 12, 20, 64, 32)])
def test_conv3x3_stride1(n, h, w, cin, cout):
    ops = _ops()
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7)
    emb = _rand(n, cout, seed=8)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.igemm(x_nhwc, ops.pack_weight(wt).cuda(), cout, taps=ops.TAPS_3X3, bias=b.cuda(), rowbias=emb.cuda())
    torch.cuda.synchronize()
    want = F.conv2d(x.float(), wt.to(ACT).float(), b, padding=1) + emb[:, :, None, None]
    _close(out.view(n, h, w, cout).permute(0, 3, 1, 2), want)


def test_conv3x3_two_sources_is_concat():
    ops = _ops()
    n, h, w, c0, c1, cout = 2, 16, 16, 128, 64, 128
    x0 = _rand(n, c0, h, w, seed=1).to(ACT)
    x1 = _rand(n, c1, h, w, seed=2).to(ACT)
    wt = _rand(cout, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5, seed=3)
    out = ops.igemm(x0.permute(0, 2, 3, 1).contiguous().cuda(), ops.pack_weight(wt, (c0, c1)).cuda(), cout,
                    a1=x1.permute(0, 2, 3, 1).contiguous().cuda(), taps=ops.TAPS_3X3)
    torch.cuda.synchronize()
    want = F.conv2d(torch.cat([x0, x1], 1).float(), wt.to(ACT).float(), None, padding=1)
    _close(out.view(n, h, w, cout).permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (4, 8, 8, 128), (1, 64, 64, 64)])
def test_conv3x3_stride2_parity_planes(n, h, w, c):
    ops = _ops()
    cout = 64
    x = _rand(n, c, h, w, seed=1).to(ACT)
    wt = _rand(cout, c, 3, 3, scale=(9 * c) ** -0.5, seed=2)
    b = _rand(cout, seed=3)
    xs = ops.parity_split(x.permute(0, 2, 3, 1).contiguous().cuda())
    out = ops.igemm(xs.view(4 * n, h // 2, w // 2, c), ops.pack_weight(wt).cuda(), cout, out_grid=(n, h // 2, w // 2),
                    taps=ops.taps_3x3_stride2(n), bias=b.cuda())
    torch.cuda.synchronize()
    want = F.conv2d(x.float(), wt.to(ACT).float(), b, stride=2, padding=1)
    _close(out.view(n, h // 2, w // 2, cout).permute(0, 3, 1, 2), want)


def test_geglu_epilogue():
    ops = _ops()
    m, dim, inner = 512, 64, 256
    a = _rand(m, dim, seed=1).to(ACT)
    w = _rand(2 * inner, dim, scale=dim ** -0.5, seed=2)
    b = _rand(2 * inner, seed=3)
    bn = 128
    wq, bq = ops.pack_geglu(w, b, bn)
    out = ops.igemm(a.cuda(), ops.pack_weight(wq).cuda(), inner, bias=bq.cuda(), mode=ops.EPI_GEGLU, bn=bn)
    torch.cuda.synchronize()
    y = a.float() @ w.to(ACT).float().t() + b
    xh, gate = y.chunk(2, dim=-1)
    _close(out, xh * F.gelu(gate))


def test_heads_epilogue_scatter():
    ops = _ops()
    batch, tokens, dim, heads, d, dpad = 2, 96, 64, 2, 40, 64
    inner = heads * d
    a = _rand(batch * tokens, dim, seed=1).to(ACT)
    w = _rand(3 * inner, dim, scale=dim ** -0.5, seed=2)
    qkv = torch.zeros(3, batch * heads, tokens, dpad, dtype=ACT, device="cuda")
    ops.igemm(a.cuda(), ops.pack_weight(w).cuda(), 3 * inner, mode=ops.EPI_HEADS, out=qkv,
              heads=(d, dpad, heads, tokens, batch * heads * tokens * dpad))
    torch.cuda.synchronize()
    y = (a.float() @ w.to(ACT).float().t()).view(batch, tokens, 3, heads, d).permute(2, 0, 3, 1, 4)
    got = qkv.view(3, batch, heads, tokens, dpad).float().cpu()
    _close(got[..., :d], y)
    assert got[..., d:].abs().max().item() == 0.0


@pytest.mark.parametrize("m,k,n", [(128, 64, 32), (1000, 320, 320), (154, 768, 640), (4096, 1280, 1280), (300, 96, 72),
                                   (65536, 320, 320)])
@pytest.mark.parametrize("what", ["plain", "res", "rowbias"])
def test_staged_epilogue_equals_direct(m, k, n, what):
    """The staged (TMA in / TMA out) and the direct epilogue are two schedules of the same arithmetic: bit-identical
    results, ragged M / N clipped by the tensor map."""
    ops = _ops()
    a = _rand(m, k, seed=1).to(ACT).cuda()
    w = _rand(n, k, scale=k ** -0.5, seed=2)
    b = _rand(n, seed=3).cuda()
    wp = ops.pack_weight(w).cuda()
    kw = {}
    if what == "res":
        kw["residual"] = _rand(m, n, seed=4).to(ACT).cuda()
    if what == "rowbias":
        kw["rowbias"] = _rand(1, n, seed=5).cuda()
    guard = torch.full((m + 64, n), 7.0, dtype=ACT, device="cuda")   # rows past m must stay untouched
    o_dir = ops.igemm(a, wp, n, bias=b, epilogue=ops.EPILOGUE_DIRECT, **kw)
    o_stg = ops.igemm(a, wp, n, bias=b, epilogue=ops.EPILOGUE_STAGED, out=guard[:m], **kw)
    torch.cuda.synchronize()
    assert torch.equal(o_dir, o_stg)
    assert (guard[m:] == 7.0).all()
    want = a.float().cpu() @ w.to(ACT).float().t() + b.cpu()
    if what == "res":
        want = want + kw["residual"].float().cpu()
    if what == "rowbias":
        want = want + kw["rowbias"].cpu()
    _close(o_stg, want)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 128), (3, 8, 8, 128, 64), (1, 64, 64, 64, 160), (1, 12, 20, 64, 32),
                                            (5, 4, 4, 64, 96)])
def test_staged_epilogue_conv_tiles(n, h, w, cin, cout):
    """Staged epilogue over 2-D / 3-D pixel tiles (the warp's 32 rows form a {w, h, n} box) incl. overhanging tiles."""
    ops = _ops()
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7)
    res = _rand(n, h, w, cout, seed=9).to(ACT)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.igemm(x_nhwc, ops.pack_weight(wt).cuda(), cout, taps=ops.TAPS_3X3, bias=b.cuda(),
                    residual=res.view(-1, cout).cuda(), epilogue=ops.EPILOGUE_STAGED)
    ref = ops.igemm(x_nhwc, ops.pack_weight(wt).cuda(), cout, taps=ops.TAPS_3X3, bias=b.cuda(),
                    residual=res.view(-1, cout).cuda(), epilogue=ops.EPILOGUE_DIRECT)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    want = F.conv2d(x.float(), wt.to(ACT).float(), b, padding=1) + res.float().permute(0, 3, 1, 2)
    _close(out.view(n, h, w, cout).permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("m,dim", [(512, 64), (4096, 320), (1000, 128)])
def test_geglu_epilogue_shapes(m, dim):
    ops = _ops()
    inner = 4 * dim
    a = _rand(m, dim, seed=1).to(ACT)
    w = _rand(2 * inner, dim, scale=dim ** -0.5, seed=2)
    b = _rand(2 * inner, seed=3)
    wq, bq = ops.pack_geglu(w, b, 128)
    out = ops.igemm(a.cuda(), ops.pack_weight(wq).cuda(), inner, bias=bq.cuda(), mode=ops.EPI_GEGLU, bn=128)
    torch.cuda.synchronize()
    y = a.float() @ w.to(ACT).float().t() + b
    xh, gate = y.chunk(2, dim=-1)
    _close(out, xh * F.gelu(gate))


@pytest.mark.parametrize("n,h,w,cin,cout,bn", [(2, 16, 16, 64, 128, 128), (3, 8, 8, 128, 64, 64), (1, 64, 64, 64, 160, 160),
                                               (2, 32, 32, 192, 256, 256), (1, 12, 20, 64, 32, 32), (5, 4, 4, 64, 96, 32)])
@pytest.mark.parametrize("epilogue", [1, 2])
def test_cta_pair_conv_equals_single(n, h, w, cin, cout, bn, epilogue):
    """CTA-pair mode (cta_group::2, 256-row tiles over a cluster of 2) vs the single-CTA kernel: same dot products in
    the same order -> bit-identical; odd tile counts leave the pair's second tile fully out of bounds."""
    ops = _ops()
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7)
    res = _rand(n * h * w, cout, seed=9).to(ACT).cuda()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = ops.pack_weight(wt).cuda()
    single = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b.cuda(), residual=res, pair=False, epilogue=epilogue)
    paired = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b.cuda(), residual=res, pair=True, bn=bn, epilogue=epilogue)
    torch.cuda.synchronize()
    want = F.conv2d(x.float(), wt.to(ACT).float(), b, padding=1) + res.float().cpu().view(n, h, w, cout).permute(0, 3, 1, 2)
    _close(paired.view(n, h, w, cout).permute(0, 3, 1, 2), want)
    assert torch.equal(single, paired)


@pytest.mark.parametrize("m,k,n", [(4096, 1280, 1280), (1000, 320, 320), (65536, 2560, 640)])
def test_cta_pair_gemm(m, k, n):
    ops = _ops()
    a = _rand(m, k, seed=1).to(ACT).cuda()
    w = _rand(n, k, scale=k ** -0.5, seed=2)
    b = _rand(n, seed=3).cuda()
    wp = ops.pack_weight(w).cuda()
    o1 = ops.igemm(a, wp, n, bias=b, pair=False)
    o2 = ops.igemm(a, wp, n, bias=b, pair=True)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    _close(o2[:2048], a[:2048].float().cpu() @ w.to(ACT).float().t() + b.cpu())


@pytest.mark.parametrize("n,h,w,cin,cout,bn", [(2, 16, 16, 64, 320, 160), (1, 32, 32, 128, 224, 32), (3, 8, 8, 64, 96, 32),
                                               (1, 64, 64, 64, 640, 160), (2, 16, 16, 64, 160, 32)])
def test_cta_pair_dual_n_subtiles_equal_single(n, h, w, cin, cout, bn):
    """Pair mode with two N sub-tiles per A stage (3-slot TMEM accumulator ring), odd sub-tile counts included."""
    ops = _ops()
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7).cuda()
    emb = _rand(n, cout, seed=8).cuda()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = ops.pack_weight(wt).cuda()
    single = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, rowbias=emb, pair=False)
    dual = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, rowbias=emb, pair=True, bn=bn)
    nodual = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, rowbias=emb, pair=True, bn=bn, nsub=1)
    staged = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, rowbias=emb, pair=True, bn=bn, epilogue=ops.EPILOGUE_STAGED)
    torch.cuda.synchronize()
    assert torch.equal(single, dual) and torch.equal(single, nodual) and torch.equal(single, staged)
    want = F.conv2d(x.float(), wt.to(ACT).float(), b.cpu(), padding=1) + emb.cpu()[:, :, None, None]
    _close(dual.view(n, h, w, cout).permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("n,h,w,cin,cout", [(16, 8, 8, 128, 256), (4, 8, 8, 192, 96), (3, 16, 16, 64, 160), (16, 8, 8, 320, 1280)])
@pytest.mark.parametrize("what", ["bias", "rowbias", "res"])
@pytest.mark.parametrize("pair", [True, False])
def test_split_k_conv(n, h, w, cin, cout, what, pair):
    """Split-K by tap groups (fp32 partials + cb_splitk_reduce) for the few-tile / long-K convs of the 8x8 level."""
    ops = _ops()
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7).cuda()
    kw = {}
    if what == "rowbias":
        kw["rowbias"] = _rand(n, cout, seed=8).cuda()
    if what == "res":
        kw["residual"] = _rand(n * h * w, cout, seed=9).to(ACT).cuda()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = ops.pack_weight(wt).cuda()
    one = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, pair=pair, ksplit=1, **kw)
    for ks in (3, 9):
        split = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, pair=pair, ksplit=ks, **kw)
        again = ops.igemm(x_nhwc, wp, cout, taps=ops.TAPS_3X3, bias=b, pair=pair, ksplit=ks, **kw)
        torch.cuda.synchronize()
        assert torch.equal(split, again)                                   # deterministic
        from tests._models import tol
        assert (split.float() - one.float()).abs().max().item() <= tol(2e-2)   # same sum, different fp32 association
    want = F.conv2d(x.float(), wt.to(ACT).float(), b.cpu(), padding=1)
    if what == "rowbias":
        want = want + kw["rowbias"].cpu()[:, :, None, None]
    if what == "res":
        want = want + kw["residual"].float().cpu().view(n, h, w, cout).permute(0, 3, 1, 2)
    _close(split.view(n, h, w, cout).permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 4, 128), (1, 64, 64, 4, 320), (2, 32, 32, 3, 128), (1, 8, 8, 9, 64)])
def test_conv3x3_tiny_cin_by_channel_oob_fill(n, h, w, cin, cout):
    """conv_in layers (4 latent / 3 RGB channels): the NHWC input carries ceil8(cin) channels and the 64-channel TMA box
    reads the rest as out-of-bounds zeros; the weights are zero padded to 64 by pack_weight."""
    ops = _ops()
    cpad = -(-cin // 8) * 8
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7)
    x_nhwc = torch.zeros(n, h, w, cpad, dtype=ACT)
    x_nhwc[..., :cin] = x.permute(0, 2, 3, 1)
    out = ops.igemm(x_nhwc.cuda(), ops.pack_weight(wt).cuda(), cout, taps=ops.TAPS_3X3, bias=b.cuda())
    torch.cuda.synchronize()
    want = F.conv2d(x.float(), wt.to(ACT).float(), b, padding=1)
    _close(out.view(n, h, w, cout).permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 64, 64), (1, 16, 16, 128, 96), (2, 32, 32, 64, 128), (1, 12, 20, 64, 64), (3, 16, 16, 320, 320)])
def test_conv3x3_after_nearest_upsample_folded(n, h, w, cin, cout):
    """conv3x3(F.interpolate(x, 2, 'nearest')) as four 2x2 convs over the low-resolution tensor, each writing its output
    parity class through a strided TMA-store map (openaimodel.py:113-123, model.py:60-64)."""
    ops = _ops()
    x = _rand(n, cin, h, w, seed=5).to(ACT)
    wt = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=6)
    b = _rand(cout, seed=7)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.conv3x3_up2x(x_nhwc, ops.pack_weight_up2x(wt.cuda()), cout, b.cuda())
    torch.cuda.synchronize()
    assert out.shape == (n, 2 * h, 2 * w, cout)
    want = F.conv2d(F.interpolate(x.float(), scale_factor=2, mode="nearest"), wt, b, padding=1)
    _close(out.permute(0, 3, 1, 2), want)
    # and against the unfolded form on the same kernels (upsample2x + conv3x3): only the weight rounding differs
    ref = ops.igemm(ops.upsample2x(x_nhwc), ops.pack_weight(wt).cuda(), cout, taps=ops.TAPS_3X3, bias=b.cuda())
    _close(out.view(-1, cout), ref)


@pytest.mark.parametrize("rows,dim,nout,geglu", [(4096, 320, 960, False), (1000, 640, 640, False), (512, 1280, 320, False),
                                                 (2048, 320, 1280, True), (300, 640, 2560, True)])
def test_layernorm_folded_into_producer_and_consumer(rows, dim, nout, geglu):
    """BasicTransformerBlock's LayerNorm without a LayerNorm launch (ldm/modules/attention.py:900-912): the producer
    GEMM (+ residual) writes per-row partial sums of its 16-bit output, the consumer GEMM reads the un-normalised rows
    with gamma folded into its (row-centred) weights and applies rstd * acc + b' in the epilogue."""
    from cremage_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(rows + dim)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc)
    a = r(rows, dim).to(ops.ACT).cuda()
    res = (r(rows, dim, sc=2.0) + 1.5).to(ops.ACT).cuda()            # residual stream with a non-zero mean
    w0, b0 = r(dim, dim, sc=dim ** -0.5), r(dim, sc=0.1)
    gamma, beta = 1.0 + r(dim, sc=0.2), r(dim, sc=0.2)
    w1, b1 = r(2 * nout if geglu else nout, dim, sc=dim ** -0.5), r(2 * nout if geglu else nout, sc=0.1)
    x = ops.igemm(a, ops.pack_weight(w0.cuda()), dim, bias=b0.cuda(), residual=res, ln_stats=True)
    part = getattr(x, "_ln_part", None)
    assert part is not None and part.shape[0] == rows and part.shape[2] == 2
    # the partial sums are those of the 16-bit output
    xs = x.float()
    assert torch.allclose(part[:, :, 0].sum(1), xs.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(part[:, :, 1].sum(1), (xs * xs).sum(1), rtol=1e-4, atol=1e-2)
    wf, bf = ops.fold_layernorm(w1.cuda(), b1.cuda(), gamma.cuda(), beta.cuda())
    ln = ops.LnFold(part, dim, 1e-5)
    if geglu:
        wq, bq = ops.pack_geglu(wf, bf, ops.GEGLU_BN)
        got = ops.igemm(x, ops.pack_weight(wq), nout, bias=bq.contiguous(), mode=ops.EPI_GEGLU, bn=ops.GEGLU_BN, ln_in=ln)
        y = torch.nn.functional.layer_norm(xs, (dim,), gamma.cuda(), beta.cuda(), 1e-5) @ w1.cuda().t() + b1.cuda()
        want = y[:, :nout] * torch.nn.functional.gelu(y[:, nout:])
    else:
        got = ops.igemm(x, ops.pack_weight(wf), nout, bias=bf, ln_in=ln)
        want = torch.nn.functional.layer_norm(xs, (dim,), gamma.cuda(), beta.cuda(), 1e-5) @ w1.cuda().t() + b1.cuda()
    err = (got.float() - want).abs().max().item()
    ref = want.abs().max().item()
    # the stand-alone form (cb_layernorm -> 16-bit -> plain GEMM) against the same fp32 result, for comparison
    h = ops.layernorm(x, gamma.cuda().float().contiguous(), beta.cuda().float().contiguous(), 1e-5)
    if geglu:
        wq0, bq0 = ops.pack_geglu(w1.cuda(), b1.cuda(), ops.GEGLU_BN)
        alone = ops.igemm(h, ops.pack_weight(wq0), nout, bias=bq0.contiguous(), mode=ops.EPI_GEGLU, bn=ops.GEGLU_BN)
    else:
        alone = ops.igemm(h, ops.pack_weight(w1.cuda()), nout, bias=b1.cuda().float().contiguous())
    err_alone = (alone.float() - want).abs().max().item()
    print(f"[parity] LayerNorm fold rows={rows} dim={dim} nout={nout} geglu={geglu}: max_abs_err={err:.3e} "
          f"(stand-alone LayerNorm + GEMM: {err_alone:.3e}) ref_absmax={ref:.2f}")
    from tests._models import tol
    assert err <= tol(2e-2) * max(ref, 1.0)
