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
