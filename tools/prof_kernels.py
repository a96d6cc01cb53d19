#!/usr/bin/env python
"""Per-shape timing of the hot kernels on the SD1.5 shapes (UNet batch 16 = 8 images with CFG), CUDA events.

    python tools/prof_kernels.py [--only attn|igemm|norm] [--iters 10] [--batch 16]

Prints one line per shape: ms, achieved TFLOP/s or GB/s, fraction of the measured peak. Also the target of the ncu
captures kept under profiles/ (run with --iters 1 under ncu).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cremage_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["hbm_gbs"]
    return 1590.0, 6650.0


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def act(*shape):
    return (torch.randn(*shape, device="cuda") * 0.5).to(ops.ACT)


def bench_attn(B, iters, warm):
    tf_peak, _ = peaks()
    print("== attention (cb_attention) ==")
    for name, heads, nq, nk, d in [("self 64x64 d40", 8, 4096, 4096, 40), ("self 32x32 d80", 8, 1024, 1024, 80),
                                   ("self 16x16 d160", 8, 256, 256, 160), ("self 8x8 d160", 8, 64, 64, 160),
                                   ("cross 64x64 d40", 8, 4096, 77, 40), ("cross 32x32 d80", 8, 1024, 77, 80),
                                   ("cross 16x16 d160", 8, 256, 77, 160)]:
        inner = heads * d
        if nq == nk:   # self-attention: q | k | v column slices of one projection output, as the modules run it
            qkv = act(B * nq, 3 * inner)
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
        else:
            q, kv = act(B * nq, inner), act(B * nk, 2 * inner)
            k, v = kv[:, :inner], kv[:, inner:]
        ms = timeit(lambda: ops.attention(q, k, v, B, heads, nq, nk, d, d ** -0.5), iters, warm)
        fl = 4.0 * B * heads * nq * nk * d
        print(f"{name:22s} bh={B * heads:4d} nq={nq:5d} nk={nk:5d} d={d:3d}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
              f"{fl / ms / 1e9 / tf_peak:6.1%} of burst peak")


def bench_igemm(B, iters, warm, filt=""):
    tf_peak, _ = peaks()
    print("== implicit GEMM (cb_igemm) ==")
    shapes = [
        # (name, n, h, w, cin, cout, taps)
        ("conv3x3 320->320 @64", B, 64, 64, 320, 320, 9), ("conv3x3 640->640 @32", B, 32, 32, 640, 640, 9),
        ("conv3x3 1280->1280 @16", B, 16, 16, 1280, 1280, 9), ("conv3x3 1280->1280 @8", B, 8, 8, 1280, 1280, 9),
        ("conv3x3 2560->1280 @8", B, 8, 8, 2560, 1280, 9), ("conv3x3 2560->1280 @16", B, 16, 16, 2560, 1280, 9),
        ("conv3x3 1920->640 @32", B, 32, 32, 1920, 640, 9), ("conv3x3 960->320 @64", B, 64, 64, 960, 320, 9),
        ("conv3x3 640->320 @64", B, 64, 64, 640, 320, 9), ("conv3x3 640->640 @64", B, 64, 64, 640, 640, 9),
        ("linear 320->320 @4096", 1, 1, B * 4096, 320, 320, 1), ("linear 640->640 @1024", 1, 1, B * 1024, 640, 640, 1),
        ("linear 1280->1280 @256", 1, 1, B * 256, 1280, 1280, 1), ("linear 320->960 qkv", 1, 1, B * 4096, 320, 960, 1),
        ("linear 1280->320 ff2", 1, 1, B * 4096, 1280, 320, 1), ("linear 5120->1280 ff2", 1, 1, B * 256, 5120, 1280, 1),
        ("conv3x3 512->512 @64 vae", B // 2, 64, 64, 512, 512, 9), ("conv3x3 512->512 @128 vae", B // 2, 128, 128, 512, 512, 9),
        ("conv3x3 256->256 @256 vae", B // 2, 256, 256, 256, 256, 9), ("conv3x3 128->128 @512 vae", B // 2, 512, 512, 128, 128, 9),
    ]
    for name, n, h, w, cin, cout, taps in shapes:
        if filt and filt not in name:
            continue
        x = act(n, h, w, cin)
        wt = ops.pack_weight(torch.randn(cout, cin, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device="cuda") * (taps * cin) ** -0.5)
        bias = torch.zeros(cout, device="cuda")
        tp = ops.TAPS_3X3 if taps == 9 else ops.TAPS_1X1
        out = torch.empty(n * h * w, cout, dtype=ops.ACT, device="cuda")
        res = act(n * h * w, cout) if "+res" in name else None
        ms = timeit(lambda: ops.igemm(x, wt, cout, taps=tp, bias=bias, out=out, residual=res), iters, warm)
        fl = 2.0 * n * h * w * taps * cin * cout
        print(f"{name:28s} M={n * h * w:7d} K={taps * cin:6d} N={cout:5d}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
              f"{fl / ms / 1e9 / tf_peak:6.1%} of burst peak")
    # GEGLU
    for name, m, dim in [("geglu 320->2560 @4096", B * 4096, 320), ("geglu 640->5120 @1024", B * 1024, 640),
                         ("geglu 1280->10240 @256", B * 256, 1280)]:
        if filt and filt not in name:
            continue
        x = act(m, dim)
        w = torch.randn(8 * dim, dim, device="cuda") * dim ** -0.5
        b = torch.zeros(8 * dim, device="cuda")
        wq, bq = ops.pack_geglu(w, b, ops.GEGLU_BN)
        wq = ops.pack_weight(wq)
        ms = timeit(lambda: ops.igemm(x, wq, 4 * dim, bias=bq, mode=ops.EPI_GEGLU, bn=ops.GEGLU_BN), iters, warm)
        fl = 2.0 * m * dim * 8 * dim
        print(f"{name:28s} M={m:7d} K={dim:6d} N={8 * dim:5d}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
              f"{fl / ms / 1e9 / tf_peak:6.1%} of burst peak")


def bench_norm(B, iters, warm):
    _, bw_peak = peaks()
    print("== GroupNorm+SiLU / LayerNorm ==")
    for name, n, h, w, c in [("gn 320 @64", B, 64, 64, 320), ("gn 640 @32", B, 32, 32, 640), ("gn 1280 @16", B, 16, 16, 1280),
                             ("gn 2560 @8", B, 8, 8, 2560), ("gn 960 @64", B, 64, 64, 960),
                             ("gn 128 @512 vae", B // 2, 512, 512, 128), ("gn 256 @256 vae", B // 2, 256, 256, 256),
                             ("gn 512 @128 vae", B // 2, 128, 128, 512)]:
        x = act(n, h, w, c)
        g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ms = timeit(lambda: ops.groupnorm(x, g, b, 1e-5, True), iters, warm)
        by = 4.0 * n * h * w * c
        print(f"{name:22s} elems={n * h * w * c / 1e6:8.1f} M: {ms * 1e3:9.1f} us  {by / ms / 1e6:8.1f} GB/s  {by / ms / 1e6 / bw_peak:6.1%} of HBM peak")
    for name, rows, c in [("ln 320 @4096", B * 4096, 320), ("ln 640 @1024", B * 1024, 640), ("ln 1280 @256", B * 256, 1280)]:
        x = act(rows, c)
        g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ms = timeit(lambda: ops.layernorm(x, g, b), iters, warm)
        by = 4.0 * rows * c
        print(f"{name:22s} elems={rows * c / 1e6:8.1f} M: {ms * 1e3:9.1f} us  {by / ms / 1e6:8.1f} GB/s  {by / ms / 1e6 / bw_peak:6.1%} of HBM peak")


def bench_gnfused(B, iters, warm):
    """GroupNorm statistics fused into the producing conv's epilogue: conv without / with partials, then the
    GroupNorm from partials (fold + streaming pass) against the stand-alone GroupNorm on the same tensor."""
    _, bw_peak = peaks()
    print("== fused GroupNorm statistics: conv (plain | +partials) and GroupNorm (stand-alone | from partials) ==")
    saved = ops.GN_FUSE_MIN_K_CHUNKS, ops.GN_FUSE_MIN_BYTES
    ops.GN_FUSE_MIN_K_CHUNKS, ops.GN_FUSE_MIN_BYTES = 1, 0
    for name, n, h, w, cin, cout, k3, res in [
            ("unet 320@64 3x3", B, 64, 64, 320, 320, True, True), ("unet 640@32 3x3", B, 32, 32, 640, 640, True, True),
            ("unet 1280@16 3x3", B, 16, 16, 1280, 1280, True, True), ("unet 1280@8 3x3", B, 8, 8, 1280, 1280, True, False),
            ("unet 320@64 1x1+res", B, 64, 64, 320, 320, False, True), ("unet 640@32 1x1+res", B, 32, 32, 640, 640, False, True),
            ("unet 1280@16 1x1+res", B, 16, 16, 1280, 1280, False, True),
            ("vae 512@128 3x3", B // 2, 128, 128, 512, 512, True, True), ("vae 256@256 3x3", B // 2, 256, 256, 256, 256, True, True),
            ("vae 128@512 3x3", B // 2, 512, 512, 128, 128, True, True)]:
        x = act(n, h, w, cin)
        kk = 3 if k3 else 1
        wt = ops.pack_weight(torch.randn(cout, cin, kk, kk, device="cuda") * (kk * kk * cin) ** -0.5)
        bias = torch.zeros(cout, device="cuda")
        r = act(n * h * w, cout) if res else None
        taps = ops.TAPS_3X3 if k3 else ops.TAPS_1X1
        xin = x if k3 else x.view(n, 1, h * w, cin)
        t0 = timeit(lambda: ops.igemm(xin, wt, cout, taps=taps, bias=bias, residual=r), iters, warm)
        t1 = timeit(lambda: ops.igemm(xin, wt, cout, taps=taps, bias=bias, residual=r, gn_stats=True), iters, warm)
        y = ops.nhwc(ops.igemm(xin, wt, cout, taps=taps, bias=bias, residual=r, gn_stats=True), n, h, w, cout)
        if getattr(y, "_gn_part", None) is None:
            print(f"{name:22s} conv {t0 * 1e3:8.1f} us: no partials (split-K launch)")
            continue
        yp = y.clone()
        g, b = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
        t2 = timeit(lambda: ops.groupnorm(yp, g, b, 1e-5, True), iters, warm)
        t3 = timeit(lambda: ops.groupnorm(y, g, b, 1e-5, True), iters, warm)
        by = 4.0 * n * h * w * cout
        print(f"{name:22s} conv {t0 * 1e3:8.1f} -> {t1 * 1e3:8.1f} us | gn {t2 * 1e3:8.1f} us ({by / t2 / 1e6 / bw_peak:5.1%}) -> "
              f"{t3 * 1e3:8.1f} us ({by / t3 / 1e6:7.1f} GB/s, {by / t3 / 1e6 / bw_peak:5.1%} of HBM peak) | "
              f"net {(t1 + t3 - t0 - t2) * 1e3:+8.1f} us")
    ops.GN_FUSE_MIN_K_CHUNKS, ops.GN_FUSE_MIN_BYTES = saved


def bench_sampler(iters, warm):
    """Per-step latent update kernels (fused CFG mix + update) at sizes where bandwidth, not launch latency, decides:
    algorithmic bytes = 4 B x (streams read + written) per element."""
    _, bw_peak = peaks()
    print("== sampler step kernels (fp32 latents) ==")
    for logn in (14 + 3, 22, 24, 26):   # 2^17 = batch 8 of 4x64x64; the large ones measure bandwidth
        b = 8
        per = (1 << logn) // b
        x = torch.randn(b, per, device="cuda")
        eps2 = torch.randn(2 * b, per, device="cuda")
        noise = torch.randn(b, per, device="cuda")
        old = torch.randn(b, per, device="cuda")
        cases = [
            ("euler_ancestral (x, eps_u, eps_c, noise -> x)", 5, lambda: ops.step_euler_ancestral(x, eps2, noise, 7.5, 5.0, 3.0, 2.0)),
            ("dpmpp_2m (x, eps_u, eps_c, old -> x, denoised)", 6, lambda: ops.step_dpmpp_2m(x, eps2, old, 7.5, 5.0, 0.7, -0.3, 1.5, 0.5)),
            ("ddim (x, eps_u, eps_c -> x, pred_x0)", 5, lambda: ops.step_ddim(x, eps2, None, 7.5, 0.9, 0.43, 0.95, 0.3, 0.0)),
            ("cfg_mix (u, c -> out)", 3, lambda: ops.cfg_mix(eps2[:b], eps2[b:], 7.5)),
        ]
        for name, streams, fn in cases:
            ms = timeit(fn, iters, warm)
            by = 4.0 * streams * (1 << logn)
            print(f"{name:52s} elems=2^{logn}: {ms * 1e3:9.1f} us  {by / ms / 1e6:8.1f} GB/s  {by / ms / 1e6 / bw_peak:6.1%} of HBM peak")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--filter", default="", help="substring filter on igemm shape names")
    a = ap.parse_args()
    if a.only in ("all", "attn"):
        bench_attn(a.batch, a.iters, a.warm)
    if a.only in ("all", "igemm"):
        bench_igemm(a.batch, a.iters, a.warm, a.filter)
    if a.only in ("all", "norm"):
        bench_norm(a.batch, a.iters, a.warm)
    if a.only in ("all", "gnfused"):
        bench_gnfused(a.batch, a.iters, a.warm)
    if a.only in ("all", "sampler"):
        bench_sampler(a.iters, a.warm)
