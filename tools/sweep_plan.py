#!/usr/bin/env python
"""Sweep (pair, bn, ksplit) of cb_igemm for the UNet shapes where the planner's pick is in doubt, against the pick.

    python tools/sweep_plan.py [--batch 16]
Times include cb_splitk_reduce for split launches.  Output: one line per shape with the planner's time and the best."""
import argparse, os, sys, itertools, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cremage_b200 import ops
from tools.prof_kernels import timeit, act

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=16); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--only", default="")
    a = ap.parse_args(); B = a.batch
    shapes = [("conv 1280->1280 @8", B, 8, 8, 1280, 1280, 9), ("conv 2560->1280 @8", B, 8, 8, 2560, 1280, 9),
              ("conv 1280->1280 @16", B, 16, 16, 1280, 1280, 9), ("conv 640->640 @32", B, 32, 32, 640, 640, 9),
              ("conv 320->320 @64", B, 64, 64, 320, 320, 9),
              ("lin 1280->1280 @256", 1, 1, B * 256, 1280, 1280, 1), ("lin 5120->1280 @256", 1, 1, B * 256, 5120, 1280, 1),
              ("lin 640->640 @1024", 1, 1, B * 1024, 640, 640, 1), ("lin 320->320 @4096", 1, 1, B * 4096, 320, 320, 1),
              ("lin 1280->320 @4096", 1, 1, B * 4096, 1280, 320, 1), ("lin 320->960 @4096", 1, 1, B * 4096, 320, 960, 1),
              ("lin 640->1920 @1024", 1, 1, B * 1024, 640, 1920, 1), ("lin 1280->3840 @256", 1, 1, B * 256, 1280, 3840, 1),
              ("lin 768->2560 ctx", 1, 1, B * 77, 768, 2560, 1), ("lin 768->1280 ctx", 1, 1, B * 77, 768, 1280, 1),
              ("lin 768->640 ctx", 1, 1, B * 77, 768, 640, 1), ("lin 2560->640 @1024", 1, 1, B * 1024, 2560, 640, 1),
              ("lin 640->640 @4096 xl", 1, 1, B * 4096, 640, 640, 1), ("lin 640->1920 @4096 xl", 1, 1, B * 4096, 640, 1920, 1),
              ("lin 1280->3840 @1024 xl", 1, 1, B * 1024, 1280, 3840, 1)]
    if a.only:
        shapes = [s for s in shapes if a.only in s[0]]
    for name, n, h, w, cin, cout, taps in shapes:
        x = act(n, h, w, cin)
        k = 3 if taps == 9 else 1
        wt = ops.pack_weight(torch.randn(cout, cin, k, k, device="cuda") * (taps * cin) ** -0.5)
        bias = torch.zeros(cout, device="cuda")
        tp = ops.TAPS_3X3 if taps == 9 else ops.TAPS_1X1
        out = torch.empty(n * h * w, cout, dtype=ops.ACT, device="cuda")
        base = timeit(lambda: ops.igemm(x, wt, cout, taps=tp, bias=bias, out=out), a.iters, 3)
        res = []
        for pair, bn, ks in itertools.product((False, True), (32, 64, 128, 160, 256), (1, 3, 9) if taps == 9 else (1,)):
            try:
                ms = timeit(lambda: ops.igemm(x, wt, cout, taps=tp, bias=bias, out=out, pair=pair, bn=bn, ksplit=ks), a.iters, 3)
                res.append((ms, pair, bn, ks))
            except Exception as e:
                pass
        res.sort()
        best = ", ".join(f"{ms*1e3:.1f}us pair={int(p)} bn={bn} ks={ks}" for ms, p, bn, ks in res[:4])
        print(f"{name:22s} M={n*h*w:6d} K={taps*cin:6d} N={cout:5d}: planner {base*1e3:7.1f} us | best: {best}", flush=True)

if __name__ == "__main__":
    main()
