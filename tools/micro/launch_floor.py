#!/usr/bin/env python
"""Fixed cost of one library launch inside a CUDA graph: chains of N identical launches (each reads the previous one's
output), replayed; per-launch time = replay time / N.  Shapes from tiny to the batch-2 UNet linears."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cremage_b200 import ops

def chain(fn, n=64, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): g.replay()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3   # us

def main():
    torch.manual_seed(0)
    for (m, k, n_) in [(128, 64, 64), (128, 320, 320), (512, 1280, 1280), (2048, 640, 640), (8192, 320, 320), (512, 5120, 1280), (8192, 320, 960)]:
        w = (torch.randn(n_, k, device="cuda") * 0.05).to(ops.ACT)
        wq = ops.pack_weight(w.view(n_, k, 1, 1)) if hasattr(ops, "pack_weight") else w
        b = torch.zeros(n_, device="cuda")
        x = (torch.randn(m, k, device="cuda") * 0.5).to(ops.ACT)
        out = torch.empty(m, n_, device="cuda", dtype=ops.ACT)
        us = chain(lambda: ops.igemm(x, wq, n_, bias=b, out=out))
        print(f"igemm M={m:5d} K={k:5d} N={n_:5d}: {us:7.2f} us per launch in a graph")
    x = (torch.randn(2, 64, 64, 320, device="cuda")).to(ops.ACT)
    for name, fn in [("layernorm 8192x320", lambda: ops.layernorm(x.view(-1, 320), torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")))]:
        try:
            print(f"{name}: {chain(fn):7.2f} us")
        except Exception as e:
            print(name, "skipped:", e)

if __name__ == "__main__":
    main()
