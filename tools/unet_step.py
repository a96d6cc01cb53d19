#!/usr/bin/env python
"""Graph-replayed UNet step latency (SD1.5, UNet batch 16) + VAE decode time, for A/B runs of tuning knobs
(environment variables read by cremage_b200.ops):   CB_GN_FUSE=0 python tools/unet_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    wl = next((a for a in sys.argv[1:] if a in bench.WORKLOADS), "ddim50_b8")   # e.g. euler20_b1: UNet batch 2
    sdxl = bench.WORKLOADS[wl][1] == "sdxl"
    if sdxl:
        pipe = bench.build_sdxl_pipeline()
        unet, vae = pipe.model.diffusion_model, pipe.first_stage_model
    else:
        pipe = bench.build_pipeline()
        unet, vae = pipe.model.diffusion_model, pipe.first_stage_model
    pa, pkw = bench.unet_probe_inputs(wl)
    lat = 128 if sdxl else 64
    z = torch.randn(bench.WORKLOADS[wl][0], 4, lat, lat, device="cuda")
    if "--quick" in sys.argv:   # under ncu: capture + two replays, nothing else
        with torch.no_grad():
            for _ in range(4):
                unet(*pa, **pkw)
        torch.cuda.synchronize()
        return
    with torch.no_grad():
        for _ in range(3):
            unet(*pa, **pkw)
            vae.decode(z)
        u = min(timed(lambda: unet(*pa, **pkw), 20) for _ in range(3))
        v = min(timed(lambda: vae.decode(z), 5) for _ in range(2))
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CB_"))
    print(f"[{tag or 'defaults'}] {wl}: unet step {u:.3f} ms | vae decode {v:.3f} ms")


if __name__ == "__main__":
    main()
