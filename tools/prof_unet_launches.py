#!/usr/bin/env python
"""Per-shape time budget of one UNet forward (eager, CUDA events around every launch of this library).

    python tools/prof_unet_launches.py [--model sd15|sdxl] [--batch 16] [--latent 64]

Prints (kernel, shape) rows sorted by total time: launches, ms, share, achieved TFLOP/s or GB/s.  Event-per-launch
timing serialises the stream, so the sum is a little above the graph-replayed step; the SHARES are what matter.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cremage_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="sd15", choices=["sd15", "sdxl"])
    ap.add_argument("--batch", type=int, default=16, help="UNet batch (2 x images with CFG)")
    ap.add_argument("--latent", type=int, default=0)
    ap.add_argument("--top", type=int, default=60)
    args = ap.parse_args()
    n = args.batch
    if args.model == "sdxl":
        from cremage_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
        with torch.device("meta"):
            unet = UNetModel(**bench.SDXL_UNET)
        lat = args.latent or 128
        kw = dict(context=torch.randn(n, 77, 2048, device="cuda"), y=torch.randn(n, 2816, device="cuda"))
    else:
        from cremage_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
        with torch.device("meta"):
            unet = UNetModel(**bench.SD15_UNET)
        lat = args.latent or 64
        kw = dict(context=torch.randn(n, 77, 768, device="cuda"))
    unet = unet.to_empty(device="cuda")
    bench.init_random_(unet, 0)
    unet.eval()
    unet.use_cuda_graph = False
    x = torch.randn(n, 4, lat, lat, device="cuda")
    t = torch.full((n,), 500.0, device="cuda")
    reps = 3
    with torch.no_grad():
        for _ in range(2):
            unet(x, t, **kw)
        with ops.LaunchProfile() as prof:
            for _ in range(reps):
                unet(x, t, **kw)
    rows = prof.by_shape()
    total = sum(v["ms"] for v in rows.values()) / reps
    print(f"== {args.model} UNet forward, batch {n}, latent {lat}x{lat}: {total:.3f} ms summed over launches ==")
    print(f"{'kernel':22s} {'shape':58s} {'n':>4s} {'ms':>8s} {'share':>6s} {'rate':>14s}")
    for (name, tag), v in sorted(rows.items(), key=lambda kv: -kv[1]["ms"])[:args.top]:
        ms = v["ms"] / reps
        rate = ""
        if v["flops"]:
            rate = f"{v['flops'] / reps / ms / 1e9:8.1f} TF/s"
        elif v["bytes"]:
            rate = f"{v['bytes'] / reps / ms / 1e6:8.1f} GB/s"
        print(f"{name:22s} {tag:58s} {v['launches'] // reps:4d} {ms:8.3f} {ms / total:6.1%} {rate:>14s}")


if __name__ == "__main__":
    main()
