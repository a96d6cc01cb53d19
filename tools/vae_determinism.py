#!/usr/bin/env python
"""Bit-reproducibility of AutoencoderKL.decode on one latent batch: N repeats, count of outputs that differ from the first.
    python tools/vae_determinism.py [--reps 300] [--batch 8] [--float]    (knobs through the environment)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    reps, b = arg("--reps", 300), arg("--batch", 8)
    pipe = bench.build_pipeline()
    torch.manual_seed(5)
    z = torch.randn(b, 4, 64, 64, device="cuda")
    as_float = "--float" in sys.argv
    dec = (lambda: pipe.decode_first_stage(z)) if as_float else (lambda: pipe.decode_first_stage(z, to_uint8=True))
    with torch.no_grad():
        ref = dec().clone()
        bad, worst, where = 0, 0.0, set()
        for i in range(reps):
            o = dec()
            if not torch.equal(o, ref):
                bad += 1
                d = (o.float() - ref.float()).abs()
                worst = max(worst, float(d.max()))
                idx = torch.nonzero(d.reshape(b, -1).amax(dim=1) > 0).flatten().tolist()
                where.update(idx)
                npx = int((d > 0).sum())
                if bad <= 3:
                    print(f"  repeat {i}: {npx} elements differ, images {idx}")
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CB_") or k.startswith("CREMAGE_B200_"))
    print(f"[{tag or 'defaults'}] VAE decode batch {b}: {bad}/{reps} repeats differ (max |d| {worst:g}, images {sorted(where)})")


if __name__ == "__main__":
    main()
