#!/usr/bin/env python
"""Run-to-run bit-reproducibility of the pipeline on one GPU: the same batch through the same graph several times.
    python tools/determinism_check.py [workload] [--reps 6]      (knobs through the environment, one process per setting)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    wl = next((a for a in sys.argv[1:] if a in bench.WORKLOADS), "euler20_b8")
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 6
    sdxl = bench.WORKLOADS[wl][1] == "sdxl"
    pipe = bench.build_sdxl_pipeline() if sdxl else bench.build_pipeline()
    b = bench.WORKLOADS[wl][0]
    run = bench.make_runner(pipe, wl, b)
    host, resident = bench.make_host_inputs(wl, 0, 1, b)
    dev = {k: v.cuda() for k, v in host.items()}
    dev.update(resident)
    unet = pipe.model.diffusion_model
    pa, pkw = bench.unet_probe_inputs(wl)
    with torch.no_grad():
        ref_u = None
        bad_u = 0
        for i in range(reps * 3):
            o = unet(*pa, **pkw).clone()
            if ref_u is None:
                ref_u = o
            elif not torch.equal(o, ref_u):
                bad_u += 1
        if not sdxl and bench.WORKLOADS[wl][1] == "euler_a":   # localise: latents of the sampler loop, then the decode of ONE latent
            from cremage_b200.k_diffusion.external import CompVisDenoiser
            from cremage_b200.k_diffusion.sampling import sample_euler_ancestral
            from cremage_b200.ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
            steps = bench.WORKLOADS[wl][2]
            den = CompVisDenoiser(pipe, False).cuda()
            sigmas = den.get_sigmas(steps)

            def latents():
                trace = []
                wrapper = LDMWrapperForKDiffusion(den, dev["cond"], dev["uncond"], bench.CFG_SCALE)
                it = iter(range(steps))
                z = sample_euler_ancestral(wrapper, dev["x_T"] * sigmas[0], sigmas, disable=True,
                                           noise_sampler=lambda s_, sn: dev["noise"][next(it)],
                                           callback=lambda d: trace.append(d["x"].clone()))
                return z.clone(), trace
            z0, t0 = latents()
            first_bad = {}
            for i in range(reps - 1):
                z1, t1 = latents()
                for st, (a, c) in enumerate(zip(t0, t1)):
                    if not torch.equal(a, c):
                        first_bad[i] = (st, float((a - c).abs().max()))
                        break
            print("sampler loop: repeats whose latents differ -> (first differing step, max |dx|):", first_bad or "none")
            i0 = pipe.decode_first_stage(z0, to_uint8=True).clone()
            badv = sum(0 if torch.equal(pipe.decode_first_stage(z0, to_uint8=True), i0) else 1 for _ in range(reps * 2))
            print(f"VAE decode of one latent: {badv}/{reps * 2} repeats differ")
        ref = run(dev).clone()
        bad, worst = 0, 0
        for i in range(reps - 1):
            img = run(dev)
            if not torch.equal(img, ref):
                bad += 1
                worst = max(worst, int((img.int() - ref.int()).abs().max().item()))
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CB_") or k.startswith("CREMAGE_B200_"))
    print(f"[{tag or 'defaults'}] {wl}: unet forward {bad_u}/{reps * 3 - 1} repeats differ; pipeline {bad}/{reps - 1} repeats differ (max |d| u8 {worst})")


if __name__ == "__main__":
    main()
