#!/usr/bin/env python
"""Benchmark of the SD1.5 denoising hot path (UNet + sampler + VAE decode) on B200, driver contract.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ddim50_b8|euler20_b8|euler20_b1]

One "step" = one pass of the hot path over one batch: BASELINE.json configs[1] -- SD1.5 txt2img 512x512, batch 8 per
GPU, 50-step DDIM (eta 0), CFG 7.5 (UNet batch 16), random-init weights, synthetic context / latents -- followed by the
AutoencoderKL decode to uint8 images.  `value` = images/s over all ranks with inputs resident in HBM; `e2e` = the same
through host buffers (pinned H2D of x_T / context / uncond, D2H of the uint8 images) inside the timed region.
Multi-GPU: one rank per GPU, the batch is sharded (8 images per rank, weak scaling), the only collective is the NCCL
all_gather of the decoded uint8 images; time = max over ranks.

`--impl reference` times the reference's CPU path (the oracle port of it: /root/reference does not exist on the GPU
box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (batch per GPU, sampler, steps)
    "ddim50_b8": (8, "ddim", 50),
    "euler20_b8": (8, "euler_a", 20),
    "euler20_b1": (1, "euler_a", 20),
    "dpmpp2m30_b8": (8, "dpmpp_2m", 30),
    # BASELINE.json configs[3]: hires-fix 512 -> 1024, denoise 0.5 (20-step DDIM, 10 steps at 128x128 latents), 4 / GPU
    "hires20_b4": (4, "hires", 20),
}
CFG_SCALE = 7.5
# algorithmic work per image (BASELINE.md section 3, 2*MACs of the reference graph)
GF_UNET_PER_SAMPLE_FWD = 803.27
GF_VAE_PER_IMAGE = 2514.5

SD15_UNET = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                 num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                 transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)
SD15_VAE = dict(embed_dim=4, lossconfig=None,
                ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                              ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0))


def _dtype_name():
    from cremage_b200 import _lib
    return _lib.DTYPE  # fp16 (the reference's GPU precision: model.half() + autocast) or bf16; fp32 accumulate


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def init_random_(module: torch.nn.Module, seed: int):
    """Random-init weights of the SD1.5 architecture (no checkpoints offline): N(0, 1/fan_in) matrices, small biases,
    norm gains near one -- including the tensors the reference zero-initialises, else the UNet output is zero."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * 0.05)
            elif p.dim() == 1:
                p.copy_(1.0 + torch.randn(p.shape, generator=g, device="cuda") * 0.1)
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") / math.sqrt(fan_in))


def build_pipeline():
    from cremage_b200.ldm.models.autoencoder import AutoencoderKL
    from cremage_b200.ldm.models.diffusion.ddpm import LatentDiffusion
    from cremage_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device("meta"):
        unet = UNetModel(**SD15_UNET)
        vae = AutoencoderKL(**SD15_VAE)
    unet = unet.to_empty(device="cuda")
    vae = vae.to_empty(device="cuda")
    init_random_(unet, 0)
    init_random_(vae, 1)
    return LatentDiffusion(unet, vae).cuda().eval()


def make_runner(ldm, workload):
    """Returns run(x_T, cond, uncond) -> uint8 images [b, 512, 512, 3] through the repo's public (reference-mirroring) API."""
    from cremage_b200.k_diffusion.external import CompVisDenoiser
    from cremage_b200.k_diffusion.sampling import get_sigmas_karras, sample_dpmpp_2m, sample_euler_ancestral
    from cremage_b200.ldm.models.diffusion.ddim import DDIMSampler
    from cremage_b200.ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    b, sampler, steps = WORKLOADS[workload]
    if sampler == "hires":
        from cremage_b200.hires import hires_fix_latent
        smp = DDIMSampler(ldm)

        def run(x_T, cond, uncond, noise):
            z, _ = smp.sample(S=steps, batch_size=b, shape=[4, 64, 64], conditioning=cond, eta=0.0, x_T=x_T,
                              unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=uncond, verbose=False)
            z = hires_fix_latent(smp, z, cond, uncond, CFG_SCALE, sampling_steps=steps, strength=0.5, noise=noise)
            return ldm.decode_first_stage(z, to_uint8=True)
        return run
    if sampler == "ddim":
        smp = DDIMSampler(ldm)

        def run(x_T, cond, uncond, noise):
            z, _ = smp.sample(S=steps, batch_size=b, shape=[4, 64, 64], conditioning=cond, eta=0.0, x_T=x_T,
                              unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=uncond, verbose=False)
            return ldm.decode_first_stage(z, to_uint8=True)
        return run
    den = CompVisDenoiser(ldm, False).cuda()
    if sampler == "euler_a":
        sigmas = den.get_sigmas(steps)

        def run(x_T, cond, uncond, noise):
            wrapper = LDMWrapperForKDiffusion(den, cond, uncond, CFG_SCALE)
            it = iter(range(steps))
            z = sample_euler_ancestral(wrapper, x_T * sigmas[0], sigmas, disable=True,
                                       noise_sampler=lambda s, sn: noise[next(it)])
            return ldm.decode_first_stage(z, to_uint8=True)
        return run
    sigmas = get_sigmas_karras(steps, 0.0316386, 14.5521805, device="cuda")

    def run(x_T, cond, uncond, noise):
        wrapper = LDMWrapperForKDiffusion(den, cond, uncond, CFG_SCALE)
        z = sample_dpmpp_2m(wrapper, x_T * sigmas[0], sigmas, disable=True)
        return ldm.decode_first_stage(z, to_uint8=True)
    return run


def kernel_breakdown(ldm, b):
    """One eager UNet forward (CFG batch 2b) + one VAE decode with CUDA events around every launch of this library:
    per-kernel time shares and achieved rates for the roofline section."""
    from cremage_b200 import ops
    unet = ldm.model.diffusion_model
    x = torch.randn(2 * b, 4, 64, 64, device="cuda")
    t = torch.full((2 * b,), 500.0, device="cuda")
    ctx = torch.randn(2 * b, 77, 768, device="cuda")
    z = torch.randn(b, 4, 64, 64, device="cuda")
    saved = unet.use_cuda_graph
    unet.use_cuda_graph = False
    out = {}
    try:
        with torch.no_grad():
            for _ in range(2):
                unet(x, t, context=ctx)
            with ops.LaunchProfile() as prof:
                for _ in range(3):
                    unet(x, t, context=ctx)
            out["unet_fwd"] = {k: {kk: vv / 3 for kk, vv in v.items()} for k, v in prof.summary().items()}
            ldm.first_stage_model.decode(z)
            with ops.LaunchProfile() as prof:
                ldm.first_stage_model.decode(z)
            out["vae_decode"] = prof.summary()
    finally:
        unet.use_cuda_graph = saved
    return out


def run_ours(args):
    import torch.distributed as dist
    from cremage_b200 import engine
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b, sampler, steps = WORKLOADS[args.workload]
    ldm = build_pipeline()
    run = make_runner(ldm, args.workload)

    gen = torch.Generator().manual_seed(1000 + rank)
    h_cond = torch.randn(b, 77, 768, generator=gen).pin_memory()
    h_uncond = torch.randn(b, 77, 768, generator=gen).pin_memory()
    h_xT = torch.randn(b, 4, 64, 64, generator=gen).pin_memory()
    noise = torch.randn(steps, b, 4, 64, 64, generator=gen).cuda() if sampler == "euler_a" else None
    if sampler == "hires":
        noise = torch.randn(b, 4, 128, 128, generator=gen).cuda()
    out_px = 1024 if sampler == "hires" else 512
    h_img = torch.empty(b, out_px, out_px, 3, dtype=torch.uint8).pin_memory()
    d_cond, d_uncond, d_xT = h_cond.cuda(), h_uncond.cuda(), h_xT.cuda()
    gathered = torch.empty(world * b, out_px, out_px, 3, dtype=torch.uint8, device="cuda") if world > 1 else None

    def step_resident():
        img = run(d_xT, d_cond, d_uncond, noise)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        return img

    def step_e2e():
        x, c, u = h_xT.cuda(non_blocking=True), h_cond.cuda(non_blocking=True), h_uncond.cuda(non_blocking=True)
        img = run(x, c, u, noise)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        h_img.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return img

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step_resident()
        launches0 = engine.total_launches()
        clocks = ClockSampler(local)
        clocks.start()
        ms = timed(step_resident, args.steps)
        launches = engine.total_launches() - launches0
        ms_e2e = timed(step_e2e, args.steps)
        clk = clocks.stop()
        # UNet step latency (one CFG-doubled forward through the public API, graph replay)
        unet = ldm.model.diffusion_model
        x2 = torch.randn(2 * b, 4, 64, 64, device="cuda")
        t2 = torch.full((2 * b,), 500.0, device="cuda")
        c2 = torch.randn(2 * b, 77, 768, device="cuda")
        unet(x2, t2, context=c2)
        unet_ms = timed(lambda: unet(x2, t2, context=c2), 10) / 10
        brk = kernel_breakdown(ldm, b) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    imgs = b * world * args.steps
    value = imgs / (ms / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)
    # roofline of the dominant kernel (tcgen05 implicit GEMM) over one UNet forward
    ig = brk["unet_fwd"]["cb_igemm"]
    achieved = ig["flops"] / (ig["ms"] * 1e-3) / 1e12
    total_ms = sum(v["ms"] for v in brk["unet_fwd"].values())
    kernels = {}
    for phase, rec in brk.items():
        tot = sum(v["ms"] for v in rec.values())
        for name, v in rec.items():
            e = {"launches": round(v["launches"], 1), "ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 4)}
            if v["flops"]:
                e["tflops"] = round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)
                e["frac_of_bf16_sustained"] = round(e["tflops"] / pk["bf16_tflops_sustained"], 4)
            if v["bytes"]:
                e["gbs"] = round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)
                e["frac_of_hbm"] = round(e["gbs"] / pk["hbm_gbs"], 4)
            kernels[f"{phase}/{name}"] = e
    gf_per_image = steps * 2 * GF_UNET_PER_SAMPLE_FWD + GF_VAE_PER_IMAGE
    if sampler == "hires":  # + 10 CFG steps at 128x128 latents (9348 GF each) and a 1024x1024 decode (~10.5 TF), BASELINE.md section 3
        gf_per_image = steps * 2 * GF_UNET_PER_SAMPLE_FWD + int(0.5 * steps) * 9348.0 + 10500.0
    cpu = cpu_baseline_sample(steps) if world == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": "SD1.5 512x512 images/sec (UNet + sampler + VAE decode)", "value": round(value, 4),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": _dtype_name(), "data": "synthetic",
        "config": {"workload": f"SD1.5 txt2img 512x512, batch {b}/GPU, {steps}-step {sampler}, CFG {CFG_SCALE}, "
                               f"random-init weights, + AutoencoderKL decode to uint8 ({args.workload})",
                   "global_batch": b * world, "parallelism": f"dp{world} (batch sharded, NCCL all_gather of uint8 images)",
                   "l2": "no flush: weights (1.8 GB) + activations per step exceed the 126 MB L2",
                   "algorithmic_gflop_per_image": round(gf_per_image, 1)},
        "e2e": {"value": round(e2e_value, 4), "unit": "images/s",
                "h2d_bytes_per_step": int(h_xT.numel() * 4 + h_cond.numel() * 4 + h_uncond.numel() * 4),
                "d2h_bytes_per_step": int(h_img.numel())},
        "gpu_launches": int(launches),
        "unet_step_ms": round(unet_ms, 3),
        "model_tflops": round(value * gf_per_image / 1e3, 1),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": pk["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops_sustained"], 4), "traffic": None,
                     "kernel": "igemm_kernel (tcgen05 implicit GEMM: conv3x3 / conv1x1 / linear), all launches of one "
                               f"UNet forward at batch {2 * b}; share of UNet kernel time {ig['ms'] / total_ms:.3f}",
                     "peak_source": pk["source"] + " bf16_tflops_sustained"},
        "kernels": kernels,
        "clocks": clk,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle = port of the reference's path; the only places bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_models():
    from oracle import sd_oracle as O
    usd = O.make_weights(O.unet_param_shapes(O.SD15_UNET), seed=0)
    vsd = O.make_weights(O.decoder_param_shapes(O.SD15_VAE), seed=1)
    return O, usd, vsd


def _cpu_unet_pair(O, usd, x, t, ctx):
    with torch.no_grad():
        return O.unet_forward(usd, O.SD15_UNET, x, t, ctx)


def cpu_baseline_sample(steps: int):
    """Oracle (fp32 torch port of the reference path) on the host cores: one CFG-pair UNet forward (B=1) and one VAE
    decode are timed and extrapolated to the workload's per-image cost."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, usd, vsd = _cpu_models()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 64, 64, generator=g)
    t = torch.tensor([500.0, 500.0])
    ctx = torch.randn(2, 77, 768, generator=g)
    _cpu_unet_pair(O, usd, x, t, ctx)
    t0 = time.perf_counter()
    _cpu_unet_pair(O, usd, x, t, ctx)
    t_unet = time.perf_counter() - t0
    z = torch.randn(1, 4, 64, 64, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.vae_decode(vsd, O.SD15_VAE, z)
    t_vae = time.perf_counter() - t0
    per_image = steps * t_unet + t_vae
    return {"value": round(1.0 / per_image, 6), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"1 CFG-pair UNet forward (B=1, {t_unet:.2f} s) + 1 VAE decode ({t_vae:.2f} s), fp32 torch on "
                      f"{cores} host threads, extrapolated to {steps} steps + decode per image"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    b, sampler, steps = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, usd, vsd = _cpu_models()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 64, 64, generator=g)
    t = torch.tensor([500.0, 500.0])
    ctx = torch.randn(2, 77, 768, generator=g)
    for _ in range(min(args.warmup, 1)):
        _cpu_unet_pair(O, usd, x, t, ctx)
    k = min(args.steps, 5)
    t0 = time.perf_counter()
    for _ in range(k):
        _cpu_unet_pair(O, usd, x, t, ctx)
    t_unet = (time.perf_counter() - t0) / k
    z = torch.randn(1, 4, 64, 64, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.vae_decode(vsd, O.SD15_VAE, z)
    t_vae = time.perf_counter() - t0
    per_image = steps * t_unet + t_vae
    value = 1.0 / per_image
    sample = (f"each step = 1 CFG-pair UNet forward at B=1 ({t_unet:.2f} s, mean of {k}) ; + 1 VAE decode ({t_vae:.2f} s); "
              f"images/s extrapolated to {steps} sampler steps + decode per image; oracle port of the reference path "
              f"(the reference tree is not present on the GPU box), fp32, {cores} host threads")
    line = {"impl": "reference", "metric": "SD1.5 512x512 images/sec (UNet + sampler + VAE decode)",
            "value": round(value, 6), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(per_image * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SD1.5 txt2img 512x512, {steps}-step {sampler}, CFG {CFG_SCALE}, random-init weights, "
                                   f"+ AutoencoderKL decode ({args.workload}); CPU, bounded sample at batch 1"},
            "cpu_baseline": {"value": round(value, 6), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 6), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ddim50_b8", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the cremage_b200 path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
