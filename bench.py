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
    # BASELINE.json configs[4]: SDXL base 1024x1024, 30-step DPM++ 2M Karras (sgm DPMPP2MSampler + EDMDiscretization),
    # VanillaCFG, dual text-encoder context [77, 2048] + pooled/size vector [2816]; batch 8 over 8 GPUs = 1 / GPU
    "sdxl30_b1": (1, "sdxl", 30),
    "sdxl30_b4": (4, "sdxl", 30),
}
CFG_SCALE = 7.5
# dram__bytes_read.sum + dram__bytes_write.sum summed over the 222 igemm launches of one SD1.5 UNet forward at batch 16
# (the same launches `roofline.achieved` aggregates): ncu --metrics capture profiles/r1_igemm_dram_traffic_unet_b16_v2.csv
# (9.56 GB read + 1.56 GB written; weights alone are 1.72 GB, every activation is written and read once more by its
# consumer, the rest is A-tile re-reads that miss the 126 MB L2)
IGEMM_DRAM_TRAFFIC_NOTE = 11.12e9
# algorithmic work per image (BASELINE.md section 3, 2*MACs of the reference graph)
GF_UNET_PER_SAMPLE_FWD = 803.27
GF_VAE_PER_IMAGE = 2514.5

SD15_UNET = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                 num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                 transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)
SDXL_UNET = dict(adm_in_channels=2816, num_classes="sequential", use_checkpoint=True, in_channels=4, out_channels=4,
                 model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4],
                 num_head_channels=64, use_linear_in_transformer=True, transformer_depth=[1, 2, 10], context_dim=2048,
                 spatial_transformer_attn_type="softmax-xformers")   # sdxl/configs/inference/sd_xl_base.yaml:17-33
SDXL_CFG_SCALE = 5.0
GF_SDXL_UNET_PER_CFG_PAIR = 13522.5
GF_SDXL_VAE_PER_IMAGE = 10500.0
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


def build_sdxl_pipeline():
    """DiffusionEngine of sd_xl_base.yaml (UNet + DiscreteDenoiser/EpsScaling + the same AutoencoderKL decoder graph)."""
    from cremage_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    from cremage_b200.sgm.models.diffusion import DiffusionEngine
    from cremage_b200.sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from cremage_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device("meta"):
        unet = UNetModel(**SDXL_UNET)
        vae = AutoencoderKLInferenceWrapper(**SD15_VAE)
    unet = unet.to_empty(device="cuda")
    vae = vae.to_empty(device="cuda")
    init_random_(unet, 0)
    init_random_(vae, 1)
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"})
    return DiffusionEngine(unet, den, vae, scale_factor=0.13025, disable_first_stage_autocast=True).cuda().eval()


def make_sdxl_runner(eng, workload):
    from cremage_b200.sgm.modules.diffusionmodules.sampling import DPMPP2MSampler
    b, _, steps = WORKLOADS[workload]
    smp = DPMPP2MSampler(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                                                "params": {"sigma_min": 0.0292, "sigma_max": 14.6146, "rho": 3.0}},
                         num_steps=steps, guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                                         "params": {"scale": SDXL_CFG_SCALE}})
    denoiser = lambda inp, sigma, c: eng.denoiser(eng.model, inp, sigma, c)

    def run(inp):
        smp.guider._cat.clear()   # a new batch is a new prompt: its context is projected to K/V again (once per batch)
        cond = {"crossattn": inp["cond"], "vector": inp["cond_vec"]}
        uc = {"crossattn": inp["uncond"], "vector": inp["uncond_vec"]}
        z = smp(denoiser, inp["x_T"], cond=cond, uc=uc)
        return eng.decode_first_stage(z, to_uint8=True)
    return run


def make_runner(pipe, workload):
    """Returns run(inputs: dict of device tensors) -> uint8 images [b, H, W, 3] through the repo's public
    (reference-mirroring) API."""
    if WORKLOADS[workload][1] == "sdxl":
        return make_sdxl_runner(pipe, workload)
    f = _make_sd15_runner(pipe, workload)
    return lambda inp: f(inp["x_T"], inp["cond"], inp["uncond"], inp.get("noise"))


def _make_sd15_runner(ldm, workload):
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


def unet_probe_inputs(workload):
    """One CFG-doubled UNet call of the workload (public forward signature) for the step-latency / breakdown probes."""
    b, sampler, _ = WORKLOADS[workload]
    if sampler == "sdxl":
        return (torch.randn(2 * b, 4, 128, 128, device="cuda"), torch.full((2 * b,), 500.0, device="cuda")), \
               dict(context=torch.randn(2 * b, 77, 2048, device="cuda"), y=torch.randn(2 * b, 2816, device="cuda"))
    return (torch.randn(2 * b, 4, 64, 64, device="cuda"), torch.full((2 * b,), 500.0, device="cuda")), \
           dict(context=torch.randn(2 * b, 77, 768, device="cuda"))


def kernel_breakdown(unet, vae, workload):
    """One eager UNet forward (CFG batch 2b) + one VAE decode with CUDA events around every launch of this library:
    per-kernel time shares and achieved rates for the roofline section."""
    from cremage_b200 import ops
    b, sampler, _ = WORKLOADS[workload]
    args, kw = unet_probe_inputs(workload)
    lat = 128 if sampler in ("sdxl", "hires") else 64
    z = torch.randn(b, 4, lat, lat, device="cuda")
    saved = unet.use_cuda_graph
    unet.use_cuda_graph = False
    out = {}
    try:
        with torch.no_grad():
            for _ in range(2):
                unet(*args, **kw)
            with ops.LaunchProfile() as prof:
                for _ in range(3):
                    unet(*args, **kw)
            out["unet_fwd"] = {k: {kk: vv / 3 for kk, vv in v.items()} for k, v in prof.summary().items()}
            vae.decode(z)
            with ops.LaunchProfile() as prof:
                vae.decode(z)
            out["vae_decode"] = prof.summary()
    finally:
        unet.use_cuda_graph = saved
    return out


def make_host_inputs(workload, rank):
    """Pinned host buffers of one batch (what a caller hands over): latents, context (and SDXL vector conditioning)."""
    b, sampler, steps = WORKLOADS[workload]
    gen = torch.Generator().manual_seed(1000 + rank)
    r = lambda *shape: torch.randn(*shape, generator=gen).pin_memory()
    if sampler == "sdxl":
        return {"cond": r(b, 77, 2048), "uncond": r(b, 77, 2048), "cond_vec": r(b, 2816), "uncond_vec": r(b, 2816),
                "x_T": r(b, 4, 128, 128)}, {}
    host = {"cond": r(b, 77, 768), "uncond": r(b, 77, 768), "x_T": r(b, 4, 64, 64)}
    resident = {}
    if sampler == "euler_a":   # injected ancestral noise (seeded, resident: the reference draws it on the device)
        resident["noise"] = torch.randn(steps, b, 4, 64, 64, generator=gen).cuda()
    if sampler == "hires":
        resident["noise"] = torch.randn(b, 4, 128, 128, generator=gen).cuda()
    return host, resident


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
    sdxl = sampler == "sdxl"
    if sdxl:
        pipe = build_sdxl_pipeline()
        unet, vae = pipe.model.diffusion_model, pipe.first_stage_model
    else:
        pipe = build_pipeline()
        unet, vae = pipe.model.diffusion_model, pipe.first_stage_model
    run = make_runner(pipe, args.workload)

    host, resident = make_host_inputs(args.workload, rank)
    out_px = 1024 if sampler in ("hires", "sdxl") else 512
    h_img = torch.empty(b, out_px, out_px, 3, dtype=torch.uint8).pin_memory()
    dev = {k: v.cuda() for k, v in host.items()}
    dev.update(resident)
    gathered = torch.empty(world * b, out_px, out_px, 3, dtype=torch.uint8, device="cuda") if world > 1 else None

    def step_resident():
        img = run(dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        return img

    def step_e2e():
        inp = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        inp.update(resident)
        img = run(inp)
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
        pa, pkw = unet_probe_inputs(args.workload)
        unet(*pa, **pkw)
        unet_ms = timed(lambda: unet(*pa, **pkw), 10) / 10
        brk = kernel_breakdown(unet, vae, args.workload) if rank == 0 else None
        also = {}
        if rank == 0 and world == 1 and args.workload == "ddim50_b8" and not args.no_extra:
            # BASELINE.json's metric is quoted as "20-step images/sec": the 20-step Euler-ancestral sampler of
            # configs[0] (k-diffusion path) at the same batch 8 and at the reference's own batch 1, same pipeline
            for wl in ("euler20_b8", "euler20_b1"):
                h2, r2 = make_host_inputs(wl, rank)
                d2 = {k: v.cuda() for k, v in h2.items()}
                d2.update(r2)
                run2 = make_runner(pipe, wl)
                for _ in range(3):
                    run2(d2)
                k2 = max(args.steps, 3)
                ms2 = timed(lambda: run2(d2), k2)
                also[wl] = {"images_per_s": round(WORKLOADS[wl][0] * k2 / (ms2 / 1e3), 4),
                            "ms_per_batch": round(ms2 / k2, 3), "batch": WORKLOADS[wl][0], "steps_timed": k2}

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
    if sdxl:
        gf_per_image = steps * GF_SDXL_UNET_PER_CFG_PAIR + GF_SDXL_VAE_PER_IMAGE
    cpu = cpu_baseline_sample(steps) if world == 1 and not args.no_cpu_baseline and not sdxl else None
    model = "SDXL base txt2img 1024x1024" if sdxl else "SD1.5 txt2img 512x512"
    scale = SDXL_CFG_SCALE if sdxl else CFG_SCALE
    wgt_gb = "5.1 GB" if sdxl else "1.8 GB"
    line = {
        "metric": ("SDXL 1024x1024 images/sec (UNet + sampler + VAE decode)" if sdxl else
                   "SD1.5 512x512 images/sec (UNet + sampler + VAE decode)"), "value": round(value, 4),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": _dtype_name(), "data": "synthetic",
        "config": {"workload": f"{model}, batch {b}/GPU, {steps}-step {sampler}, CFG {scale}, "
                               f"random-init weights, + AutoencoderKL decode to uint8 ({args.workload})",
                   "global_batch": b * world, "parallelism": f"dp{world} (batch sharded, NCCL all_gather of uint8 images)",
                   "l2": f"no flush: weights ({wgt_gb}) + activations per step exceed the 126 MB L2",
                   "algorithmic_gflop_per_image": round(gf_per_image, 1)},
        "e2e": {"value": round(e2e_value, 4), "unit": "images/s",
                "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
                "d2h_bytes_per_step": int(h_img.numel())},
        "gpu_launches": int(launches),
        "unet_step_ms": round(unet_ms, 3),
        "model_tflops": round(value * gf_per_image / 1e3, 1),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": pk["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops_sustained"], 4),
                     "traffic": IGEMM_DRAM_TRAFFIC_NOTE if (not sdxl and b == 8 and sampler != "hires") else None,
                     "kernel": "igemm_kernel (tcgen05 implicit GEMM: conv3x3 / conv1x1 / linear), all launches of one "
                               f"UNet forward at batch {2 * b}; share of UNet kernel time {ig['ms'] / total_ms:.3f}",
                     "peak_source": pk["source"] + " bf16_tflops_sustained"},
        "kernels": kernels,
        "clocks": clk,
        "cpu_baseline": cpu,
    }
    if also:
        line["metric_20step"] = also
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
    if sampler == "sdxl":
        return run_reference_sdxl(args, world, steps, cores)
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


def run_reference_sdxl(args, world, steps, cores):
    """CPU arm of the SDXL workload: oracle port of the sgm UNet (one CFG pair at 128x128 latents) + one 1024x1024
    VAE decode, extrapolated to the 30-step image."""
    from oracle import sd_oracle as O
    from oracle import sgm_oracle as S
    usd = O.make_weights(S.sgm_unet_param_shapes(S.SDXL_UNET), seed=0)
    vsd = O.make_weights(O.decoder_param_shapes(O.SD15_VAE), seed=1)
    g = torch.Generator().manual_seed(5)
    x, t = torch.randn(2, 4, 128, 128, generator=g), torch.tensor([500, 500])
    ctx, y = torch.randn(2, 77, 2048, generator=g), torch.randn(2, 2816, generator=g)
    k = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(k):
            S.sgm_unet_forward(usd, S.SDXL_UNET, x, t, ctx, y)
    t_unet = (time.perf_counter() - t0) / k
    z = torch.randn(1, 4, 128, 128, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.vae_decode(vsd, O.SD15_VAE, z)
    t_vae = time.perf_counter() - t0
    per_image = steps * t_unet + t_vae
    value = 1.0 / per_image
    sample = (f"each step = 1 CFG-pair sgm UNet forward at B=1, 128x128 latents ({t_unet:.1f} s, mean of {k}); + 1 VAE decode "
              f"to 1024x1024 ({t_vae:.1f} s); extrapolated to {steps} steps + decode per image; oracle port, fp32, {cores} host threads")
    line = {"impl": "reference", "metric": "SDXL 1024x1024 images/sec (UNet + sampler + VAE decode)",
            "value": round(value, 6), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(per_image * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SDXL base txt2img 1024x1024, {steps}-step DPM++ 2M, CFG {SDXL_CFG_SCALE}, random-init "
                                   f"weights, + decode ({args.workload}); CPU, bounded sample at batch 1"},
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
    ap.add_argument("--no-extra", action="store_true", help="skip the additional 20-step Euler-a measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the cremage_b200 path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
