#!/usr/bin/env python
"""Benchmark of the SD1.5 denoising hot path (UNet + sampler + VAE decode) on B200, driver contract.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--scaling weak|strong]

One "step" = one pass of the hot path over one batch.  Default workload `euler20_b8` = BASELINE.json's metric as named
("SD1.5 512x512 20-step images/sec"): the 20-step Euler-ancestral sampler of configs[0] (k_diffusion path, CFG 7.5) at
configs[1]'s batch 8 per GPU and 16-bit precision, random-init weights, synthetic context / latents / injected noise,
followed by the AutoencoderKL decode to uint8 images.  `value` = images/s over all ranks with inputs resident in HBM;
`e2e` = the same through host buffers (pinned H2D of x_T / context / uncond, D2H of the uint8 images) inside the timed
region.  configs[1] verbatim (50-step DDIM, batch 8) is measured in the same run at every N (`configs1_ddim50_b8`).
Multi-GPU: one rank per GPU, the global batch (one global seed, `cremage_b200.dist.full_batch_noise`) is sharded --
8 images per rank (weak scaling; `strong_scaling` adds the fixed global batch 8 split 8/N) -- the only collective is
the NCCL all_gather of the decoded uint8 images; time = max over ranks.  Before timing, rank 0 recomputes every other
rank's shard itself and checks the gathered N-GPU images bit-for-bit (`multi_gpu_parity`).
At N = 1 the line also carries: `torch_gpu_baseline` (the reference's path as plain PyTorch on this GPU: fp16 weights
+ autocast + cuDNN / cuBLAS / SDPA -- BASELINE.md section 4's "real bar"), `vae16` (configs[2]), `euler20_b1`.
At N = 8 it adds configs[3] / configs[4] (`hires20_b4`, `sdxl30_b1`), both "sharded over 8 B200"
(CREMAGE_BENCH_CONFIGS=none skips them, =hires20_b4,sdxl30_b1,vae16 forces them at any N).

`--impl reference` times the reference's CPU path (the oracle port of it: /root/reference does not exist on the GPU
box) on the host cores: each step is one bounded sample of the same workload (one CFG-pair UNet forward at batch 1 +
its share of the VAE decode).
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
    # BASELINE.json configs[2]: AutoencoderKL decode only, 64x64x4 latent -> 512x512 RGB, batch 16
    "vae16": (16, "vae", 0),
}
CFG_SCALE = 7.5


def igemm_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum summed over the igemm launches of one SD1.5 UNet forward at batch 16
    (the same launches `roofline.achieved` aggregates), read from the NEWEST ncu --metrics capture committed under
    profiles/ (`*traffic*.csv`).  Returns (bytes, file name) or (None, None)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*traffic*.csv")), key=os.path.getmtime)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reversed(files):
        total, hdr = 0.0, None
        with open(path, errors="ignore") as f:
            for r in csv.reader(f):
                if "Kernel Name" in r:
                    hdr = r
                    continue
                if hdr is None or len(r) != len(hdr):
                    continue
                d = dict(zip(hdr, r))
                if "igemm_kernel" in d["Kernel Name"] and d["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(d["Metric Value"].replace(",", "")) * unit.get(d["Metric Unit"], 1.0)
        if total > 0:
            return total, os.path.relpath(path, ROOT)
    return None, None


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


def make_sdxl_runner(eng, workload, b=None):
    from cremage_b200.sgm.modules.diffusionmodules.sampling import DPMPP2MSampler
    _, _, steps = WORKLOADS[workload]
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


def make_runner(pipe, workload, b=None):
    """Returns run(inputs: dict of device tensors) -> uint8 images [b, H, W, 3] through the repo's public
    (reference-mirroring) API.  `b`: images per rank (default: the workload's)."""
    b = WORKLOADS[workload][0] if b is None else b
    if WORKLOADS[workload][1] == "sdxl":
        return make_sdxl_runner(pipe, workload, b)
    if WORKLOADS[workload][1] == "vae":
        return lambda inp: pipe.decode_first_stage(inp["z"], to_uint8=True)
    f = _make_sd15_runner(pipe, workload, b)
    return lambda inp: f(inp["x_T"], inp["cond"], inp["uncond"], inp.get("noise"))


def _make_sd15_runner(ldm, workload, b):
    from cremage_b200.k_diffusion.external import CompVisDenoiser
    from cremage_b200.k_diffusion.sampling import get_sigmas_karras, sample_dpmpp_2m, sample_euler_ancestral
    from cremage_b200.ldm.models.diffusion.ddim import DDIMSampler
    from cremage_b200.ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    _, sampler, steps = WORKLOADS[workload]
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


def unet_probe_inputs(workload, b=None):
    """One CFG-doubled UNet call of the workload (public forward signature) for the step-latency / breakdown probes."""
    _, sampler, _ = WORKLOADS[workload]
    b = WORKLOADS[workload][0] if b is None else b
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
    if sampler == "vae":
        b = 8          # the UNet share of this line is informational: probe it at the metric's batch
    args, kw = unet_probe_inputs(workload, b)
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


GLOBAL_SEED = 20260


def make_host_inputs(workload, rank, world, b):
    """Pinned host buffers of rank `rank`'s shard of ONE global batch (what a caller hands over): every tensor is drawn
    for the full global batch from one global seed on every rank and sliced (cremage_b200.dist.full_batch_noise), so the
    N-GPU job generates exactly the images a 1-GPU job with the same seed and the same per-GPU batches would."""
    from cremage_b200.dist import full_batch_noise
    _, sampler, steps = WORKLOADS[workload]
    gb = b * world

    def r(k, *shape):
        return full_batch_noise((gb, *shape), GLOBAL_SEED + k, rank, world).contiguous().pin_memory()
    if sampler == "vae":
        return {"z": r(0, 4, 64, 64)}, {}
    if sampler == "sdxl":
        return {"cond": r(1, 77, 2048), "uncond": r(2, 77, 2048), "cond_vec": r(3, 2816), "uncond_vec": r(4, 2816),
                "x_T": r(0, 4, 128, 128)}, {}
    host = {"cond": r(1, 77, 768), "uncond": r(2, 77, 768), "x_T": r(0, 4, 64, 64)}
    resident = {}
    if sampler == "euler_a":   # injected ancestral noise (seeded, resident: the reference draws it on the device)
        resident["noise"] = r(5, steps, 4, 64, 64).transpose(0, 1).contiguous().cuda()
    if sampler == "hires":
        resident["noise"] = r(5, 4, 128, 128).cuda()
    return host, resident


def describe(workload, b, world, scaling):
    """`metric`, `config` and the algorithmic work per image of a workload -- shared by our arm and the reference arm so
    the two lines carry the same config."""
    _, sampler, steps = WORKLOADS[workload]
    sdxl = sampler == "sdxl"
    gf = steps * 2 * GF_UNET_PER_SAMPLE_FWD + GF_VAE_PER_IMAGE
    if sampler == "hires":  # + 10 CFG steps at 128x128 latents (9348 GF each) and a 1024x1024 decode (~10.5 TF), BASELINE.md section 3
        gf = steps * 2 * GF_UNET_PER_SAMPLE_FWD + int(0.5 * steps) * 9348.0 + 10500.0
    if sdxl:
        gf = steps * GF_SDXL_UNET_PER_CFG_PAIR + GF_SDXL_VAE_PER_IMAGE
    if sampler == "vae":
        gf = GF_VAE_PER_IMAGE
    model = "SDXL base txt2img 1024x1024" if sdxl else "SD1.5 txt2img 512x512"
    scale = SDXL_CFG_SCALE if sdxl else CFG_SCALE
    if sampler == "vae":
        what = f"SD1.5 AutoencoderKL decode only, 64x64x4 latent -> 512x512 uint8, batch {b}/GPU ({workload})"
        metric = "SD1.5 AutoencoderKL decode 512x512 images/sec"
    else:
        what = (f"{model}, batch {b}/GPU, {steps}-step {sampler}, CFG {scale}, random-init weights, "
                f"+ AutoencoderKL decode to uint8 ({workload})")
        metric = ("SDXL 1024x1024 images/sec (UNet + sampler + VAE decode)" if sdxl else
                  f"SD1.5 512x512 {steps}-step images/sec (UNet + sampler + VAE decode)")
    config = {"workload": what, "global_batch": b * world,
              "parallelism": f"dp{world} (one global seed, batch rows sharded, NCCL all_gather of uint8 images)",
              "l2": f"no flush: weights ({'5.1 GB' if sdxl else '1.8 GB'}) + activations per step exceed the 126 MB L2",
              "algorithmic_gflop_per_image": round(gf, 1)}
    return metric, config, gf


class Timer:
    def __init__(self, world):
        self.world = world

    def __call__(self, fn, k):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())


def measure(pipe, workload, b, rank, world, k, warmup, timed, e2e=True, parity=False):
    """Times `k` batches of `workload` (b images per rank) resident and end to end; optional multi-GPU bit-parity check."""
    import torch.distributed as dist
    from cremage_b200 import engine
    _, sampler, _ = WORKLOADS[workload]
    run = make_runner(pipe, workload, b)
    host, resident = make_host_inputs(workload, rank, world, b)
    out_px = 1024 if sampler in ("hires", "sdxl") else 512
    h_img = torch.empty(b, out_px, out_px, 3, dtype=torch.uint8).pin_memory()
    dev = {kk: v.cuda() for kk, v in host.items()}
    dev.update(resident)
    gathered = torch.empty(world * b, out_px, out_px, 3, dtype=torch.uint8, device="cuda") if world > 1 else None

    def step_resident():
        img = run(dev)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        return img

    def step_e2e():
        inp = {kk: v.cuda(non_blocking=True) for kk, v in host.items()}
        inp.update(resident)
        img = run(inp)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        h_img.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return img

    res = {}
    for _ in range(warmup):
        step_resident()
    if parity and world > 1:
        # outside the timed region: rank 0 recomputes every rank's shard of the global batch itself (same per-GPU batch
        # shape, hence the same kernels) and compares with what the ranks produced and all_gather delivered
        torch.cuda.synchronize()
        if rank == 0:
            worst, equal = 0, True
            for r in range(world):
                h_r, res_r = make_host_inputs(workload, r, world, b)
                d_r = {kk: v.cuda() for kk, v in h_r.items()}
                d_r.update(res_r)
                mine = run(d_r)
                theirs = gathered[r * b:(r + 1) * b]
                if not torch.equal(mine, theirs):
                    equal = False
                    worst = max(worst, int((mine.int() - theirs.int()).abs().max().item()))
            res["multi_gpu_parity"] = {"checked": True, "bit_identical": equal, "max_abs_diff_u8": worst, "ranks": world,
                                       "global_batch": b * world, "global_seed": GLOBAL_SEED,
                                       "what": "rank 0 recomputed every rank's shard of the one-seed global batch; "
                                               "compared with the all_gather'ed N-GPU uint8 images"}
        dist.barrier()
    launches0 = engine.total_launches()
    res["ms"] = timed(step_resident, k)
    res["launches"] = engine.total_launches() - launches0
    if e2e:
        res["ms_e2e"] = timed(step_e2e, k)
        res["h2d"] = int(sum(v.numel() * v.element_size() for v in host.values()))
        res["d2h"] = int(h_img.numel())
    return res


def side_line(res, b, world, k):
    out = {"images_per_s": round(b * world * k / (res["ms"] / 1e3), 4), "ms_per_batch": round(res["ms"] / k, 3),
           "batch_per_gpu": b, "global_batch": b * world, "steps_timed": k}
    if "ms_e2e" in res:
        out["e2e_images_per_s"] = round(b * world * k / (res["ms_e2e"] / 1e3), 4)
    return out


def extra_configs(world):
    env = os.environ.get("CREMAGE_BENCH_CONFIGS", "auto").strip().lower()
    if env == "none":
        return []
    if env == "auto":
        return ["hires20_b4", "sdxl30_b1"] if world == 8 else []
    return [w for w in env.split(",") if w in WORKLOADS]


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b, sampler, steps = WORKLOADS[args.workload]
    strong = args.scaling == "strong"
    if strong:
        if 8 % world:
            raise SystemExit("--scaling strong splits a global batch of 8: --gpus must divide 8")
        b = 8 // world
    sdxl = sampler == "sdxl"
    pipe = build_sdxl_pipeline() if sdxl else build_pipeline()
    unet, vae = pipe.model.diffusion_model, pipe.first_stage_model
    timed = Timer(world)
    warm = max(args.warmup, 3)
    side = {}
    with torch.no_grad():
        clocks = ClockSampler(local)
        clocks.start()
        main = measure(pipe, args.workload, b, rank, world, args.steps, warm, timed, parity=True)
        clk = clocks.stop()
        # UNet step latency (one CFG-doubled forward through the public API, graph replay)
        pb = 8 if sampler == "vae" else b
        pa, pkw = unet_probe_inputs(args.workload, pb)
        unet(*pa, **pkw)
        unet_ms = timed(lambda: unet(*pa, **pkw), 10) / 10
        brk = kernel_breakdown(unet, vae, args.workload) if rank == 0 else None
        k2 = max(min(args.steps, 5), 3)
        if not sdxl and not args.no_extra and args.workload == "euler20_b8" and not strong:
            # configs[1] verbatim (50-step DDIM, batch 8 per GPU) at every N, same pipeline
            side["configs1_ddim50_b8"] = side_line(measure(pipe, "ddim50_b8", 8, rank, world, k2, 3, timed), 8, world, k2)
            if world > 1 and 8 % world == 0:
                # strong scaling (SURVEY 8d "report both"): the fixed global batch 8 split 8/N
                bs = 8 // world
                r = measure(pipe, "euler20_b8", bs, rank, world, k2, 3, timed, e2e=False)
                side["strong_scaling"] = dict(side_line(r, bs, world, k2), workload="euler20, global batch 8 split 8/N")
            if world == 1:
                side["euler20_b1"] = side_line(measure(pipe, "euler20_b1", 1, rank, world, k2, 3, timed), 1, world, k2)
                side["vae16"] = side_line(measure(pipe, "vae16", 16, rank, world, k2, 3, timed), 16, world, k2)
        if not args.no_extra and not strong:
            for wl in extra_configs(world):
                if wl == args.workload:
                    continue
                wb, ws, _ = WORKLOADS[wl]
                if (ws == "sdxl") != sdxl:      # the other model family: build it, measure, free it
                    other = build_sdxl_pipeline() if ws == "sdxl" else build_pipeline()
                    r = measure(other, wl, wb, rank, world, 2, 3, timed)
                    del other
                    torch.cuda.empty_cache()
                else:
                    r = measure(pipe, wl, wb, rank, world, 2, 3, timed)
                side[wl] = dict(side_line(r, wb, world, 2), workload=describe(wl, wb, world, "weak")[1]["workload"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms, ms_e2e = main["ms"], main["ms_e2e"]
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
    metric, config, gf_per_image = describe(args.workload, b, world, args.scaling)
    traffic, traffic_src = igemm_dram_traffic() if (not sdxl and pb == 8 and sampler != "hires") else (None, None)
    cpu = cpu_baseline_sample(steps) if world == 1 and not args.no_cpu_baseline and not sdxl and sampler != "vae" else None
    tgb = None
    if world == 1 and not args.no_extra and not sdxl and not strong and sampler in ("euler_a", "ddim", "dpmpp_2m"):
        try:
            tgb = torch_gpu_baseline(b, steps)
        except Exception as e:   # the bar is informational: never lose the bench line to it
            tgb = {"error": f"{type(e).__name__}: {e}"[:300]}
    line = {
        "metric": metric, "value": round(value, 4),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": _dtype_name(), "data": "synthetic",
        "config": config,
        "e2e": {"value": round(e2e_value, 4), "unit": "images/s", "h2d_bytes_per_step": main["h2d"],
                "d2h_bytes_per_step": main["d2h"]},
        "gpu_launches": int(main["launches"]),
        "unet_step_ms": round(unet_ms, 3),
        "model_tflops": round(value * gf_per_image / 1e3, 1),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": pk["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops_sustained"], 4),
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "igemm_kernel (tcgen05 implicit GEMM: conv3x3 / conv1x1 / linear), all launches of one "
                               f"UNet forward at batch {2 * pb}; share of UNet kernel time {ig['ms'] / total_ms:.3f}",
                     "peak_source": pk["source"] + " bf16_tflops_sustained"},
        "kernels": kernels,
        "clocks": clk,
        "cpu_baseline": cpu,
    }
    if "multi_gpu_parity" in main:
        line["multi_gpu_parity"] = main["multi_gpu_parity"]
    if tgb is not None:
        line["torch_gpu_baseline"] = tgb
    line.update(side)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# the reference's path as plain PyTorch on this GPU ("the real bar")
# ----------------------------------------------------------------------------------------------------------------------
def torch_gpu_baseline(b: int, steps: int):
    """The reference's own arithmetic (oracle port of its modules; the reference tree does not travel to this box) the
    way Cremage runs it on a GPU -- `model.half()` + `torch.autocast("cuda")` (sd/image_generator.py:489,748), cuDNN
    convolutions, cuBLAS linears, fused SDPA attention (xformers / SDPA classes) -- on the same workload: one CFG-doubled
    UNet forward at batch 2b and one VAE decode at batch b, CUDA-event timed, eager and under a CUDA graph, NCHW (as the
    reference) and channels_last; images/s = b / (steps * t_unet + t_vae) with the best of the four UNet timings."""
    from oracle import sd_oracle as O
    prev_impl, prev_bench, prev_te = O.ATTENTION_IMPL, torch.backends.cudnn.benchmark, O.timestep_embedding
    O.ATTENTION_IMPL = "sdpa"
    torch.backends.cudnn.benchmark = True
    te_cache = {}

    def cached_te(ts, dim, *a, **k):   # the frequency table is built on the host: not capturable, and constant here
        key = (ts.data_ptr(), dim)
        if key not in te_cache:
            te_cache[key] = prev_te(ts, dim, *a, **k)
        return te_cache[key]
    O.timestep_embedding = cached_te
    g = torch.Generator(device="cuda").manual_seed(3)

    def weights(shapes):
        sd = {}
        for k, shp in shapes.items():
            if k.endswith(".bias"):
                sd[k] = (torch.randn(shp, generator=g, device="cuda") * 0.05).half()
            elif len(shp) == 1:
                sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g, device="cuda")).half()
            else:
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                sd[k] = (torch.randn(shp, generator=g, device="cuda") / math.sqrt(fan_in)).half()
        return sd

    def cl(sd):
        return {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}

    def time_fn(fn, n=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def graphed(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        return gr.replay

    out = {}
    try:
        usd = weights(O.unet_param_shapes(O.SD15_UNET))
        vsd = weights(O.decoder_param_shapes(O.SD15_VAE))
        x = torch.randn(2 * b, 4, 64, 64, device="cuda")
        t = torch.full((2 * b,), 500.0, device="cuda")
        ctx = torch.randn(2 * b, 77, 768, device="cuda")
        z = torch.randn(b, 4, 64, 64, device="cuda")
        timings = {}
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            for layout in ("nchw", "channels_last"):
                sd_u = usd if layout == "nchw" else cl(usd)
                sd_v = vsd if layout == "nchw" else cl(vsd)
                xx = x if layout == "nchw" else x.contiguous(memory_format=torch.channels_last)
                zz = z if layout == "nchw" else z.contiguous(memory_format=torch.channels_last)
                f_u = lambda: O.unet_forward(sd_u, O.SD15_UNET, xx, t, ctx)
                f_v = lambda: O.vae_decode(sd_v, O.SD15_VAE, zz)
                timings[f"unet_{layout}_eager_ms"] = time_fn(f_u)
                timings[f"vae_{layout}_eager_ms"] = time_fn(f_v, 3)
                try:
                    timings[f"unet_{layout}_graph_ms"] = time_fn(graphed(f_u), 10)
                    timings[f"vae_{layout}_graph_ms"] = time_fn(graphed(f_v), 3)
                except Exception as e:
                    timings[f"{layout}_graph_error"] = f"{type(e).__name__}: {e}"[:160]
                    torch.cuda.synchronize()
        t_unet = min(v for k, v in timings.items() if k.startswith("unet_") and k.endswith("_ms"))
        t_vae = min(v for k, v in timings.items() if k.startswith("vae_") and k.endswith("_ms"))
        per_batch = steps * t_unet + t_vae
        out = {"value": round(b / (per_batch / 1e3), 4), "unit": "images/s",
               "unet_step_ms": round(t_unet, 3), "vae_decode_ms": round(t_vae, 3),
               "timings_ms": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in timings.items()},
               "what": f"oracle port of the reference modules on cuda:0, fp16 weights + torch.autocast(fp16), cuDNN "
                       f"(benchmark mode) / cuBLAS / F.scaled_dot_product_attention, torch {torch.__version__}; UNet batch "
                       f"{2 * b} (CFG pair of batch {b}) and VAE decode batch {b} timed with CUDA events; images/s = "
                       f"{b} / ({steps} x best UNet ms + best VAE ms), sampler arithmetic not charged"}
    finally:
        O.ATTENTION_IMPL = prev_impl
        O.timestep_embedding = prev_te
        torch.backends.cudnn.benchmark = prev_bench
        usd = vsd = None
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle = port of the reference's path; the only places bench.py touches oracle/ besides the torch-GPU bar)
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
    """CPU arm.  A "step" here is one bounded sample of the workload: ONE CFG-pair UNet forward at batch 1 (1/steps of an
    image's sampler work) plus the same share of the image's VAE decode (timed once, up front).  `ms_per_step` is that
    sample's measured duration, so steps x ms_per_step is the region that was really timed, and
    value = images/s = 1 / (sampler steps per image x ms_per_step)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    b, sampler, steps = WORKLOADS[args.workload]
    if args.scaling == "strong":
        b = max(8 // max(world, 1), 1)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    metric, config, _ = describe(args.workload, b, world, args.scaling)
    g = torch.Generator().manual_seed(5)
    if sampler == "sdxl":
        from oracle import sd_oracle as O
        from oracle import sgm_oracle as S
        usd = O.make_weights(S.sgm_unet_param_shapes(S.SDXL_UNET), seed=0)
        vsd = O.make_weights(O.decoder_param_shapes(O.SD15_VAE), seed=1)
        x, t = torch.randn(2, 4, 128, 128, generator=g), torch.tensor([500, 500])
        ctx, y = torch.randn(2, 77, 2048, generator=g), torch.randn(2, 2816, generator=g)
        fwd = lambda: S.sgm_unet_forward(usd, S.SDXL_UNET, x, t, ctx, y)
        z = torch.randn(1, 4, 128, 128, generator=g)
        k, w, budget = max(1, args.steps), 0, 150.0
    else:
        O, usd, vsd = _cpu_models()
        x, t = torch.randn(2, 4, 64, 64, generator=g), torch.tensor([500.0, 500.0])
        ctx = torch.randn(2, 77, 768, generator=g)
        fwd = lambda: O.unet_forward(usd, O.SD15_UNET, x, t, ctx)
        z = torch.randn(1, 4, 64, 64, generator=g)
        k, w, budget = max(1, args.steps), max(0, args.warmup), 240.0
    with torch.no_grad():
        t0 = time.perf_counter()
        O.vae_decode(vsd, O.SD15_VAE, z)
        t_vae = time.perf_counter() - t0
        if sampler == "vae":
            per_step, unet_steps, k = t_vae, 1, 1
        else:
            # K timed steps after W warm-ups as asked, cut short only by the wall-clock budget (the line reports what ran)
            t_start, done_w = time.perf_counter(), 0
            for _ in range(w):
                fwd()
                done_w += 1
                if time.perf_counter() - t_start > 0.25 * budget:
                    break
            w = done_w
            t0, done = time.perf_counter(), 0
            for _ in range(k):
                fwd()
                done += 1
                if time.perf_counter() - t0 > budget:
                    break
            k = done
            t_unet = (time.perf_counter() - t0) / k
            unet_steps = steps
            per_step = t_unet + t_vae / steps
    value = 1.0 / (unet_steps * per_step)
    sample = (f"{k} timed steps; each = 1 CFG-pair UNet forward at batch 1 (mean {per_step - t_vae / max(unet_steps, 1):.2f} s) + 1/{unet_steps} "
              f"of one VAE decode ({t_vae:.2f} s, timed once before the loop); images/s = 1 / ({unet_steps} x ms_per_step); oracle "
              f"port of the reference path (the reference tree is not present on the GPU box), fp32, {cores} host threads")
    line = {"impl": "reference", "metric": metric, "value": round(value, 6), "unit": "images/s", "n_gpus": world,
            "steps": k, "warmup": w, "ms_per_step": round(per_step * 1e3, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": round(value, 6), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 6), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="euler20_b8", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU; strong: a global batch of 8 split 8/N")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="only the headline workload (no side measurements)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the cremage_b200 path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
