"""ORACLE support -- generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
through oracle/ref_shim.py) on seeded synthetic inputs and weights.  Run in the authoring container:

    python oracle/make_golden.py [--full]

The committed vectors are what pins oracle/sd_oracle.py (tests/test_oracle_pin.py) and what the CUDA path is compared
with on the GPU box (tests/test_gpu_model_parity.py), where the reference tree does not exist.
`--full` also regenerates the SD1.5-sized fixtures (about two minutes of CPU).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import sd_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator(device="cpu").manual_seed(seed))


def schedule_tensors():
    ref_shim.install()
    from ldm.modules.diffusionmodules.util import make_beta_schedule

    betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    # ldm/models/diffusion/ddpm.py:134-160
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return f32(betas), f32(ac), f32(ac_prev)


def gen_schedules():
    ref_shim.install()
    from k_diffusion import external, sampling
    from ldm.modules.diffusionmodules.util import (make_ddim_sampling_parameters, make_ddim_timesteps,
                                                   timestep_embedding)

    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(None, betas, ac, ac_prev)
    den = external.CompVisDenoiser(ldm, quantize=False)
    out = {"betas": betas.numpy(), "alphas_cumprod": ac.numpy(), "sigmas_table": den.sigmas.numpy(),
           "log_sigmas_table": den.log_sigmas.numpy()}
    for n in (20, 30, 50):
        s = den.get_sigmas(n)
        out[f"get_sigmas_{n}"] = s.numpy()
        out[f"sigma_to_t_{n}"] = den.sigma_to_t(s[:-1]).numpy()
        steps = [sampling.get_ancestral_step(s[i], s[i + 1]) for i in range(n)]
        out[f"ancestral_down_{n}"] = np.array([float(a) for a, _ in steps], dtype=np.float32)
        out[f"ancestral_up_{n}"] = np.array([float(b) for _, b in steps], dtype=np.float32)
    out["karras_30"] = sampling.get_sigmas_karras(30, 0.0316386, 14.5521805).numpy()
    out["karras_20"] = sampling.get_sigmas_karras(20, 0.0316386, 14.5521805).numpy()
    for S in (5, 20, 50):
        ts = make_ddim_timesteps("uniform", S, 1000, verbose=False)
        sig, a, ap = make_ddim_sampling_parameters(ac.cpu(), ts, 0.0, verbose=False)
        out[f"ddim_timesteps_{S}"] = ts
        out[f"ddim_alphas_{S}"] = np.asarray(a)
        out[f"ddim_alphas_prev_{S}"] = np.asarray(ap)
        out[f"ddim_sqrt_one_minus_alphas_{S}"] = np.asarray(np.sqrt(1.0 - a))
    t = torch.tensor([999.0, 946.4210205078125, 52.57894134521484, 0.0, 981.0, 1.0])
    out["temb_t"] = t.numpy()
    out["temb_320"] = timestep_embedding(t, 320).numpy()
    np.savez_compressed(os.path.join(GOLD, "schedules.npz"), **out)
    print("schedules.npz written")


def _sampler_goldens(unet, cfg_scale, cond, uncond, x_T, noise, steps_euler, steps_2m, steps_ddim):
    """Run the reference's own wrapper chain and samplers (sampling.py:147,593; ddim.py:78) with injected noise."""
    from k_diffusion import external, sampling
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion

    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    den = external.CompVisDenoiser(ldm, quantize=False)
    wrapper = LDMWrapperForKDiffusion(den, cond, uncond, cfg_scale)
    out = {}
    with torch.no_grad():
        # Euler ancestral, sigmas as EulerAncestralSampler.compute_sigmas (k_diffusion_samplers.py:313-314)
        sig = den.get_sigmas(steps_euler)
        trace = []
        it = iter(range(steps_euler))
        x = sampling.sample_euler_ancestral(wrapper, x_T * sig[0], sig, disable=True,
                                            noise_sampler=lambda s, sn: noise[next(it)],
                                            callback=lambda d: trace.append(d["x"].clone()))
        out["euler_a_sigmas"] = sig.numpy()
        out["euler_a_trace"] = torch.stack(trace[1:] + [x]).numpy()  # latent after each step
        # DPM++ 2M Karras (k_diffusion_samplers.py:386-387)
        sig = sampling.get_sigmas_karras(steps_2m, 0.0316386, 14.5521805)
        trace = []
        x = sampling.sample_dpmpp_2m(wrapper, x_T * sig[0], sig, disable=True,
                                     callback=lambda d: trace.append(d["x"].clone()))
        out["dpmpp2m_sigmas"] = sig.numpy()
        out["dpmpp2m_trace"] = torch.stack(trace[1:] + [x]).numpy()
        # DDIM eta = 0 (ddim.py:78-190)
        DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)  # no .to("cuda") here
        smp = DDIMSampler(ldm)
        inter = []
        x, _ = smp.sample(S=steps_ddim, batch_size=x_T.shape[0], shape=list(x_T.shape[1:]), conditioning=cond,
                          eta=0.0, x_T=x_T, unconditional_guidance_scale=cfg_scale,
                          unconditional_conditioning=uncond, verbose=False,
                          img_callback=lambda pred_x0, i: inter.append(pred_x0.clone()))
        out["ddim_final"] = x.numpy()
        out["ddim_pred_x0_trace"] = torch.stack(inter).numpy()
        # hires-fix second pass, latent upscaler (sd/image_generator.py:969-999 -> img2img_sampling DDIM branch)
        import torch.nn.functional as F
        strength = 0.6
        smp.make_schedule(ddim_num_steps=steps_ddim, ddim_eta=0.0, verbose=False)
        up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        t_enc = int(strength * steps_ddim)
        hnoise = randn(tuple(up.shape), 77)
        z_enc = smp.stochastic_encode(up, torch.tensor([t_enc] * x.shape[0]), noise=hnoise)
        xh = smp.decode(z_enc, cond, t_enc, unconditional_guidance_scale=cfg_scale, unconditional_conditioning=uncond)
        out["hires_strength"] = np.float32(strength)
        out["hires_noise"] = hnoise.numpy()
        out["hires_upsampled"] = up.numpy()
        out["hires_final"] = xh.numpy()
    return out


def gen_tiny():
    ref_shim.install()
    cfg = O.TINY_UNET
    sd = O.make_weights(O.unet_param_shapes(cfg), seed=100)
    unet = ref_shim.reference_unet(cfg, sd)
    x = randn((2, 4, 16, 16), 1)
    t = torch.tensor([946.4210205078125, 13.0])
    ctx = randn((2, 7, cfg.context_dim), 2)
    with torch.no_grad():
        y = unet(x, t, context=ctx)
    out = {"x": x.numpy(), "t": t.numpy(), "context": ctx.numpy(), "out": y.numpy(),
           "weights_checksum": np.float64(O.weights_checksum(sd))}
    print("tiny unet out: mean %.4f std %.4f" % (y.mean().item(), y.std().item()))
    np.savez_compressed(os.path.join(GOLD, "tiny_unet.npz"), **out)

    vcfg = O.TINY_VAE
    vsd = O.make_weights(O.decoder_param_shapes(vcfg), seed=200)
    dec, pq = ref_shim.reference_decoder(vcfg, vsd)
    z = randn((2, 4, 8, 8), 3)
    with torch.no_grad():
        img = dec(pq(z))
    print("tiny vae out: mean %.4f std %.4f" % (img.mean().item(), img.std().item()))
    np.savez_compressed(os.path.join(GOLD, "tiny_vae.npz"), z=z.numpy(), out=img.numpy(),
                        weights_checksum=np.float64(O.weights_checksum(vsd)))

    b = 2
    cond = randn((b, 7, cfg.context_dim), 4)
    uncond = randn((b, 7, cfg.context_dim), 5)
    x_T = randn((b, 4, 16, 16), 6)
    noise = randn((6, b, 4, 16, 16), 7)
    s = _sampler_goldens(unet, 7.5, cond, uncond, x_T, noise, steps_euler=5, steps_2m=6, steps_ddim=5)
    np.savez_compressed(os.path.join(GOLD, "tiny_sampling.npz"), cond=cond.numpy(), uncond=uncond.numpy(),
                        x_T=x_T.numpy(), noise=noise.numpy(), cfg_scale=np.float32(7.5), **s)
    print("tiny fixtures written")


def gen_full():
    """SD1.5-sized fixtures: one UNet forward (CFG pair), a 3-step Euler-a trajectory and a 32x32-latent VAE decode.
    Inputs are re-derivable from seeds, so only outputs are stored."""
    ref_shim.install()
    torch.set_num_threads(os.cpu_count())
    cfg = O.SD15_UNET
    sd = O.make_weights(O.unet_param_shapes(cfg), seed=0)
    unet = ref_shim.reference_unet(cfg, sd)
    x = randn((2, 4, 64, 64), 11)
    t = torch.tensor([946.4210205078125, 261.0])
    ctx = randn((2, 77, 768), 12)
    with torch.no_grad():
        y = unet(x, t, context=ctx)
    print("sd15 unet out: mean %.4f std %.4f absmax %.3f" % (y.mean().item(), y.std().item(), y.abs().max().item()))
    np.savez_compressed(os.path.join(GOLD, "sd15_unet.npz"), out=y.numpy(), t=t.numpy(),
                        seeds=np.array([0, 11, 12]), weights_checksum=np.float64(O.weights_checksum(sd)))

    from k_diffusion import external, sampling
    from ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    den = external.CompVisDenoiser(ldm, quantize=False)
    cond, uncond = randn((1, 77, 768), 13), randn((1, 77, 768), 14)
    x_T = randn((1, 4, 64, 64), 15)
    noise = randn((3, 1, 4, 64, 64), 16)
    wrapper = LDMWrapperForKDiffusion(den, cond, uncond, 7.5)
    sig = den.get_sigmas(20)[[0, 7, 14, 20]]  # a 3-step slice of the 20-step schedule keeps CPU time bounded
    trace = []
    it = iter(range(3))
    with torch.no_grad():
        xf = sampling.sample_euler_ancestral(wrapper, x_T * sig[0], sig, disable=True,
                                             noise_sampler=lambda s, sn: noise[next(it)],
                                             callback=lambda d: trace.append(d["x"].clone()))
    np.savez_compressed(os.path.join(GOLD, "sd15_euler3.npz"), sigmas=sig.numpy(),
                        trace=torch.stack(trace[1:] + [xf]).numpy(), seeds=np.array([0, 13, 14, 15, 16]))

    vcfg = O.SD15_VAE
    vsd = O.make_weights(O.decoder_param_shapes(vcfg), seed=1)
    dec, pq = ref_shim.reference_decoder(vcfg, vsd)
    z = randn((1, 4, 32, 32), 17)
    with torch.no_grad():
        img = dec(pq(z))
    print("sd15 vae out: mean %.4f std %.4f absmax %.3f" % (img.mean().item(), img.std().item(), img.abs().max().item()))
    np.savez_compressed(os.path.join(GOLD, "sd15_vae32.npz"), out=img.numpy().astype(np.float16),
                        seeds=np.array([1, 17]), weights_checksum=np.float64(O.weights_checksum(vsd)))
    print("full-size fixtures written")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    gen_schedules()
    gen_tiny()
    if args.full:
        gen_full()
