"""ORACLE support -- goldens at the sizes BASELINE.json's configs name, from the UNMODIFIED reference on CPU (fp32).

    python oracle/make_golden_configs.py [cfg1] [cfg2] [cfg3] [hires] [sdxl] [sdxlvae]      (default: all)

cfg1     configs[0] verbatim: SD1.5 64x64 latent, batch 1, the complete 20-step `sample_euler_ancestral`
         (k_diffusion/sampling.py:147) through CompVisDenoiser + LDMWrapperForKDiffusion, CFG 7.5, injected noise;
         latent after every step + the decoded 512x512 image (Decoder, ldm/modules/diffusionmodules/model.py:542).
cfg2     configs[1] at batch 2: the complete DDIM S=50 eta=0 run (ldm/models/diffusion/ddim.py:78-190), final latent and
         pred_x0 at steps 0 / 25 / 49.
cfg3     configs[2] at batch 2: AutoencoderKL decode of a 64x64 latent.
hires    configs[3]'s defining shape: ONE SD1.5 UNet sample-forward at 128x128 (16 384-token self-attention).
sdxl     configs[4]'s network at full shape: one sd_xl_base UNet forward (CFG pair, 128x128, ctx 77x2048, y 2816)
         through the reference's vendored sgm (modules/sdxl/sgm/modules/diffusionmodules/openaimodel.py:828).
sdxlvae  a25: the SDXL first stage decode runs in fp32 in the reference (sgm/models/diffusion.py:119-137); decoder
         weights scaled (conv_in x 32768) so that activations exceed the fp16 range (|h| > 65504), 32x32 latent, plus a 128x128-latent
         decode with the plain weights.

Inputs are re-derivable from the seeds below, only outputs are stored (fp32 latents; images as fp16).
CPU cost on 8 cores: cfg1 ~3 min, cfg2 ~14 min, cfg3 15 s, hires ~1 min, sdxl ~3 min, sdxlvae ~1 min.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import sd_oracle as O  # noqa: E402
from oracle.make_golden import GOLD, randn, schedule_tensors  # noqa: E402

# seeds (weights, cond, uncond, x_T, noise); the GPU tests re-derive the same tensors
SEEDS_CFG1 = dict(unet=0, vae=1, cond=21, uncond=22, x_T=23, noise=24)
SEEDS_CFG2 = dict(unet=0, cond=31, uncond=32, x_T=33)
SEEDS_CFG3 = dict(vae=1, z=41)
SEEDS_HIRES = dict(unet=0, x=51, ctx=52)
SEEDS_SDXL = dict(unet=300, x=61, ctx=62, y=63)
SEEDS_SDXLVAE = dict(vae=2, z=71, z128=72)


def _sd15_unet():
    cfg = O.SD15_UNET
    sd = O.make_weights(O.unet_param_shapes(cfg), seed=0)
    return ref_shim.reference_unet(cfg, sd), sd


def gen_cfg1():
    from k_diffusion import external, sampling
    from ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    s = SEEDS_CFG1
    unet, sd = _sd15_unet()
    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    den = external.CompVisDenoiser(ldm, quantize=False)
    cond, uncond = randn((1, 77, 768), s["cond"]), randn((1, 77, 768), s["uncond"])
    x_T = randn((1, 4, 64, 64), s["x_T"])
    noise = randn((20, 1, 4, 64, 64), s["noise"])
    wrapper = LDMWrapperForKDiffusion(den, cond, uncond, 7.5)
    sig = den.get_sigmas(20)
    trace, den_trace = [], []
    it = iter(range(20))

    def cb(d):
        trace.append(d["x"].clone())
        den_trace.append(d["denoised"].clone())

    t0 = time.time()
    with torch.no_grad():
        xf = sampling.sample_euler_ancestral(wrapper, x_T * sig[0], sig, disable=True,
                                             noise_sampler=lambda a, b: noise[next(it)], callback=cb)
    print("cfg1: 20 steps in %.1f s" % (time.time() - t0))
    vcfg = O.SD15_VAE
    vsd = O.make_weights(O.decoder_param_shapes(vcfg), seed=s["vae"])
    dec, pq = ref_shim.reference_decoder(vcfg, vsd)
    with torch.no_grad():
        img = dec(pq(xf / 0.18215))   # ddpm.py:794-798
    np.savez_compressed(os.path.join(GOLD, "cfg1_euler20.npz"), sigmas=sig.numpy(),
                        trace=torch.stack(trace[1:] + [xf]).numpy(), denoised=torch.stack(den_trace).numpy(),
                        image=img.numpy().astype(np.float16), seeds=np.array(list(s.values())),
                        weights_checksum=np.float64(O.weights_checksum(sd)))
    print("cfg1_euler20.npz written; final |x|max %.3f, image absmax %.3f" % (xf.abs().max(), img.abs().max()))


def gen_cfg2():
    from ldm.models.diffusion.ddim import DDIMSampler
    s = SEEDS_CFG2
    unet, sd = _sd15_unet()
    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    b = 2
    cond, uncond = randn((b, 77, 768), s["cond"]), randn((b, 77, 768), s["uncond"])
    x_T = randn((b, 4, 64, 64), s["x_T"])
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    smp = DDIMSampler(ldm)
    inter = []
    t0 = time.time()
    with torch.no_grad():
        x, _ = smp.sample(S=50, batch_size=b, shape=[4, 64, 64], conditioning=cond, eta=0.0, x_T=x_T,
                          unconditional_guidance_scale=7.5, unconditional_conditioning=uncond, verbose=False,
                          img_callback=lambda pred_x0, i: inter.append(pred_x0.clone()))
    print("cfg2: 50 steps in %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLD, "cfg2_ddim50.npz"), final=x.numpy(),
                        pred_x0=torch.stack([inter[0], inter[25], inter[49]]).numpy(), pred_x0_steps=np.array([0, 25, 49]),
                        seeds=np.array(list(s.values())), weights_checksum=np.float64(O.weights_checksum(sd)))
    print("cfg2_ddim50.npz written; final |x|max %.3f" % x.abs().max())


def gen_cfg3():
    s = SEEDS_CFG3
    vcfg = O.SD15_VAE
    vsd = O.make_weights(O.decoder_param_shapes(vcfg), seed=s["vae"])
    dec, pq = ref_shim.reference_decoder(vcfg, vsd)
    z = randn((2, 4, 64, 64), s["z"])
    with torch.no_grad():
        img = dec(pq(z))
    np.savez_compressed(os.path.join(GOLD, "cfg3_vae64.npz"), out=img.numpy().astype(np.float16),
                        seeds=np.array(list(s.values())), weights_checksum=np.float64(O.weights_checksum(vsd)))
    print("cfg3_vae64.npz written; image absmax %.3f std %.3f" % (img.abs().max(), img.std()))


def gen_hires():
    s = SEEDS_HIRES
    unet, sd = _sd15_unet()
    x = randn((1, 4, 128, 128), s["x"])
    t = torch.tensor([500.0])
    ctx = randn((1, 77, 768), s["ctx"])
    t0 = time.time()
    with torch.no_grad():
        y = unet(x, t, context=ctx)
    print("hires: one 128x128 forward in %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLD, "hires_unet128.npz"), out=y.numpy(), t=t.numpy(),
                        seeds=np.array(list(s.values())), weights_checksum=np.float64(O.weights_checksum(sd)))
    print("hires_unet128.npz written; out absmax %.3f std %.3f" % (y.abs().max(), y.std()))


def gen_sdxl():
    from oracle import sgm_oracle as S
    from oracle.make_golden_sgm import install_sgm
    install_sgm()
    from sgm.modules.diffusionmodules.openaimodel import UNetModel
    s = SEEDS_SDXL
    cfg = S.SDXL_UNET
    sd = O.make_weights(S.sgm_unet_param_shapes(cfg), seed=s["unet"])
    unet = UNetModel(in_channels=4, model_channels=cfg.model_channels, out_channels=4,
                     num_res_blocks=cfg.num_res_blocks, attention_resolutions=list(cfg.attention_resolutions),
                     channel_mult=list(cfg.channel_mult), num_head_channels=cfg.num_head_channels,
                     use_linear_in_transformer=True, transformer_depth=list(cfg.transformer_depth),
                     context_dim=cfg.context_dim, num_classes="sequential", adm_in_channels=cfg.adm_in_channels,
                     use_checkpoint=False, spatial_transformer_attn_type="softmax")
    unet.load_state_dict(sd, strict=True)
    unet.eval()
    chk = O.weights_checksum(sd)
    del sd
    x = randn((2, 4, 128, 128), s["x"])
    t = torch.tensor([7, 640])
    ctx, y = randn((2, 77, cfg.context_dim), s["ctx"]), randn((2, cfg.adm_in_channels), s["y"])
    t0 = time.time()
    with torch.no_grad():
        out = unet(x, t, context=ctx, y=y)
    print("sdxl: one CFG-pair 128x128 forward in %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLD, "sdxl_unet128.npz"), out=out.numpy(), t=t.numpy(),
                        seeds=np.array(list(s.values())), weights_checksum=np.float64(chk))
    print("sdxl_unet128.npz written; out absmax %.3f std %.3f" % (out.abs().max(), out.std()))


def scale_decoder_for_overflow(vsd, gain: float):
    """Multiply conv_in by `gain`: every GroupNorm downstream is scale invariant, so the image is (up to eps) unchanged
    while the residual stream of the mid block / first up level carries values `gain` times larger -- beyond fp16's
    65504 for gain = 32768 -- which is the situation real SDXL VAE weights create and the reason the reference decodes
    in fp32 (sgm/models/diffusion.py:125, sd_xl_base.yaml:5)."""
    out = dict(vsd)
    out["decoder.conv_in.weight"] = vsd["decoder.conv_in.weight"] * gain
    out["decoder.conv_in.bias"] = vsd["decoder.conv_in.bias"] * gain
    return out


def gen_sdxlvae():
    s = SEEDS_SDXLVAE
    vcfg = O.SD15_VAE    # the SDXL VAE has the SD1.5 decoder graph (sd_xl_base.yaml first_stage_config ddconfig)
    vsd = O.make_weights(O.decoder_param_shapes(vcfg), seed=s["vae"])
    big = scale_decoder_for_overflow(vsd, 32768.0)
    dec, pq = ref_shim.reference_decoder(vcfg, big)
    z = randn((1, 4, 32, 32), s["z"])
    peak = []
    hook = dec.mid.block_1.register_forward_hook(lambda m, i, o: peak.append(float(o.abs().max())))
    with torch.no_grad():
        img = dec(pq(z))
    hook.remove()
    print("sdxlvae: mid.block_1 output absmax %.1f (fp16 max 65504)" % peak[0])
    dec2, pq2 = ref_shim.reference_decoder(vcfg, vsd)
    z128 = randn((1, 4, 128, 128), s["z128"])
    t0 = time.time()
    with torch.no_grad():
        img128 = dec2(pq2(z128))
    print("sdxlvae: 128x128-latent decode in %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(GOLD, "sdxl_vae_overflow.npz"), out=img.numpy().astype(np.float16),
                        mid_absmax=np.float32(peak[0]), gain=np.float32(32768.0), seeds=np.array(list(s.values())),
                        weights_checksum=np.float64(O.weights_checksum(vsd)))
    # 1024^2 image: keep the fixture small -- a 256x256 crop at full precision + 8x8 block means of the whole image
    full = img128[0]
    pooled = torch.nn.functional.avg_pool2d(full[None], 8)[0]
    np.savez_compressed(os.path.join(GOLD, "sdxl_vae128.npz"), crop=full[:, 384:640, 384:640].numpy().astype(np.float16),
                        pooled=pooled.numpy().astype(np.float32), seeds=np.array(list(s.values())))
    print("sdxl_vae_overflow.npz / sdxl_vae128.npz written")


ALL = {"cfg1": gen_cfg1, "cfg2": gen_cfg2, "cfg3": gen_cfg3, "hires": gen_hires, "sdxl": gen_sdxl,
       "sdxlvae": gen_sdxlvae}

if __name__ == "__main__":
    ref_shim.install()
    torch.set_num_threads(os.cpu_count())
    for name in (sys.argv[1:] or list(ALL)):
        t0 = time.time()
        ALL[name]()
        print("%s done in %.1f s" % (name, time.time() - t0), flush=True)
