"""ORACLE support -- goldens that close SURVEY 8f N4 / 8a a16, all from the UNMODIFIED reference on the tiny UNet:

* SDE family with an injected `noise_sampler` (k_diffusion/sampling.py:551-617 sample_dpmpp_sde, :619-662
  sample_dpmpp_2m_sde midpoint + heun, :664-717 sample_dpmpp_3m_sde) -- torchsde's Brownian tree is never touched;
* `s_churn > 0` in sample_euler / sample_heun / sample_dpm_2 (:118-144, :167-225): the reference draws
  `torch.randn_like(x)` itself, so the draw is injected by replacing `torch.randn_like` for the duration of the call;
* the sampler front ends: `EulerAncestralSampler.sample`, `Dpmpp2mSampler.sample(denoising_steps=...)` and
  `KDiffusionSamplerBase.stochastic_encode` (ldm/models/diffusion/k_diffusion_samplers.py:197-297).  `sample` wraps
  `do_sample` in `torch.autocast(GPU_DEVICE)`; on this CPU that would mean bf16, so `torch.autocast` is replaced by a
  null context while the golden is generated (the golden is the fp32 trajectory, like every other golden here);
* sgm `DPMPP2SAncestralSampler` (sgm/modules/diffusionmodules/sampling.py:384-457, its `noise_sampler` attribute
  replaced) and `EulerEDMSampler(s_churn=...)` (:147-220).

    python oracle/make_golden_samplers2.py   ->  tests/golden/tiny_samplers_sde.npz, tiny_sgm_samplers2.npz"""
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import sd_oracle as O  # noqa: E402
from oracle.make_golden import randn, schedule_tensors  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
STEPS = 5
CHURN = 2.0


@contextlib.contextmanager
def injected_randn_like(seq):
    """torch.randn_like(x) -> next tensor of `seq` (the reference calls it directly in the churn branch)."""
    real = torch.randn_like
    it = iter(seq)
    torch.randn_like = lambda x, *a, **k: next(it).to(x.device)
    try:
        yield
    finally:
        torch.randn_like = real


@contextlib.contextmanager
def no_autocast():
    real = torch.autocast
    torch.autocast = lambda *a, **k: contextlib.nullcontext()
    try:
        yield
    finally:
        torch.autocast = real


def main_ldm():
    ref_shim.install()
    from k_diffusion import external, sampling
    from ldm.models.diffusion import k_diffusion_samplers as K
    from ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    cfg = O.TINY_UNET
    sd = O.make_weights(O.unet_param_shapes(cfg), seed=100)
    unet = ref_shim.reference_unet(cfg, sd)
    g = np.load(os.path.join(GOLD, "tiny_sampling.npz"))
    cond, uncond = torch.from_numpy(g["cond"]), torch.from_numpy(g["uncond"])
    x_T = torch.from_numpy(g["x_T"])
    noise = randn((2 * STEPS + 2, *x_T.shape), 91)          # up to two draws per step (sample_dpmpp_sde)
    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    den = external.CompVisDenoiser(ldm, quantize=False)
    scale = float(g["cfg_scale"])
    wrapper = LDMWrapperForKDiffusion(den, cond, uncond, scale)
    sig_d = den.get_sigmas(STEPS)
    sig_k = sampling.get_sigmas_karras(STEPS, 0.0316386, 14.5521805)     # the SDE front ends (:373-411)
    out = {"noise": noise.numpy(), "sigmas_discrete": sig_d.numpy(), "sigmas_karras": sig_k.numpy(),
           "s_churn": np.float32(CHURN)}

    def ns():
        it = iter(range(noise.shape[0]))
        return lambda s, sn: noise[next(it)]

    with torch.no_grad():
        x0 = x_T * sig_k[0]
        out["dpmpp_sde"] = sampling.sample_dpmpp_sde(wrapper, x0, sig_k, disable=True, noise_sampler=ns()).numpy()
        out["dpmpp_2m_sde"] = sampling.sample_dpmpp_2m_sde(wrapper, x0, sig_k, disable=True, noise_sampler=ns()).numpy()
        out["dpmpp_2m_sde_heun"] = sampling.sample_dpmpp_2m_sde(wrapper, x0, sig_k, disable=True, noise_sampler=ns(),
                                                                solver_type="heun").numpy()
        out["dpmpp_3m_sde"] = sampling.sample_dpmpp_3m_sde(wrapper, x0, sig_k, disable=True, noise_sampler=ns()).numpy()
        out["dpmpp_2m_sde_eta0"] = sampling.sample_dpmpp_2m_sde(wrapper, x0, sig_k, disable=True, eta=0.0,
                                                                noise_sampler=ns()).numpy()
        x0 = x_T * sig_d[0]
        with injected_randn_like(noise):
            out["euler_churn"] = sampling.sample_euler(wrapper, x0, sig_d, disable=True, s_churn=CHURN).numpy()
        with injected_randn_like(noise):
            out["heun_churn"] = sampling.sample_heun(wrapper, x0, sig_d, disable=True, s_churn=CHURN, s_noise=1.003).numpy()
        with injected_randn_like(noise):
            out["dpm_2_churn"] = sampling.sample_dpm_2(wrapper, x0, sig_k, disable=True, s_churn=CHURN,
                                                       s_tmin=0.05, s_tmax=10.0).numpy()

        # front ends.  register_buffer moves to "cuda" when torch.cuda.is_available() (forced True by the shim): keep on CPU
        K.KDiffusionSamplerBase.register_buffer = lambda self, name, attr: setattr(self, name, attr)
        common = dict(batch_size=x_T.shape[0], shape=list(x_T.shape[1:]), conditioning=cond,
                      unconditional_guidance_scale=scale, unconditional_conditioning=uncond)
        with no_autocast():
            with injected_randn_like(noise):
                xa, _ = K.EulerAncestralSampler(ldm).sample(S=STEPS, x0=x_T * sig_d[0], **common)
            smp = K.Dpmpp2mSampler(ldm)
            xm, _ = smp.sample(S=6, x0=x_T * 2.0, denoising_steps=3, **common)
        out["front_euler_a"] = xa.numpy()
        out["front_dpmpp2m_img2img"] = xm.numpy()
        out["front_dpmpp2m_img2img_sigmas"] = smp.sigmas.numpy()
        enc = smp.stochastic_encode(x_T, torch.tensor([2, 2]), 5, noise=noise[0])
        out["front_stochastic_encode"] = enc.numpy()
        enc2 = smp.stochastic_encode(x_T, torch.tensor([1, 3]), 5, noise=noise[1])     # per-sample indices
        out["front_stochastic_encode_ragged"] = enc2.numpy()
    for k, v in out.items():
        print(k, np.shape(v), float(np.abs(v).max()))
    np.savez_compressed(os.path.join(GOLD, "tiny_samplers_sde.npz"), **out)


def main_sgm():
    from oracle import sgm_oracle as S
    from oracle.make_golden_sgm import install_sgm
    from oracle.make_golden_sgm_samplers import EDM
    install_sgm()
    from sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from sgm.modules.diffusionmodules.openaimodel import UNetModel
    from sgm.modules.diffusionmodules.sampling import DPMPP2SAncestralSampler, EulerAncestralSampler, EulerEDMSampler
    from sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    g = np.load(os.path.join(GOLD, "tiny_sgm.npz"))
    cfg = S.TINY_SGM_UNET
    sd = O.make_weights(S.sgm_unet_param_shapes(cfg), seed=300)
    unet = UNetModel(in_channels=4, model_channels=cfg.model_channels, out_channels=4,
                     num_res_blocks=cfg.num_res_blocks, attention_resolutions=list(cfg.attention_resolutions),
                     channel_mult=list(cfg.channel_mult), num_head_channels=cfg.num_head_channels,
                     use_linear_in_transformer=True, transformer_depth=list(cfg.transformer_depth),
                     context_dim=cfg.context_dim, num_classes="sequential", adm_in_channels=cfg.adm_in_channels,
                     use_checkpoint=False, spatial_transformer_attn_type="softmax")
    unet.load_state_dict(sd, strict=True)
    unet.eval()
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"})
    model = OpenAIWrapper(unet)
    steps = 6
    common = dict(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization", "params": EDM},
                  num_steps=steps, device="cpu",
                  guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                 "params": {"scale": float(g["cfg_scale"])}})
    cond = {"crossattn": torch.from_numpy(g["cond_crossattn"]), "vector": torch.from_numpy(g["cond_vector"])}
    uc = {"crossattn": torch.from_numpy(g["uc_crossattn"]), "vector": torch.from_numpy(g["uc_vector"])}
    x_T = torch.from_numpy(g["x_T"])
    noise = randn((steps + 1, *x_T.shape), 92)
    denoiser = lambda inp, sigma, c: den(model, inp, sigma, c)

    def inject(smp):
        it = iter(range(noise.shape[0]))
        smp.noise_sampler = lambda x: noise[next(it)]
        return smp

    with torch.no_grad():
        d2s = inject(DPMPP2SAncestralSampler(**common))(denoiser, x_T.clone(), cond=cond, uc=uc)
        ea = inject(EulerAncestralSampler(**common))(denoiser, x_T.clone(), cond=cond, uc=uc)
        with injected_randn_like(noise):
            ec = EulerEDMSampler(s_churn=CHURN, **common)(denoiser, x_T.clone(), cond=cond, uc=uc)
    print("sgm dpmpp2s_a absmax %.3f, euler_a %.3f, euler churn %.3f" % (d2s.abs().max(), ea.abs().max(), ec.abs().max()))
    np.savez_compressed(os.path.join(GOLD, "tiny_sgm_samplers2.npz"), noise=noise.numpy(), dpmpp2s_ancestral=d2s.numpy(),
                        euler_ancestral=ea.numpy(), euler_churn=ec.numpy(), steps=np.int64(steps), s_churn=np.float32(CHURN))


def main_ddim_mask():
    """DDIM inpainting branch (ldm/models/diffusion/ddim.py:171-174): `sample(mask=, x0=)` re-noises the known region with
    `model.q_sample(x0, ts)` (ddpm.py:296-299, its randn_like injected) and blends before every step."""
    ref_shim.install()
    from ldm.models.diffusion.ddim import DDIMSampler
    cfg = O.TINY_UNET
    sd = O.make_weights(O.unet_param_shapes(cfg), seed=100)
    unet = ref_shim.reference_unet(cfg, sd)
    g = np.load(os.path.join(GOLD, "tiny_sampling.npz"))
    cond, uncond, x_T = torch.from_numpy(g["cond"]), torch.from_numpy(g["uncond"]), torch.from_numpy(g["x_T"])
    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    x0 = randn(tuple(x_T.shape), 93)
    mask = (randn((x_T.shape[0], 1, *x_T.shape[2:]), 94) > 0).float()      # 1 = keep the known latent
    noise = randn((5, *x_T.shape), 95)
    with torch.no_grad(), injected_randn_like(noise):
        x, _ = DDIMSampler(ldm).sample(S=5, batch_size=x_T.shape[0], shape=list(x_T.shape[1:]), conditioning=cond, eta=0.0,
                                       x_T=x_T, mask=mask, x0=x0, unconditional_guidance_scale=float(g["cfg_scale"]),
                                       unconditional_conditioning=uncond, verbose=False)
    np.savez_compressed(os.path.join(GOLD, "tiny_ddim_mask.npz"), x0=x0.numpy(), mask=mask.numpy(), noise=noise.numpy(),
                        final=x.numpy())
    print("tiny_ddim_mask.npz written; final absmax %.3f" % x.abs().max())


if __name__ == "__main__":
    if "ddim_mask" in sys.argv[1:]:
        main_ddim_mask()
    else:
        main_ldm()
        main_sgm()
        main_ddim_mask()
