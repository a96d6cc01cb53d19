"""ORACLE support -- goldens for the remaining k-diffusion samplers (Heun, DPM-2, DPM-2 ancestral, LMS, DPM++ 2S
ancestral) from the UNMODIFIED reference functions (modules/k_diffusion/sampling.py:167-286,517-548) through the
reference's own CompVisDenoiser / LDMWrapperForKDiffusion chain, tiny UNet, injected noise.
    python oracle/make_golden_samplers.py   ->  tests/golden/tiny_samplers_extra.npz"""
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


def main():
    ref_shim.install()
    from k_diffusion import external, sampling
    from ldm.models.diffusion.ldm_wrapper_for_k_diffusion import LDMWrapperForKDiffusion
    cfg = O.TINY_UNET
    sd = O.make_weights(O.unet_param_shapes(cfg), seed=100)
    unet = ref_shim.reference_unet(cfg, sd)
    g = np.load(os.path.join(GOLD, "tiny_sampling.npz"))      # same cond / uncond / x_T / noise as the other samplers
    cond, uncond = torch.from_numpy(g["cond"]), torch.from_numpy(g["uncond"])
    x_T, noise = torch.from_numpy(g["x_T"]), torch.from_numpy(g["noise"])
    betas, ac, ac_prev = schedule_tensors()
    ldm = ref_shim.DuckLDM(unet, betas, ac, ac_prev)
    den = external.CompVisDenoiser(ldm, quantize=False)
    wrapper = LDMWrapperForKDiffusion(den, cond, uncond, float(g["cfg_scale"]))
    sig_d = den.get_sigmas(STEPS)                                             # Heun / LMS front ends (:324,:355)
    sig_k = sampling.get_sigmas_karras(STEPS, float(den.sigma_min), float(den.sigma_max))   # DPM-2 family (:335,:345,:366)
    out = {"sigmas_discrete": sig_d.numpy(), "sigmas_karras": sig_k.numpy()}

    def ns():
        it = iter(range(STEPS))
        return lambda s, sn: noise[next(it)]
    with torch.no_grad():
        out["heun"] = sampling.sample_heun(wrapper, x_T * sig_d[0], sig_d, disable=True).numpy()
        out["lms"] = sampling.sample_lms(wrapper, x_T * sig_d[0], sig_d, disable=True).numpy()
        out["dpm_2"] = sampling.sample_dpm_2(wrapper, x_T * sig_k[0], sig_k, disable=True).numpy()
        out["dpm_2_ancestral"] = sampling.sample_dpm_2_ancestral(wrapper, x_T * sig_k[0], sig_k, disable=True,
                                                                 noise_sampler=ns()).numpy()
        out["dpmpp_2s_ancestral"] = sampling.sample_dpmpp_2s_ancestral(wrapper, x_T * sig_k[0], sig_k, disable=True,
                                                                       noise_sampler=ns()).numpy()
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))
    np.savez_compressed(os.path.join(GOLD, "tiny_samplers_extra.npz"), **out)


if __name__ == "__main__":
    main()
