"""ORACLE support -- goldens for the remaining sgm samplers from the UNMODIFIED reference
sgm/modules/diffusionmodules/sampling.py: HeunEDMSampler (:147-220,321-358) and LinearMultistepSampler (:271-306),
through the reference's DiscreteDenoiser / VanillaCFG / OpenAIWrapper on the tiny sgm UNet of tiny_sgm.npz.
    python oracle/make_golden_sgm_samplers.py  ->  tests/golden/tiny_sgm_samplers.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import sd_oracle as O  # noqa: E402
from oracle import sgm_oracle as S  # noqa: E402
from oracle.make_golden_sgm import GOLD, install_sgm  # noqa: E402

EDM = {"sigma_min": 0.0292, "sigma_max": 14.6146, "rho": 3.0}
STEPS = 6


def main():
    install_sgm()
    from sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from sgm.modules.diffusionmodules.openaimodel import UNetModel
    from sgm.modules.diffusionmodules.sampling import HeunEDMSampler, LinearMultistepSampler
    from sgm.modules.diffusionmodules.wrappers import OpenAIWrapper

    g = np.load(os.path.join(GOLD, "tiny_sgm.npz"))
    cfg = S.TINY_SGM_UNET
    sd = O.make_weights(S.sgm_unet_param_shapes(cfg), seed=300)
    assert abs(O.weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6
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
    common = dict(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization", "params": EDM},
                  num_steps=STEPS, device="cpu",
                  guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                 "params": {"scale": float(g["cfg_scale"])}})
    cond = {"crossattn": torch.from_numpy(g["cond_crossattn"]), "vector": torch.from_numpy(g["cond_vector"])}
    uc = {"crossattn": torch.from_numpy(g["uc_crossattn"]), "vector": torch.from_numpy(g["uc_vector"])}
    x_T = torch.from_numpy(g["x_T"])
    denoiser = lambda inp, sigma, c: den(model, inp, sigma, c)
    with torch.no_grad():
        heun = HeunEDMSampler(**common)(denoiser, x_T.clone(), cond=cond, uc=uc)
        lms = LinearMultistepSampler(order=4, **common)(denoiser, x_T.clone(), cond=cond, uc=uc)
    print("heun absmax %.3f, lms absmax %.3f, |heun - lms| max %.3f" %
          (heun.abs().max(), lms.abs().max(), (heun - lms).abs().max()))
    np.savez_compressed(os.path.join(GOLD, "tiny_sgm_samplers.npz"), heun_final=heun.numpy(), lms_final=lms.numpy(),
                        steps=np.int64(STEPS), lms_order=np.int64(4))


if __name__ == "__main__":
    main()
