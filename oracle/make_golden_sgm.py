"""ORACLE support -- SDXL-side goldens from the UNMODIFIED vendored sgm of the reference (modules/sdxl/sgm):
tiny UNet forward with vector conditioning, sigma tables, and a DPM++ 2M trajectory through the reference's own
DiscreteDenoiser / VanillaCFG / OpenAIWrapper / DPMPP2MSampler.   python oracle/make_golden_sgm.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import sd_oracle as O  # noqa: E402
from oracle import sgm_oracle as S  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def install_sgm():
    """Import sgm submodules without running sgm/__init__.py (it pulls the CLIP encoders): namespace-style package
    stubs with the real directories as __path__, plus stubs for un-vendored imports that are never executed."""
    ref_shim.install()
    R = os.path.join(ref_shim.REFERENCE_ROOT, "modules", "sdxl", "sgm")

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    for n in ("pytorch_lightning", "kornia", "ftfy", "open_clip"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            m.__getattr__ = lambda k: _Any
            sys.modules[n] = m
    oc = sys.modules["omegaconf"]
    oc.OmegaConf = _Any
    oc.ListConfig = oc.listconfig.ListConfig
    oc.DictConfig = dict
    for name, path in (("sgm", R), ("sgm.modules", R + "/modules"), ("sgm.models", R + "/models"),
                       ("sgm.modules.diffusionmodules", R + "/modules/diffusionmodules")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    os.environ["GPU_DEVICE"] = "cpu"


def randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator(device="cpu").manual_seed(seed))


def main():
    install_sgm()
    from sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from sgm.modules.diffusionmodules.discretizer import EDMDiscretization
    from sgm.modules.diffusionmodules.openaimodel import UNetModel
    from sgm.modules.diffusionmodules.sampling import DPMPP2MSampler
    from sgm.modules.diffusionmodules.wrappers import OpenAIWrapper

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
    x, t = randn((2, 4, 16, 16), 1), torch.tensor([7, 640])
    ctx, y = randn((2, 7, cfg.context_dim), 2), randn((2, cfg.adm_in_channels), 3)
    with torch.no_grad():
        out = unet(x, t, context=ctx, y=y)
    print("tiny sgm unet out: mean %.4f std %.4f" % (out.mean().item(), out.std().item()))

    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"})
    model = OpenAIWrapper(unet)
    # "DPM++ 2M Karras" of the SDXL pipeline: EDMDiscretization(0.0292, 14.6146, rho 3.0), sdxl_pipeline/options.py:204-225
    smp = DPMPP2MSampler(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                                                "params": {"sigma_min": 0.0292, "sigma_max": 14.6146, "rho": 3.0}},
                         num_steps=6, device="cpu",
                         guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                        "params": {"scale": 5.0}})
    cond = {"crossattn": randn((2, 7, cfg.context_dim), 4), "vector": randn((2, cfg.adm_in_channels), 5)}
    uc = {"crossattn": randn((2, 7, cfg.context_dim), 6), "vector": randn((2, cfg.adm_in_channels), 7)}
    x_T = randn((2, 4, 16, 16), 8)
    denoiser = lambda inp, sigma, c: den(model, inp, sigma, c)
    with torch.no_grad():
        z = smp(denoiser, x_T.clone(), cond=cond, uc=uc)
    np.savez_compressed(
        os.path.join(GOLD, "tiny_sgm.npz"), x=x.numpy(), t=t.numpy(), context=ctx.numpy(), y=y.numpy(), out=out.numpy(),
        weights_checksum=np.float64(O.weights_checksum(sd)), denoiser_sigmas=den.sigmas.numpy(),
        edm_sigmas_6=EDMDiscretization(0.0292, 14.6146, 3.0)(6).numpy(),
        edm_sigmas_30=EDMDiscretization(0.0292, 14.6146, 3.0)(30).numpy(),
        cond_crossattn=cond["crossattn"].numpy(), cond_vector=cond["vector"].numpy(),
        uc_crossattn=uc["crossattn"].numpy(), uc_vector=uc["vector"].numpy(), x_T=x_T.numpy(),
        cfg_scale=np.float32(5.0), dpmpp2m_final=z.numpy())
    print("tiny_sgm.npz written")


if __name__ == "__main__":
    main()
