"""SDXL side of the path: the reference's vendored sgm (modules/sdxl/sgm).
CPU: oracle/sgm_oracle.py against the golden produced by the unmodified reference (oracle/make_golden_sgm.py), sigma
tables / index quantisation bit-exact, state-dict key layout.  GPU: the cremage_b200.sgm mirrors (UNetModel with vector
conditioning, DiscreteDenoiser, VanillaCFG, DPMPP2MSampler) against the same golden."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from oracle import sgm_oracle as S
from tests._models import tol  # noqa: E402
from tests._models import gold

EDM = dict(sigma_min=0.0292, sigma_max=14.6146, rho=3.0)  # "DPM++ 2M Karras", sdxl_pipeline/options.py:204-225


def _sgm_kwargs(cfg: S.SgmUNetConfig):
    return dict(in_channels=cfg.in_channels, model_channels=cfg.model_channels, out_channels=cfg.out_channels,
                num_res_blocks=cfg.num_res_blocks, attention_resolutions=list(cfg.attention_resolutions),
                channel_mult=list(cfg.channel_mult), num_head_channels=cfg.num_head_channels,
                use_linear_in_transformer=True, transformer_depth=list(cfg.transformer_depth),
                context_dim=cfg.context_dim, num_classes="sequential", adm_in_channels=cfg.adm_in_channels,
                use_checkpoint=False, spatial_transformer_attn_type="softmax")


def _weights():
    g = gold("tiny_sgm.npz")
    sd = O.make_weights(S.sgm_unet_param_shapes(S.TINY_SGM_UNET), seed=300)
    assert abs(O.weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6
    return g, sd


def test_oracle_sgm_unet_matches_reference_golden():
    g, sd = _weights()
    with torch.no_grad():
        out = S.sgm_unet_forward(sd, S.TINY_SGM_UNET, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]),
                                 torch.from_numpy(g["context"]), torch.from_numpy(g["y"]))
    assert np.abs(out.numpy() - g["out"]).max() < 2e-5


def test_oracle_sigma_tables_bit_exact():
    g = gold("tiny_sgm.npz")
    assert np.array_equal(S.legacy_ddpm_sigma_table(1000).numpy(), g["denoiser_sigmas"])
    assert np.array_equal(S.edm_sigmas(6, **EDM).numpy(), g["edm_sigmas_6"])
    assert np.array_equal(S.edm_sigmas(30, **EDM).numpy(), g["edm_sigmas_30"])


def test_oracle_dpmpp2m_trajectory_matches_reference_golden():
    g, sd = _weights()
    table = S.legacy_ddpm_sigma_table(1000)
    net = lambda x, t, c: S.sgm_unet_forward(sd, S.TINY_SGM_UNET, x, t, c["crossattn"], c["vector"])
    den = lambda x, sigma, c: S.discrete_denoise(net, table, x, sigma, c)
    cond = {"crossattn": torch.from_numpy(g["cond_crossattn"]), "vector": torch.from_numpy(g["cond_vector"])}
    uc = {"crossattn": torch.from_numpy(g["uc_crossattn"]), "vector": torch.from_numpy(g["uc_vector"])}
    with torch.no_grad():
        z = S.sample_dpmpp_2m_sgm(den, torch.from_numpy(g["x_T"]), S.edm_sigmas(6, **EDM), cond, uc,
                                  float(g["cfg_scale"]))
    assert np.abs(z.numpy() - g["dpmpp2m_final"]).max() < 5e-4


def test_host_mirrors_bit_exact():
    """Discretizations and DiscreteDenoiser index quantisation of the mirrors == the reference's (bit-exact)."""
    from cremage_b200.sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from cremage_b200.sgm.modules.diffusionmodules.discretizer import EDMDiscretization, LegacyDDPMDiscretization
    g = gold("tiny_sgm.npz")
    assert np.array_equal(EDMDiscretization(**EDM)(6).numpy(), g["edm_sigmas_6"])
    assert np.array_equal(EDMDiscretization(**EDM)(30).numpy(), g["edm_sigmas_30"])
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"})
    assert np.array_equal(den.sigmas.numpy(), g["denoiser_sigmas"])
    assert isinstance(den.discretization, LegacyDDPMDiscretization)
    sig = torch.from_numpy(g["edm_sigmas_30"][:-1].copy())
    table = torch.from_numpy(g["denoiser_sigmas"])
    want = (sig - table[:, None]).abs().argmin(dim=0)
    assert torch.equal(den.sigma_to_idx(sig), want)
    assert torch.equal(den.possibly_quantize_sigma(sig), table[want])
    assert den.possibly_quantize_c_noise(table[want]).dtype == torch.int64
    assert torch.equal(den.possibly_quantize_c_noise(table[want]), want)


def test_state_dict_keys_match_reference_layout():
    from cremage_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
    for cfg in (S.TINY_SGM_UNET, S.SDXL_UNET):
        with torch.device("meta"):
            m = UNetModel(**_sgm_kwargs(cfg))
        shapes = S.sgm_unet_param_shapes(cfg)
        have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert have == shapes
    assert sum(int(np.prod(s)) for s in S.sgm_unet_param_shapes(S.SDXL_UNET).values()) == 2567463684


def test_guider_shapes_like_reference_test():
    """test/sgm/guiders_test.py:39-87 (shape behaviour of VanillaCFG.prepare_inputs) without a device."""
    from cremage_b200.sgm.modules.diffusionmodules.guiders import VanillaCFG
    gd = VanillaCFG(scale=5.0)
    x, s = torch.zeros(3, 4, 8, 8), torch.ones(3)
    c = {"crossattn": torch.zeros(3, 7, 16), "vector": torch.zeros(3, 12)}
    uc = {"crossattn": torch.ones(3, 7, 16), "vector": torch.ones(3, 12)}
    x2, s2, c2 = gd.prepare_inputs(x, s, c, uc)
    assert x2.shape == (6, 4, 8, 8) and s2.shape == (6,)
    assert c2["crossattn"].shape == (6, 7, 16) and c2["vector"].shape == (6, 12)
    assert torch.equal(c2["crossattn"][:3], uc["crossattn"]) and torch.equal(c2["vector"][3:], c["vector"])  # uc first


def _build_sgm_unet(cfg, sd):
    from cremage_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device("meta"):
        m = UNetModel(**_sgm_kwargs(cfg))
    m = m.to_empty(device="cpu")
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.gpu
def test_sgm_unet_forward_vs_reference_golden():
    g, sd = _weights()
    m = _build_sgm_unet(S.TINY_SGM_UNET, sd)
    out = m(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda(),
            context=torch.from_numpy(g["context"]).cuda(), y=torch.from_numpy(g["y"]).cuda())
    want = torch.from_numpy(g["out"])
    err = (out.cpu() - want).abs().max().item()
    rel = ((out.cpu() - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item()
    print(f"[parity] tiny sgm UNet: max_abs_err={err:.4e} rel_rms={rel:.4e} ref_absmax={want.abs().max():.3f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)
    out2 = m(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda(),
             context=torch.from_numpy(g["context"]).cuda(), y=torch.from_numpy(g["y"]).cuda())
    assert torch.equal(out, out2)  # graph replay is bit-identical


@pytest.mark.gpu
def test_sgm_dpmpp2m_trajectory_vs_reference_golden():
    from cremage_b200.sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from cremage_b200.sgm.modules.diffusionmodules.sampling import DPMPP2MSampler
    from cremage_b200.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    g, sd = _weights()
    model = OpenAIWrapper(_build_sgm_unet(S.TINY_SGM_UNET, sd))
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}).cuda()
    smp = DPMPP2MSampler(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization",
                                                "params": EDM},
                         num_steps=6, guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                                     "params": {"scale": float(g["cfg_scale"])}})
    cond = {"crossattn": torch.from_numpy(g["cond_crossattn"]).cuda(), "vector": torch.from_numpy(g["cond_vector"]).cuda()}
    uc = {"crossattn": torch.from_numpy(g["uc_crossattn"]).cuda(), "vector": torch.from_numpy(g["uc_vector"]).cuda()}
    x_T = torch.from_numpy(g["x_T"]).cuda()
    x_keep = x_T.clone()
    z = smp(lambda inp, sigma, c: den(model, inp, sigma, c), x_T, cond=cond, uc=uc)
    assert torch.equal(x_T, x_keep)  # the caller's latent is not mutated
    want = torch.from_numpy(g["dpmpp2m_final"])
    err = (z.cpu() - want).abs().max().item()
    print(f"[parity] sgm DPM++2M 6 steps: max_abs_err={err:.4e} latent_absmax={want.abs().max():.2f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)


# ---------------------------------------------------------------------------------------------------------------
# remaining sgm samplers: HeunEDMSampler, LinearMultistepSampler -- goldens from the unmodified
# reference (oracle/make_golden_sgm_samplers.py)
# ---------------------------------------------------------------------------------------------------------------
def _cond(g, dev="cpu"):
    cond = {"crossattn": torch.from_numpy(g["cond_crossattn"]).to(dev), "vector": torch.from_numpy(g["cond_vector"]).to(dev)}
    uc = {"crossattn": torch.from_numpy(g["uc_crossattn"]).to(dev), "vector": torch.from_numpy(g["uc_vector"]).to(dev)}
    return cond, uc


@pytest.mark.parametrize("which", ["heun", "lms"])
def test_oracle_heun_lms_trajectories_match_reference_golden(which):
    g, sd = _weights()
    gs = gold("tiny_sgm_samplers.npz")
    table = S.legacy_ddpm_sigma_table(1000)
    net = lambda x, t, c: S.sgm_unet_forward(sd, S.TINY_SGM_UNET, x, t, c["crossattn"], c["vector"])
    den = lambda x, sigma, c: S.discrete_denoise(net, table, x, sigma, c)
    cond, uc = _cond(g)
    fn = S.sample_heun_sgm if which == "heun" else S.sample_lms_sgm
    with torch.no_grad():
        z = fn(den, torch.from_numpy(g["x_T"]), S.edm_sigmas(int(gs["steps"]), **EDM), cond, uc, float(g["cfg_scale"]))
    assert np.abs(z.numpy() - gs[which + "_final"]).max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["heun", "lms"])
def test_sgm_heun_lms_trajectories_vs_reference_golden(which):
    from cremage_b200.sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from cremage_b200.sgm.modules.diffusionmodules.sampling import HeunEDMSampler, LinearMultistepSampler
    from cremage_b200.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    g, sd = _weights()
    gs = gold("tiny_sgm_samplers.npz")
    model = OpenAIWrapper(_build_sgm_unet(S.TINY_SGM_UNET, sd))
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}).cuda()
    common = dict(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization", "params": EDM},
                  num_steps=int(gs["steps"]),
                  guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                 "params": {"scale": float(g["cfg_scale"])}})
    smp = HeunEDMSampler(**common) if which == "heun" else LinearMultistepSampler(order=int(gs["lms_order"]), **common)
    cond, uc = _cond(g, "cuda")
    x_T = torch.from_numpy(g["x_T"]).cuda()
    keep = x_T.clone()
    z = smp(lambda inp, sigma, c: den(model, inp, sigma, c), x_T, cond=cond, uc=uc)
    assert torch.equal(x_T, keep)
    want = torch.from_numpy(gs[which + "_final"])
    err = (z.cpu() - want).abs().max().item()
    print(f"[parity] sgm {which} {int(gs['steps'])} steps: max_abs_err={err:.4e} latent_absmax={want.abs().max():.2f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)


# ---------------------------------------------------------------------------------------------------------------
# DPMPP2SAncestralSampler, EulerAncestralSampler (noise drawn on every step), EulerEDMSampler with s_churn > 0:
# goldens from the unmodified reference with injected noise (oracle/make_golden_samplers2.py)
# ---------------------------------------------------------------------------------------------------------------
SGM2 = ["dpmpp2s_ancestral", "euler_ancestral", "euler_churn"]


@pytest.mark.parametrize("which", SGM2)
def test_oracle_ancestral_churn_trajectories_match_reference_golden(which):
    g, sd = _weights()
    gs = gold("tiny_sgm_samplers2.npz")
    table = S.legacy_ddpm_sigma_table(1000)
    net = lambda x, t, c: S.sgm_unet_forward(sd, S.TINY_SGM_UNET, x, t, c["crossattn"], c["vector"])
    den = lambda x, sigma, c: S.discrete_denoise(net, table, x, sigma, c)
    cond, uc = _cond(g)
    noise = list(torch.from_numpy(gs["noise"]))
    sig = S.edm_sigmas(int(gs["steps"]), **EDM)
    x_T, scale = torch.from_numpy(g["x_T"]), float(g["cfg_scale"])
    with torch.no_grad():
        if which == "dpmpp2s_ancestral":
            z = S.sample_dpmpp_2s_ancestral_sgm(den, x_T, sig, cond, uc, scale, noise)
        elif which == "euler_ancestral":
            z = S.sample_euler_ancestral_sgm(den, x_T, sig, cond, uc, scale, noise)
        else:
            z = S.sample_euler_edm_sgm(den, x_T, sig, cond, uc, scale, noise, s_churn=float(gs["s_churn"]))
    assert np.abs(z.numpy() - gs[which]).max() < 1e-3 * max(1.0, np.abs(gs[which]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("which", SGM2)
def test_sgm_ancestral_churn_trajectories_vs_reference_golden(which):
    from cremage_b200.sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from cremage_b200.sgm.modules.diffusionmodules import sampling as SM
    from cremage_b200.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    g, sd = _weights()
    gs = gold("tiny_sgm_samplers2.npz")
    model = OpenAIWrapper(_build_sgm_unet(S.TINY_SGM_UNET, sd))
    den = DiscreteDenoiser(scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
                           num_idx=1000,
                           discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}).cuda()
    common = dict(discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.EDMDiscretization", "params": EDM},
                  num_steps=int(gs["steps"]),
                  guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFG",
                                 "params": {"scale": float(g["cfg_scale"])}})
    noise = torch.from_numpy(gs["noise"]).cuda()
    it = iter(range(noise.shape[0]))
    draws = []

    def sampler_noise(x):
        draws.append(1)
        return noise[next(it)]

    cond, uc = _cond(g, "cuda")
    x_T = torch.from_numpy(g["x_T"]).cuda()
    denoiser = lambda inp, sigma, c: den(model, inp, sigma, c)
    if which == "euler_churn":
        smp = SM.EulerEDMSampler(s_churn=float(gs["s_churn"]), **common)
        real = torch.randn_like
        torch.randn_like = lambda x, *a, **k: sampler_noise(x)
        try:
            z = smp(denoiser, x_T, cond=cond, uc=uc)
        finally:
            torch.randn_like = real
    else:
        smp = (SM.DPMPP2SAncestralSampler if which == "dpmpp2s_ancestral" else SM.EulerAncestralSampler)(**common)
        smp.noise_sampler = sampler_noise
        z = smp(denoiser, x_T, cond=cond, uc=uc)
        assert len(draws) == int(gs["steps"])      # the reference draws on every step, also the last (RNG-stream parity)
    want = torch.from_numpy(gs[which])
    err = (z.cpu() - want).abs().max().item()
    print(f"[parity] sgm {which}: max_abs_err={err:.4e} latent_absmax={want.abs().max():.2f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)
