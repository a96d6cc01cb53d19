"""Drop-in seam: the reference's YAML `target:` strings and import paths resolve to the
cremage_b200 mirrors, with the reference's own hyper-parameters (configs/ldm/configs/stable-diffusion/
v1-inference.yaml:29-67, modules/sdxl/configs/inference/sd_xl_base.yaml:17-33 -- data, copied as literals)."""
import importlib
import sys

import pytest
import torch

V1_UNET = {"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel",
           "params": dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                          num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                          transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)}
V1_VAE = {"target": "ldm.models.autoencoder.AutoencoderKL",
          "params": dict(embed_dim=4, monitor="val/rec_loss",
                         ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                       ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0),
                         lossconfig={"target": "torch.nn.Identity"})}
SDXL_UNET = {"target": "sgm.modules.diffusionmodules.openaimodel.UNetModel",
             "params": dict(adm_in_channels=2816, num_classes="sequential", use_checkpoint=True, in_channels=4,
                            out_channels=4, model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2,
                            channel_mult=[1, 2, 4], num_head_channels=64, use_linear_in_transformer=True,
                            transformer_depth=[1, 2, 10], context_dim=2048,
                            spatial_transformer_attn_type="softmax-xformers")}


def test_instantiate_from_config_resolves_reference_targets():
    from cremage_b200.ldm.util import instantiate_from_config, resolve_target
    assert resolve_target(V1_UNET["target"]) == "cremage_b200.ldm.modules.diffusionmodules.openaimodel.UNetModel"
    assert resolve_target("torch.nn.Identity") == "torch.nn.Identity"
    assert resolve_target("ldm.modules.encoders.modules.FrozenCLIPEmbedder") == "ldm.modules.encoders.modules.FrozenCLIPEmbedder"
    with torch.device("meta"):
        unet = instantiate_from_config(V1_UNET)
        vae = instantiate_from_config(V1_VAE)
        xl = instantiate_from_config(SDXL_UNET)
    assert type(unet).__module__.startswith("cremage_b200.ldm") and sum(p.numel() for p in unet.parameters()) == 859520964
    assert type(vae).__module__.startswith("cremage_b200.ldm")
    assert type(xl).__module__.startswith("cremage_b200.sgm") and sum(p.numel() for p in xl.parameters()) == 2567463684
    # the names pinned by the reference's only hot-path test (test/ldm/ldm_instantiation_test.py:23-26)
    keys = unet.state_dict().keys()
    for k in ("input_blocks.1.0.in_layers.0.weight", "input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight",
              "middle_block.1.proj_out.weight", "out.2.bias"):
        assert k in keys


def test_dropin_install_aliases_reference_import_paths():
    from cremage_b200 import dropin
    before = {n: sys.modules.get(n) for n in dropin.ALIASES}
    done = dropin.install()
    try:
        assert set(done) == set(dropin.ALIASES)
        from ldm.models.diffusion.ddim import DDIMSampler                      # the reference's own import lines
        from ldm.models.diffusion.k_diffusion_samplers import EulerAncestralSampler, Dpmpp2mSampler, HeunSampler
        from k_diffusion.sampling import sample_euler_ancestral, sample_dpmpp_2m, get_sigmas_karras
        from sgm.modules.diffusionmodules.sampling import DPMPP2MSampler
        for obj in (DDIMSampler, EulerAncestralSampler, Dpmpp2mSampler, HeunSampler, sample_euler_ancestral,
                    sample_dpmpp_2m, get_sigmas_karras, DPMPP2MSampler):
            assert obj.__module__.startswith("cremage_b200.")
        mod = importlib.import_module("ldm.modules.diffusionmodules.openaimodel")
        assert mod.UNetModel.__module__.startswith("cremage_b200.")
    finally:
        dropin.uninstall()
    assert {n: sys.modules.get(n) for n in dropin.ALIASES} == before
    with pytest.raises(ValueError):
        dropin.install(only=["ldm.modules.encoders.modules"])


def test_reference_import_lines_resolve_after_install():
    """The reference's NON-mirrored code imports names from the aliased modules that the hot path never defines
    (ldm/models/diffusion/ddpm.py:38-41, cldm/cldm.py:12-24, cremage/utils/sampler_utils.py:6-18).  After install()
    every one of those lines must still import: mirrored names from cremage_b200, the rest from the reference's own
    module through the mirror's module-level __getattr__."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    ref_shim.install_lightning_stub()
    from cremage_b200 import dropin
    dropin.install()
    try:
        ns = {}
        # ldm/models/diffusion/ddpm.py:38-41
        exec("from ldm.models.autoencoder import VQModelInterface, IdentityFirstStage, AutoencoderKL", ns)
        exec("from ldm.modules.diffusionmodules.util import make_beta_schedule, extract_into_tensor, noise_like", ns)
        exec("from ldm.models.diffusion.ddim import DDIMSampler", ns)
        # cldm/cldm.py:12-24
        exec("from ldm.modules.diffusionmodules.util import (conv_nd, linear, zero_module, timestep_embedding)", ns)
        exec("from ldm.modules.attention import SpatialTransformer", ns)
        exec("from ldm.modules.diffusionmodules.openaimodel import UNetModel, TimestepEmbedSequential, ResBlock, "
             "Downsample, AttentionBlock", ns)
        # cremage/utils/sampler_utils.py:6-18
        for name in ("EulerSampler", "EulerAncestralSampler", "HeunSampler", "Dpm2Sampler", "Dpm2AncestralSampler",
                     "LmsSampler", "Dpmpp2sAncestralSampler", "DpmppSdeSampler", "Dpmpp2mSampler", "Dpmpp2mSdeSampler",
                     "Dpmpp3mSdeSampler"):
            exec(f"from ldm.models.diffusion.k_diffusion_samplers import {name}", ns)
            assert ns[name].__module__.startswith("cremage_b200."), name
        for mirrored in ("AutoencoderKL", "DDIMSampler", "UNetModel", "ResBlock", "SpatialTransformer",
                         "make_beta_schedule"):
            assert ns[mirrored].__module__.startswith("cremage_b200."), mirrored
        for fallback in ("VQModelInterface", "IdentityFirstStage", "noise_like", "AttentionBlock",
                         "timestep_embedding"):
            assert not ns[fallback].__module__.startswith("cremage_b200."), fallback
        assert tuple(ns["noise_like"]((2, 3), "cpu").shape) == (2, 3)          # the reference's own function runs
        import ldm.modules.diffusionmodules.util as u
        with pytest.raises(AttributeError):
            u.no_such_name_anywhere
    finally:
        dropin.uninstall()


def test_unmirrored_name_degrades_to_a_placeholder_when_the_reference_module_cannot_load(monkeypatch):
    """No reference tree (or its dependencies missing): the import line still succeeds, the name raises on USE."""
    from cremage_b200 import dropin

    def boom(name):
        raise ImportError("no reference here")
    monkeypatch.setattr(dropin, "_load_reference_module", boom)
    dropin.install(only=["ldm.models.autoencoder"])
    try:
        ns = {}
        exec("from ldm.models.autoencoder import VQModelInterface, AutoencoderKL", ns)
        assert ns["AutoencoderKL"].__module__.startswith("cremage_b200.")
        with pytest.raises(NotImplementedError, match="not part of the B200 hot path"):
            ns["VQModelInterface"]()
    finally:
        dropin.uninstall()


def test_subtree_cast_and_data_swap_invalidate_packs_and_graphs():
    """`blocks.half()` (UNetModel.convert_to_fp16) and `p.data = ...` bump no Parameter._version and never
    pass through the root's _apply; the process-wide pack epoch + the (address, dtype) fingerprint must catch both."""
    from cremage_b200 import engine
    from cremage_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    with torch.device("meta"):
        m = UNetModel(image_size=32, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[1],
                      num_res_blocks=1, channel_mult=[1], num_heads=2, use_spatial_transformer=True,
                      transformer_depth=1, context_dim=16, use_checkpoint=False, legacy=False)
    m = m.to_empty(device="cpu")
    e0 = engine.PACK_EPOCH
    m.convert_to_fp16()                                   # sub-tree _apply only
    assert engine.PACK_EPOCH > e0
    p = next(m.input_blocks.parameters())
    fp0 = engine.param_fingerprint([p])
    p.data = p.data.clone()                               # same version counter, new storage
    assert engine.param_fingerprint([p]) != fp0
    e1 = engine.PACK_EPOCH
    m.load_state_dict(m.state_dict())
    assert engine.PACK_EPOCH > e1


@pytest.mark.gpu
def test_convert_to_fp16_after_a_graphed_forward_repacks():
    """A captured graph must not replay pointers into packs freed by a sub-tree cast (engine.py)."""
    from oracle import sd_oracle as O
    from tests._models import build_unet, gold, tol
    g = gold("tiny_unet.npz")
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    unet = build_unet(O.TINY_UNET, sd)
    x, t, ctx = (torch.from_numpy(g[k]).cuda() for k in ("x", "t", "context"))
    y0 = unet(x, t, context=ctx)
    y0b = unet(x, t, context=ctx)                          # graph replay
    assert torch.equal(y0, y0b)
    unet.convert_to_fp16()                                 # frees the sub-modules' packs; root _apply never runs
    torch.cuda.empty_cache()
    junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(64)]   # reuse whatever was freed
    y1 = unet(x, t, context=ctx)
    del junk
    want = torch.from_numpy(g["out"])
    assert torch.isfinite(y1).all() and (y1.float().cpu() - want).abs().max().item() < tol(3e-2)
    with torch.no_grad():                                  # weight patch through .data: new storage, same version
        w = unet.out[2].weight
        w.data = (w.data.float() * 2.0).to(w.dtype)
        unet.out[2].bias.data = (unet.out[2].bias.data.float() * 2.0).to(w.dtype)
    y2 = unet(x, t, context=ctx)
    assert (y2.float() - 2.0 * y1.float()).abs().max().item() < tol(2e-2)


@pytest.mark.gpu
def test_weights_survive_half_and_device_round_trips():
    """The reference does load_state_dict -> .half() -> .to(device), and low_vram_shift moves the UNet to the CPU and back
    between phases (sd/image_generator.py:345,489,493; ldm/models/diffusion/ddpm.py:1460-1498)."""
    from oracle import sd_oracle as O
    from tests._models import build_unet, gold
    g = gold("tiny_unet.npz")
    sd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    unet = build_unet(O.TINY_UNET, sd)
    x, t, ctx = (torch.from_numpy(g[k]).cuda() for k in ("x", "t", "context"))
    y0 = unet(x, t, context=ctx)
    unet.half()
    y1 = unet(x.half(), t, context=ctx.half()).float()
    unet.cpu()
    with pytest.raises(RuntimeError):
        unet(x.cpu(), t.cpu(), context=ctx.cpu())          # no CPU fallback
    unet.cuda()
    y2 = unet(x.half(), t, context=ctx.half()).float()
    assert torch.equal(y1, y2)
    want = torch.from_numpy(g["out"])
    from tests._models import tol
    assert (y0.cpu() - want).abs().max().item() < tol(2e-2) and (y2.cpu() - want).abs().max().item() < tol(3e-2)
