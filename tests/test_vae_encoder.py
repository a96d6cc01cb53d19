"""VAE encoder: Encoder + quant_conv + DiagonalGaussianDistribution.  CPU: oracle restatement against the
golden from the reference's own Encoder (oracle/make_golden_encoder.py) and the state-dict layout.  GPU: the
cremage_b200 AutoencoderKL.encode / LatentDiffusion.get_first_stage_encoding against the same golden, and an
encode -> decode round trip through both directions of the first stage."""
import numpy as np
import pytest
import torch

from oracle import sd_oracle as O
from tests._models import tol  # noqa: E402
from tests._models import ENCODER_SEED, build_ldm, build_vae, gold, vae_kwargs


def _weights():
    g = gold("tiny_vae_encoder.npz")
    sd = O.make_weights(O.encoder_param_shapes(O.TINY_VAE), seed=ENCODER_SEED)
    assert abs(O.weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6
    return g, sd


def test_oracle_encoder_matches_reference_golden():
    g, sd = _weights()
    with torch.no_grad():
        m = O.vae_encode_moments(sd, O.TINY_VAE, torch.from_numpy(g["x"]))
        s = O.gaussian_sample(m, torch.from_numpy(g["noise"]))
    assert np.abs(m.numpy() - g["moments"]).max() < 2e-5
    assert np.abs(s.numpy() - g["sample"]).max() < 2e-5


def test_state_dict_layout_has_checkpoint_keys():
    from cremage_b200.ldm.models.autoencoder import AutoencoderKL
    for cfg in (O.TINY_VAE, O.SD15_VAE):
        with torch.device("meta"):
            m = AutoencoderKL(**vae_kwargs(cfg))
        want = dict(O.encoder_param_shapes(cfg))
        want.update(O.decoder_param_shapes(cfg))
        have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert have == want
    assert "encoder.down.0.downsample.conv.weight" in have and "quant_conv.weight" in have


def test_asymmetric_stride2_tap_table():
    """Downsample = pad (0,1,0,1) + conv stride 2 padding 0: input row 2*oy + kh, so kh = 2 reads parity plane 0 one
    row further down (model.py:79-84)."""
    from cremage_b200 import ops
    dw, dh, dn = ops.taps_3x3_stride2_asym(3)
    assert dh == [0, 0, 0, 0, 0, 0, 1, 1, 1] and dw == [0, 0, 1] * 3
    assert dn == [0, 3, 0, 6, 9, 6, 0, 3, 0]


@pytest.mark.gpu
def test_encode_vs_reference_golden():
    from cremage_b200.ldm.models.autoencoder import DiagonalGaussianDistribution
    g, _ = _weights()
    vsd = O.make_weights(O.decoder_param_shapes(O.TINY_VAE), seed=200)
    vae = build_vae(O.TINY_VAE, vsd)
    post = vae.encode(torch.from_numpy(g["x"]).cuda())
    assert isinstance(post, DiagonalGaussianDistribution)
    want = torch.from_numpy(g["moments"])
    err = (post.parameters.cpu() - want).abs().max().item()
    print(f"[parity] VAE encoder moments: max_abs_err={err:.4e} ref_absmax={want.abs().max():.3f}")
    assert err <= tol(2e-2) * max(want.abs().max().item(), 1.0)
    s = post.sample(noise=torch.from_numpy(g["noise"]).cuda())
    assert (s.cpu() - torch.from_numpy(g["sample"])).abs().max().item() <= 3e-2 * max(np.abs(g["sample"]).max(), 1.0)
    assert (post.mode().cpu() - torch.from_numpy(g["mode"])).abs().max().item() <= tol(2e-2) * max(np.abs(g["mode"]).max(), 1.0)
    assert torch.allclose(post.mean, post.mode()) and (post.std > 0).all()


@pytest.mark.gpu
def test_first_stage_encoding_scale_and_round_trip_shapes():
    g, _ = _weights()
    usd = O.make_weights(O.unet_param_shapes(O.TINY_UNET), seed=100)
    vsd = O.make_weights(O.decoder_param_shapes(O.TINY_VAE), seed=200)
    ldm = build_ldm(O.TINY_UNET, usd, O.TINY_VAE, vsd)
    x = torch.from_numpy(g["x"]).cuda()
    post = ldm.encode_first_stage(x)
    noise = torch.from_numpy(g["noise"]).cuda()
    z = ldm.get_first_stage_encoding(post, noise=noise)
    assert torch.allclose(z, ldm.scale_factor * post.sample(noise=noise), rtol=1e-6, atol=1e-6)
    assert tuple(z.shape) == (2, 4, 16, 24)
    img = ldm.decode_first_stage(z)
    assert tuple(img.shape) == tuple(x.shape) and torch.isfinite(img).all()
