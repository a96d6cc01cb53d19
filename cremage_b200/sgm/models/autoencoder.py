"""sgm first stage on the decode side (modules/sdxl/sgm/models/autoencoder.py:490-505, AutoencodingEngineLegacy.decode:
post_quant_conv then decoder; AutoencoderKL / AutoencoderKLInferenceWrapper are the same decode path)."""
from ...ldm.models.autoencoder import AutoencoderKL as _AutoencoderKL


class AutoencodingEngineLegacy(_AutoencoderKL):
    def __init__(self, embed_dim: int, **kwargs):
        ddconfig = kwargs.pop("ddconfig")
        kwargs.pop("ckpt_path", None)
        kwargs.pop("ckpt_engine", None)
        super().__init__(ddconfig=ddconfig, lossconfig=kwargs.get("lossconfig"), embed_dim=embed_dim)


class AutoencoderKL(AutoencodingEngineLegacy):
    pass


class AutoencoderKLInferenceWrapper(AutoencoderKL):
    pass
