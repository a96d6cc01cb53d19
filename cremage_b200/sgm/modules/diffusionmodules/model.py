"""sgm VAE decoder (modules/sdxl/sgm/modules/diffusionmodules/model.py:614-763): the same graph as the ldm Decoder
(AttnBlock via SDPA computes the same function) -- shared implementation."""
from ....ldm.modules.diffusionmodules.model import AttnBlock, Decoder, ResnetBlock, Upsample, make_attn  # noqa: F401
