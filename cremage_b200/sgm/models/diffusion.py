"""Inference-only mirror of `sgm.models.diffusion.DiffusionEngine` (modules/sdxl/sgm/models/diffusion.py:19-137):
`model` = OpenAIWrapper(UNetModel), `denoiser`, `first_stage_model`, `scale_factor`, `decode_first_stage` (:119-137,
z / scale_factor then decode). No conditioner (context / vector are inputs), no Lightning.

`disable_first_stage_autocast` (True in sd_xl_base.yaml:5 / sd_xl_refiner.yaml): the reference then decodes OUTSIDE
fp16 autocast, i.e. in fp32, because the SDXL VAE's activations exceed fp16's 65504.  The kernels store activations in
16 bits, so the flag routes the first stage to the bf16 build of the library (fp32's exponent range, fp32 accumulation
and normalisation statistics) instead of the process default fp16; tests/test_gpu_config_parity.py holds the
large-magnitude fixture (|h| up to 2e5) this must pass."""
import torch
import torch.nn as nn

from ... import ops
from ..modules.diffusionmodules.wrappers import OpenAIWrapper
from ..util import instantiate_from_config


class DiffusionEngine(nn.Module):
    def __init__(self, network_config, denoiser_config, first_stage_config=None, scale_factor: float = 1.0,
                 disable_first_stage_autocast: bool = False, **ignored):
        super().__init__()
        self.model = OpenAIWrapper(instantiate_from_config(network_config))
        self.denoiser = instantiate_from_config(denoiser_config)
        self.first_stage_model = instantiate_from_config(first_stage_config) if first_stage_config is not None else None
        self.scale_factor = scale_factor
        self.disable_first_stage_autocast = disable_first_stage_autocast

    @torch.no_grad()
    def decode_first_stage(self, z, to_uint8: bool = False):
        if self.disable_first_stage_autocast:
            with ops.precision("bf16"):
                return self.first_stage_model.decode_first_stage(z, self.scale_factor, to_uint8=to_uint8)
        return self.first_stage_model.decode_first_stage(z, self.scale_factor, to_uint8=to_uint8)
