"""Mirror of sgm/modules/diffusionmodules/guiders.py: VanillaCFG (:24-65), IdentityGuider. The guidance mix
`x_u + scale * (x_c - x_u)` runs in the cfg_mix kernel."""
from typing import Dict, Tuple

import torch

from .... import ops


class Guider:
    def __call__(self, x: torch.Tensor, sigma: float) -> torch.Tensor:
        raise NotImplementedError

    def prepare_inputs(self, x: torch.Tensor, s: float, c: Dict, uc: Dict) -> Tuple[torch.Tensor, float, Dict]:
        raise NotImplementedError


class VanillaCFG(Guider):
    def __init__(self, scale: float):
        self.scale = scale
        self._cat = {}   # key -> (uc tensor, c tensor, versions, concatenation): the same conditioning every step

    def __call__(self, x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        x_u, x_c = x.chunk(2)
        return ops.cfg_mix(x_u.float().contiguous(), x_c.float().contiguous(), float(self.scale)).to(x.dtype)

    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"]:
                # the conditioning tensors are the same objects on every sampler step: hand the UNet the SAME doubled
                # tensor each time (values as torch.cat((uc, c), 0) of the reference) so its per-context K/V cache holds
                hit = self._cat.get(k)
                if (hit is None or hit[0] is not uc[k] or hit[1] is not c[k]
                        or hit[2] != (uc[k]._version, c[k]._version)):
                    hit = (uc[k], c[k], (uc[k]._version, c[k]._version), torch.cat((uc[k], c[k]), 0))
                    self._cat[k] = hit
                c_out[k] = hit[3]
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out


class IdentityGuider(Guider):
    def __call__(self, x: torch.Tensor, sigma: float) -> torch.Tensor:
        return x

    def prepare_inputs(self, x: torch.Tensor, s: float, c: Dict, uc: Dict) -> Tuple[torch.Tensor, float, Dict]:
        c_out = dict()
        for k in c:
            c_out[k] = c[k]
        return x, s, c_out
