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

    def __call__(self, x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        x_u, x_c = x.chunk(2)
        return ops.cfg_mix(x_u.float().contiguous(), x_c.float().contiguous(), float(self.scale)).to(x.dtype)

    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"]:
                c_out[k] = torch.cat((uc[k], c[k]), 0)
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
