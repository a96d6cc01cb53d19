"""Mirror of sgm/modules/diffusionmodules/denoiser.py: Denoiser.forward (:23-39), DiscreteDenoiser (:42-75) -- sigma is
quantised to the nearest of `num_idx` table entries and c_noise becomes an INTEGER table index (:61-75; index math
bit-exact), then  network(input * c_in, c_noise, cond) * c_out + input * c_skip  with the two scalings applied by the
axpby kernel."""
from typing import Dict, Union

import torch
import torch.nn as nn

from .... import ops
from ...util import append_dims, instantiate_from_config


def _axpby_rows(x, a, y, b):
    """a[i] * x[i] + b[i] * y[i] per batch row (one launch when the scalars agree across the batch)."""
    al = a.reshape(-1).tolist()
    bl = b.reshape(-1).tolist() if b is not None else None
    if all(v == al[0] for v in al) and (bl is None or all(v == bl[0] for v in bl)):
        return ops.axpby(x, al[0], y, 0.0 if bl is None else bl[0])
    return torch.cat([ops.axpby(x[i:i + 1].contiguous(), al[i], None if y is None else y[i:i + 1].contiguous(),
                                0.0 if bl is None else bl[i]) for i in range(x.shape[0])])


class Denoiser(nn.Module):
    def __init__(self, scaling_config: Dict):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)

    def possibly_quantize_sigma(self, sigma: torch.Tensor) -> torch.Tensor:
        return sigma

    def possibly_quantize_c_noise(self, c_noise: torch.Tensor) -> torch.Tensor:
        return c_noise

    def forward(self, network: nn.Module, input: torch.Tensor, sigma: torch.Tensor, cond: Dict,
                **additional_model_inputs) -> torch.Tensor:
        if not input.is_cuda:
            raise RuntimeError("cremage_b200 has no CPU path: move the latents to a CUDA device")
        sigma = self.possibly_quantize_sigma(sigma)
        sigma_shape = sigma.shape
        sigma = append_dims(sigma, input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(sigma_shape))
        x = input.float().contiguous()
        net = network(_axpby_rows(x, c_in, None, None).to(input.dtype), c_noise, cond, **additional_model_inputs)
        return _axpby_rows(net.float().contiguous(), c_out, x, c_skip).to(input.dtype)


class DiscreteDenoiser(Denoiser):
    def __init__(self, scaling_config: Dict, num_idx: int, discretization_config: Dict, do_append_zero: bool = False,
                 quantize_c_noise: bool = True, flip: bool = True):
        super().__init__(scaling_config)
        self.discretization = instantiate_from_config(discretization_config)
        sigmas = self.discretization(num_idx, do_append_zero=do_append_zero, flip=flip)
        self.register_buffer("sigmas", sigmas)
        self.quantize_c_noise = quantize_c_noise
        self.num_idx = num_idx

    def sigma_to_idx(self, sigma: torch.Tensor) -> torch.Tensor:
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx: Union[torch.Tensor, int]) -> torch.Tensor:
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma: torch.Tensor) -> torch.Tensor:
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise: torch.Tensor) -> torch.Tensor:
        if self.quantize_c_noise:
            return self.sigma_to_idx(c_noise)
        return c_noise
