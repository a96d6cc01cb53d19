"""Drop-in installer: make the reference's own dotted names resolve to the B200 modules.

    import cremage_b200.dropin as dropin
    dropin.install()          # before the reference builds its model

After `install()` the module names the reference's YAML `target:` strings and import statements use for the hot path --
`ldm.modules.diffusionmodules.openaimodel`, `ldm.modules.diffusionmodules.model`, `ldm.modules.attention`,
`ldm.models.autoencoder`, `ldm.models.diffusion.ddim`, `ldm.models.diffusion.k_diffusion_samplers`,
`ldm.models.diffusion.ldm_wrapper_for_k_diffusion`, `k_diffusion.sampling`, `k_diffusion.external`,
`sgm.modules.diffusionmodules.{openaimodel,sampling,guiders,denoiser,denoiser_scaling,discretizer,wrappers}`,
`sgm.modules.attention`, `sgm.models.autoencoder` -- are entries of `sys.modules` that point at the cremage_b200 mirrors,
so `instantiate_from_config` (ldm/util.py:81-96) and `from ldm.models.diffusion.ddim import DDIMSampler`
(cremage/utils/sampler_utils.py) pick them up with no edit to the reference tree.  Names that are NOT mirrored (text
encoders, ControlNet, the Lightning `LatentDiffusion`) keep resolving to the reference's own modules.

`install(only=...)` restricts the aliasing; `uninstall()` restores the previous entries.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, Iterable, Optional

ALIASES = [
    "ldm.modules.diffusionmodules.openaimodel",
    "ldm.modules.diffusionmodules.model",
    "ldm.modules.diffusionmodules.util",
    "ldm.modules.attention",
    "ldm.models.autoencoder",
    "ldm.models.diffusion.ddim",
    "ldm.models.diffusion.k_diffusion_samplers",
    "ldm.models.diffusion.ldm_wrapper_for_k_diffusion",
    "k_diffusion.sampling",
    "k_diffusion.external",
    "sgm.modules.diffusionmodules.openaimodel",
    "sgm.modules.diffusionmodules.model",
    "sgm.modules.diffusionmodules.sampling",
    "sgm.modules.diffusionmodules.sampling_utils",
    "sgm.modules.diffusionmodules.guiders",
    "sgm.modules.diffusionmodules.denoiser",
    "sgm.modules.diffusionmodules.denoiser_scaling",
    "sgm.modules.diffusionmodules.discretizer",
    "sgm.modules.diffusionmodules.wrappers",
    "sgm.modules.attention",
    "sgm.models.autoencoder",
]
_SAVED: Dict[str, Optional[object]] = {}


def install(only: Optional[Iterable[str]] = None) -> Dict[str, str]:
    """Alias the reference module names to the cremage_b200 mirrors; returns {reference name: mirror name}."""
    done = {}
    for name in (ALIASES if only is None else list(only)):
        if name not in ALIASES:
            raise ValueError(f"cremage_b200 has no mirror of '{name}'")
        mirror = importlib.import_module("cremage_b200." + name)
        if name not in _SAVED:
            _SAVED[name] = sys.modules.get(name)
        sys.modules[name] = mirror
        parent, _, leaf = name.rpartition(".")
        if parent in sys.modules and not parent.startswith("cremage_b200"):
            setattr(sys.modules[parent], leaf, mirror)    # `import ldm.modules.attention as a` style access
        done[name] = mirror.__name__
    return done


def uninstall() -> None:
    for name, prev in list(_SAVED.items()):
        if prev is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = prev
        del _SAVED[name]
