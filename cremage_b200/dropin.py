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

The reference's non-mirrored code also imports names from the aliased modules that the hot path never uses
(`noise_like`, `VQModelInterface`, `IdentityFirstStage` in ldm/models/diffusion/ddpm.py:39-40; `AttentionBlock`,
`conv_nd`, `linear` in cldm/cldm.py:12-22).  Every installed mirror therefore gets a module-level `__getattr__`
(PEP 562) that resolves a missing name from the reference's OWN module of the same dotted name, loaded lazily from
the reference tree on `sys.path` under a private name; if that module cannot be imported, the name resolves to a
placeholder class that raises `NotImplementedError` on USE, so the import line itself never fails.

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
_REFERENCE: Dict[str, object] = {}     # dotted name -> the reference's own module, loaded on demand by _fallback


def _load_reference_module(name: str):
    """The reference's module `name`, executed from its source file WITHOUT touching sys.modules[name] (which is the
    mirror): found through the parent package's __path__ (the parent is never aliased), or through sys.path."""
    if name in _REFERENCE:
        return _REFERENCE[name]
    prev = _SAVED.get(name)
    if prev is not None and not getattr(prev, "__name__", "").startswith("cremage_b200"):
        _REFERENCE[name] = prev          # it was already imported before install()
        return prev
    import importlib.machinery
    import importlib.util
    parent, _, leaf = name.rpartition(".")
    search = None
    if parent:
        pkg = sys.modules.get(parent)
        if pkg is None or getattr(pkg, "__name__", "").startswith("cremage_b200"):
            pkg = importlib.import_module(parent)
        search = list(getattr(pkg, "__path__", []))
    spec = importlib.machinery.PathFinder.find_spec(leaf, search)
    if spec is None or spec.origin is None or "cremage_b200" in spec.origin:
        raise ImportError(f"the reference's own '{name}' is not on sys.path")
    spec = importlib.util.spec_from_file_location(name, spec.origin)
    mod = importlib.util.module_from_spec(spec)
    _REFERENCE[name] = mod               # registered first: a cycle through the mirror's __getattr__ terminates
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        _REFERENCE.pop(name, None)
        raise
    return mod


def _placeholder(modname: str, attr: str, why: BaseException):
    msg = (f"{modname}.{attr} is not part of the B200 hot path; cremage_b200 does not mirror it and the reference's own "
           f"module could not be imported ({type(why).__name__}: {why})")

    class _Unavailable:
        def __init__(self, *a, **k):
            raise NotImplementedError(msg)

    _Unavailable.__name__ = _Unavailable.__qualname__ = attr
    _Unavailable.__doc__ = msg
    return _Unavailable


def _fallback(modname: str):
    def __getattr__(attr: str):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        try:
            ref = _load_reference_module(modname)
        except Exception as e:                      # reference tree absent / its dependencies missing
            return _placeholder(modname, attr, e)
        try:
            return getattr(ref, attr)
        except AttributeError:
            raise AttributeError(f"neither cremage_b200.{modname} nor the reference's {modname} defines '{attr}'") from None
    return __getattr__


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
        mirror.__dict__["__getattr__"] = _fallback(name)
        parent, _, leaf = name.rpartition(".")
        if parent in sys.modules and not parent.startswith("cremage_b200"):
            setattr(sys.modules[parent], leaf, mirror)    # `import ldm.modules.attention as a` style access
        done[name] = mirror.__name__
    return done


def uninstall() -> None:
    for name in list(_SAVED):
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__name__", "").startswith("cremage_b200"):
            m.__dict__.pop("__getattr__", None)
    _REFERENCE.clear()
    for name, prev in list(_SAVED.items()):
        if prev is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = prev
        del _SAVED[name]
