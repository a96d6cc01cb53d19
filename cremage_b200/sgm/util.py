"""Mirror of the `sgm.util` helpers the SDXL sampling path uses (modules/sdxl/sgm/util.py)."""
import importlib

import torch


def append_dims(x, target_dims):
    """Appends dimensions to the end of a tensor until it has target_dims dimensions."""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) and not isinstance(d, (dict, str)) else d


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    """`{"target": "pkg.mod.Class", "params": {...}}` -> object (sgm/util.py); already-built objects pass through.
    Reference `sgm....` targets resolve to their cremage_b200 mirrors."""
    if not isinstance(config, dict):
        return config
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    target = config["target"]
    if target.startswith("sgm."):
        target = "cremage_b200." + target
    return get_obj_from_str(target)(**config.get("params", dict()))
