"""Mirror of the plugin seam of the reference: `ldm.util.instantiate_from_config` / `get_obj_from_str`
(HowToSD/cremage modules/ldm/util.py:81-96).  Reference `ldm.` / `sgm.` / `k_diffusion.` targets resolve to their
cremage_b200 mirrors, so the reference's own YAML trees instantiate the B200 modules unchanged."""
import importlib

MIRRORED_ROOTS = ("ldm.", "sgm.", "k_diffusion.")


def resolve_target(target: str) -> str:
    """`ldm.modules.diffusionmodules.openaimodel.UNetModel` -> `cremage_b200.ldm....UNetModel` when a mirror exists."""
    if target.startswith(MIRRORED_ROOTS):
        module, cls = ("cremage_b200." + target).rsplit(".", 1)
        try:
            if hasattr(importlib.import_module(module), cls):
                return module + "." + cls
        except ImportError:
            pass
    return target


def get_obj_from_str(string, reload=False):
    module, cls = resolve_target(string).rsplit(".", 1)
    if reload:
        module_imp = importlib.import_module(module)
        importlib.reload(module_imp)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config == '__is_first_stage__':
            return None
        elif config == "__is_unconditional__":
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))
