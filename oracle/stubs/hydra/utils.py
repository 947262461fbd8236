import copy
import importlib
import os


def _locate(path):
    mod, name = path.rsplit(".", 1)
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg, *args, **kwargs):
    """Minimal `hydra.utils.instantiate`: `_target_` + kwargs, `_recursive_` aware."""
    if cfg is None or (isinstance(cfg, dict) and len(cfg) == 0):
        return None
    cfg = dict(copy.deepcopy(cfg))
    cfg.update(kwargs)
    recursive = cfg.pop("_recursive_", True)
    cfg.pop("_convert_", None)
    target = cfg.pop("_target_")
    if recursive:
        for k, v in list(cfg.items()):
            if isinstance(v, dict) and "_target_" in v:
                cfg[k] = instantiate(v)
    cls = _locate(target) if isinstance(target, str) else target
    return cls(*args, **cfg)


def get_original_cwd():
    return os.getcwd()
