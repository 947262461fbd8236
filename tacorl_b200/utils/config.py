"""Hydra-compatible instantiate for the `_target_` configs of the reference
(config/networks/**, config/module/*.yaml).  Uses hydra when it is installed, else a minimal
re-implementation with the same `_target_` / `_recursive_` semantics.  Reference class paths
(`tacorl.…`) are transparently remapped to their B200 mirrors (`tacorl_b200.…`)."""
import copy
import importlib

REMAP_PREFIX = ("tacorl.", "tacorl_b200.")


def remap_target(path: str) -> str:
    if path.startswith(REMAP_PREFIX[0]) and not path.startswith(REMAP_PREFIX[1]):
        return REMAP_PREFIX[1] + path[len(REMAP_PREFIX[0]):]
    return path


def locate(path: str):
    mod, name = remap_target(path).rsplit(".", 1)
    return getattr(importlib.import_module(mod), name)


def to_container(cfg):
    """OmegaConf.to_container(resolve=True) when given an OmegaConf node, deepcopy otherwise."""
    try:
        from omegaconf import OmegaConf  # type: ignore
        if OmegaConf.is_config(cfg):
            return OmegaConf.to_container(cfg, resolve=True)
    except Exception:
        pass
    return copy.deepcopy(cfg)


def instantiate(cfg, *args, **kwargs):
    """hydra.utils.instantiate semantics for dict configs (None / {} -> None)."""
    if cfg is None:
        return None
    cfg = to_container(cfg)
    if isinstance(cfg, dict) and len(cfg) == 0:
        return None
    cfg = dict(cfg)
    cfg.update(kwargs)
    recursive = cfg.pop("_recursive_", True)
    cfg.pop("_convert_", None)
    target = cfg.pop("_target_")
    if recursive:
        for k, v in list(cfg.items()):
            if isinstance(v, dict) and "_target_" in v:
                cfg[k] = instantiate(v)
    cls = locate(target) if isinstance(target, str) else target
    return cls(*args, **cfg)
