"""Checkpoint discovery mirrored from /root/reference/src/tacorl/utils/networks.py:90-142
(run dir -> last.ckpt / *_epoch_NN_* -> .hydra/config.yaml -> module class -> load_from_checkpoint)."""
from pathlib import Path
from typing import Union

import torch
import yaml

from .misc import load_class


def get_checkpoint_i_from_dir(dir, i: int = -1):
    ckpt_paths = list(Path(dir).rglob("*.ckpt"))
    if i == -1:
        for p in ckpt_paths:
            if p.stem == "last":
                return p
    for p in ckpt_paths:
        words = str(p).split("_")
        for k, w in enumerate(words):
            if w == "epoch" and k + 1 < len(words) and words[k + 1].isdigit() and int(words[k + 1]) == i:
                return p
    ckpt_paths = sorted(ckpt_paths, key=lambda f: f.stat().st_mtime)
    return ckpt_paths[i]


def get_config_from_dir(dir):
    config_yaml = list(Path(dir).rglob("*config.yaml"))[0]
    try:
        from omegaconf import OmegaConf  # type: ignore
        return OmegaConf.to_container(OmegaConf.load(config_yaml), resolve=True)
    except ImportError:
        with open(config_yaml) as f:
            return yaml.safe_load(f)


def load_state_into(module, ckpt_path):
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    module.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt, strict=True)
    return module


def load_pl_module_from_checkpoint(filepath: Union[Path, str], epoch: int = -1, overwrite_cfg: dict = {}):
    filepath = Path(filepath).expanduser()
    if filepath.is_dir():
        filedir, ckpt_path = filepath, get_checkpoint_i_from_dir(filepath, epoch)
    elif filepath.is_file():
        assert filepath.suffix == ".ckpt", "File must have .ckpt extension"
        ckpt_path, filedir = filepath, filepath.parents[0]
    else:
        raise ValueError(f"not valid file path: {str(filepath)}")
    config = get_config_from_dir(filedir)
    module_cfg = dict(config["module"])
    class_name = module_cfg.pop("_target_")
    module_cfg.pop("_recursive_", None)
    module_class = load_class(class_name)
    load_cfg = {**module_cfg, **overwrite_cfg}
    if hasattr(module_class, "load_from_checkpoint") and getattr(module_class, "_is_real_lightning", False):
        return module_class.load_from_checkpoint(ckpt_path, **load_cfg)
    return load_state_into(module_class(**load_cfg), ckpt_path)
