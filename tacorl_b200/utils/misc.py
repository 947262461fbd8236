"""Small helpers mirrored from /root/reference/src/tacorl/utils/{misc,networks}.py."""
import importlib

import torch


def load_class(name):                                            # utils/misc.py:262-265
    from .config import remap_target
    mod, cls = remap_target(name).rsplit(".", 1)
    return getattr(importlib.import_module(mod), cls)


def expand_array(obs, n_samples, reshape=True):                  # utils/misc.py:132-138
    e = obs.expand(n_samples, *obs.shape)
    return e.reshape(-1, *obs.shape[1:]) if reshape else e


def expand_obs(obs, n_samples, reshape=True):                    # utils/misc.py:141-153
    if isinstance(obs, dict):
        return {k: expand_obs(v, n_samples, reshape) for k, v in obs.items()}
    return expand_array(obs, n_samples, reshape)


def get_batch_size_from_input(inp):                              # utils/networks.py:18-29
    while isinstance(inp, dict):
        inp = list(inp.values())[0]
    shape = inp.shape
    return 1 if len(shape) in (1, 3) else shape[0]


def set_parameter_requires_grad(model, requires_grad):           # utils/networks.py:59-62
    for _, child in model.named_children():
        for p in child.parameters():
            p.requires_grad = requires_grad
