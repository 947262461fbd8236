"""Noise draws of the hot path, in the reference's torch-call pattern.

By default each helper issues the same torch RNG call the reference issues (same shape, same
order), so a seeded run on the same device type consumes the generator identically.  Tests and
multi-GPU equivalence runs can instead replay an explicit tape of pre-drawn tensors.
"""
import collections
import contextlib

import torch
from torch.distributions.utils import _standard_normal

_TAPE = None


@contextlib.contextmanager
def noise_tape(tensors):
    """Replay `tensors` (in consumption order) instead of drawing from the torch generator."""
    global _TAPE
    prev = _TAPE
    _TAPE = collections.deque(tensors)
    try:
        yield _TAPE
    finally:
        _TAPE = prev


def _pop(shape, device):
    t = _TAPE.popleft()
    assert tuple(t.shape) == tuple(shape), f"noise tape shape {tuple(t.shape)} != requested {tuple(shape)}"
    return t.to(device=device, dtype=torch.float32).contiguous()


def standard_normal(shape, device):
    """eps of Normal.rsample (torch.distributions.utils._standard_normal)."""
    if _TAPE is not None:
        return _pop(shape, device)
    return _standard_normal(torch.Size(shape), dtype=torch.float32, device=torch.device(device))


def normal_noise(shape, device):
    """N(0,1) underlying torch.normal(mean, std) in Normal.sample."""
    if _TAPE is not None:
        return _pop(shape, device)
    return torch.empty(shape, device=device, dtype=torch.float32).normal_()


def rand(shape, device):
    if _TAPE is not None:
        return _pop(shape, device)
    return torch.rand(shape, device=device)


def uniform(shape, lo, hi, device):
    if _TAPE is not None:
        return _pop(shape, device)
    return torch.empty(shape, device=device, dtype=torch.float32).uniform_(lo, hi)


def randint(shape, lo, hi, device):
    """torch.randint(lo, hi, shape) as drawn by RandomShiftsAug (utils/transforms.py:287-289), as int32."""
    if _TAPE is not None:
        t = _TAPE.popleft()
        assert tuple(t.shape) == tuple(shape), f"noise tape shape {tuple(t.shape)} != requested {tuple(shape)}"
        return t.to(device=device, dtype=torch.int32).contiguous()
    return torch.randint(lo, hi, shape, device=device).to(torch.int32)


def dropout_mask(shape, p, device, training=True):
    """Pre-scaled keep mask (keep / (1-p)) of F.dropout, or None when dropout is inactive."""
    if not training or p == 0.0:
        return None
    if _TAPE is not None:
        return _pop(shape, device)
    return torch.nn.functional.dropout(torch.ones(shape, device=device), p, True)
