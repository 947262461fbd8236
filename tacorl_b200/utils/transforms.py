"""Device-side mirrors of the reference's training transforms and window collation
(/root/reference/src/tacorl/utils/transforms.py:87-101, 265-330; datamodule/dataset/play_dataset.py:115-169, 282-330;
config/datamodule/transform_manager/transforms/rl_train.yaml): the per-sample work of the DataLoader workers, run as
streaming kernels over uint8 frames that already live in HBM.

`FramePipeline` is the fused entry: window gather + pad_sequence + RandomShiftsAug in one kernel (uint8 out, which the
encoder's first kernel scales / normalises itself), plus the optional colour jitter (float32 out, already normalised)."""
import numpy as np
import torch

from .. import _lib as L
from . import rng

OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_HUE, OP_NONE = 0, 1, 2, 3, 0xF


def window_gather(store, start, window, T, shift=None, pad=0):
    """store: uint8 (F, C, H, W) on the device; start / window: int32 (B,); shift: int32 (B, T, 2) in [0, 2 pad] or None.
    Returns uint8 (B, T, C, H, W)."""
    assert store.dtype == torch.uint8 and store.is_cuda and store.is_contiguous()
    F, C, H, W = store.shape
    B = start.numel()
    start = start.to(device=store.device, dtype=torch.int32).contiguous()
    window = None if window is None else window.to(device=store.device, dtype=torch.int32).contiguous()
    shift = None if shift is None else shift.to(device=store.device, dtype=torch.int32).contiguous()
    out = torch.empty(B, T, C, H, W, device=store.device, dtype=torch.uint8)
    L.call("tacorl_window_gather_u8", L.ptr_any(store), F, C, H, W, L.ptr_any(start), L.ptr_any(window), L.ptr_any(shift),
           int(pad), B, T, L.ptr_any(out), L.stream())
    return out


def actions_gather_pad(store, start, window, T, zero_pad=True):
    """store: float32 (F, A); returns (B, T, A) padded like PlayDataset.pad_sequence."""
    assert store.dtype == torch.float32 and store.is_cuda and store.is_contiguous()
    F, A = store.shape
    B = start.numel()
    start = start.to(device=store.device, dtype=torch.int32).contiguous()
    window = None if window is None else window.to(device=store.device, dtype=torch.int32).contiguous()
    out = torch.empty(B, T, A, device=store.device, dtype=torch.float32)
    L.call("tacorl_actions_gather_pad", L.ptr(store), F, A, L.ptr_any(start), L.ptr_any(window), B, T, int(zero_pad),
           L.ptr(out), L.stream())
    return out


def pack_order(ops):
    """ColorJitter's op order (ids as in torchvision: 0 brightness, 1 contrast, 2 saturation, 3 hue) -> kernel code."""
    ops = [o for o in ops if o != OP_SATURATION][:3]
    ops = ops + [OP_NONE] * (3 - len(ops))
    return ops[0] | (ops[1] << 4) | (ops[2] << 8)


def color_jitter(x, order=None, factors=None, mean=0.5, std=0.5):
    """x: uint8 (N, 3, H, W) on the device -> float32 normalised (N, 3, H, W).
    order: int32 (N,) packed op order (pack_order) or None; factors: float32 (N, 3) = brightness, contrast, hue."""
    assert x.dtype == torch.uint8 and x.is_cuda and x.is_contiguous() and x.shape[1] == 3
    N, _, H, W = x.shape
    out = torch.empty(N, 3, H, W, device=x.device, dtype=torch.float32)
    ws = torch.empty(max(N, 1), device=x.device, dtype=torch.float32)
    if order is not None:
        order = order.to(device=x.device, dtype=torch.int32).contiguous()
        factors = factors.to(device=x.device, dtype=torch.float32).contiguous()
    L.call("tacorl_color_jitter_u8", L.ptr_any(x), N, H, W, L.ptr_any(order), L.ptr(factors), float(mean), float(std),
           L.ptr(ws), L.ptr(out), L.stream())
    return out


def draw_color_jitter_params(n, brightness, contrast, hue, prob=1.0):
    """Per-image parameters in the reference's draw order: ColorTransform.apply_transform draws np.random.rand() < prob
    (utils/transforms.py:311-313), then torchvision ColorJitter.get_params draws torch.randperm(4) and one uniform_ per
    enabled op on the CPU generator.  Returns (order int32 (n,), factors float32 (n, 3))."""
    order = torch.empty(n, dtype=torch.int32)
    factors = torch.ones(n, 3)
    factors[:, 2] = 0.0
    for i in range(n):
        if not (np.random.rand() < prob):
            order[i] = pack_order([])
            continue
        perm = torch.randperm(4).tolist()
        if brightness:
            factors[i, 0] = float(torch.empty(1).uniform_(max(0.0, 1 - brightness), 1 + brightness))
        if contrast:
            factors[i, 1] = float(torch.empty(1).uniform_(max(0.0, 1 - contrast), 1 + contrast))
        if hue:
            factors[i, 2] = float(torch.empty(1).uniform_(-hue, hue))
        enabled = {OP_BRIGHTNESS: bool(brightness), OP_CONTRAST: bool(contrast), OP_HUE: bool(hue)}
        order[i] = pack_order([o for o in perm if enabled.get(o, False)])
    return order, factors


class FramePipeline:
    """rl_train.yaml's image branch for frames already at the training resolution: gather the windows of a batch from a
    device-resident uint8 frame store, pad short windows by repetition, RandomShiftsAug(pad), and -- when any jitter
    strength is non-zero -- ScaleImageTensor + ColorTransform + Normalize.  Returns uint8 (B, T, 3, H, W) (no jitter;
    the encoder normalises on load) or float32 normalised (B, T, 3, H, W)."""

    def __init__(self, pad=6, contrast=0.1, brightness=0.1, hue=0.02, prob=1.0, mean=0.5, std=0.5):
        self.pad, self.contrast, self.brightness, self.hue, self.prob = pad, contrast, brightness, hue, prob
        self.mean, self.std = mean, std

    def __call__(self, store, start, window, T):
        B = start.numel()
        shift = None
        if self.pad > 0:      # one draw per frame, x then y (torch.randint(0, 2 pad + 1, (n, 1, 1, 2)), :287-289)
            shift = rng.randint((B * T, 1, 1, 2), 0, 2 * self.pad + 1, store.device).reshape(B, T, 2)
        frames = window_gather(store, start, window, T, shift, self.pad)
        if not (self.contrast or self.brightness or self.hue):
            return frames
        order, factors = draw_color_jitter_params(B * T, self.brightness, self.contrast, self.hue, self.prob)
        out = color_jitter(frames.view(B * T, *frames.shape[2:]), order, factors, self.mean, self.std)
        return out.view(B, T, *out.shape[1:])
