"""TEST INFRASTRUCTURE — CPU restatement of the reference's per-sample input pipeline (window collation + image
transforms), the checker of tacorl_b200/csrc/data_pipeline.cu.  Citations relative to /root/reference/src/tacorl/.

The colour ops live in the third-party dependency torchvision (un-pinned in setup.cfg; 0.26 in this image):
`torchvision.transforms.ColorJitter` = brightness / contrast / saturation / hue applied in a random permutation
(`get_params`), each via `torchvision.transforms.functional.adjust_*`; they are called here directly.
Pinned against the unmodified reference classes by tests/test_host_logic.py (when the reference is present)."""
import torch
import torch.nn.functional as F


def random_shifts(x, pad, shift):
    """RandomShiftsAug.forward, utils/transforms.py:270-299, with the `torch.randint` draw given:
    x (n, c, h, w) float; shift (n, 1, 1, 2) integer-valued in [0, 2 pad]."""
    n, c, h, w = x.shape
    assert h == w
    x = F.pad(x, (pad,) * 4, "replicate")
    eps = 1.0 / (h + 2 * pad)
    arange = torch.linspace(-1.0 + eps, 1.0 - eps, h + 2 * pad, dtype=x.dtype)[:h]
    arange = arange.unsqueeze(0).repeat(h, 1).unsqueeze(2)
    base_grid = torch.cat([arange, arange.transpose(1, 0)], dim=2).unsqueeze(0).repeat(n, 1, 1, 1)
    sh = shift.to(x.dtype) * (2.0 / (h + 2 * pad))
    return F.grid_sample(x, base_grid + sh, padding_mode="zeros", align_corners=False)


def scale_image(x):
    """ScaleImageTensor, utils/transforms.py:87-101 (uint8 input)."""
    return x.float().div(255).clip(0.0, 1.0)


def color_jitter(img, ops, brightness, contrast, hue):
    """torchvision ColorJitter.forward for one image (3, h, w) in [0, 1]: `ops` = the permutation `fn_idx` restricted
    to the enabled ops (0 brightness, 1 contrast, 3 hue)  — ColorTransform.apply_transform, utils/transforms.py:310-314."""
    import torchvision.transforms.functional as TF
    for op in ops:
        if op == 0:
            img = TF.adjust_brightness(img, brightness)
        elif op == 1:
            img = TF.adjust_contrast(img, contrast)
        elif op == 3:
            img = TF.adjust_hue(img, hue)
    return img


def normalize(x, mean=0.5, std=0.5):
    """torchvision Normalize(mean=[0.5], std=[0.5]) of rl_train.yaml:12-14."""
    return (x - mean) / std


def pad_frames(frames, window, T):
    """PlayDataset.pad_with_repetition, datamodule/dataset/play_dataset.py:312-318: frames (w, ...) -> (T, ...)."""
    w = int(window)
    return torch.cat([frames[:w], frames[w - 1:w].repeat_interleave(T - w, dim=0)], dim=0) if w < T else frames[:T]


def pad_rel_actions(actions, window, T):
    """pad_sequence for the "rel" action modalities, play_dataset.py:291-301: zeros except the repeated gripper channel."""
    w = int(window)
    if w >= T:
        return actions[:T]
    tail = torch.zeros(T - w, actions.shape[-1], dtype=actions.dtype)
    tail[:, -1] = actions[w - 1, -1]
    return torch.cat([actions[:w], tail], dim=0)
