"""Synthetic CALVIN-shaped batches: the reference's batch contract
(datamodule/dataset/play_dataset.py:115-169) filled with U(-1,1) images / actions."""
import torch


def play_batch(B, T=16, H=200, W=200, seed=0, modalities=("rgb_static",), gripper_hw=(84, 84), with_goal=False,
               goal_modalities=("rgb_static",)):
    g = torch.Generator().manual_seed(seed)
    actions = torch.rand(B, T, 7, generator=g) * 2 - 1
    actions[..., -1] = torch.where(actions[..., -1] > 0, 1.0, -1.0)
    states = {}
    for m in modalities:
        h, w = (H, W) if m == "rgb_static" else gripper_hw
        states[m] = torch.randint(0, 256, (B, T, 3, h, w), generator=g, dtype=torch.uint8).float() / 127.5 - 1.0
    batch = {"states": states, "actions": actions}
    if with_goal:
        batch["goal"] = {m: torch.randint(0, 256, (B, 3) + ((H, W) if m == "rgb_static" else tuple(gripper_hw)),
                                          generator=g, dtype=torch.uint8).float() / 127.5 - 1.0 for m in goal_modalities}
        u = torch.rand(B, generator=g)
        disp = torch.floor(torch.log(1 - u) / torch.log(torch.tensor(0.7))).long() + 1
        disp[torch.rand(B, generator=g) < 0.1] = -1
        batch["disp"] = disp
    return batch


def init_like_reference(module, seed=0):
    """Deterministic re-initialisation (uniform +-1/sqrt(fan_in)) so every rank starts identical."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("temperature"):
                p.fill_(1.0)
            elif name in ("log_alpha", "log_alpha_prime"):
                p.zero_()
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) / fan_in ** 0.5)
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.05)
