"""TEST INFRASTRUCTURE — deterministic synthetic parameters, CALVIN-shaped batches and
tensor fingerprints shared by oracle/make_golden.py, tests/ and bench.py.

Everything is drawn from CPU `torch.Generator`s keyed by (seed, name), so a fixture only
has to store seeds + shapes + fingerprints, not megabytes of weights.  Batch layout follows
the reference's batch contract (datamodule/dataset/play_dataset.py:115-169, pad_sequence
:282-310; SURVEY.md §8 A0 and §8d "Synthetic inputs").
"""
import zlib

import torch

BUFFER_SUFFIXES = ("one_hot_embedding_eye", "ones", "gripper_bounds", "action_max_bound",
                   "action_min_bound")


def _gen(seed, name):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))


def synth_param(name, shape, seed):
    """Deterministic value for one state_dict entry."""
    leaf = name.split(".")[-1]
    shape = tuple(shape)
    if leaf == "one_hot_embedding_eye":
        return torch.eye(shape[0])
    if leaf == "ones":
        return torch.ones(shape)
    if leaf == "gripper_bounds":
        return torch.tensor([-1.0, 1.0])
    if leaf == "action_max_bound":
        return torch.ones(shape)
    if leaf == "action_min_bound":
        return -torch.ones(shape)
    if leaf == "temperature":
        return torch.ones(shape)
    if leaf in ("log_alpha", "log_alpha_prime"):
        return torch.zeros(shape)
    g = _gen(seed, name)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = 1.0 / fan_in ** 0.5
        if "norm" in name and leaf == "weight":
            return torch.ones(shape)
    else:
        if "norm" in name and leaf == "weight":
            return 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        bound = 0.05
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def synth_state_dict(shapes, seed):
    """shapes: {name: shape}.  Returns {name: fp32 tensor} in the given key order."""
    return {k: synth_param(k, s, seed) for k, s in shapes.items()}


def synth_images(shape, seed, name="img"):
    """U(-1,1)-ish images exactly representable from uint8: (u8/127.5 - 1)
    (the post-`Normalize` range of rl_train.yaml:12-14)."""
    g = _gen(seed, name)
    u8 = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    return u8.float() / 127.5 - 1.0


def synth_play_batch(B, T, H, W, seed, modalities=("rgb_static",), gripper_hw=(84, 84),
                     with_goal=False, pad=False, goal_modalities=("rgb_static",)):
    """CALVIN-shaped PlayLMP / TACO-RL batch.  actions U(-1,1) with gripper channel ±1;
    `pad=True` mimics pad_sequence: window w~U{T/2..T}, frames repeat the last valid one and
    relative actions are zero after w (gripper kept)."""
    g = _gen(seed, "batch")
    actions = torch.rand(B, T, 7, generator=g) * 2 - 1
    actions[..., -1] = torch.where(actions[..., -1] > 0, 1.0, -1.0)
    states = {}
    for m in modalities:
        h, w = (H, W) if m == "rgb_static" else gripper_hw
        states[m] = synth_images((B, T, 3, h, w), seed, m)
    batch = {"states": states, "actions": actions,
             "idx": torch.arange(B), "window_size": torch.full((B,), T)}
    if pad:
        ws = torch.randint(max(T // 2, 2), T + 1, (B,), generator=g)
        for b in range(B):
            w_ = int(ws[b])
            for m in states:
                states[m][b, w_:] = states[m][b, w_ - 1]
            actions[b, w_:, :6] = 0.0
            actions[b, w_:, 6] = actions[b, w_ - 1, 6]
        batch["window_size"] = ws
    if with_goal:
        batch["goal"] = {m: synth_images((B, 3) + ((H, W) if m == "rgb_static" else tuple(gripper_hw)), seed,
                                         "goal" if m == "rgb_static" else "goal_" + m) for m in goal_modalities}
        # disp ~ Geometric(0.3) with 10% = -1 (config/datamodule/dataset/tacorl.yaml:11-14)
        u = torch.rand(B, generator=g)
        disp = torch.floor(torch.log(1 - u) / torch.log(torch.tensor(0.7))).long() + 1
        neg = torch.rand(B, generator=g) < 0.1
        disp[neg] = -1
        batch["disp"] = disp
    return batch


def synth_cql_batch(B, H, W, seed, modalities=("rgb_static",), goal_modalities=("rgb_static",), gripper_hw=(84, 84)):
    """Transition batch of the flat-CQL baseline (GoalCondReplayBufferDataset.get_transition,
    datamodule/dataset/goal_cond_replay_buffer_dataset.py:277-296): reward = terminal = [goal is the next step]."""
    g = _gen(seed, "cql_batch")
    hw = lambda m: (H, W) if m == "rgb_static" else tuple(gripper_hw)
    obs = {m: synth_images((B, 3) + hw(m), seed, "obs_" + m) for m in modalities}
    nxt = {m: synth_images((B, 3) + hw(m), seed, "next_" + m) for m in modalities}
    goal = {m: synth_images((B, 3) + hw(m), seed, "cqlgoal_" + m) for m in goal_modalities}
    actions = torch.rand(B, 7, generator=g) * 2 - 1
    actions[..., -1] = torch.where(actions[..., -1] > 0, 1.0, -1.0)
    hit = (torch.rand(B, generator=g) < 0.3).long()
    hit[0], hit[-1] = 1, 0
    return {"observations": {"observation": obs, "goal": goal}, "actions": actions,
            "next_observations": {"observation": nxt, "goal": goal}, "rewards": hit.clone(), "terminals": hit.clone()}


def clone_batch(batch):
    """Fresh dict per call (the reference mutates batch['states'] in place,
    play_lmp_for_rl.py:188-190)."""
    return {k: clone_batch(v) if isinstance(v, dict) else v.clone() for k, v in batch.items()}


def fingerprint(t, seed=7, name="probe"):
    """(sum, l2, probe-dot) in fp64 — a compact, element-sensitive signature of a tensor."""
    t = t.detach().double().cpu().reshape(-1)
    g = _gen(seed, name + str(t.numel()))
    probe = torch.rand(t.numel(), generator=g, dtype=torch.float64) * 2 - 1
    return [float(t.sum()), float(t.norm()), float((t * probe).sum())]


def fingerprint_close(fp_a, fp_b, rtol, atol=1e-7):
    """Compare two fingerprints; scale tolerance by the l2 norm."""
    scale = max(abs(fp_a[1]), abs(fp_b[1]), 1e-30)
    ok_sum = abs(fp_a[0] - fp_b[0]) <= rtol * max(scale * 50, abs(fp_a[0])) + atol
    ok_l2 = abs(fp_a[1] - fp_b[1]) <= rtol * scale + atol
    ok_dot = abs(fp_a[2] - fp_b[2]) <= rtol * scale * 50 + atol
    return ok_sum and ok_l2 and ok_dot
