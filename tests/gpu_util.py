"""Helpers shared by the GPU parity tests (CUDA path vs CPU oracle)."""
import json
import os

import torch

from oracle import ref_loader_cfg as RC
from oracle import synth as S
from oracle import tacorl_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def rel_err(got, want):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    return float((got - want).norm() / (want.norm() + 1e-30))


def assert_close(tag, got, want, rtol, atol=1e-7):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    assert got.shape == want.shape, (tag, got.shape, want.shape)
    err = float((got - want).norm())
    ref = float(want.norm())
    assert err <= rtol * ref + atol * max(1.0, want.numel() ** 0.5), f"{tag}: rel L2 err {err / (ref + 1e-30):.3e} (|ref|={ref:.3e})"


def load_golden(name):
    return json.load(open(os.path.join(GOLD, name + ".json")))


def build_play_lmp(pr_kind="tanh_net", modalities=("rgb_static",), rnn_hidden=2048, latent=16, max_window=16,
                   dropout_p=0.0):
    from tacorl_b200.utils.config import instantiate
    cfg = RC.play_lmp_cfg(pr_kind=pr_kind, modalities=modalities, rnn_hidden=rnn_hidden, latent_plan_dim=latent,
                          max_window=max_window, dropout_p=dropout_p)
    cfg["_target_"] = "tacorl.modules.play_lmp.play_lmp_for_rl.PlayLMP"   # reference path, remapped
    cfg["_recursive_"] = False
    return instantiate(cfg)


def to_dev(batch):
    out = {}
    for k, v in batch.items():
        out[k] = {kk: vv.to(DEV) for kk, vv in v.items()} if isinstance(v, dict) else v.to(DEV)
    return out


def play_lmp_tape(noise, B, goal_dim=32):
    return [noise["eps_pr"], noise["u1"], noise["u2"], noise["random_plan"], torch.zeros(B, goal_dim),
            noise["u1_rp"], noise["u2_rp"]]


def double_params(sd, frozen=()):
    return O.params_from({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}, frozen)
