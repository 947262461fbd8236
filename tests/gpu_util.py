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
                   dropout_p=0.0, goal_modalities=None):
    from tacorl_b200.utils.config import instantiate
    cfg = RC.play_lmp_cfg(pr_kind=pr_kind, modalities=modalities, rnn_hidden=rnn_hidden, latent_plan_dim=latent,
                          max_window=max_window, dropout_p=dropout_p, goal_modalities=goal_modalities)
    cfg["_target_"] = "tacorl.modules.play_lmp.play_lmp_for_rl.PlayLMP"   # reference path, remapped
    cfg["_recursive_"] = False
    return instantiate(cfg)


def to_dev(batch):
    return {k: to_dev(v) if isinstance(v, dict) else v.to(DEV) for k, v in batch.items()}


def play_lmp_tape(noise, B, goal_dim=32):
    return [noise["eps_pr"], noise["u1"], noise["u2"], noise["random_plan"], torch.zeros(B, goal_dim),
            noise["u1_rp"], noise["u2_rp"]]


def double_params(sd, frozen=()):
    return O.params_from({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}, frozen)


def build_tacorl(lmp, precision="fp32", **overrides):
    from tacorl_b200 import ops
    from tacorl_b200.utils.config import instantiate
    ops.set_precision(precision)
    cfg = RC.tacorl_cfg()
    cfg.update(overrides)
    cfg["_target_"] = "tacorl.modules.tacorl.tacorl.TACORL"                # reference path, remapped
    cfg["_recursive_"] = False
    return instantiate(cfg, play_lmp=lmp)


def build_cql_flat(precision="fp32", **overrides):
    from tacorl_b200 import ops
    from tacorl_b200.utils.config import instantiate
    ops.set_precision(precision)
    cfg = RC.cql_offline_cfg()
    cfg.update(overrides)
    cfg["_target_"] = "tacorl.modules.cql.cql_offline_lightning.CQL_Offline"   # reference path, remapped
    cfg["_recursive_"] = False
    return instantiate(cfg)


def cql_tape(noise):
    return [noise[k] for k in O.CQL_NOISE_ORDER]


def tacorl_tape(noise):
    return [noise[k] for k in ("plan_noise", "eps_actor", "eps_next", "rand_actions", "eps_curr", "eps_nextn")]


TACORL_KEYS = ["action_loss", "alpha", "alpha_loss", "actor_loss", "q1_loss", "q2_loss", "bellman_q1_loss",
               "bellman_q2_loss", "conservative_q1_loss", "conservative_q2_loss", "alpha_prime", "alpha_prime_loss",
               "q1_data", "q1_random", "q1_policy", "q2_data", "q2_random", "q2_policy"]


def parity_report(name, rows):
    """Per-tensor / per-scalar numbers of a parity test, written to gpurun_out/parity/<name>.json (gpurun brings the
    directory back; the round's copy is committed under profiles/)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out", "parity")
    try:
        os.makedirs(out, exist_ok=True)
        json.dump({"test": name, "rows": rows}, open(os.path.join(out, name + ".json"), "w"), indent=1)
    except OSError:      # read-only checkout: the assertions still run
        pass
