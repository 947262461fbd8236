"""GPU parity tests: tacorl_b200.modules.tacorl.TACORL training_step (decoder update, CQL actor / twin-Q /
Lagrange update, Polyak) against the numbers recorded from the unmodified reference (tests/golden/tacorl_*)."""
import pytest
import torch

from oracle import ref_loader_cfg as RC
from oracle import synth as S
from oracle import tacorl_oracle as O
from tests.gpu_util import DEV, build_play_lmp, load_golden, to_dev

pytestmark = pytest.mark.gpu
KEYS = ["action_loss", "alpha", "alpha_loss", "actor_loss", "q1_loss", "q2_loss", "bellman_q1_loss",
        "bellman_q2_loss", "conservative_q1_loss", "conservative_q2_loss", "alpha_prime", "alpha_prime_loss",
        "q1_data", "q1_random", "q1_policy", "q2_data", "q2_random", "q2_policy"]


def _build_tacorl(rec):
    from tacorl_b200 import ops
    from tacorl_b200.utils.config import instantiate
    ops.set_precision("fp32")
    lmp = build_play_lmp(rec["pr_kind"], ("rgb_static",), rec["rnn_hidden"], 16, rec["T"])
    cfg = RC.tacorl_cfg()
    cfg["_target_"] = "tacorl.modules.tacorl.tacorl.TACORL"
    cfg["_recursive_"] = False
    return instantiate(cfg, play_lmp=lmp)


def _tape(noise):
    return [noise[k] for k in ("plan_noise", "eps_actor", "eps_next", "rand_actions", "eps_curr", "eps_nextn")]


@pytest.mark.parametrize("name", ["tacorl_bc_84", "tacorl_q_84", "tacorl_defaultpr_84"])
def test_tacorl_steps_match_reference_golden(name):
    from tacorl_b200.utils.rng import noise_tape
    rec = load_golden(name)
    t = _build_tacorl(rec)
    assert {k: list(v.shape) for k, v in t.state_dict().items()} == rec["shapes"]
    t.load_state_dict(S.synth_state_dict(rec["shapes"], rec["seed"]), strict=True)
    t.to(DEV)
    t.current_epoch = rec["epoch"]
    t.optimizers()
    batch = S.synth_play_batch(rec["B"], rec["T"], rec["H"], rec["W"], rec["seed"], with_goal=True)
    batch["disp"] = torch.tensor(rec["disp"])
    for s, step in enumerate(rec["steps"]):
        torch.manual_seed(rec["noise_seed_base"] + s)
        noise = O.draw_tacorl_noise(rec["B"])
        with noise_tape(_tape(noise)) as tape:
            t.training_step(to_dev(S.clone_batch(batch)))
            assert len(tape) == 0, "every reference draw must be consumed, in order"
        for k in KEYS:
            got, want = float(t.logged["train/" + k]), step["scalars"][k]
            assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (name, s, k, got, want)
        sd = t.state_dict()
        bad = [k for k, fp in step["params"].items() if not S.fingerprint_close(S.fingerprint(sd[k]), fp, 2e-4)]
        assert not bad, (name, s, bad[:8])


def test_tacorl_validation_step_changes_nothing():
    rec = load_golden("tacorl_q_84")
    t = _build_tacorl(rec)
    t.load_state_dict(S.synth_state_dict(rec["shapes"], rec["seed"]), strict=True)
    t.to(DEV)
    t.optimizers()
    before = {k: v.clone() for k, v in t.state_dict().items()}
    batch = S.synth_play_batch(rec["B"], rec["T"], rec["H"], rec["W"], rec["seed"], with_goal=True)
    t.validation_step(to_dev(batch))
    for k, v in t.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert "validation/q1_loss" in t.logged
