"""GPU parity tests: tacorl_b200.modules.tacorl.TACORL training_step (decoder update, CQL actor / twin-Q /
Lagrange update, Polyak) against the numbers recorded from the unmodified reference (tests/golden/tacorl_*)."""
import pytest
import torch

from oracle import synth as S
from oracle import tacorl_oracle as O
from tests.gpu_util import DEV, build_play_lmp, build_tacorl, load_golden, tacorl_tape, to_dev

pytestmark = pytest.mark.gpu
KEYS = ["action_loss", "alpha", "alpha_loss", "actor_loss", "q1_loss", "q2_loss", "bellman_q1_loss",
        "bellman_q2_loss", "conservative_q1_loss", "conservative_q2_loss", "alpha_prime", "alpha_prime_loss",
        "q1_data", "q1_random", "q1_policy", "q2_data", "q2_random", "q2_policy"]


def _build_tacorl(rec):
    mods = tuple(rec.get("modalities", ["rgb_static"]))
    lmp = build_play_lmp(rec["pr_kind"], mods, rec["rnn_hidden"], rec.get("latent_plan_dim", 16), rec["T"],
                         dropout_p=rec.get("dropout_p", 0.0), goal_modalities=rec.get("goal_modalities"))
    return build_tacorl(lmp)


def _batch(rec):
    mods = rec.get("modalities", ["rgb_static"])
    batch = S.synth_play_batch(rec["B"], rec["T"], rec["H"], rec["W"], rec["seed"], modalities=mods, with_goal=True,
                               goal_modalities=rec.get("goal_modalities", mods[:1]))
    batch["disp"] = torch.tensor(rec["disp"])
    return batch


_tape = tacorl_tape


@pytest.mark.parametrize("name", ["tacorl_bc_84", "tacorl_q_84", "tacorl_defaultpr_84", "tacorl_transformer_84",
                                  "tacorl_multiview_bc", "tacorl_multiview_q"])
def test_tacorl_steps_match_reference_golden(name):
    """tacorl_transformer_84: transformer recogniser with dropout 0.1 -- the frozen LMP must run in eval mode, i.e. the
    noise tape (which holds no dropout masks) is consumed exactly.  tacorl_multiview_*: BASELINE configs[3], static +
    gripper views for observation and goal, latent plan 32."""
    from tacorl_b200.utils.rng import noise_tape
    rec = load_golden(name)
    t = _build_tacorl(rec)
    t.train()
    assert {k: list(v.shape) for k, v in t.state_dict().items()} == rec["shapes"]
    t.load_state_dict(S.synth_state_dict(rec["shapes"], rec["seed"]), strict=True)
    t.to(DEV)
    t.current_epoch = rec["epoch"]
    t.optimizers()
    batch = _batch(rec)
    for s, step in enumerate(rec["steps"]):
        torch.manual_seed(rec["noise_seed_base"] + s)
        noise = O.draw_tacorl_noise(rec["B"], latent=rec.get("latent_plan_dim", 16))
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
    batch = _batch(rec)
    t.validation_step(to_dev(batch))
    for k, v in t.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert "validation/q1_loss" in t.logged
