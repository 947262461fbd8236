"""GPU tests of the forward-only / deployment entry points: the inner loop of the reference's
rollout (/root/reference/src/tacorl/evaluation/rollout_manager.py:364-386 -- deterministic latent plan from the visual
actor, then `plan_duration` stateful single-step decoder calls) and the validation steps, against the CPU oracle."""
import pytest
import torch

from oracle import synth as S
from oracle import tacorl_oracle as O
from tests.gpu_util import DEV, assert_close, build_play_lmp, build_tacorl, to_dev

pytestmark = pytest.mark.gpu


def _tacorl(hidden, T=8, seed=21):
    lmp = build_play_lmp("tanh_net", ("rgb_static",), hidden, 16, T)
    t = build_tacorl(lmp)
    shapes = {k: list(v.shape) for k, v in t.state_dict().items()}
    sd = S.synth_state_dict(shapes, seed)
    t.load_state_dict(sd, strict=True)
    t.to(DEV)
    t.eval()
    return t, sd


@pytest.mark.parametrize("hidden,hw,precision", [(64, 84, "fp32"), (2048, 200, "fp32"), (2048, 200, "bf16")])
def test_rollout_inner_loop_matches_oracle(hidden, hw, precision):
    """pp_actor.get_actions(obs, deterministic=True) -> clear_hidden_state() -> plan_duration x act() on one frame at a
    time (batch 1, T = 1, hidden state carried inside the module) == the oracle's decoder run over the whole sequence."""
    from tacorl_b200 import ops
    from tacorl_b200.utils.rng import noise_tape
    steps = 5
    t, sd = _tacorl(hidden)
    try:
        ops.set_precision(precision)
        P = O.params_from(sd)
        frames = S.synth_images((steps, 3, hw, hw), 3, "rollout")
        goal = S.synth_images((3, hw, hw), 3, "rollout_goal")
        g = torch.Generator().manual_seed(9)
        u1 = torch.rand(steps, 1, 1, 6, 10, generator=g)
        u2 = torch.rand(steps, 1, 1, 6, generator=g)
        with torch.no_grad():
            obs = {"observation": {"rgb_static": frames[0].to(DEV)}, "goal": {"rgb_static": goal.to(DEV)}}
            plan, zero_lp = t.actor.get_actions(obs, deterministic=True, reparameterize=False)
            assert plan.shape == (16,) and float(zero_lp.abs().sum()) == 0.0
            # oracle: tanh(mean) of the policy on cat(encoder(obs), goal_encoder(encoder(goal)))
            a_emb = O.visual_emb(P, "actor.", {"rgb_static": frames[:1]}, {"rgb_static": goal[None]})
            mu, _ = O.mlp_policy(P, "actor.actor.policy.", a_emb)
            tol = 1e-4 if precision == "fp32" else 3e-2
            assert_close("deterministic plan", plan[None], torch.tanh(mu), tol, atol=1e-5)
            t.action_decoder.clear_hidden_state()
            got = []
            for s in range(steps):
                ad_state = t.perceptual_encoder.get_state_from_observation(
                    observation={"rgb_static": frames[s].to(DEV)}, modalities=t.action_decoder_modalities)
                assert ad_state.shape == (32,)
                with noise_tape([u1[s], u2[s]]) as tape:
                    a = t.action_decoder.act(latent_plan=plan.unsqueeze(0), perceptual_emb=ad_state.unsqueeze(0).unsqueeze(0))
                    assert len(tape) == 0
                assert a.shape == (1, 1, 7) and t.action_decoder.hidden_state.shape == (2, 1, hidden)
                got.append(a.squeeze().cpu())
            got = torch.stack(got)
            emb = O.lmp_encoder(P, "perceptual_encoder.networks.rgb_static.", frames)[None]          # (1, steps, 32)
            lp, ls, mm, grip, _ = O.action_decoder_forward(P, "action_decoder.", plan.cpu()[None], emb)
            want = O.dlm_sample(lp, ls, mm, grip, u1[:, 0, 0][None], u2[:, 0, 0][None])[0]
            if precision == "fp32":
                assert_close("actions over the rollout", got, want, 1e-4, atol=1e-5)
            else:   # the Gumbel-max mixture pick can flip on a near-tie under bf16 rounding: gripper + most dims must agree
                assert torch.equal(got[:, -1], want[:, -1])
                close = ((got[:, :6] - want[:, :6]).abs() < 5e-2).float().mean()
                assert close > 0.8, close
            # clear_hidden_state() restarts the recurrence
            t.action_decoder.clear_hidden_state()
            assert t.action_decoder.hidden_state is None
            with noise_tape([u1[0], u2[0]]):
                ad0 = t.perceptual_encoder.get_state_from_observation(observation={"rgb_static": frames[0].to(DEV)},
                                                                      modalities=t.action_decoder_modalities)
                a0 = t.action_decoder.act(latent_plan=plan.unsqueeze(0), perceptual_emb=ad0.unsqueeze(0).unsqueeze(0))
            assert rel(a0.squeeze().cpu(), got[0]) < (1e-6 if precision == "fp32" else 1e-2)
    finally:
        ops.set_precision("fp32")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_stochastic_get_actions_and_log_prob_entry_points():
    """Actor API used by the CQL agent outside training_step: get_actions(reparameterize), sample_n_with_log_prob,
    log_prob (actor.py:98-156) against the oracle's TanhNormal."""
    from tacorl_b200 import ops
    from tacorl_b200.utils.rng import noise_tape
    ops.set_precision("fp32")
    t, sd = _tacorl(64)
    P = O.params_from(sd)
    B, hw = 3, 84
    obs = {"observation": {"rgb_static": S.synth_images((B, 3, hw, hw), 5, "o")}, "goal": {"rgb_static": S.synth_images((B, 3, hw, hw), 5, "g")}}
    g = torch.Generator().manual_seed(2)
    eps, eps_n, acts = torch.randn(B, 16, generator=g), torch.randn(4, B, 16, generator=g), torch.rand(B, 16, generator=g) * 1.6 - 0.8
    with torch.no_grad():
        dobs = {k: {kk: vv.to(DEV) for kk, vv in v.items()} for k, v in obs.items()}
        a_emb = O.visual_emb(P, "actor.", obs["observation"], obs["goal"])
        mu, std = O.mlp_policy(P, "actor.actor.policy.", a_emb)
        with noise_tape([eps]):
            a, lp = t.actor.get_actions(dobs, deterministic=False, reparameterize=True)
        z = mu + std * eps
        assert_close("rsample", a, torch.tanh(z), 1e-4, atol=1e-6)
        assert_close("log_prob", lp, O.tanh_normal_log_prob(mu, std, pre_tanh=z), 1e-4, atol=1e-5)
        with noise_tape([eps_n]):
            an, lpn = t.actor.sample_n_with_log_prob(dobs, 4)
        zn = mu + std * eps_n
        assert_close("sample_n", an, torch.tanh(zn), 1e-4, atol=1e-6)
        assert_close("sample_n log_prob", lpn, O.tanh_normal_log_prob(mu, std, pre_tanh=zn), 1e-4, atol=1e-5)
        lpa = t.actor.log_prob(dobs, acts.to(DEV))
        assert_close("log_prob(value)", lpa, O.tanh_normal_log_prob(mu, std, value=acts), 1e-4, atol=1e-5)
        q = t.q1(dobs, acts.to(DEV))
        q_emb = O.visual_emb(P, "q1.", obs["observation"], obs["goal"])
        assert_close("critic forward", q, O.q_value(P, "q1.", q_emb, acts), 1e-4, atol=1e-6)


def test_play_lmp_validation_step_matches_oracle():
    """PlayLMP.validation_step (play_lmp_for_rl.py:319-348): no gradient, same losses as the oracle, returns the sampled
    proposal plan."""
    from tacorl_b200 import ops
    from tacorl_b200.utils.rng import noise_tape
    from tests.gpu_util import play_lmp_tape
    ops.set_precision("fp32")
    B, T = 3, 8
    m = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, T)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    sd = S.synth_state_dict(shapes, 8)
    m.load_state_dict(sd)
    m.to(DEV)
    m.eval()
    batch = S.synth_play_batch(B, T, 84, 84, 8)
    torch.manual_seed(5)
    noise = O.draw_play_lmp_noise(B, T)
    pp_noise = torch.randn(B, 16)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    with noise_tape(play_lmp_tape(noise, B) + [pp_noise]) as tape:
        out = m.validation_step(to_dev(S.clone_batch(batch)), 0)
        assert len(tape) == 0
    P = O.params_from(sd)
    ref = O.play_lmp_forward(P, S.clone_batch(batch), noise)
    for k in ("kl_loss", "action_loss", "total_loss", "random_plan_action_loss"):
        got, want = float(m.logged["validation/" + k]), float(ref[k])
        assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (k, got, want)
    assert_close("sampled_plan_pp", out["sampled_plan_pp"], torch.tanh(ref["mu_p"] + ref["std_p"] * pp_noise).detach(), 1e-4, atol=1e-6)
    assert torch.equal(out["idx"].cpu(), batch["idx"])
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert all(p.grad is None for p in m.parameters())
