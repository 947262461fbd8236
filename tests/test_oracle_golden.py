"""Pins oracle/tacorl_oracle.py against the golden fixtures generated from the UNMODIFIED
reference (oracle/make_golden.py).  CPU only; the reference itself is not needed."""
import glob
import json
import os

import pytest
import torch

from oracle import synth as S
from oracle import tacorl_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 2e-4


def _load(name):
    return json.load(open(os.path.join(GOLD, name + ".json")))


def _check_fp(tag, got, want, rtol=RTOL):
    assert S.fingerprint_close(S.fingerprint(got), want, rtol), (tag, S.fingerprint(got), want)


def _check_scalars(got, want, keys, rtol=RTOL):
    for k in keys:
        g, w = float(got[k]), want[k]
        assert abs(g - w) <= rtol * max(1.0, abs(w)), (k, g, w)


@pytest.mark.parametrize("name", ["playlmp_birnn_84", "playlmp_birnn_pad_128", "playlmp_multiview",
                                  "playlmp_transformer_84"])
def test_play_lmp_steps_match_reference(name):
    rec = _load(name)
    P = O.params_from(S.synth_state_dict(rec["shapes"], rec["seed"]))
    batch = S.synth_play_batch(rec["B"], rec["T"], rec["H"], rec["W"], rec["seed"],
                               modalities=rec["modalities"], pad=rec["pad"])
    cfg = {"modalities": rec["modalities"], "pr_kind": rec["pr_kind"]}
    latent = rec["shapes"]["plan_recognition.mean_fc.weight"][0]
    opt = {}
    for s, step in enumerate(rec["steps"]):
        torch.manual_seed(rec["noise_seed_base"] + s)
        noise = O.draw_play_lmp_noise(rec["B"], rec["T"], latent=latent)
        out, grads = O.play_lmp_training_step(P, opt, S.clone_batch(batch), noise, cfg)
        _check_scalars(out, step["scalars"], ["kl_loss", "kl_loss_scaled", "action_loss", "total_loss",
                                              "gripper_accuracy", "random_plan_action_loss",
                                              "random_plan_gripper_accuracy"])
        assert set(grads) == set(step["grads"])
        for k, fp in step["grads"].items():
            _check_fp(f"{name}/step{s}/grad/{k}", grads[k], fp)
        for k, fp in step["params"].items():
            _check_fp(f"{name}/step{s}/param/{k}", P[k], fp)


TACORL_FIXTURES = ["tacorl_bc_84", "tacorl_q_84", "tacorl_defaultpr_84", "tacorl_transformer_84",
                   "tacorl_multiview_bc", "tacorl_multiview_q"]


@pytest.mark.parametrize("name", TACORL_FIXTURES)
def test_tacorl_steps_match_reference(name):
    rec = _load(name)
    mods = rec.get("modalities", ["rgb_static"])
    goal_mods = rec.get("goal_modalities", mods[:1])
    latent = rec.get("latent_plan_dim", 16)
    P = O.params_from(S.synth_state_dict(rec["shapes"], rec["seed"]), O.TACORL_FROZEN)
    batch = S.synth_play_batch(rec["B"], rec["T"], rec["H"], rec["W"], rec["seed"], modalities=mods, with_goal=True,
                               goal_modalities=goal_mods)
    batch["disp"] = torch.tensor(rec["disp"])
    opt = O.new_tacorl_opt_state(P)
    cfg = {"pr_kind": rec["pr_kind"], "target_entropy": rec["target_entropy"], "modalities": mods,
           "goal_modalities": goal_mods}
    keys = ["action_loss", "alpha", "alpha_loss", "actor_loss", "q1_loss", "q2_loss", "bellman_q1_loss",
            "bellman_q2_loss", "conservative_q1_loss", "conservative_q2_loss", "alpha_prime",
            "alpha_prime_loss", "q1_data", "q1_random", "q1_policy", "q2_data", "q2_random", "q2_policy"]
    for s, step in enumerate(rec["steps"]):
        torch.manual_seed(rec["noise_seed_base"] + s)
        noise = O.draw_tacorl_noise(rec["B"], latent=latent)
        logged, _ = O.tacorl_training_step(P, opt, S.clone_batch(batch), noise, cfg, rec["epoch"])
        _check_scalars(logged, step["scalars"], keys)
        for k, fp in step["params"].items():
            _check_fp(f"{name}/step{s}/param/{k}", P[k], fp)


@pytest.mark.parametrize("name", ["cql_flat_bc", "cql_flat_q"])
def test_flat_cql_steps_match_reference(name):
    """the flat-CQL baseline (discrete-gripper actor, entropy-regularised backup)."""
    rec = _load(name)
    P = O.params_from(S.synth_state_dict(rec["shapes"], rec["seed"]))
    batch = S.synth_cql_batch(rec["B"], rec["H"], rec["W"], rec["seed"])
    opt = O.new_tacorl_opt_state(P)
    cfg = {"target_entropy": rec["target_entropy"], "deterministic_backup": rec["deterministic_backup"]}
    keys = ["alpha", "alpha_loss", "actor_loss", "q1_loss", "q2_loss", "bellman_q1_loss", "bellman_q2_loss",
            "conservative_q1_loss", "conservative_q2_loss", "alpha_prime", "alpha_prime_loss", "q1_data", "q1_random",
            "q1_policy", "q2_data", "q2_random", "q2_policy"]
    for s, step in enumerate(rec["steps"]):
        torch.manual_seed(rec["noise_seed_base"] + s)
        noise = O.draw_cql_noise(rec["B"])
        logged, _ = O.cql_training_step(P, opt, S.clone_batch(batch), noise, cfg, rec["epoch"])
        _check_scalars(logged, step["scalars"], keys)
        for k, fp in step["params"].items():
            _check_fp(f"{name}/step{s}/param/{k}", P[k], fp)


def test_encoder_shapes_match_reference():
    rec = _load("encoder_shapes")
    sd = S.synth_state_dict(rec["shapes"], rec["seed"])
    sd["model.6.temperature"] = torch.tensor([rec["temperature"]])
    for case in rec["cases"]:
        P = O.params_from(sd)
        n, h, w = case["n"], case["h"], case["w"]
        x = S.synth_images((n, 3, h, w), rec["seed"], f"enc{h}x{w}")
        y = O.lmp_encoder(P, "", x)
        cot = torch.rand(y.shape, generator=S._gen(rec["seed"], f"cot{h}x{w}")) * 2 - 1
        (y * cot).sum().backward()
        _check_fp("out", y, case["out"])
        assert torch.allclose(y[0, :8], torch.tensor(case["out_head"]), rtol=1e-4, atol=1e-5)
        feat = O.spatial_softargmax(O.lmp_encoder_convs(P, "", x), P["model.6.temperature"])
        _check_fp("softargmax", feat, case["softargmax"])
        for k, fp in case["grads"].items():
            _check_fp(f"{h}x{w}/grad/{k}", P[k].grad, fp)


def test_ops_known_answers():
    rec = _load("ops_kat")
    t = lambda k: torch.tensor(rec[k])
    lp, ls, mu, grip, act = t("lp"), t("ls"), t("mu"), t("grip"), t("act")
    assert abs(float(O.dlm_loss(lp, ls, mu, grip, act)) - rec["dlm_loss"]) < 1e-4
    assert abs(float(O.dlm_logistic_loss(lp, ls, mu, act[:, :, :-1])) - rec["logistic_loss"]) < 1e-4
    samp = O.dlm_sample(lp, torch.clamp(ls, min=-5.0), mu, grip, t("u1"), t("u2"))
    assert torch.allclose(samp, t("sample"), rtol=1e-5, atol=1e-6)
    mean, std, z, val = t("mean"), t("std"), t("z"), t("val")
    assert torch.allclose(O.tanh_normal_log_prob(mean, std, pre_tanh=z), t("logp_pre"), rtol=1e-5, atol=1e-5)
    assert torch.allclose(O.tanh_normal_log_prob(mean, std, value=val), t("logp_val"), rtol=1e-5, atol=1e-5)


def test_rnn_restatement_matches_torch_rnn():
    torch.manual_seed(3)
    for bidir in (False, True):
        rnn = torch.nn.RNN(12, 20, num_layers=2, nonlinearity="relu", bidirectional=bidir, batch_first=True)
        P = {"r." + k: v for k, v in rnn.state_dict().items()}
        x = torch.randn(4, 7, 12)
        h0 = torch.randn(2 * (2 if bidir else 1), 4, 20)
        want, hn = rnn(x, h0)
        got, hn2 = O.rnn_stack(P, "r.", x, 2, bidir, h0)
        assert torch.allclose(got, want, atol=1e-5) and torch.allclose(hn, hn2, atol=1e-5)


def test_adam_and_clip_restatements_match_torch():
    torch.manual_seed(4)
    ps = [torch.randn(5, 3, requires_grad=True), torch.randn(7, requires_grad=True)]
    mine = [p.detach().clone() for p in ps]
    opt = torch.optim.Adam(ps, lr=3e-4)
    st = O.new_adam_state(mine)
    for _ in range(4):
        gs = [torch.randn_like(p) * 3 for p in ps]
        gm = [g.clone() for g in gs]
        for p, g in zip(ps, gs):
            p.grad = g
        n_ref = torch.nn.utils.clip_grad_norm_(ps, 1.0)
        n_mine = O.clip_grad_norm(gm, 1.0)
        assert torch.allclose(n_ref, n_mine, rtol=1e-6)
        opt.step()
        O.adam_step(mine, gm, st, 3e-4)
    for a, b in zip(ps, mine):
        assert torch.allclose(a.detach(), b, rtol=1e-6, atol=1e-7)


def test_all_fixtures_are_covered():
    names = {os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, "*.json"))}
    assert names >= {"ops_kat", "encoder_shapes", "playlmp_birnn_84", "tacorl_bc_84", "tacorl_q_84"}
