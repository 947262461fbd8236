"""TEST INFRASTRUCTURE — generates tests/golden/*.json by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_loader.py) on deterministic
synthetic parameters / batches / seeds.  Run in the build container only:

    python -m oracle.make_golden

The fixtures hold seeds, shapes, the reference's logged scalars and fingerprints
(sum, l2, probe-dot; oracle/synth.py) of every gradient and of every parameter after a few
optimiser steps.  tests/test_oracle_golden.py replays them against oracle/tacorl_oracle.py
without the reference being present (it does not exist on the GPU box).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as R  # noqa: E402
from oracle import synth as S  # noqa: E402
from oracle import tacorl_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _shapes(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def _load_synth(module, seed):
    sd = S.synth_state_dict(_shapes(module), seed)
    module.load_state_dict(sd, strict=True)
    return sd


def _scalars(logged, prefix="train/"):
    return {k[len(prefix):]: float(v) for k, v in logged.items() if k.startswith(prefix)}


def golden_play_lmp(name, pr_kind, modalities, B, T, H, W, rnn_hidden, steps=3, seed=11,
                    pad=False, dropout_p=0.0):
    torch.manual_seed(0)
    m = R.build_reference_play_lmp(pr_kind=pr_kind, modalities=modalities, rnn_hidden=rnn_hidden,
                                   dropout_p=dropout_p, max_window=T)
    shapes = _shapes(m)
    _load_synth(m, seed)
    batch = S.synth_play_batch(B, T, H, W, seed, modalities=modalities, pad=pad)
    opt = m.configure_optimizers()
    rec = {"kind": "play_lmp", "pr_kind": pr_kind, "modalities": list(modalities), "B": B, "T": T,
           "H": H, "W": W, "rnn_hidden": rnn_hidden, "seed": seed, "pad": pad, "shapes": shapes,
           "noise_seed_base": 1000, "steps": []}
    for s in range(steps):
        torch.manual_seed(1000 + s)
        opt.zero_grad()
        loss = m.training_step(S.clone_batch(batch), s)
        loss.backward()
        step = {"scalars": _scalars(m.logged),
                "grads": {k: S.fingerprint(p.grad) for k, p in m.named_parameters()
                          if p.grad is not None}}
        opt.step()
        step["params"] = {k: S.fingerprint(v) for k, v in m.state_dict().items()
                          if v.dtype.is_floating_point}
        rec["steps"].append(step)
    json.dump(rec, open(os.path.join(OUT, name + ".json"), "w"))
    print("wrote", name, rec["steps"][0]["scalars"])


def golden_tacorl(name, pr_kind, B, T, H, W, rnn_hidden, epoch, steps=2, seed=13, modalities=("rgb_static",),
                  goal_modalities=None, latent_plan_dim=16, dropout_p=0.0):
    """dropout_p > 0 with the transformer recogniser pins the eval-mode call of the frozen LMP (tacorl.py:237-238):
    the reference must NOT draw dropout masks inside get_pr_latent_plan."""
    torch.manual_seed(0)
    lmp = R.build_reference_play_lmp(pr_kind=pr_kind, rnn_hidden=rnn_hidden, dropout_p=dropout_p,
                                     max_window=T, modalities=modalities, goal_modalities=goal_modalities,
                                     latent_plan_dim=latent_plan_dim)
    t = R.build_reference_tacorl(lmp)
    t.train()
    t.current_epoch = epoch
    shapes = _shapes(t)
    _load_synth(t, seed)
    # targets start as copies of q1/q2 in the reference (tacorl.py:121-122); synthetic params
    # are independent per key, which exercises Polyak more strongly — keep them independent.
    goal_mods = list(goal_modalities) if goal_modalities is not None else list(modalities)[:1]
    batch = S.synth_play_batch(B, T, H, W, seed, modalities=modalities, with_goal=True, goal_modalities=goal_mods)
    batch["disp"][0] = 1
    batch["disp"][1] = -1
    rec = {"kind": "tacorl", "pr_kind": pr_kind, "B": B, "T": T, "H": H, "W": W,
           "rnn_hidden": rnn_hidden, "seed": seed, "epoch": epoch, "shapes": shapes,
           "disp": batch["disp"].tolist(), "noise_seed_base": 2000,
           "modalities": list(modalities), "goal_modalities": goal_mods, "latent_plan_dim": latent_plan_dim,
           "dropout_p": dropout_p, "target_entropy": float(t.target_entropy), "steps": []}
    for s in range(steps):
        torch.manual_seed(2000 + s)
        t.training_step(S.clone_batch(batch))
        rec["steps"].append({
            "scalars": _scalars(t.logged),
            "params": {k: S.fingerprint(v) for k, v in t.state_dict().items()
                       if v.dtype.is_floating_point}})
    json.dump(rec, open(os.path.join(OUT, name + ".json"), "w"))
    print("wrote", name, {k: round(v, 5) for k, v in rec["steps"][0]["scalars"].items()})


def golden_cql(name, B, H, W, epoch, steps=2, seed=29, **overrides):
    """Flat-CQL baseline: CQL_Offline of config/experiment/cql_real_world.yaml, discrete-gripper actor.
    (seed 29: with seeds 23 / 31 the clipped gripper-bias / conv-bias gradients of consecutive steps nearly cancel inside
    Adam's first moment, and fp32 summation order alone moves the parameter by 2e-4 -- an ill-conditioned fixture.)"""
    torch.manual_seed(0)
    m = R.build_reference_cql(**overrides)
    m.train()
    m.current_epoch = epoch
    shapes = _shapes(m)
    _load_synth(m, seed)
    batch = S.synth_cql_batch(B, H, W, seed)
    rec = {"kind": "cql_flat", "B": B, "H": H, "W": W, "seed": seed, "epoch": epoch, "shapes": shapes,
           "noise_seed_base": 3000, "target_entropy": float(m.target_entropy),
           "deterministic_backup": bool(m.deterministic_backup), "steps": []}
    for s in range(steps):
        torch.manual_seed(3000 + s)
        m.training_step(S.clone_batch(batch), s)
        rec["steps"].append({
            "scalars": _scalars(m.logged),
            "params": {k: S.fingerprint(v) for k, v in m.state_dict().items() if v.dtype.is_floating_point}})
    json.dump(rec, open(os.path.join(OUT, name + ".json"), "w"))
    print("wrote", name, {k: round(v, 5) for k, v in rec["steps"][0]["scalars"].items()})


def golden_encoder(name, sizes, seed=17):
    R.import_reference()
    from tacorl.networks.visual_encoders.encoder import LMPVisionEncoder
    torch.manual_seed(0)
    enc = LMPVisionEncoder()
    shapes = _shapes(enc)
    sd = S.synth_state_dict(shapes, seed)
    sd["model.6.temperature"] = torch.tensor([0.7])   # exercise the temperature path
    enc.load_state_dict(sd)
    rec = {"kind": "encoder", "seed": seed, "shapes": shapes, "temperature": 0.7, "cases": []}
    for (n, h, w) in sizes:
        x = S.synth_images((n, 3, h, w), seed, f"enc{h}x{w}")
        enc.zero_grad()
        y = enc(x)
        cot = torch.rand(y.shape, generator=S._gen(seed, f"cot{h}x{w}")) * 2 - 1
        (y * cot).sum().backward()
        feat = enc.model(x)
        rec["cases"].append({"n": n, "h": h, "w": w, "out": S.fingerprint(y),
                             "softargmax": S.fingerprint(feat),
                             "out_head": y[0, :8].tolist(),
                             "grads": {k: S.fingerprint(p.grad) for k, p in enc.named_parameters()}})
    json.dump(rec, open(os.path.join(OUT, name + ".json"), "w"))
    print("wrote", name)


def golden_ops(name, seed=19):
    """Known-answer vectors for the loss / distribution primitives, incl. the edge branches of
    the discretised logistic mixture (action_decoder_logistic.py:220-232)."""
    R.import_reference()
    from tacorl.networks.action_decoders.action_decoder_logistic import ActionDecoderLogistic
    from tacorl.utils.distributions import TanhNormal
    g = S._gen(seed, "ops")
    dec = ActionDecoderLogistic(state_dim=32, latent_plan_dim=16, hidden_size=8)
    B, T = 3, 5
    lp = torch.randn(B, T, 6, 10, generator=g)
    ls = torch.randn(B, T, 6, 10, generator=g) * 2 - 2      # some below the -5 clamp
    ls[0, 0] = -7.0
    mu = torch.randn(B, T, 6, 10, generator=g) * 0.5
    grip = torch.randn(B, T, 2, generator=g)
    act = torch.rand(B, T, 7, generator=g) * 2 - 1
    act[0, 0, 0] = -1.0          # lower edge branch
    act[0, 0, 1] = 1.0           # upper edge branch
    act[0, 1, 2] = 0.9995        # upper edge (> 1 - 1e-3)
    mu[1, 0] = 5.0               # far mean -> cdf_delta < 1e-5 -> mid-pdf branch
    ls[1, 0] = -4.0
    act[..., -1] = torch.where(act[..., -1] > 0, 1.0, -1.0)
    loss = dec._loss(lp, ls, mu, grip, act)
    u1 = torch.rand(B, T, 6, 10, generator=g)
    u2 = torch.rand(B, T, 6, generator=g)
    mean = torch.randn(4, 16, generator=g)
    std = torch.rand(4, 16, generator=g) + 0.1
    z = torch.randn(4, 16, generator=g) * 2
    tn = TanhNormal(mean, std)
    val = torch.rand(4, 16, generator=g) * 2.2 - 1.1        # some outside ±0.999
    rec = {
        "kind": "ops", "lp": lp.tolist(), "ls": ls.tolist(), "mu": mu.tolist(),
        "grip": grip.tolist(), "act": act.tolist(), "dlm_loss": float(loss),
        "logistic_loss": float(dec._logistic_loss(lp, ls, mu, act[:, :, :-1])),
        "u1": u1.tolist(), "u2": u2.tolist(),
        "mean": mean.tolist(), "std": std.tolist(), "z": z.tolist(), "val": val.tolist(),
        "logp_pre": tn.log_prob(torch.tanh(z), pre_tanh_value=z).tolist(),
        "logp_val": tn.log_prob(val).tolist(),
    }
    # _sample consumes torch.rand twice; feed the same numbers through a patched torch.rand
    seq = [u1, u2]
    orig = torch.rand
    torch.rand = lambda *a, **k: seq.pop(0)
    try:
        # forward() clamps log_scales at -5 before _sample sees them (action_decoder_logistic.py:292)
        rec["sample"] = dec._sample(lp, torch.clamp(ls, min=-5.0), mu, grip).tolist()
    finally:
        torch.rand = orig
    json.dump(rec, open(os.path.join(OUT, name + ".json"), "w"))
    print("wrote", name, rec["dlm_loss"])


FIXTURES = {
    "ops_kat": lambda: golden_ops("ops_kat"),
    "encoder_shapes": lambda: golden_encoder("encoder_shapes", [(2, 84, 84), (2, 128, 128), (1, 150, 200), (1, 200, 200)]),
    "playlmp_birnn_84": lambda: golden_play_lmp("playlmp_birnn_84", "tanh_net", ("rgb_static",), 3, 8, 84, 84, 64),
    "playlmp_birnn_pad_128": lambda: golden_play_lmp("playlmp_birnn_pad_128", "tanh_net", ("rgb_static",), 2, 16, 128,
                                                     128, 96, steps=2, pad=True),
    "playlmp_multiview": lambda: golden_play_lmp("playlmp_multiview", "tanh_net", ("rgb_static", "rgb_gripper"), 2, 8,
                                                 96, 128, 64, steps=2),
    "playlmp_transformer_84": lambda: golden_play_lmp("playlmp_transformer_84", "transformer", ("rgb_static",), 3, 8,
                                                      84, 84, 64, steps=2, dropout_p=0.0),
    "tacorl_bc_84": lambda: golden_tacorl("tacorl_bc_84", "tanh_net", 4, 8, 84, 84, 64, epoch=0),
    "tacorl_q_84": lambda: golden_tacorl("tacorl_q_84", "tanh_net", 4, 8, 84, 84, 64, epoch=7),
    "tacorl_defaultpr_84": lambda: golden_tacorl("tacorl_defaultpr_84", "default", 3, 8, 84, 84, 64, epoch=7),
    # round 2: the shipped plan recogniser (transformer, dropout 0.1) under TACORL: the frozen LMP runs in eval mode
    "tacorl_transformer_84": lambda: golden_tacorl("tacorl_transformer_84", "transformer", 3, 8, 84, 84, 64, epoch=7,
                                                   dropout_p=0.1),
    # round 2: BASELINE configs[3] (tacorl_real_world on a play_lmp_gripper_real_world LMP): static + gripper views
    # for observation AND goal, latent plan 32 (experiment/play_lmp_real_world.yaml:10)
    "tacorl_multiview_bc": lambda: golden_tacorl("tacorl_multiview_bc", "tanh_net", 3, 8, 96, 128, 64, epoch=0,
                                                 modalities=("rgb_static", "rgb_gripper"),
                                                 goal_modalities=("rgb_static", "rgb_gripper"), latent_plan_dim=32),
    "tacorl_multiview_q": lambda: golden_tacorl("tacorl_multiview_q", "tanh_net", 3, 8, 96, 128, 64, epoch=7,
                                                modalities=("rgb_static", "rgb_gripper"),
                                                goal_modalities=("rgb_static", "rgb_gripper"), latent_plan_dim=32),
    "cql_flat_bc": lambda: golden_cql("cql_flat_bc", 4, 84, 84, epoch=0),
    "cql_flat_q": lambda: golden_cql("cql_flat_q", 4, 84, 84, epoch=7),
}


def main():
    """python -m oracle.make_golden [name ...]   (no names: every fixture)"""
    os.makedirs(OUT, exist_ok=True)
    for name in (sys.argv[1:] or list(FIXTURES)):
        FIXTURES[name]()


if __name__ == "__main__":
    main()
