"""GPU parity at the size bench.py times: TACORL.training_step with the full networks (RNN hidden 2048, 3x200x200
frames, 16-frame windows; BASELINE configs[2]) against the CPU oracle — fp32 path at 1e-4 against an fp64 run of the
oracle (BC epoch and Q epoch), bf16 path at 1e-2 on every logged scalar over 3 optimiser steps.

Mirrors the update order of /root/reference/src/tacorl/modules/cql/cql_offline_lightning.py:470-542 and
modules/tacorl/tacorl.py:254-273 (decoder Adam -> alpha Adam -> losses with the new alpha -> alpha' Adam -> actor /
q1 / q2 clip + Adam -> Polyak).  Per-tensor numbers are written to gpurun_out/parity/ (copied to profiles/)."""
import pytest
import torch

from oracle import synth as S
from oracle import tacorl_oracle as O
from tests.gpu_util import (DEV, TACORL_KEYS, build_play_lmp, build_tacorl, double_params, parity_report, tacorl_tape,
                            to_dev)

pytestmark = pytest.mark.gpu
B, T, H, W, HID = 8, 16, 200, 200, 2048
GROUPS = ("alpha", "actor", "q1", "q2", "alpha_prime", "decoder")      # optimizers() order (cql…py:553-574 + tacorl.py:289-300)


def _build(precision, epoch, seed=9):
    lmp = build_play_lmp("tanh_net", ("rgb_static",), HID, 16, T)
    t = build_tacorl(lmp, precision)
    shapes = {k: list(v.shape) for k, v in t.state_dict().items()}
    sd = S.synth_state_dict(shapes, seed)
    t.load_state_dict(sd, strict=True)
    t.to(DEV)
    t.train()
    t.current_epoch = epoch
    t.optimizers()
    batch = S.synth_play_batch(B, T, H, W, seed, with_goal=True)
    batch["disp"][0], batch["disp"][1], batch["disp"][2] = 1, -1, 1
    return t, sd, batch


def _clip(gs, max_norm=1.0):
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in gs))
    return float(torch.clamp(max_norm / (total + 1e-6), max=1.0))


@pytest.mark.parametrize("epoch", [0, 7])
def test_tacorl_full_size_fp32_vs_fp64_oracle(epoch):
    from tacorl_b200 import ops
    from tacorl_b200.utils.rng import noise_tape
    t, sd, batch = _build("fp32", epoch)
    torch.manual_seed(300 + epoch)
    noise = O.draw_tacorl_noise(B)
    with noise_tape(tacorl_tape(noise)) as tape:
        t.training_step(to_dev(S.clone_batch(batch)))
        assert len(tape) == 0
    cfg = {"pr_kind": "tanh_net", "target_entropy": t.target_entropy}
    # fp64 arbiter and the reference's own arithmetic (fp32 on the CPU)
    P64 = double_params(sd, O.TACORL_FROZEN)
    b64 = S.clone_batch(batch)
    b64["states"] = {k: v.double() for k, v in b64["states"].items()}
    b64["goal"] = {k: v.double() for k, v in b64["goal"].items()}
    b64["actions"] = b64["actions"].double()
    n64 = {k: v.double() for k, v in noise.items()}
    log64, g64 = O.tacorl_training_step(P64, O.new_tacorl_opt_state(P64), b64, n64, cfg, epoch)
    P32 = O.params_from(sd, O.TACORL_FROZEN)
    log32, g32 = O.tacorl_training_step(P32, O.new_tacorl_opt_state(P32), S.clone_batch(batch), noise, cfg, epoch)
    rows = []
    for k in TACORL_KEYS:
        got, want = float(t.logged["train/" + k]), float(log64[k])
        rows.append({"scalar": k, "cuda_fp32": got, "oracle_fp64": want, "oracle_fp32": float(log32[k]),
                     "rel_err": abs(got - want) / max(1.0, abs(want))})
    # gradients: the optimisers' flat gradient buffers still hold this step's (unclipped) gradients
    names = {id(p): n for n, p in t.named_parameters()}
    groups = O.tacorl_param_groups(P64)
    bad = []
    for gi, (gname, opt) in enumerate(zip(GROUPS, t.optimizers())):
        ps = opt.param_groups[0]["params"]
        mine = [gv.detach().double().cpu() for gv in opt.grad_views]
        assert [names[id(p)] for p in ps] == groups[gname], gname
        coef = _clip(mine) if gname in ("actor", "q1", "q2") else 1.0       # the oracle returns clipped gradients
        for p, g, w64, w32 in zip(ps, mine, g64[gname], g32[gname]):
            n = names[id(p)]
            err = float((g * coef - w64).norm() / (w64.norm() + 1e-30))
            ref_err = float((w32.double() - w64).norm() / (w64.norm() + 1e-30))
            rows.append({"grad": n, "group": gname, "rel_err_vs_fp64": err, "fp32_cpu_oracle_rel_err_vs_fp64": ref_err})
            if err > max(1e-4, 1.5 * ref_err):
                bad.append((n, err, ref_err))
    # parameters after the step (Adam / clip / Polyak), every entry of the state_dict
    new_sd = t.state_dict()
    for k, v in P64.items():
        if not v.dtype.is_floating_point:
            continue
        err = float((new_sd[k].double().cpu() - v.detach()).norm() / (v.detach().norm() + 1e-30))
        rows.append({"param_after_step": k, "rel_err_vs_fp64": err})
        if err > 1e-4:
            bad.append((k, err))
    parity_report(f"tacorl_full_fp32_epoch{epoch}", rows)
    worst_scalar = max(r["rel_err"] for r in rows if "scalar" in r)
    assert worst_scalar <= 1e-4, [r for r in rows if "scalar" in r and r["rel_err"] > 1e-4]
    assert not bad, bad[:10]
    ops.set_precision("fp32")


@pytest.mark.parametrize("epoch", [0, 7])
def test_tacorl_full_size_bf16_three_steps_within_1e2(epoch):
    """The configuration bench.py's TACO-RL workload times (bf16 tensor-core operands, fp32 accumulate), 3 steps."""
    from tacorl_b200 import ops
    from tacorl_b200.utils.rng import noise_tape
    try:
        t, sd, batch = _build("bf16", epoch)
        P = O.params_from(sd, O.TACORL_FROZEN)
        opt = O.new_tacorl_opt_state(P)
        cfg = {"pr_kind": "tanh_net", "target_entropy": t.target_entropy}
        rows = []
        for s in range(3):
            torch.manual_seed(400 + 10 * epoch + s)
            noise = O.draw_tacorl_noise(B)
            with noise_tape(tacorl_tape(noise)) as tape:
                t.training_step(to_dev(S.clone_batch(batch)))
                assert len(tape) == 0
            logged, _ = O.tacorl_training_step(P, opt, S.clone_batch(batch), noise, cfg, epoch)
            for k in TACORL_KEYS:
                got, want = float(t.logged["train/" + k]), float(logged[k])
                rows.append({"step": s, "scalar": k, "cuda_bf16": got, "oracle_fp32": want,
                             "rel_err": abs(got - want) / max(1.0, abs(want))})
        parity_report(f"tacorl_full_bf16_epoch{epoch}", rows)
        bad = [r for r in rows if r["rel_err"] > 1e-2]
        assert not bad, bad
    finally:
        ops.set_precision("fp32")
