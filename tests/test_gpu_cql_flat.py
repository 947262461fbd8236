"""GPU parity tests of the flat-CQL baseline: tacorl_b200.modules.cql.cql_offline_lightning.CQL_Offline
with the discrete-gripper actor of config/module/cql_offline_goal_cond.yaml, against the numbers recorded from the
unmodified reference (tests/golden/cql_flat_*) and against the fp64 oracle."""
import pytest
import torch

from oracle import synth as S
from oracle import tacorl_oracle as O
from tests.gpu_util import DEV, build_cql_flat, cql_tape, double_params, load_golden, parity_report, to_dev

pytestmark = pytest.mark.gpu
KEYS = ["alpha", "alpha_loss", "actor_loss", "q1_loss", "q2_loss", "bellman_q1_loss", "bellman_q2_loss",
        "conservative_q1_loss", "conservative_q2_loss", "alpha_prime", "alpha_prime_loss", "q1_data", "q1_random",
        "q1_policy", "q2_data", "q2_random", "q2_policy"]


def _module(rec, **kw):
    m = build_cql_flat(**kw)
    m.train()
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == rec["shapes"]
    m.load_state_dict(S.synth_state_dict(rec["shapes"], rec["seed"]), strict=True)
    m.to(DEV)
    m.current_epoch = rec["epoch"]
    m.optimizers()
    return m


@pytest.mark.parametrize("name", ["cql_flat_bc", "cql_flat_q"])
def test_flat_cql_steps_match_reference_golden(name):
    from tacorl_b200.utils.rng import noise_tape
    rec = load_golden(name)
    m = _module(rec)
    batch = S.synth_cql_batch(rec["B"], rec["H"], rec["W"], rec["seed"])
    for s, step in enumerate(rec["steps"]):
        torch.manual_seed(rec["noise_seed_base"] + s)
        noise = O.draw_cql_noise(rec["B"])
        with noise_tape(cql_tape(noise)) as tape:
            m.training_step(to_dev(S.clone_batch(batch)), s)
            assert len(tape) == 0, "every reference draw must be consumed, in order"
        for k in KEYS:
            got, want = float(m.logged["train/" + k]), step["scalars"][k]
            assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (name, s, k, got, want)
        sd = m.state_dict()
        bad = [k for k, fp in step["params"].items() if not S.fingerprint_close(S.fingerprint(sd[k]), fp, 2e-4)]
        assert not bad, (name, s, bad[:8])


@pytest.mark.parametrize("epoch", [0, 7])
def test_flat_cql_gradients_vs_fp64_oracle(epoch):
    """One step at 200x200 with 8 transitions: every logged scalar within 1e-4 of the fp64 oracle, every gradient the
    optimisers consume within 1e-4 (or no further from fp64 than 3x the fp32 CPU oracle is, see below)."""
    from tacorl_b200.utils.rng import noise_tape
    rec = load_golden("cql_flat_bc")
    B, H, W, seed = 8, 200, 200, 41
    rec = dict(rec, epoch=epoch)
    m = _module(rec)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    batch = S.synth_cql_batch(B, H, W, seed)
    torch.manual_seed(77)
    noise = O.draw_cql_noise(B)
    cfg = {"target_entropy": m.target_entropy, "deterministic_backup": False}

    def oracle(dtype):
        P = double_params(sd) if dtype == torch.float64 else O.params_from(sd)
        cast = lambda t: t.to(dtype) if t.dtype.is_floating_point else t
        bt = {k: ({kk: {m_: cast(x) for m_, x in vv.items()} for kk, vv in v.items()} if isinstance(v, dict) else cast(v))
              for k, v in S.clone_batch(batch).items()}
        nz = {k: v.to(dtype) for k, v in noise.items()}
        # alpha after its own Adam step is what the actor loss / backup read
        out0 = O.cql_losses(P, bt, nz, cfg, epoch)
        g_alpha = torch.autograd.grad(out0["alpha_loss"], P["log_alpha"])[0]
        with torch.no_grad():
            st = O.new_adam_state([P["log_alpha"]])
            O.adam_step([P["log_alpha"]], [g_alpha], st, 1e-4)
        out = O.cql_losses(P, bt, nz, cfg, epoch)
        out["alpha_loss"] = out0["alpha_loss"]               # logged before log_alpha steps (:446-457)
        groups = O.tacorl_param_groups(P)
        grads = {}
        for grp, loss in (("actor", "actor_loss"), ("q1", "q1_loss"), ("q2", "q2_loss")):
            ps = [P[k] for k in groups[grp]]
            gs = torch.autograd.grad(out[loss], ps, retain_graph=True, allow_unused=True)
            grads.update({k: (torch.zeros_like(P[k]) if g is None else g) for k, g in zip(groups[grp], gs)})
        return out, grads

    out64, g64 = oracle(torch.float64)
    out32, g32 = oracle(torch.float32)
    # CUDA: keep the gradients autograd deposits (p.grad) by turning the three clipped Adam steps into no-ops
    for o in m.optimizers()[1:4]:
        o.step = lambda *a, **k: None
    with noise_tape(cql_tape(noise)) as tape:
        m.training_step(to_dev(S.clone_batch(batch)), 0)
        assert len(tape) == 0
    rows = []
    for k in KEYS:
        got, want = float(m.logged["train/" + k]), float(out64[k])
        rows.append({"scalar": k, "cuda": got, "fp64": want})
        assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (epoch, k, got, want)
    named = dict(m.named_parameters())
    worst = 0.0
    for k, g in g64.items():
        p = named[k]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        ref = g.to(DEV)
        err = float((got.double() - ref).norm() / (ref.norm() + 1e-30))
        err32 = float((g32[k].double() - g).norm() / (g.norm() + 1e-30))
        rows.append({"grad": k, "rel_err_vs_fp64": err, "fp32_cpu_oracle_vs_fp64": err32, "norm": float(g.norm())})
        if float(g.norm()) < 1e-12:
            assert float(got.norm()) < 1e-9, k
            continue
        # tensors the fp32 CPU evaluation itself misses by more than 1e-4 are cancellation-dominated (BC epoch, gripper
        # head: alpha * (onehot_sampled - p) - (onehot_data - p) with alpha = 0.9999 and mostly equal one-hots; conv
        # stack: long fp32 sums): there the bar is the fp32 reference's own distance from fp64, times 3
        assert err <= max(1e-4, 3.0 * err32), (epoch, k, err, err32)
        worst = max(worst, err)
    parity_report(f"cql_flat_fp32_epoch{epoch}", rows)


def test_flat_cql_validation_step_changes_nothing():
    rec = load_golden("cql_flat_q")
    m = _module(rec)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    m.validation_step(to_dev(S.synth_cql_batch(rec["B"], rec["H"], rec["W"], rec["seed"])))
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert "validation/q1_loss" in m.logged


def test_discrete_gripper_actor_entry_points():
    """Actor.get_actions / sample_n_with_log_prob / log_prob (actor.py:66-156) against the oracle's restatement."""
    from tacorl_b200.utils.rng import noise_tape
    rec = load_golden("cql_flat_bc")
    m = _module(rec)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    P = O.params_from(sd)
    B = 16
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 64, generator=g)
    mu, std, lg = O.mlp_policy_gripper(P, "actor.actor.policy.", x)
    actor = m.actor.actor
    xd = x.to(DEV)
    # deterministic
    a, lp = actor.get_actions(xd, deterministic=True)
    want = torch.cat([torch.tanh(mu), torch.argmax(lg, -1, keepdim=True).float() * 2 - 1], -1)
    assert torch.allclose(a.cpu(), want, atol=1e-5) and float(lp.abs().sum()) == 0.0
    # stochastic, both paths
    eps, u = torch.randn(B, 6, generator=g), torch.rand(B, 2, generator=g)
    for reparam in (True, False):
        with noise_tape([eps, u]) as tape:
            a, lp = actor.get_actions(xd, deterministic=False, reparameterize=reparam)
            assert len(tape) == 0
        z = mu + std * eps
        gi = O.gumbel_argmax(lg, u, clamp=reparam)
        want_a = torch.cat([torch.tanh(z), gi.unsqueeze(-1).float() * 2 - 1], -1)
        want_lp = O.tanh_normal_log_prob(mu, std, pre_tanh=z) + O.gripper_log_prob(lg, gi)
        assert torch.allclose(a.detach().cpu(), want_a, atol=1e-5)
        assert torch.allclose(lp.detach().cpu(), want_lp.detach(), rtol=1e-4, atol=1e-4)
    # n samples
    n = 4
    epsn, un = torch.randn(n, B, 6, generator=g), torch.rand(n, B, 2, generator=g)
    with noise_tape([epsn, un]) as tape, torch.no_grad():       # (the reference's only call site: :265-268, no_grad)
        a, lp = actor.sample_n_with_log_prob(xd, n_actions=n)
        assert len(tape) == 0
    zz = mu + std * epsn
    gi = O.gumbel_argmax(lg.unsqueeze(0).expand(n, B, 2), un, clamp=False)
    assert a.shape == (n, B, 7) and lp.shape == (n, B, 1)
    assert torch.allclose(a.cpu(), torch.cat([torch.tanh(zz), gi.unsqueeze(-1).float() * 2 - 1], -1).detach(), atol=1e-5)
    want_lp = O.tanh_normal_log_prob(mu, std, pre_tanh=zz) + O.gripper_log_prob(lg, gi)
    assert torch.allclose(lp.cpu(), want_lp.detach(), rtol=1e-4, atol=1e-4)
    # log_prob of data actions, with the gradient w.r.t. the gripper head
    acts = torch.rand(B, 7, generator=g) * 2 - 1
    acts[:, -1] = torch.where(acts[:, -1] > 0, 1.0, -1.0)
    lp = actor.log_prob(xd, acts.to(DEV))
    want = O.tanh_normal_log_prob(mu, std, value=acts[:, :-1]) + O.gripper_log_prob(lg, acts[:, -1] / 2 + 0.5)
    assert torch.allclose(lp.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-4)
    lp.sum().backward()
    gw = torch.autograd.grad(want.sum(), P["actor.actor.policy.gripper_action.weight"])[0]
    got = actor.policy.gripper_action.weight.grad
    assert float((got.cpu() - gw).norm() / gw.norm()) < 1e-4
