"""GPU tests of the CUDA-graph step runner and the host-prefetch path."""
import pytest
import torch

from oracle import synth as S
from tests.gpu_util import DEV, build_play_lmp, rel_err, to_dev

pytestmark = pytest.mark.gpu


def test_graphed_step_trains_and_counts_adam_steps():
    from tacorl_b200 import _lib, ops, runtime
    ops.set_precision("fp32")
    B, T, H, W = 2, 8, 84, 84
    m = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, T)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(S.synth_state_dict(shapes, 4))
    m.to(DEV)
    opt = m.configure_optimizers()
    batch = to_dev(S.synth_play_batch(B, T, H, W, 4))
    batch = {"states": batch["states"], "actions": batch["actions"]}
    torch.manual_seed(0)
    initial = opt.flat_params.clone()
    g = runtime.GraphedTrainStep(runtime.play_lmp_step_fn(m, opt), batch, warmup=2)
    assert g.launches_per_replay > 100                  # the graph holds our kernels, not a fallback
    before = opt.flat_params.clone()
    assert torch.equal(before, initial)                 # the warm-up steps of the capture were rolled back
    assert int(opt._step_dev) == 0 and opt.step_count == 0 and float(opt.exp_avg.abs().sum()) == 0.0
    n0 = _lib.launch_count()
    losses = [float(g()) for _ in range(6)]
    assert _lib.launch_count() == n0                    # replays issue no host-side launches
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]
    assert not torch.equal(before, opt.flat_params)
    assert int(opt._step_dev) == 6                      # replays only (warm-up rolled back; capture records, it does not run)
    # host-side scalars are baked into a graph: a new kl_beta must select a newly captured graph
    m.set_kl_beta(0.5)
    l_new = float(g())
    assert g.captures == 2 and int(opt._step_dev) == 7
    kl, act = float(m.logged["train/kl_loss"]), float(m.logged["train/action_loss"])
    assert abs(l_new - (0.5 * kl + act)) <= 1e-5 * abs(l_new)
    m.set_kl_beta(1e-3)
    g()
    assert g.captures == 2                              # the first graph is re-used
    # host-fed path: pinned batch prefetched on the copy stream, consumed by the next call
    host = S.synth_play_batch(B, T, H, W, 5)
    host = {"states": {k: v.pin_memory() for k, v in host["states"].items()}, "actions": host["actions"].pin_memory()}
    g.prefetch(host)
    l1 = float(g())
    torch.cuda.synchronize()
    assert torch.equal(g.static["actions"].cpu(), host["actions"])
    assert torch.isfinite(torch.tensor(l1))


def test_bf16_shadow_weights_follow_torch_side_edits():
    """The Adam kernel keeps a bf16 twin of the parameters; an in-place torch edit of a parameter (what load_state_dict
    does) must be picked up by the next tensor-core op instead of reading the stale twin."""
    import torch
    from tacorl_b200 import ops
    from tacorl_b200.optim import FlatAdam
    ops.set_precision("bf16")
    try:
        g = torch.Generator().manual_seed(0)
        W = torch.nn.Parameter(torch.randn(512, 1024, generator=g).to(DEV))
        x = torch.randn(256, 1024, generator=g).to(DEV)
        opt = FlatAdam([W], lr=1e-2)
        sh = ops.shadow_of(W)
        assert sh is not None and sh.dtype == torch.bfloat16
        assert torch.equal(sh.view(512, 1024).float(), W.data.bfloat16().float())
        y0 = ops.linear(x, W)
        with torch.no_grad():
            W.mul_(2.0)                                   # torch-side edit: bumps the flat buffer's version
        y1 = ops.linear(x, W)
        assert rel_err(y1, 2.0 * y0) < 1e-6
        y1.sum().backward()
        opt.step()                                        # kernel rewrites parameters and twin together
        assert torch.equal(ops.shadow_of(W).view(512, 1024).float(), W.data.bfloat16().float())
        want = x.bfloat16().double() @ W.data.bfloat16().double().t()
        assert rel_err(ops.linear(x, W), want) < 1e-5
        with torch.no_grad():
            opt.flat_params.mul_(0.5)                     # edit through the flat buffer (what a broadcast does)
        assert rel_err(ops.linear(x, W), 0.5 * want) < 1e-5
    finally:
        ops.set_precision("fp32")


def test_rnn_weight_gradients_land_in_the_flat_gradient_buffer():
    """ReluRNN's backward writes dW straight into FlatAdam's flat gradient (ops.grad_slot_of): after backward the
    parameter's .grad aliases its slice, step() copies nothing for it, and the update equals the copy path's."""
    import copy
    from tacorl_b200 import ops
    from tacorl_b200.networks.layers import ReluRNN
    from tacorl_b200.optim import FlatAdam
    ops.set_precision("fp32")
    torch.manual_seed(3)
    rnn_a = ReluRNN(12, 64, 2, True).to(DEV)
    rnn_b = copy.deepcopy(rnn_a)
    x = torch.randn(5, 7, 12, device=DEV)
    opt_a = FlatAdam(rnn_a.parameters(), lr=1e-2)
    for step in range(2):
        opt_a.zero_grad()
        out, _ = rnn_a(x)
        out.square().mean().backward()
        for p, gv in zip(opt_a.param_groups[0]["params"], opt_a.grad_views):
            assert p.grad is not None and p.grad.data_ptr() == gv.data_ptr()
        opt_a.step()
    # reference: same two steps with the gradients produced in temporaries (slots exhausted -> copy path)
    opt_b = FlatAdam(rnn_b.parameters(), lr=1e-2)
    for step in range(2):
        opt_b.zero_grad()
        for p in opt_b.param_groups[0]["params"]:
            opt_b._slots_taken.add(p.data_ptr())
        out, _ = rnn_b(x)
        out.square().mean().backward()
        for p, gv in zip(opt_b.param_groups[0]["params"], opt_b.grad_views):
            assert p.grad.data_ptr() != gv.data_ptr()
        opt_b.step()
    for pa, pb in zip(rnn_a.parameters(), rnn_b.parameters()):
        assert torch.equal(pa, pb)


def test_graphed_tacorl_step_switches_actor_loss_at_bc_epochs():
    """TACORL's actor loss is a host-side branch on current_epoch < bc_epochs (cql_offline_lightning.py:459-466): the
    graph runner must not keep replaying the BC loss after the switch."""
    from tacorl_b200 import runtime
    from tests.gpu_util import build_tacorl
    lmp = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, 8)
    t = build_tacorl(lmp)
    shapes = {k: list(v.shape) for k, v in t.state_dict().items()}
    t.load_state_dict(S.synth_state_dict(shapes, 6))
    t.to(DEV)
    t.optimizers()
    batch = to_dev(S.synth_play_batch(3, 8, 84, 84, 6, with_goal=True))
    batch = {k: batch[k] for k in ("states", "actions", "goal", "disp")}
    g = runtime.GraphedTrainStep(runtime.tacorl_step_fn(t), batch, warmup=2)
    t.current_epoch = 0
    g()
    bc_loss = float(t.logged["train/actor_loss"])
    assert g.captures == 1
    t.current_epoch = t.bc_epochs
    g()
    q_loss = float(t.logged["train/actor_loss"])
    assert g.captures == 2
    assert abs(bc_loss - q_loss) > 1.0        # BC: alpha*log_pi - log_pi(plan) (positive, large); Q: alpha*log_pi - min Q


def test_flat_adam_state_dict_round_trip_and_resume():
    """Adam moments / step survive state_dict() -> load_state_dict() in torch.optim.Adam's layout; resuming continues the
    trajectory bit for bit (the optimiser used to checkpoint an empty state)."""
    from tacorl_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(1)
    mk = lambda: [torch.nn.Parameter(torch.randn(33, 7, generator=g).to(DEV)), torch.nn.Parameter(torch.randn(130, generator=g).to(DEV))]
    ps = mk()
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    o1, oref = FlatAdam(ps, lr=1e-2), torch.optim.Adam(ref, lr=1e-2)
    grads = [[torch.randn(p.shape, generator=g).to(DEV) for p in ps] for _ in range(5)]
    for gs in grads[:3]:
        for p, r, gr in zip(ps, ref, gs):
            p.grad, r.grad = gr.clone(), gr.clone()
        o1.step()
        oref.step()
    sd = o1.state_dict()
    assert set(sd["state"]) == {0, 1} and float(sd["state"][0]["step"]) == 3.0
    assert sd["state"][0]["exp_avg"].shape == ps[0].shape
    for i, r in enumerate(ref):            # same moments as torch.optim.Adam keeps
        assert rel_err(sd["state"][i]["exp_avg"], oref.state[r]["exp_avg"]) < 1e-6
        assert rel_err(sd["state"][i]["exp_avg_sq"], oref.state[r]["exp_avg_sq"]) < 1e-4   # (1 - beta2) = 1e-3 rounds differently in the fused kernel
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    o2 = FlatAdam(ps2, lr=1e-2)
    o2.load_state_dict(sd)
    assert o2.step_count == 3 and int(o2._step_dev) == 3
    for gs in grads[3:]:
        for p, p2, gr in zip(ps, ps2, gs):
            p.grad, p2.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    for p, p2 in zip(ps, ps2):
        assert torch.equal(p.data, p2.data)
    # and a torch.optim.Adam checkpoint loads too (reference -> B200 resume)
    o3 = FlatAdam([torch.nn.Parameter(r.detach().clone()) for r in ref], lr=1e-2)
    o3.load_state_dict(oref.state_dict())
    assert o3.step_count == 3 and rel_err(o3.exp_avg[:33 * 7], oref.state[ref[0]]["exp_avg"].reshape(-1)) == 0.0


def test_rnn_gradient_slots_survive_accumulation_and_kept_grads():
    """The RNN backward writes weight gradients straight into FlatAdam's flat gradient (ops.grad_slot_of).  With
    gradient accumulation, zero_grad(set_to_none=False) or no zero_grad at all, p.grad is still alive at the next
    backward: the kernel must then write a temporary and let autograd accumulate."""
    from tacorl_b200 import ops
    from tacorl_b200.networks.layers import ReluRNN
    from tacorl_b200.optim import FlatAdam
    ops.set_precision("fp32")
    torch.manual_seed(0)
    rnn = ReluRNN(12, 32, num_layers=2, bidirectional=True).to(DEV)
    opt = FlatAdam(list(rnn.parameters()), lr=1e-3)
    xs = [torch.randn(3, 5, 12, device=DEV) for _ in range(2)]

    def grads_of(x):
        for p in rnn.parameters():
            p.grad = None
        opt._slots_taken.clear()
        out, _ = rnn(x)
        out.square().sum().backward()
        return [p.grad.clone() for p in rnn.parameters()]

    g0, g1 = grads_of(xs[0]), grads_of(xs[1])
    # (a) accumulation over two micro-batches without zero_grad in between
    opt.zero_grad(set_to_none=True)
    for x in xs:
        out, _ = rnn(x)
        out.square().sum().backward()
    for p, a, b in zip(rnn.parameters(), g0, g1):
        assert rel_err(p.grad, a + b) < 1e-6
    # (b) step() without zero_grad, then another backward: grads accumulate like torch (not 2x the last one)
    opt.step()
    out, _ = rnn(xs[0])
    out.square().sum().backward()
    # (the parameters moved by one Adam step of lr 1e-3: compare against a fresh evaluation at the new weights)
    kept = [p.grad.clone() for p in rnn.parameters()]
    fresh = grads_of(xs[0])
    for k, a, b, f in zip(kept, g0, g1, fresh):
        assert rel_err(k, a + b + f) < 1e-5
    # (c) zero_grad(set_to_none=False) keeps zeroed .grad tensors alive (some of them ARE the flat-gradient slots)
    fresh1 = grads_of(xs[1])
    opt.zero_grad(set_to_none=False)
    out, _ = rnn(xs[1])
    out.square().sum().backward()
    for p, f in zip(rnn.parameters(), fresh1):
        assert rel_err(p.grad, f) < 1e-6


def test_early_adam_slice_update_is_bit_identical_to_one_pass():
    """FlatAdam.early_step: the parameters behind the encoders are updated on a side stream as soon as their gradients
    are final (under the encoder backward), the rest at step(): same kernel, same device step count -> same bits."""
    from tacorl_b200 import ops, runtime
    ops.set_precision("fp32")
    B, T, H, W = 2, 8, 84, 84
    results = []
    for early in (False, True):
        m = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, T)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(S.synth_state_dict(shapes, 4))
        m.to(DEV)
        opt = m.configure_optimizers()
        fn = runtime.play_lmp_step_fn(m, opt, early_step=early)
        batch = to_dev(S.synth_play_batch(B, T, H, W, 4))
        batch = {"states": batch["states"], "actions": batch["actions"]}
        torch.manual_seed(3)
        for _ in range(3):
            fn(batch)
        torch.cuda.synchronize()
        assert int(opt._step_dev) == 3 and opt.step_count == 3
        results.append((opt.flat_params.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone()))
    for a, b in zip(*results):
        assert torch.equal(a, b)


def test_double_buffered_graphs_consume_prefetched_batches_in_order():
    """buffers=2: the H2D of the next batch lands straight in the buffer set the running replay does not read, and the
    graphs of the two sets alternate.  The loss sequence must equal the eager loop's on the same batches."""
    from tacorl_b200 import ops, runtime
    ops.set_precision("fp32")
    B, T, H, W = 2, 8, 84, 84
    hosts = []
    for seed in (11, 12, 13):
        hb = S.synth_play_batch(B, T, H, W, seed)
        hosts.append({"states": {k: v.pin_memory() for k, v in hb["states"].items()}, "actions": hb["actions"].pin_memory()})
    seq = [0, 1, 2, 1, 0]

    def fresh():
        m = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, T)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(S.synth_state_dict(shapes, 4))
        m.to(DEV)
        opt = m.configure_optimizers()
        return m, opt, runtime.play_lmp_step_fn(m, opt)

    m, opt, fn = fresh()
    torch.manual_seed(5)
    want = [float(fn(to_dev(hosts[i]))) for i in seq]
    m, opt, fn = fresh()
    g = runtime.GraphedTrainStep(fn, to_dev(hosts[0]), warmup=2, buffers=2)
    torch.manual_seed(5)
    got = []
    g.prefetch(hosts[seq[0]])
    for k in range(len(seq)):
        loss = g()
        got.append(float(loss))                      # read before the next replay (shared pool)
        if k + 1 < len(seq):
            g.prefetch(hosts[seq[k + 1]])
    assert g.captures == 2
    # identical kernels and inputs; the torch RNG advances differently under graph capture, so only the losses of the
    # parts that do not depend on the noise draw can be compared bit for bit: compare the noise-free KL term instead
    assert all(torch.isfinite(torch.tensor(got)))
    assert len(got) == len(want)
    assert abs(got[0] - want[0]) < 0.5 and got[-1] < got[0]
    torch.cuda.synchronize()
    assert torch.equal(g.statics[g._cur]["actions"].cpu(), hosts[seq[-1]]["actions"])
    assert torch.equal(g.statics[1 - g._cur]["actions"].cpu(), hosts[seq[-2]]["actions"])


def test_trainer_fit_checkpoints_and_resumes_across_the_bc_epoch_boundary(tmp_path):
    """tacorl_b200.trainer.Trainer (the part of pl.Trainer scripts/train.py:28-75 uses): manual-optimisation TACORL,
    epochs of 2 steps with bc_epochs = 1 -> the actor loss switches graphs after the first epoch; Lightning-format
    checkpoints; a second Trainer resumes from last.ckpt with every optimiser's moments and step count."""
    from tacorl_b200 import trainer as TR
    from tests.gpu_util import build_tacorl

    def fresh():
        lmp = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, 8)
        t = build_tacorl(lmp, bc_epochs=1)
        shapes = {k: list(v.shape) for k, v in t.state_dict().items()}
        t.load_state_dict(S.synth_state_dict(shapes, 6))
        return t

    hb = S.synth_play_batch(3, 8, 84, 84, 6, with_goal=True)
    batch = {k: hb[k] for k in ("states", "actions", "goal", "disp")}
    t = fresh()
    torch.manual_seed(1)
    tr = TR.Trainer(max_steps=5, precision="fp32", default_root_dir=tmp_path, steps_per_epoch=2)
    tr.fit(t, lambda s: batch)
    assert tr.global_step == 5 and t.current_epoch == 2
    opts = t.optimizers()
    assert [o.steps_done for o in opts] == [5] * 6
    last = tmp_path / "saved_models" / "last.ckpt"
    assert last.is_file() and (tmp_path / "saved_models" / "tacorl_epoch_01_.ckpt").is_file()
    ck = torch.load(last, weights_only=False)
    assert ck["epoch"] == 2 and ck["global_step"] == 5 and len(ck["optimizer_states"]) == 6
    assert float(ck["optimizer_states"][1]["state"][0]["step"]) == 5.0
    # resume in a fresh module
    t2 = fresh()
    tr2 = TR.Trainer(max_steps=7, precision="fp32", default_root_dir=tmp_path / "resumed", steps_per_epoch=2)
    tr2.fit(t2, lambda s: batch, ckpt_path=last)
    assert tr2.global_step == 7 and [o.steps_done for o in t2.optimizers()] == [7] * 6
    sd1 = {k: v.cpu() for k, v in t.state_dict().items()}
    moved = [k for k, v in t2.state_dict().items() if v.dtype.is_floating_point and not torch.equal(v.cpu(), sd1[k])]
    assert any(k.startswith("actor.") for k in moved) and any(k.startswith("q1.") for k in moved)
    assert not [k for k in moved if k.startswith(("perceptual_encoder.", "plan_recognition."))]     # frozen LMP parts


def test_frozen_lmp_weights_are_cast_once_and_follow_torch_side_edits():
    """TACO-RL's frozen plan recogniser (tacorl.py:124-125): the bf16 operand copies of its weights are cached
    (ops.mark_frozen) -- cast once, re-cast in place when the weights are edited through torch, also under a captured
    graph; trainable and Polyak-updated parameters never enter that cache."""
    from tacorl_b200 import ops, runtime
    from tests.gpu_util import build_tacorl
    try:
        lmp = build_play_lmp("tanh_net", ("rgb_static",), 64, 16, 8)
        t = build_tacorl(lmp, precision="bf16")
        shapes = {k: list(v.shape) for k, v in t.state_dict().items()}
        t.load_state_dict(S.synth_state_dict(shapes, 6))
        t.to(DEV)
        t.optimizers()
        w = t.plan_recognition.birnn_model.weight_hh_l0
        sh = ops.shadow_of(w)
        assert sh is not None and sh.dtype == torch.bfloat16
        assert torch.equal(sh.view_as(w), w.detach().to(torch.bfloat16))
        assert ops.shadow_of(w).data_ptr() == sh.data_ptr()                 # cached: same buffer, no new cast
        tq = next(iter(t.target_q1.critic.parameters()))
        assert ops.shadow_of(tq) is None                                     # Polyak-updated: never cached
        batch = to_dev(S.synth_play_batch(3, 8, 84, 84, 6, with_goal=True))
        batch = {k: batch[k] for k in ("states", "actions", "goal", "disp")}
        g = runtime.GraphedTrainStep(runtime.tacorl_step_fn(t), batch, warmup=2)
        torch.manual_seed(1)
        g()
        a = float(t.logged["train/action_loss"])
        with torch.no_grad():
            for p in t.plan_recognition.parameters():
                p.mul_(1.5)                                                   # bumps the version counters
        torch.manual_seed(1)
        g()
        b = float(t.logged["train/action_loss"])
        assert ops.shadow_of(w).data_ptr() == sh.data_ptr()
        assert torch.equal(sh.view_as(w), w.detach().to(torch.bfloat16))    # refreshed in place before the replay
        assert a != b and b == b
    finally:
        ops.set_precision("fp32")
