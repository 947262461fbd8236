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
    g = runtime.GraphedTrainStep(runtime.play_lmp_step_fn(m, opt), batch, warmup=2)
    assert g.launches_per_replay > 100                  # the graph holds our kernels, not a fallback
    before = opt.flat_params.clone()
    n0 = _lib.launch_count()
    losses = [float(g()) for _ in range(6)]
    assert _lib.launch_count() == n0                    # replays issue no host-side launches
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]
    assert not torch.equal(before, opt.flat_params)
    assert int(opt._step_dev) == 2 + 6                  # warm-up + replays (capture records, it does not run)
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
