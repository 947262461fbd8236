"""GPU tests of the CUDA-graph step runner and the host-prefetch path."""
import pytest
import torch

from oracle import synth as S
from tests.gpu_util import DEV, build_play_lmp, to_dev

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
