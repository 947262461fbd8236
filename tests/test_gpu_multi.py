"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a single-GPU box): N ranks over NCCL reproduce the 1-rank
global-batch gradients and parameters after the step, through the overlapped, graph-captured all-reduce path."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(workload, precision, world=2):
    port = str(29600 + os.getpid() % 300)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, WORLD_SIZE=str(world), NCCL_DEBUG="WARN")
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "workers", "dp_equivalence_worker.py"),
                                       workload, precision], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append(p.communicate()[0] + "\n[timeout]")
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    assert "DP_EQUIV_OK" in outs[0], outs[0][-3000:]
    line = [l for l in outs[0].splitlines() if l.startswith("DP_EQUIV_OK")][0]
    out = os.path.join(ROOT, "gpurun_out", "parity")
    try:
        os.makedirs(out, exist_ok=True)
        open(os.path.join(out, f"dp_equivalence_{workload}_{precision}_n{world}.json"), "w").write(line[len("DP_EQUIV_OK "):])
    except OSError:
        pass


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("workload,precision", [("play_lmp", "fp32"), ("tacorl", "fp32"), ("play_lmp", "bf16")])
def test_two_ranks_reproduce_the_global_batch_step(workload, precision):
    _run(workload, precision, 2)
