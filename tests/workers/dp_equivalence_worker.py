"""Worker of tests/test_gpu_multi.py (one process per GPU, NCCL): the data-parallel step -- batch shard per rank,
gradient all-reduce overlapped with the encoder backward, captured in a CUDA graph, 1/world folded into the Adam
kernel -- must reproduce the 1-rank global-batch step (reference: Lightning DDP,
/root/reference/config/trainer/default.yaml:1-4, scripts/train.py:73-75).

argv: <workload play_lmp|tacorl> <precision>.  Prints DP_EQUIV_OK <json> on rank 0."""
import gc
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import synth as S  # noqa: E402
from oracle import tacorl_oracle as O  # noqa: E402
from tacorl_b200 import ops, parallel, runtime  # noqa: E402
from tacorl_b200.utils.rng import noise_tape  # noqa: E402
from tests.gpu_util import build_play_lmp, build_tacorl, play_lmp_tape, tacorl_tape  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def on(dev, batch):
    return {k: ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v.to(dev)) for k, v in batch.items()}


def main():
    workload, precision = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "16")
    os.environ.setdefault("NCCL_MIN_NCHANNELS", "16")
    dist.init_process_group("nccl", device_id=dev)
    ops.set_precision(precision)
    Bg, T, HW, HID, seed = 4 * world, 8, 84, 256, 31
    Bl = Bg // world
    lo, hi = rank * Bl, (rank + 1) * Bl

    def build():
        lmp = build_play_lmp("tanh_net", ("rgb_static",), HID, 16, T)
        m = lmp if workload == "play_lmp" else build_tacorl(lmp, precision)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(S.synth_state_dict(shapes, seed))
        m.to(dev)
        m.train()
        return m

    gbatch = S.synth_play_batch(Bg, T, HW, HW, seed, with_goal=(workload == "tacorl"))
    keys = ("states", "actions") + (("goal", "disp") if workload == "tacorl" else ())
    gbatch = {k: gbatch[k] for k in keys}
    lbatch = parallel.shard_batch(gbatch, rank, world)
    torch.manual_seed(77)
    if workload == "play_lmp":
        gn = O.draw_play_lmp_noise(Bg, T)
        ln = {k: v[lo:hi] for k, v in gn.items()}
        gtape = [t.to(dev) for t in play_lmp_tape(gn, Bg)]
        ltape = [t.to(dev) for t in play_lmp_tape(ln, Bl)]
    else:
        n = 4
        gn = O.draw_tacorl_noise(Bg, n=n)
        ln = {k: (v[lo:hi] if k in ("plan_noise", "eps_actor", "eps_next") else
                  v.view(n, Bg, -1)[:, lo:hi].reshape(-1, v.shape[-1]) if k == "rand_actions" else v[:, lo:hi])
              for k, v in gn.items()}
        gtape = [t.to(dev).contiguous() for t in tacorl_tape(gn)]
        ltape = [t.to(dev).contiguous() for t in tacorl_tape(ln)]

    def with_tape(fn, tape):
        def step(batch):
            with noise_tape(list(tape)):
                return fn(batch)
        step.mode_key = getattr(fn, "mode_key", None)
        step.preserve = getattr(fn, "preserve", None)
        return step

    # ---- (1) this rank alone on the GLOBAL batch, eagerly, no data parallelism
    ref = build()
    ref_opts = ref.optimizers()
    fn = runtime.play_lmp_step_fn(ref, ref_opts[0]) if workload == "play_lmp" else runtime.tacorl_step_fn(ref)
    with_tape(fn, gtape)(on(dev, gbatch))
    torch.cuda.synchronize()

    # ---- (2) data parallel: local shard, overlapped all-reduce, replayed from a captured graph
    m = build()
    opts = m.optimizers()
    for o in opts:
        parallel.attach_data_parallel(o, world)
    fn = runtime.play_lmp_step_fn(m, opts[0]) if workload == "play_lmp" else runtime.tacorl_step_fn(m)
    g = runtime.GraphedTrainStep(with_tape(fn, ltape), on(dev, lbatch), device=dev, warmup=2)
    g()
    torch.cuda.synchronize()
    out = {"workload": workload, "precision": precision, "world": world, "captured_launches": g.launches_per_replay}
    worst_g = worst_p = 0.0
    for o, r in zip(opts, ref_opts):
        worst_g = max(worst_g, rel(o.flat_grad * o.grad_scale, r.flat_grad))
        worst_p = max(worst_p, rel(o.flat_params, r.flat_params))
    out["flat_grad_rel_err"], out["params_after_step_rel_err"] = worst_g, worst_p
    if workload == "tacorl":
        out["target_rel_err"] = max(rel(a.flat, b.flat) for a, b in zip(m._target_bufs, ref._target_bufs))
    # every rank must hold identical parameters after the step
    mine = torch.cat([o.flat_params for o in opts])
    other = mine.clone()
    dist.broadcast(other, src=0)
    out["rank_divergence"] = float((mine - other).abs().max())
    tol = 2e-5 if precision == "fp32" else 2e-2
    ok = worst_g < tol and worst_p < tol and out["rank_divergence"] == 0.0 and out.get("target_rel_err", 0.0) < tol
    # ---- teardown: release the captured graphs (they reference the communicator's streams) BEFORE the group
    del g
    gc.collect()
    torch.cuda.synchronize()
    out["native_nccl"] = bool(parallel.NativeAllReduce._ready)
    parallel.NativeAllReduce.shutdown()
    done = threading.Event()

    def destroy():
        dist.destroy_process_group()
        done.set()

    th = threading.Thread(target=destroy, daemon=True)
    th.start()
    th.join(timeout=60)
    out["destroy_process_group_returned"] = done.is_set()
    if rank == 0:
        print(("DP_EQUIV_OK " if ok else "DP_EQUIV_FAIL ") + json.dumps(out), flush=True)
    sys.stdout.flush()
    os._exit(0 if ok and done.is_set() else 1)


if __name__ == "__main__":
    main()
