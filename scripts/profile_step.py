"""Runs whole training steps eagerly (no CUDA graph) so that ncu can list / capture their kernels:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/profile_step.py play_lmp 2
    ncu --set full --clock-control none --import-source on -k regex:"rnn_wave|mlp_chain" -c 12 -o out python scripts/profile_step.py play_lmp 1

argv: workload (bench.py's names), timed steps (after one un-profiled... ncu sees every launch: use -s to skip warm-up)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "play_lmp"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    sys.argv = sys.argv[:1]
    args = bench.parse()
    ctx = bench.Ctx(args)
    wl = bench.WORKLOADS[name]
    m, opts, fn = bench.build_ours(wl, ctx.dev, 1, prec)
    batch = bench.to_device(bench.host_batch(wl, args.batch, 1, True), ctx.dev)
    from tacorl_b200 import _lib
    for s in range(steps + 1):
        n0 = _lib.launch_count()
        fn(batch)
        torch.cuda.synchronize()
        print(f"step {s}: {_lib.launch_count() - n0} library launches", flush=True)


if __name__ == "__main__":
    main()
