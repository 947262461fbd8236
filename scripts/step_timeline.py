"""Kernel timeline of ONE replay of the graph-captured training step (CUPTI through torch.profiler; no nsys in the
image).  Prints, per kernel name: launches, busy time, and the wall-clock picture of the step: span, union of busy
intervals (any stream), idle gaps, and the largest gaps with their neighbours.  Un-profiled numbers come from bench.py;
this is for shares and gaps only.

    python scripts/step_timeline.py [play_lmp|tacorl] [bf16|fp32] [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "play_lmp"
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    out = sys.argv[3] if len(sys.argv) > 3 else None
    sys.argv = sys.argv[:1]
    args = bench.parse()
    ctx = bench.Ctx(args)
    from tacorl_b200 import runtime
    wl = bench.WORKLOADS[name]
    m, opts, fn = bench.build_ours(wl, ctx.dev, 1, prec)
    host = bench.host_batch(wl, args.batch, 1, True)
    g = runtime.GraphedTrainStep(fn, bench.to_device(host, ctx.dev), device=ctx.dev, warmup=3)
    for _ in range(5):
        g()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    if not ks:
        print("no kernel events captured")
        return
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    span = t1 - t0
    busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
    gaps = []
    last_name = ks[0][2]
    for s, e, n in ks[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, last_name, n, cur_e - t0))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
        last_name = n if e >= cur_e else last_name
    busy += cur_e - cur_s
    per = {}
    for s, e, n in ks:
        short = n.split("(")[0].replace("void ", "").replace("tacorl::", "")[:60]
        d = per.setdefault(short, [0, 0.0])
        d[0] += 1
        d[1] += e - s
    print(f"{name}/{prec}: {len(ks)} device activities, span {span:.1f} us, busy (union over streams) {busy:.1f} us, "
          f"idle {span - busy:.1f} us in {len(gaps)} gaps (median {sorted(g_[0] for g_ in gaps)[len(gaps) // 2]:.2f} us)")
    print("| kernel | launches | busy us | share of span |")
    print("|---|---|---|---|")
    for n, (c, d) in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{n}` | {c} | {d:.1f} | {100 * d / span:.1f}% |")
    print("largest gaps (us, after kernel -> before kernel, at us):")
    for gdur, a, b, at in sorted(gaps, key=lambda t: -t[0])[:12]:
        print(f"  {gdur:7.2f}  {a.split('(')[0][-40:]} -> {b.split('(')[0][-40:]}  @{at:.0f}")
    if out:
        json.dump([{"start_us": s - t0, "dur_us": e - s, "name": n.split("(")[0]} for s, e, n in ks], open(out, "w"))


if __name__ == "__main__":
    main()
