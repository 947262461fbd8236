"""CUDA-event timing of the fused MLP chain (ops.mlp_chain) on the shapes the two training steps use."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacorl_b200 import ops

dev = "cuda"
ops.set_precision("bf16")


def bench(name, rows, ins, widths, acts, two_seg):
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(rows, i, generator=g).to(dev).requires_grad_(True) for i in ins]
    dims = [sum(ins)] + list(widths)
    layers = []
    for l in range(len(widths)):
        mk = lambda n: ((torch.rand(n, dims[l], generator=g) - 0.5).to(dev).requires_grad_(True), torch.zeros(n, device=dev, requires_grad=True))
        layers.append([mk(widths[l] // 2), mk(widths[l] - widths[l] // 2)] if (two_seg and l == len(widths) - 1) else mk(widths[l]))
    cot = torch.randn(rows, widths[-1], device=dev)

    def fwd():
        return ops.mlp_chain(xs if len(xs) > 1 else xs[0], layers, acts)

    def fb():
        out = fwd()
        out.backward(cot)

    for fn, tag in ((lambda: fwd(), "fwd"), (fb, "fwd+bwd")):
        with torch.no_grad() if tag == "fwd" else torch.enable_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"{name:28s} rows {rows:4d} {tag:8s} {1e3 * e0.elapsed_time(e1) / 20:8.1f} us (eager launches, incl. host overhead)")


bench("goal encoder 32-256-256-32", 64, (32,), (256, 256, 32), ("relu", "relu"), False)
bench("policy (32|32)-256x3-(16|16)", 64, (32, 32), (256, 256, 256, 32), ("silu",) * 3, True)
bench("q (64|16)-256x3-1", 832, (64, 16), (256, 256, 256, 1), ("silu",) * 3, False)
bench("q (64|16)-256x3-1", 64, (64, 16), (256, 256, 256, 1), ("silu",) * 3, False)
