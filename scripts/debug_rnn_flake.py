"""Diagnostic: fp32 ReLU-RNN op after a bf16 PlayLMP step in the same process (tests/test_gpu_ops.py flake)."""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import tacorl_oracle as O  # noqa: E402
from tacorl_b200 import _lib, ops  # noqa: E402
from tests import test_gpu_bf16 as TB  # noqa: E402

DEV = "cuda"
variant = sys.argv[1] if len(sys.argv) > 1 else "asis"


def rnn_check(bidir, last_only, use_h0, B=64, T=16, I=32, H=256):
    ops.set_precision("fp32")
    g = torch.Generator().manual_seed(B + T + I + H + bidir)
    rnn = torch.nn.RNN(I, H, num_layers=2, nonlinearity="relu", bidirectional=bidir, batch_first=True)
    sd = {k: v.detach().clone() for k, v in rnn.state_dict().items()}
    names = list(sd.keys())
    x = torch.randn(B, T, I, generator=g)
    D = 2 if bidir else 1
    h0 = torch.randn(2 * D, B, H, generator=g) * 0.3 if use_h0 else None
    P = {"r." + k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    out, hn = O.rnn_stack(P, "r.", x64, 2, bidir, None if h0 is None else h0.double())
    want = out[:, -1] if last_only else out
    cot = torch.randn(want.shape, generator=g)
    (want * cot.double()).sum().backward()
    ws = [sd[k].clone().to(DEV).requires_grad_(True) for k in names]
    xd = x.to(DEV).requires_grad_(True)
    got, hn_d = ops.relu_rnn(xd.transpose(0, 1), ws, 2, bidir, last_only, None if h0 is None else h0.to(DEV))
    if not last_only:
        got = got.transpose(0, 1)
    (got * cot.to(DEV)).sum().backward()
    rel = lambda a, b: float((a.detach().double().cpu() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-30))
    res = {"out": rel(got, want), "dx": rel(xd.grad, x64.grad)}
    for k, w in zip(names, ws):
        res[k] = rel(w.grad, P["r." + k].grad)
    bad = {k: v for k, v in res.items() if v > 1e-4}
    print(f"  rnn bidir={bidir} last={last_only} h0={use_h0}: worst {max(res.values()):.2e}  bad: {bad}")
    if bad and B == 64:
        first = {k: w.grad.clone() for k, w in zip(names, ws)}
        first_dx = xd.grad.clone()
        for again in range(3):
            for w in ws:
                w.grad = None
            xd.grad = None
            got2, _ = ops.relu_rnn(xd.transpose(0, 1), ws, 2, bidir, last_only, None if h0 is None else h0.to(DEV))
            if not last_only:
                got2 = got2.transpose(0, 1)
            (got2 * cot.to(DEV)).sum().backward()
            r2 = {"out": rel(got2, want), "dx": rel(xd.grad, x64.grad)}
            for k, w in zip(names, ws):
                r2[k] = rel(w.grad, P["r." + k].grad)
            print(f"     again {again}: worst {max(r2.values()):.2e}; dx vs first run {rel(xd.grad, first_dx.cpu()):.2e}")
        # the ReLU kink: units whose sign differs between the device's fp32 forward and the fp64 reference
        flips = ((got.detach().cpu() > 0) != (want.detach() > 0)).sum().item()
        tiny = (want.detach().abs() < 1e-6).sum().item()
        print(f"     ReLU gates that differ (top layer output): {flips}; |reference output| < 1e-6 but nonzero-able: {tiny}")
        d = (first_dx - xd.grad).abs()
        nz = (d > 1e-5 * first_dx.abs().max()).nonzero()
        print("     differing dx elements:", nz.shape[0], "of", d.numel(), "first few", nz[:5].tolist(), "prec", ops.get_precision())


print("variant", variant)
print("before the bf16 step:")
rnn_check(False, False, True)
if variant == "bigws":
    _lib.workspace(591383296, torch.device("cuda", 0))
if variant not in ("noprestep", "bigws"):
  ops.set_precision("bf16")
  TB.test_play_lmp_bf16_step_within_1e2_of_fp64_oracle.__wrapped__(ops) if hasattr(
      TB.test_play_lmp_bf16_step_within_1e2_of_fp64_oracle, "__wrapped__") else TB.test_play_lmp_bf16_step_within_1e2_of_fp64_oracle(ops)
ops.set_precision("fp32")
if variant == "gc":
    gc.collect()
if variant == "ws":
    _lib._WS.clear()
if variant == "owners":
    ops._GRAD_OWNERS.clear(); ops._SHADOW_OWNERS.clear()
if variant == "sync":
    torch.cuda.synchronize()
print("after the bf16 step:", "grad owners", len(ops._GRAD_OWNERS), "ws", {k: v.numel() for k, v in _lib._WS.items()})
for rep in range(1 if variant in ('noprestep', 'bigws') else 6):
    for cfg in [(False, False, False), (False, False, True), (True, False, False), (True, True, False)]:
        rnn_check(*cfg, B=4, T=7, I=12, H=40)
        rnn_check(*cfg)
