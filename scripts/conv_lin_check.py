"""Checks the linear-shift implicit-GEMM convolutions (shifted UMMA descriptors) against torch's conv in fp64 and times
them against the per-tap-box kernels:  python scripts/conv_lin_check.py"""
import ctypes, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacorl_b200 import _lib as L
DEV = "cuda"
def bf(t): return t.bfloat16().float()
def run(op, in0, in1, Wt, bias, N, H, W, out_shape, reps=0):
    out = torch.full(out_shape, float("nan"), device=DEV)
    ws = L.workspace(3 << 30, torch.device(DEV), tag="convdbg")
    args = (op, L.ptr(in0), L.ptr(in1), L.ptr(Wt), L.ptr(bias), N, H, W, L.ptr(out), ctypes.c_void_p(ws.data_ptr()), ws.numel(), L.stream())
    L.call("tacorl_conv_tc_debug", *args)
    torch.cuda.synchronize()
    return out
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
g = torch.Generator().manual_seed(0)
for (N, H, W) in ((3, 84, 84), (2, 200, 200), (2, 150, 200)):
    H1, W1 = (H - 8) // 4 + 1, (W - 8) // 4 + 1
    H2, W2 = (H1 - 4) // 2 + 1, (W1 - 4) // 2 + 1
    H3, W3 = H2 - 2, W2 - 2
    x = bf(torch.rand(N, 3, H, W, generator=g) * 2 - 1).to(DEV)
    W1t = bf(torch.randn(32, 3, 8, 8, generator=g) * 0.1).to(DEV); b1 = (torch.randn(32, generator=g) * 0.1).to(DEV)
    W3t = bf(torch.randn(64, 64, 3, 3, generator=g) * 0.06).to(DEV); b3 = (torch.randn(64, generator=g) * 0.1).to(DEV)
    y1 = F.relu(F.conv2d(x.double(), W1t.double(), b1.double(), stride=4)).permute(0, 2, 3, 1)
    y2 = bf(torch.rand(N, H2, W2, 64, generator=g)).to(DEV)
    y3 = F.relu(F.conv2d(y2.permute(0, 3, 1, 2).double(), W3t.double(), b3.double())).permute(0, 2, 3, 1)
    for op in (1, 11):
        print(f"{N}x{H}x{W} conv1 op {op}: rel err {rel(run(op, x, None, W1t, b1, N, H, W, (N, H1, W1, 32)), y1):.2e}", flush=True)
    for op in (3, 13):
        print(f"{N}x{H}x{W} conv3 op {op}: rel err {rel(run(op, y2, None, W3t, b3, N, H, W, (N, H3, W3, 64)), y3):.2e}", flush=True)

# timing at the bench shape (1024 frames 200x200): whole debug op (includes the staging casts) - compare pairs
N, H, W = 1024, 200, 200
x = (torch.rand(N, 3, H, W, device=DEV) * 2 - 1)
y2 = torch.rand(N, 23, 23, 64, device=DEV)
o1 = (N, 49, 49, 32); o3 = (N, 21, 21, 64)
def t(op, in0, Wt, b, oshape):
    run(op, in0, None, Wt, b, N, H, W, oshape)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run(op, in0, None, Wt, b, N, H, W, oshape)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
for op, in0, Wt, b, osh in ((1, x, W1t, b1, o1), (11, x, W1t, b1, o1), (3, y2, W3t, b3, o3), (13, y2, W3t, b3, o3)):
    print(f"op {op}: {t(op, in0, Wt, b, osh):.3f} ms (debug op incl. staging)")
