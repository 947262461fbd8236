import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_conv_tc import _run, _bf, _geom, nhwc
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
N, H, W = 2, 84, 84
g = torch.Generator().manual_seed(1)
H1, W1, H2, W2, H3, W3 = _geom(H, W)
x = _bf(torch.rand(N, 3, H, W, generator=g) * 2 - 1)
W1t, W2t, W3t = (_bf(torch.randn(s, generator=g) * sc) for s, sc in (((32, 3, 8, 8), 0.1), ((64, 32, 4, 4), 0.06), ((64, 64, 3, 3), 0.06)))
b1, b2, b3 = (torch.zeros(n) for n in (32, 64, 64))
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
def report(tag, got, want):
    got, want = got.double(), want.double()
    print(f"== {tag}: rel {rel(got, want):.3e}")
    err = (got - want).abs()
    # per-pixel error map for frame 0 (sum over channels)
    em = err[0].sum(-1)
    print("  err map rows (frame0, first 12x12):")
    print((em[:12, :12] > 1e-2 * want.abs().mean() * want.shape[-1]).int())
    print("  per-channel err:", [round(float(e), 3) for e in err.sum((0, 1, 2))[:16]])
xd = x.double()
y1 = F.conv2d(xd, W1t.double(), None, stride=4)           # no relu/bias to see raw sums
# individual tap contributions
got = _run(1, x, None, W1t, b1, N, H, W, (N, H1, W1, 32))
report("conv1 fwd (relu)", got, nhwc(F.relu(y1)))
for dy in range(2):
    for dx in range(2):
        Wm = torch.zeros_like(W1t); Wm[:, :, 4*dy:4*dy+4, 4*dx:4*dx+4] = W1t[:, :, 4*dy:4*dy+4, 4*dx:4*dx+4]
        got = _run(1, x, None, Wm, b1, N, H, W, (N, H1, W1, 32))
        want = nhwc(F.relu(F.conv2d(xd, Wm.double(), None, stride=4)))
        print(f"tap ({dy},{dx}) only: rel {rel(got, want):.3e}")
for c in range(3):
    Wm = torch.zeros_like(W1t); Wm[:, c] = W1t[:, c]
    got = _run(1, x, None, Wm, b1, N, H, W, (N, H1, W1, 32))
    want = nhwc(F.relu(F.conv2d(xd, Wm.double(), None, stride=4)))
    print(f"channel {c} only: rel {rel(got, want):.3e}")
for py in range(4):
    Wm = torch.zeros_like(W1t); Wm[:, :, py::4, :] = W1t[:, :, py::4, :]
    got = _run(1, x, None, Wm, b1, N, H, W, (N, H1, W1, 32))
    want = nhwc(F.relu(F.conv2d(xd, Wm.double(), None, stride=4)))
    print(f"py {py} only: rel {rel(got, want):.3e}")
for px in range(4):
    Wm = torch.zeros_like(W1t); Wm[:, :, :, px::4] = W1t[:, :, :, px::4]
    got = _run(1, x, None, Wm, b1, N, H, W, (N, H1, W1, 32))
    want = nhwc(F.relu(F.conv2d(xd, Wm.double(), None, stride=4)))
    print(f"px {px} only: rel {rel(got, want):.3e}")
y1b = _bf(nhwc(F.relu(y1)).float()); y1n = y1b.permute(0, 3, 1, 2).double()
y2 = F.relu(F.conv2d(y1n, W2t.double(), None, stride=2))
report("conv2 fwd", _run(2, y1b, None, W2t, b2, N, H, W, (N, H2, W2, 64)), nhwc(y2))
y2b = _bf(nhwc(y2).float()); y2n = y2b.permute(0, 3, 1, 2).double()
y3 = F.relu(F.conv2d(y2n, W3t.double(), None, stride=1))
report("conv3 fwd", _run(3, y2b, None, W3t, b3, N, H, W, (N, H3, W3, 64)), nhwc(y3))
dy3 = _bf(torch.randn(N, H3, W3, 64, generator=g)); dy3n = dy3.permute(0, 3, 1, 2).double()
dy2 = F.conv_transpose2d(dy3n, W3t.double(), stride=1) * (y2n > 0)
report("conv3 dgrad", _run(4, dy3, y2b, W3t, None, N, H, W, (N, H2, W2, 64)), nhwc(dy2))
dy2b = _bf(torch.randn(N, H2, W2, 64, generator=g)); dy2n = dy2b.permute(0, 3, 1, 2).double()
full = F.conv_transpose2d(dy2n, W2t.double(), stride=2)
full = F.pad(full, (0, W1 - full.shape[3], 0, H1 - full.shape[2]))
report("conv2 dgrad", _run(5, dy2b, y1b, W2t, None, N, H, W, (N, H1, W1, 32)), nhwc(full * (y1n > 0)))
w = W3t.double().requires_grad_(True); (F.conv2d(y2n, w, stride=1) * dy3n).sum().backward()
got = _run(6, dy3, y2b, None, None, N, H, W, (64, 64, 3, 3)); print("conv3 wgrad rel", rel(got, w.grad))
print("  per-tap rel:", [[round(rel(got[:, :, a, b], w.grad[:, :, a, b]), 4) for b in range(3)] for a in range(3)])
w = W2t.double().requires_grad_(True); (F.conv2d(y1n, w, stride=2) * dy2n).sum().backward()
got = _run(7, dy2b, y1b, None, None, N, H, W, (64, 32, 4, 4)); print("conv2 wgrad rel", rel(got, w.grad))
print("  per-tap rel:", [[round(rel(got[:, :, a, b], w.grad[:, :, a, b]), 4) for b in range(4)] for a in range(4)])
dy1b = _bf(torch.randn(N, H1, W1, 32, generator=g))
w = W1t.double().requires_grad_(True); (F.conv2d(xd, w, stride=4) * dy1b.permute(0, 3, 1, 2).double()).sum().backward()
got = _run(8, dy1b, x, None, None, N, H, W, (32, 3, 8, 8)); print("conv1 wgrad rel", rel(got, w.grad))
print("  per-tap rel:", [[round(rel(got[:, :, a, b], w.grad[:, :, a, b]), 3) for b in range(8)] for a in range(8)])
