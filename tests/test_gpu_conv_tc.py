"""Tight per-kernel checks of the implicit-GEMM tcgen05 convolutions (conv_tc.cu) against torch's conv in fp64 on
bf16-exact operands: forward (bias+ReLU), data gradients (ReLU-gated, incl. the stride-2 parity classes) and weight
gradients of all three encoder layers, at every image size the encoder must support."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import DEV, assert_close

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def _run(op, in0, in1, Wt, bias, N, H, W, out_shape):
    from tacorl_b200 import _lib as L
    out = torch.full(out_shape, float("nan"), device=DEV)
    ws = L.workspace(1 << 30, torch.device(DEV), tag="convdbg")
    # keep every device tensor referenced until the kernels have run (a freed temporary would be recycled by the
    # caching allocator for the next H2D copy before the launch executes)
    d0 = in0.to(DEV).contiguous()
    d1 = in1.to(DEV).contiguous() if in1 is not None else None
    dw = Wt.to(DEV).contiguous() if Wt is not None else None
    db = bias.to(DEV).contiguous() if bias is not None else None
    L.call("tacorl_conv_tc_debug", op, L.ptr(d0), L.ptr(d1), L.ptr(dw), L.ptr(db), N, H, W, L.ptr(out),
           ctypes.c_void_p(ws.data_ptr()), ws.numel(), L.stream())
    torch.cuda.synchronize()
    del d0, d1, dw, db
    return out.cpu()


def _geom(H, W):
    H1, W1 = (H - 8) // 4 + 1, (W - 8) // 4 + 1
    H2, W2 = (H1 - 4) // 2 + 1, (W1 - 4) // 2 + 1
    return H1, W1, H2, W2, H2 - 2, W2 - 2


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


# (400, 84, 84): more tiles than CTAs in every kernel -- each persistent CTA runs several rounds of its stage / accumulator
# rings (the single-round shapes above cannot catch a ring-protocol error)
@pytest.mark.parametrize("N,H,W", [(3, 84, 84), (2, 200, 200), (2, 150, 200), (5, 128, 128), (70, 84, 84), (400, 84, 84)])
def test_implicit_gemm_convs_match_torch(N, H, W):
    g = torch.Generator().manual_seed(N * 1000 + H + W)
    H1, W1, H2, W2, H3, W3 = _geom(H, W)
    x = _bf(torch.rand(N, 3, H, W, generator=g) * 2 - 1)
    W1t, W2t, W3t = (_bf(torch.randn(s, generator=g) * sc) for s, sc in (((32, 3, 8, 8), 0.1), ((64, 32, 4, 4), 0.06),
                                                                         ((64, 64, 3, 3), 0.06)))
    b1, b2, b3 = (torch.randn(n, generator=g) * 0.1 for n in (32, 64, 64))
    xd = x.double()
    y1 = F.relu(F.conv2d(xd, W1t.double(), b1.double(), stride=4))
    got = _run(1, x, None, W1t, b1, N, H, W, (N, H1, W1, 32))
    assert_close("conv1 fwd", got, nhwc(y1), 6e-3)
    # linear-shift form (one input window per tile, taps = UMMA descriptors with shifted start addresses): what the
    # encoder runs; must agree with the per-tap-box kernel to the last bit (same products, same accumulation order)
    got_lin = _run(11, x, None, W1t, b1, N, H, W, (N, H1, W1, 32))
    assert_close("conv1 fwd (linear-shift)", got_lin, nhwc(y1), 6e-3)
    assert torch.equal(got_lin, got)
    y1b = _bf(nhwc(y1).float())                         # what the next layer actually consumes
    y1n = y1b.permute(0, 3, 1, 2).double()
    y2 = F.relu(F.conv2d(y1n, W2t.double(), b2.double(), stride=2))
    got = _run(2, y1b, None, W2t, b2, N, H, W, (N, H2, W2, 64))
    assert_close("conv2 fwd", got, nhwc(y2), 6e-3)
    y2b = _bf(nhwc(y2).float())
    y2n = y2b.permute(0, 3, 1, 2).double()
    y3 = F.relu(F.conv2d(y2n, W3t.double(), b3.double(), stride=1))
    got = _run(3, y2b, None, W3t, b3, N, H, W, (N, H3, W3, 64))
    assert_close("conv3 fwd", got, nhwc(y3), 1e-4)
    got_lin = _run(13, y2b, None, W3t, b3, N, H, W, (N, H3, W3, 64))
    assert_close("conv3 fwd (linear-shift)", got_lin, nhwc(y3), 1e-4)
    assert torch.equal(got_lin, got)
    # ---- data gradients
    dy3 = _bf(torch.randn(N, H3, W3, 64, generator=g))
    dy3n = dy3.permute(0, 3, 1, 2).double()
    dy2 = F.conv_transpose2d(dy3n, W3t.double(), stride=1) * (y2n > 0)
    got = _run(4, dy3, y2b, W3t, None, N, H, W, (N, H2, W2, 64))
    assert_close("conv3 dgrad", got, nhwc(dy2), 6e-3)
    dy2b = _bf(torch.randn(N, H2, W2, 64, generator=g))
    dy2n = dy2b.permute(0, 3, 1, 2).double()
    full = F.conv_transpose2d(dy2n, W2t.double(), stride=2)
    pad_h, pad_w = H1 - full.shape[2], W1 - full.shape[3]          # rows/cols no output pixel reaches
    full = F.pad(full, (0, pad_w, 0, pad_h))
    dy1 = full * (y1n > 0)
    got = _run(5, dy2b, y1b, W2t, None, N, H, W, (N, H1, W1, 32))
    assert_close("conv2 dgrad", got, nhwc(dy1), 6e-3)
    # the encoder's form: all four stride-parity classes in one kernel (one haloed dy2 patch per tile, N = 128)
    got_fused = _run(15, dy2b, y1b, W2t, None, N, H, W, (N, H1, W1, 32))
    assert_close("conv2 dgrad (fused classes)", got_fused, nhwc(dy1), 6e-3)
    assert_close("conv2 dgrad fused vs per-class", got_fused, got, 1e-6)
    # ---- weight gradients (fp32 accumulation, fp32 output)
    w = W3t.double().requires_grad_(True)
    (F.conv2d(y2n, w, stride=1) * dy3n).sum().backward()
    got = _run(6, dy3, y2b, None, None, N, H, W, (64 * 64 * 9 + 64,))
    assert_close("conv3 wgrad", got[:-64].view(64, 64, 3, 3), w.grad, 2e-4)
    assert_close("conv3 bias grad", got[-64:], dy3n.sum((0, 2, 3)), 2e-4)
    # linear-shift form (what the encoder runs): dy3 stored at y2's pitch with zero margins, taps = shifted descriptors
    dy3p = F.pad(dy3, (0, 0, 0, 2, 0, 2))
    got_lin = _run(16, dy3p, y2b, None, None, N, H, W, (64 * 64 * 9 + 64,))
    assert_close("conv3 wgrad (linear-shift)", got_lin[:-64].view(64, 64, 3, 3), w.grad, 2e-4)
    assert_close("conv3 bias grad (linear-shift)", got_lin[-64:], dy3n.sum((0, 2, 3)), 2e-4)
    w = W2t.double().requires_grad_(True)
    (F.conv2d(y1n, w, stride=2) * dy2n).sum().backward()
    got = _run(7, dy2b, y1b, None, None, N, H, W, (64 * 32 * 16 + 64,))
    assert_close("conv2 wgrad", got[:-64].view(64, 32, 4, 4), w.grad, 2e-4)
    assert_close("conv2 bias grad", got[-64:], dy2n.sum((0, 2, 3)), 2e-4)
    dy1b = _bf(torch.randn(N, H1, W1, 32, generator=g))
    w = W1t.double().requires_grad_(True)
    (F.conv2d(xd, w, stride=4) * dy1b.permute(0, 3, 1, 2).double()).sum().backward()
    got = _run(8, dy1b, x, None, None, N, H, W, (32 * 3 * 64 + 32,))
    assert_close("conv1 wgrad", got[:-32].view(32, 3, 8, 8), w.grad, 2e-4)
    assert_close("conv1 bias grad", got[-32:], dy1b.double().sum((0, 1, 2)), 2e-4)
    # linear-shift form: dy1 at the s2d image's pitch (one zero row / column of margin), 64-byte swizzled N operand
    got_lin = _run(18, F.pad(dy1b, (0, 0, 0, 1, 0, 1)), x, None, None, N, H, W, (32 * 3 * 64 + 32,))
    assert_close("conv1 wgrad (linear-shift)", got_lin[:-32].view(32, 3, 8, 8), w.grad, 2e-4)
    assert_close("conv1 bias grad (linear-shift)", got_lin[-32:], dy1b.double().sum((0, 1, 2)), 2e-4)
