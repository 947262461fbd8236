"""Times the recurrent stacks alone (CUDA events, eager): BiRNN plan recogniser and the 2-layer decoder RNN at the
bench shapes.  python scripts/rnn_time.py [prec]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacorl_b200 import ops
from tacorl_b200.networks.layers import ReluRNN
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16")
dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
bi = ReluRNN(32, 2048, 2, True).to(dev)
x = torch.randn(64, 16, 32, device=dev, requires_grad=True)
def bi_fb():
    for p in bi.parameters(): p.grad = None
    o, _ = bi(x, last_only=True); o.sum().backward()
def bi_f():
    with torch.no_grad(): bi(x, last_only=True)
dec = ReluRNN(48, 2048, 2, False).to(dev)
for B in (64, 128):
    xd = torch.randn(B, 15, 48, device=dev, requires_grad=True)
    def dec_fb():
        for p in dec.parameters(): p.grad = None
        o, _ = dec(xd); o.sum().backward()
    def dec_f():
        with torch.no_grad(): dec(xd)
    print(f"decoder RNN B={B}: fwd {timeit(dec_f):.3f} ms, fwd+bwd {timeit(dec_fb):.3f} ms  (30 fwd + 30 bwd dependent steps)")
print(f"BiRNN: fwd {timeit(bi_f):.3f} ms, fwd+bwd {timeit(bi_fb):.3f} ms  (33 fwd + 32 bwd dependent steps, directions overlapped)")
# graph-captured variant of the BiRNN fwd+bwd to remove launch overhead
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): bi_fb()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(g): bi_fb()
print(f"BiRNN fwd+bwd (CUDA graph replay): {timeit(g.replay):.3f} ms")
