"""First-contact check of the persistent recurrence kernel under a hard timeout (a hang must never reach gpurun's own
limit): python scripts/rnn_seq_check.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacorl_b200 import _lib, ops
ops.set_precision("bf16")
dev = "cuda"
torch.manual_seed(0)
T, B, I, H = 16, 64, 48, 2048
ws = [torch.empty(H, I).uniform_(-.02, .02), torch.empty(H, H).uniform_(-.02, .02), torch.zeros(H), torch.zeros(H)]
x = torch.randn(T, B, I)
outs = {}
for mode in (0, 1):
    _lib.lib().tacorl_rnn_seq_enable(mode)
    wd = [w.to(dev).requires_grad_(True) for w in ws]
    xd = x.to(dev).requires_grad_(True)
    out, _ = ops.relu_rnn(xd, wd, 1, False, False, None)
    out.sum().backward()
    torch.cuda.synchronize()
    outs[mode] = (out.detach().clone(), wd[1].grad.clone(), xd.grad.clone())
    print("mode", mode, "ok", float(out.abs().mean()), flush=True)
print("timeouts", _lib.lib().tacorl_rnn_seq_timeouts())
for a, b in zip(outs[0], outs[1]):
    print("max abs diff", float((a - b).abs().max()), "equal", torch.equal(a, b))
