"""Runs only the vision encoder forward+backward (1024 frames of 3x200x200, bf16 tensor-core path) — the op the
bench's `roofline` object times — so ncu can list / capture its kernels:  ncu ... python scripts/profile_encoder.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacorl_b200 import configs, ops  # noqa: E402
from tacorl_b200.utils import synthetic  # noqa: E402
from tacorl_b200.utils.config import instantiate  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops.set_precision(prec)
enc = instantiate(configs.lmp_vision_encoder()).cuda()
synthetic.init_like_reference(enc, 0)
x = synthetic.play_batch(64, 16, 200, 200, seed=1)["states"]["rgb_static"].view(1024, 3, 200, 200).cuda()
for _ in range(iters):
    for p in enc.parameters():
        p.grad = None
    e = enc(x)
    e.backward(torch.ones_like(e))
torch.cuda.synchronize()
print("done", float(e.sum()))
