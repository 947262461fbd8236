"""Times the encoder fwd+bwd (1024 frames 200x200) with CUDA events: python scripts/enc_time.py [prec]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tacorl_b200 import configs, ops
from tacorl_b200.utils import synthetic
from tacorl_b200.utils.config import instantiate
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16")
enc = instantiate(configs.lmp_vision_encoder()).cuda()
synthetic.init_like_reference(enc, 0)
x = synthetic.play_batch(64, 16, 200, 200, seed=1)["states"]["rgb_static"].view(1024, 3, 200, 200).cuda()
def fb():
    for p in enc.parameters(): p.grad = None
    e = enc(x); e.backward(torch.ones_like(e))
def fwd():
    with torch.no_grad(): enc(x)
for fn, name in ((fb, "fwd+bwd"), (fwd, "fwd(no_grad)")):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"chunk={os.environ.get('TACORL_ENC_CHUNK','64')} {name}: {e0.elapsed_time(e1)/10:.3f} ms")
