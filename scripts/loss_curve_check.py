"""Diagnostic / evidence script: PlayLMP loss curves of the CUDA path (fp32 and bf16) against the fp32 CPU oracle
on identical batches and noise.  Writes a table to stdout.  Usage: python scripts/loss_curve_check.py [steps] [lr]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth as S  # noqa: E402
from oracle import tacorl_oracle as O  # noqa: E402
from tests.gpu_util import build_play_lmp, play_lmp_tape, to_dev  # noqa: E402
from tacorl_b200 import ops  # noqa: E402
from tacorl_b200.utils.rng import noise_tape  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
B, T, H, W = 4, 8, 84, 84


def make(prec):
    ops.set_precision(prec)
    m = build_play_lmp("tanh_net", ("rgb_static",), 128, 16, T)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    sd = S.synth_state_dict(shapes, 9)
    m.load_state_dict(sd)
    m.lr = lr
    m.to("cuda")
    return m, m.configure_optimizers(), sd


mods = {p: make(p) for p in ("fp32", "bf16")}
P = O.params_from(mods["fp32"][2])
ost = {}
batches = [S.synth_play_batch(B, T, H, W, 100 + i) for i in range(4)]
dev_batches = [to_dev(b) for b in batches]
hist = {"oracle": [], "fp32": [], "bf16": []}
for s in range(steps):
    torch.manual_seed(5000 + s)
    noise = O.draw_play_lmp_noise(B, T)
    for prec, (m, opt, _) in mods.items():
        ops.set_precision(prec)
        opt.zero_grad()
        with noise_tape(play_lmp_tape(noise, B)):
            loss = m.training_step(S.clone_batch(dev_batches[s % 4]), s)
        loss.backward()
        opt.step()
        hist[prec].append(float(loss.detach()))
    out, _ = O.play_lmp_training_step(P, ost, S.clone_batch(batches[s % 4]), noise, lr=lr)
    hist["oracle"].append(float(out["total_loss"].detach()))
ops.set_precision("fp32")
print(f"steps={steps} lr={lr}")
print("step  oracle      fp32(cuda)  bf16(cuda)  rel_fp32   rel_bf16   | window-20 mean rel: fp32, bf16")
w = 20
for s in list(range(0, 10)) + list(range(10, steps, max(1, steps // 30))):
    o, a, b = hist["oracle"][s], hist["fp32"][s], hist["bf16"][s]
    lo = max(0, s - w + 1)
    mo = sum(hist["oracle"][lo:s + 1]) / (s + 1 - lo)
    ma = sum(hist["fp32"][lo:s + 1]) / (s + 1 - lo)
    mb = sum(hist["bf16"][lo:s + 1]) / (s + 1 - lo)
    print(f"{s:4d}  {o:10.5f}  {a:10.5f}  {b:10.5f}  {abs(a-o)/max(1,abs(o)):.2e}  {abs(b-o)/max(1,abs(o)):.2e}   | "
          f"{abs(ma-mo)/max(1,abs(mo)):.2e} {abs(mb-mo)/max(1,abs(mo)):.2e}")
worst = {k: max(abs(x - o) / max(1, abs(o)) for x, o in zip(hist[k], hist["oracle"])) for k in ("fp32", "bf16")}
print("worst per-step rel deviation:", worst)
