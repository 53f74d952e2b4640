"""One ActivatedBaseConv (conv_0 -> RootTanh -> conv_1) forward / backward on the tensor-core path, timed with CUDA events.
usage: pair_micro.py [batch] [cin] [cout] [hw] [kind: convT|conv3|conv5]   (also the command ncu wraps for a per-kernel capture)"""
import sys
sys.path.insert(0, '.')
import torch
from torch import nn
import locate_b200 as L
from locate_b200 import layers, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 96
cout = int(sys.argv[3]) if len(sys.argv) > 3 else 48
hw = int(sys.argv[4]) if len(sys.argv) > 4 else 64
kind = sys.argv[5] if len(sys.argv) > 5 else "convT"
dev = 'cuda:0'
L.configure(PRECISION="bf16")
torch.manual_seed(0)
if kind == "convT":
    m = layers.ActivatedBaseConv(cin, cout, nn.ConvTranspose2d, kernel=4, stride=2, pad=1)
elif kind == "conv3":
    m = layers.ActivatedBaseConv(cin, cout, nn.Conv2d, kernel=3, stride=1, pad=1)
else:
    m = layers.ActivatedBaseConv(cin, cout, nn.Conv2d, kernel=5, stride=2, pad=2)
m = m.to(dev)
x = torch.randn((B, cin, hw, hw), device=dev).contiguous(memory_format=torch.channels_last).bfloat16().requires_grad_(True)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fwd_nograd():
    with torch.no_grad():
        return m(x)


y = m(x)
g = torch.randn_like(y)


def fwd_bwd():
    x.grad = None
    m(x).backward(g)


with ops.KernelTimer() as t:
    fwd_nograd()
    fwd_bwd()
torch.cuda.synchronize()
for (fam, label), v in sorted(t.summary(by_label=True).items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{v['ms']:8.3f} ms  {fam:10s} {label}  {v['flops'] / v['ms'] / 1e9:8.1f} TF/s")
print(f"forward (no grad) {timed(fwd_nograd):.3f} ms; forward + backward {timed(fwd_bwd):.3f} ms")
