import sys
sys.path.insert(0, '.')
import torch
import locate_b200 as L
from locate_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = 'cuda:0'
L.configure(IMAGE_SIZE=128)
torch.manual_seed(999)
gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, dev)
dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, dev)
tr = L.GanTrainer(gen, dis, g_opt, d_opt)
g = torch.Generator().manual_seed(0)
real = ops._as_act(torch.randn((B,3,128,128), generator=g).clamp_(-1,1).to(dev)); aug = ops._as_act((real.cpu()+0.05*torch.randn((B,3,128,128), generator=g)).clamp_(-1,1).to(dev)); z = torch.randn((B,128), generator=g).to(dev)
for _ in range(3): tr.step(real, aug, z)
torch.cuda.synchronize()
with ops.KernelTimer() as t:
    tr.step(real, aug, z)
torch.cuda.synchronize()
rows = sorted(t.summary(by_label=True).items(), key=lambda kv: -kv[1]['ms'])
tot = sum(v['ms'] for _, v in rows)
print(f"B={B} GEMM-class total {tot:.2f} ms")
for (fam, label), v in rows[:45]:
    print(f"{v['ms']:8.3f} ms n={v['launches']:3d} {v['flops']/(v['ms']*1e-3)/1e12:8.1f} TF/s  {fam:10s} {label}")
