import sys, json, collections, re
sys.path.insert(0, '.')
import torch
import locate_b200 as L
from locate_b200 import ops
from torch.profiler import profile, ProfilerActivity
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
import time
t0=time.perf_counter()
for _ in range(3): tr.step(real, aug, z)
torch.cuda.synchronize(); wall=(time.perf_counter()-t0)/3*1e3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(real, aug, z); torch.cuda.synchronize()
tot = collections.defaultdict(lambda:[0,0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r'\(anonymous namespace\)::','',ev.name); name = re.sub(r'\(.*','',name)[:70]
        tot[name][0]+=1; tot[name][1]+=ev.device_time/1e3
s = sum(v[1] for v in tot.values())
print(f"B={B} wall {wall:.2f} ms/step; GPU kernel time {s:.2f} ms; launches {sum(v[0] for v in tot.values())}")
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][1])[:32]:
    print(f"{v[1]:9.3f} ms {100*v[1]/s:5.1f}% n={v[0]:5d} {k}")
