"""One eager G+D step after two warm-up steps (for ncu launch lists / per-kernel captures). usage: one_step.py [batch]"""
import sys
sys.path.insert(0, '.')
import torch
import locate_b200 as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 384
dev = 'cuda:0'
L.configure(IMAGE_SIZE=128)
torch.manual_seed(999)
gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, dev)
dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, dev)
tr = L.GanTrainer(gen, dis, g_opt, d_opt)
g = torch.Generator().manual_seed(0)
real = torch.randn((B, 3, 128, 128), generator=g).clamp_(-1, 1).to(dev)
aug = (real.cpu() + 0.05 * torch.randn((B, 3, 128, 128), generator=g)).clamp_(-1, 1).to(dev)
z = torch.randn((B, 128), generator=g).to(dev)
for _ in range(3):
    out = tr.step(real, aug, z)
torch.cuda.synchronize()
print("ok", [float(v) for v in out[0]], float(out[1][0]))
