"""Yard-stick for the bf16 tolerance: the UNMODIFIED reference under torch.autocast(cpu, bfloat16) against
itself in fp64, one discriminator loss mean(hinge(D(x))) + mean(hinge(-D(G(z)))) and one generator loss, relative
gradient-norm error of all D / G gradients.  Runs in the build container only (needs /root/reference)."""
import copy, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ref_loader
size, batch = int(sys.argv[1]), int(sys.argv[2])
libs = ref_loader.load_ref(IMAGE_SIZE=size)
torch.manual_seed(999)
with ref_loader.quiet():
    G = libs.Generator(); D = libs.Discriminator()
    G, _ = libs.get_model(G, 5e-4, 'cpu'); D, _ = libs.get_model(D, 2e-3, 'cpu')
g = torch.Generator().manual_seed(0)
real = torch.randn((batch, 3, size, size), generator=g).clamp_(-1, 1); z = torch.randn((batch, size), generator=g)
def run(G, D, dtype, autocast):
    G = copy.deepcopy(G).to(dtype); D = copy.deepcopy(D).to(dtype); G.noise = G.noise.to(dtype)
    ctx = torch.autocast('cpu', torch.bfloat16) if autocast else torch.autocast('cpu', enabled=False)
    with ctx:
        fake = G(z.to(dtype))
        d_loss = (libs.hinge(D(real.to(dtype)).view(-1)) + libs.hinge(-D(fake.detach()).view(-1))).mean()
    d_loss.backward()
    gd = torch.cat([p.grad.double().reshape(-1) for p in D.parameters() if p.grad is not None])
    D.zero_grad()
    with ctx:
        g_loss = libs.hinge(D(fake).view(-1)).mean()
    g_loss.backward()
    gg = torch.cat([p.grad.double().reshape(-1) for p in G.parameters() if p.grad is not None])
    return gd, gg
d64, g64 = run(G, D, torch.float64, False)
d16, g16 = run(G, D, torch.float32, True)
print(f"size {size} batch {batch}: reference bf16-autocast vs fp64: D grad rel err {((d16-d64).norm()/d64.norm()).item():.3e}  G grad rel err {((g16-g64).norm()/g64.norm()).item():.3e}")
