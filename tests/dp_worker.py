"""torchrun worker for tests/test_gpu_dp.py: N ranks on a sliced global batch must reproduce the single-process
run on the whole batch (losses, gradients) -- SURVEY.md section 4 item 4 / section 8e."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                  # noqa: E402
import torch.distributed as td                # noqa: E402

import locate_b200 as L                       # noqa: E402
from locate_b200 import dist                  # noqa: E402


def build(dev):
    torch.manual_seed(999)
    gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, dev)
    dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, dev)
    return L.GanTrainer(gen, dis, g_opt, d_opt), gen, dis, g_opt, d_opt


def run(trainer, d_opt, g_opt, real, aug, z):
    grads = {}
    for tag, opt in (("d", d_opt), ("g", g_opt)):
        def spy(closure=None, tag=tag, opt=opt):
            grads[tag] = opt.flat_grads[0].clone()
        opt.step = spy
    # three steps (the spy keeps the weights where they are; u / v advance identically in both runs): the first backward
    # pass of each kind only LEARNS the per-parameter contribution counts, the later ones fire buckets during the pass
    for _ in range(3):
        d_out, g_out = trainer.step(real, aug, z)
    torch.cuda.synchronize()
    return d_out.clone(), g_out.clone(), grads


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    rank, world = dist.init_from_env()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    L.configure(IMAGE_SIZE=32, BASE_FEATURE_FACTOR=4, PRECISION=precision)
    per = 3
    g = torch.Generator().manual_seed(0)
    real = torch.randn((per * world, 3, 32, 32), generator=g).clamp_(-1, 1).to(dev)
    aug = (real.cpu() + 0.05 * torch.randn((per * world, 3, 32, 32), generator=g)).clamp_(-1, 1).to(dev)
    z = torch.randn((per * world, 32), generator=g).to(dev)

    # whole batch, single process semantics (exchanges switched off)
    dist.disable()
    tr, gen, dis, g_opt, d_opt = build(dev)
    ref_d, ref_g, ref_grads = run(tr, d_opt, g_opt, real, aug, z)

    # data parallel on the slices (small buckets: many of them complete in the middle of the backward pass)
    dist.enable(td.group.WORLD)
    dist.GRAD_BUCKET_ELEMS = 16384
    tr, gen, dis, g_opt, d_opt = build(dev)
    sl = slice(rank * per, (rank + 1) * per)
    d_out, g_out, grads = run(tr, d_opt, g_opt, real[sl].contiguous(), aug[sl].contiguous(), z[sl].contiguous())
    early = sum(t.fired_early for t in getattr(tr, "_overlap", {}).values())
    if early < 4:
        print(f"rank {rank}: only {early} gradient buckets were all-reduced before the end of their backward pass")
        ok_overlap = False
    else:
        ok_overlap = True
    hinge = d_out[0:1].clone()
    td.all_reduce(hinge)                       # local shares of the global mean add up
    gl = g_out.clone()
    td.all_reduce(gl)
    tol = 2e-4 if precision == "fp32" else 2e-2
    ok = ok_overlap
    for name, a, b in (("d hinge", hinge[0], ref_d[0]), ("penalty", d_out[1], ref_d[1]), ("g loss", gl[0], ref_g[0])):
        if abs(a.item() - b.item()) > tol * max(1.0, abs(b.item())):
            print(f"rank {rank}: {name} {a.item()} vs {b.item()}")
            ok = False
    for tag in ("d", "g"):
        err = ((grads[tag] - ref_grads[tag]).norm() / ref_grads[tag].norm()).item()
        if err > (1e-3 if precision == "fp32" else 3e-2):
            print(f"rank {rank}: {tag} gradient relative error {err:.3e}")
            ok = False
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    td.barrier()
    td.destroy_process_group()
    if rank == 0:
        print("DP_INVARIANCE_OK" if flag.item() == 1.0 else "DP_INVARIANCE_FAILED")
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
