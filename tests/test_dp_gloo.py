"""Data-parallel host logic on CPU with gloo, world_size 2.

The CUDA kernels cannot run here, so the invariant is checked on the exchange protocol itself with the
CPU oracle standing in for the per-rank arithmetic: merging per-rank (sum, sumsq, N) reproduces the
whole-batch mean / unbiased std, the two backward scalars merge linearly, the penalty's global means
merge from per-rank sums, and bucketed gradient all-reduce reproduces the whole-batch gradient."""
import os

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from locate_b200 import dist
    dist.enable()
    assert dist.active() and dist.world_size() == world and dist.rank() == rank
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((6, 5, 4, 4), generator=gen, dtype=torch.float64) * 2 + 3
    g = torch.randn((6, 5, 4, 4), generator=gen, dtype=torch.float64)
    gain = torch.randn((1, 5, 1, 1), generator=gen, dtype=torch.float64)
    shard = slice(rank * 3, rank * 3 + 3)
    xs, gs_ = x[shard], g[shard]
    # forward statistics
    sums = torch.stack([xs.sum(), (xs * xs).sum()])
    n = xs.numel() * dist.all_reduce_sum_(sums)
    mean = sums[0] / n
    std = ((sums[1] - sums[0] * mean) / (n - 1)).sqrt()
    ok = torch.allclose(mean, x.mean()) and torch.allclose(std, x.std())
    # backward scalars
    sc = torch.stack([(gain * gs_).sum(), (gain * gs_ * (xs - mean)).sum()])
    dist.all_reduce_sum_(sc)
    dx = gain * gs_ / std - sc[0] / (std * n) - sc[1] * (xs - mean) / ((n - 1) * std ** 3)
    xr = x.clone().requires_grad_(True)
    ((xr - xr.mean()) * gain / xr.std()).backward(g)
    ok = ok and torch.allclose(dx, xr.grad[shard], atol=1e-10)
    # penalty means
    d_true = torch.arange(6, dtype=torch.float64) * 0.3
    d_aug = torch.arange(6, dtype=torch.float64) * 0.1 + 1
    ps = torch.stack([d_true[shard].sum(), d_aug[shard].sum()])
    ng = 3 * dist.all_reduce_sum_(ps, norm_stat=False)
    ok = ok and torch.allclose(100 * ((ps[0] - ps[1]) / ng) ** 2, 100 * (d_true.mean() - d_aug.mean()) ** 2)
    # bucketed gradient all-reduce over a flat arena
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    for h in dist.all_reduce_grads_(flat, bucket_elems=4):
        h.wait()
    ok = ok and torch.equal(flat, torch.arange(10, dtype=torch.float32) * 3)
    # overlapped bucketing: buckets fire as soon as every parameter in them has its learned number of contributions
    params = [torch.zeros(n) for n in (3, 5, 2, 6)]
    offsets = [0, 4, 12, 16]                                     # arena slots aligned to 4 elements
    arena = torch.zeros(24)
    order = [3, 2, 3, 1, 0, 2, 1, 0]                             # "backward": two contributions per parameter, last layer first
    tr = dist.GradOverlap(arena, params, offsets, bucket_elems=8)

    def backward_pass(scale):
        arena.zero_()
        tr.begin("d_step")
        fired_at = []
        for i in order:
            arena[offsets[i]:offsets[i] + params[i].numel()] += scale * (i + 1) * (rank + 1)
            tr.written(params[i])
            fired_at.append(tr.fired_early)
        early = tr.fired_early
        tr.finish()
        return early, fired_at
    e0, _ = backward_pass(1.0)                                   # learning pass: nothing fires early
    want = torch.zeros(24)
    for i in range(4):
        want[offsets[i]:offsets[i] + params[i].numel()] = 2 * (i + 1) * 3
    ok = ok and e0 == 0 and torch.equal(arena, want)
    e1, fired_at = backward_pass(2.0)                            # learned: bucket 2 (param 3) fires after its 2nd contribution
    ok = ok and e1 == 3 and fired_at[2] == 1 and torch.equal(arena, 2 * want)
    tr.begin("d_step")                                           # a pass that changed shape must not produce a silent wrong sum
    for i in (3, 3):
        tr.written(params[i])
    try:
        tr.written(params[3])
        ok = False
    except RuntimeError:
        pass
    tr.tag = None
    td.barrier()
    # non-parity mode: per-replica norm statistics
    dist.enable(sync_norm=False)
    t = torch.ones(2, dtype=torch.float64)
    ok = ok and dist.all_reduce_sum_(t) == 1 and float(t[0]) == 1.0
    ret[rank] = bool(ok)
    td.destroy_process_group()


@pytest.mark.timeout(120)
def test_dp_exchanges_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert ret.get(0) and ret.get(1)
