"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The hot path shards over the batch.  Three exchanges exist:
  1. gradient all-reduce over the flat Nadam arena, in buckets (optim.Nadam / GradBucketer);
  2. the whole-tensor norm's global statistics: (sum, sumsq) forward, two scalars backward;
  3. the penalty's two global means.
Spectral-norm u/v need no exchange (deterministic in W, replicas identical).
"""
import os

import torch
import torch.distributed as td

_STATE = {"group": None, "world": 1, "rank": 0, "sync_norm": True}


def init_from_env(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    if not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group(backend=backend)
    enable(td.group.WORLD)
    return _STATE["rank"], _STATE["world"]


def enable(group=None, sync_norm=True):
    _STATE["group"] = group if group is not None else td.group.WORLD
    _STATE["world"] = td.get_world_size(_STATE["group"])
    _STATE["rank"] = td.get_rank(_STATE["group"])
    _STATE["sync_norm"] = sync_norm


def disable():
    _STATE.update(group=None, world=1, rank=0)


def world_size():
    return _STATE["world"]


def rank():
    return _STATE["rank"]


def active():
    return _STATE["group"] is not None and _STATE["world"] > 1


def all_reduce_sum_(t, norm_stat=True):
    """In-place sum over ranks of a small statistics tensor; returns the number of ranks summed.
    With sync_norm off (documented non-parity mode) norm statistics stay per replica."""
    if not active() or (norm_stat and not _STATE["sync_norm"]):
        return 1
    td.all_reduce(t, op=td.ReduceOp.SUM, group=_STATE["group"])
    return _STATE["world"]


def all_reduce_grads_(flat, bucket_elems=16 * 1024 * 1024):
    """Sum the flat gradient arena over ranks in buckets (async handles returned to the caller)."""
    if not active():
        return []
    handles = []
    for start in range(0, flat.numel(), bucket_elems):
        chunk = flat[start:start + bucket_elems]
        handles.append(td.all_reduce(chunk, op=td.ReduceOp.SUM, group=_STATE["group"], async_op=True))
    return handles


def bucket_bounds(n_elems, bucket_elems):
    """[(start, stop)] covering n_elems (host logic, unit-tested on CPU)."""
    return [(s, min(n_elems, s + bucket_elems)) for s in range(0, n_elems, bucket_elems)]
