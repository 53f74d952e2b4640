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
GRAD_BUCKET_ELEMS = 8 * 1024 * 1024      # fp32 elements per gradient all-reduce bucket (32 MB)


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


class GradOverlap:
    """Launches each gradient bucket's all-reduce as soon as its last contributor has run, so the exchange overlaps the
    rest of the backward pass (torch's NCCL process group runs it on its own stream behind an
    event of the compute stream, a fork that CUDA-graph capture records as such).

    A bucket is a slice of the flat gradient arena; a parameter contributes to every bucket its slice touches.  How
    many gradient contributions a parameter receives per backward pass is a property of the step (the discriminator is
    applied to three batches per update: three contributions per weight), so it is LEARNED: the first backward pass
    with a given tag only counts and reduces at the end; later passes fire a bucket when all its parameters have
    received their learned number of contributions.  A contribution that arrives after its bucket was reduced raises --
    a silent wrong sum is never produced.  `finish()` reduces whatever did not fire and waits for everything."""

    def __init__(self, flat, params, offsets, bucket_elems=None, reduce_fn=None):
        bucket_elems = bucket_elems or GRAD_BUCKET_ELEMS
        self.flat = flat
        self.bounds = bucket_bounds(flat.numel(), bucket_elems)
        self.buckets_of = {}
        for p, off in zip(params, offsets):
            first, last = off // bucket_elems, max(off, off + p.numel() - 1) // bucket_elems
            self.buckets_of[id(p)] = list(range(first, min(last, len(self.bounds) - 1) + 1))
        self.expected = {}            # tag -> {id(param): contributions per backward}
        self._reduce = reduce_fn or (lambda chunk: td.all_reduce(chunk, op=td.ReduceOp.SUM, group=_STATE["group"], async_op=True))
        self.tag = None
        self.fired_early = 0          # buckets launched before finish() in the last pass (observability / tests)

    def begin(self, tag):
        self.tag = tag
        self.counts = {}
        self.handles = []
        self.fired = set()
        self.fired_early = 0
        exp = self.expected.get(tag)
        self.pending = None
        if exp is not None:
            self.pending = [0] * len(self.bounds)
            for pid, n in exp.items():
                for b in self.buckets_of[pid]:
                    self.pending[b] += 1

    def written(self, param):
        """Called after the kernels that wrote this parameter's gradient have been enqueued."""
        pid = id(param)
        if self.tag is None or pid not in self.buckets_of:
            return
        c = self.counts.get(pid, 0) + 1
        self.counts[pid] = c
        if self.pending is None:
            return
        want = self.expected[self.tag].get(pid)
        if want is None or c > want:
            if any(b in self.fired for b in self.buckets_of[pid]):
                raise RuntimeError("a gradient was produced after its bucket had been all-reduced: the backward pass changed "
                                   f"shape since it was learned (tag {self.tag!r}); call GradOverlap.expected.clear()")
            return
        if c == want:
            for b in self.buckets_of[pid]:
                self.pending[b] -= 1
                if self.pending[b] == 0:
                    self._fire(b)
                    self.fired_early += 1

    def _fire(self, b):
        s, e = self.bounds[b]
        self.fired.add(b)
        self.handles.append(self._reduce(self.flat[s:e]))

    def finish(self):
        if self.tag is None:
            return
        if self.pending is None or self.expected[self.tag] != self.counts:
            if self.pending is not None and self.fired and self.expected[self.tag] != self.counts:
                # fewer contributions than learned for some parameter: its buckets never fired (pending > 0), nothing
                # was reduced early that is now stale; relearn for the next pass
                pass
            self.expected[self.tag] = dict(self.counts)
        for b in range(len(self.bounds)):
            if b not in self.fired:
                self._fire(b)
        for h in self.handles:
            h.wait()
        self.tag = None


_TRACKER = [None]


def grad_written(param):
    """ops / conv_fn call this after enqueueing the kernels that accumulate into `param`'s gradient."""
    t = _TRACKER[0]
    if t is not None:
        t.written(param)


def bucket_bounds(n_elems, bucket_elems):
    """[(start, stop)] covering n_elems (host logic, unit-tested on CPU)."""
    return [(s, min(n_elems, s + bucket_elems)) for s in range(0, n_elems, bucket_elems)]
