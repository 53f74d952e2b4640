"""The G+D training step (main.py:142-172 with miniter = MINIBATCHES = DITERS = 1)
and the reference's loss / model helpers (utils.py:116-150, grad_penalty.py:1-2)."""
import torch
from torch import nn

from . import dist, ops
from ._lib import call, ptr
from .config import CFG
from .optim import Nadam


def hinge(output_tensor):
    """(1 - t).clamp(min=0)  (utils.py:133-134)."""
    return ops.HingeFn.apply(output_tensor)


def penalty(d_true, aug_data, dis, device, gamma=100):
    """gamma * (mean D(real) - mean D(aug))^2 (grad_penalty.py:1-2).  Composition of [B]-sized torch
    scalars for API compatibility; train_step uses the fused lb_d_loss kernel instead."""
    return gamma * (d_true.mean() - dis(aug_data.to(device)).view(-1).mean()) ** 2


def init(module: nn.Module):
    """utils.py:116-130: only InPlaceNorm.weight ~ U(0.998, 1.002) and bias = 0 take effect, because
    every conv/linear `weight` was re-registered as weight_bar by SpectralNorm."""
    if "norm" not in module.__class__.__name__.lower():
        w = getattr(module, "weight", None)
        if isinstance(w, torch.Tensor):
            nn.init.orthogonal_(w.data)
    else:
        w = getattr(module, "weight", None)
        if isinstance(w, torch.Tensor):
            nn.init.uniform_(w.data, 0.998, 1.002)
    b = getattr(module, "bias", None)
    if isinstance(b, torch.Tensor):
        nn.init.constant_(b.data, 0)


def parameter_count(net):
    return sum(p.numel() for p in net.parameters() if p.requires_grad)


def get_model(model, learning_rate, device):
    """(model on device with reference init, Nadam)  (utils.py:146-150).  Init draws happen on the CPU
    copy so the RNG stream equals the reference's CPU stream, then the model moves to `device`."""
    model.apply(init)
    model = model.to(device)
    opt = Nadam(model.parameters(), lr=learning_rate, betas=(CFG.BETA_1, CFG.BETA_2))
    if hasattr(model, "zero_grad") and hasattr(type(model), "_lb_optimizer"):
        model.__dict__["_lb_optimizer"] = opt
        opt._lb_model = model
    return model, opt


class GanTrainer:
    """One object per process (per GPU).  `step(real, aug, z)` = one discriminator update followed by
    one generator update; in data parallel the three exchanges of locate_b200.dist happen inside."""

    def __init__(self, gen, dis, g_opt, d_opt, penalty_gamma=100.0):
        self.gen, self.dis, self.g_opt, self.d_opt = gen, dis, g_opt, d_opt
        self.penalty_gamma = float(penalty_gamma)
        self._graph = self._static_in = self._static_out = None
        self._copy_stream = self._stage_bufs = self._staged = self._consumed = None

    def _overlap_begin(self, opt, tag):
        """Data parallel: from here on every gradient bucket of `opt`'s main arena is all-reduced as soon as its last
        contributor of this backward pass has run (dist.GradOverlap), overlapping the exchange with the backward pass."""
        if not dist.active():
            return
        trackers = self.__dict__.setdefault("_overlap", {})
        tr = trackers.get(id(opt))
        if tr is None:
            main = [a for a in opt.live_arenas() if not a["late"]]
            if len(main) != 1:
                return
            a = main[0]
            tr = trackers[id(opt)] = dist.GradOverlap(a["grad"], a["params"], a["offsets"])
        tr.begin(tag)
        dist._TRACKER[0] = tr

    def _reduce_and_step(self, opt):
        model = getattr(opt, "_lb_model", None)
        if model is not None:
            model._finish_uv_grads()       # complete the gradients of trainable spectral-norm v's BEFORE they are reduced
        tr, dist._TRACKER[0] = dist._TRACKER[0], None
        done = None
        if tr is not None and tr.flat.data_ptr() in [f.data_ptr() for f in opt.flat_grads]:
            tr.finish()                    # buckets that fired during the backward pass are (being) reduced already
            done = tr.flat.data_ptr()
        for h in [h for flat in opt.flat_grads if flat.data_ptr() != done for h in dist.all_reduce_grads_(flat)]:
            h.wait()
        opt.step()

    def d_step(self, real, aug, z, update=True):
        """main.py:142-156 (+ :158-159 when `update`): D forward on real / generated / augmented, loss, backward."""
        dev = real.device
        with torch.no_grad():
            fake = self.gen(z)
        self.dis.zero_grad()
        if update:
            self._overlap_begin(self.d_opt, "d_step")
        d_true = self.dis(real).view(-1)
        d_fake = self.dis(fake).view(-1)
        d_aug = self.dis(aug).view(-1)
        n = d_true.numel()
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        call("lb_loss_sums", ptr(d_true), ptr(d_aug), n, ptr(sums))
        n_global = float(n * dist.all_reduce_sum_(sums, norm_stat=False))
        out = torch.empty(3, dtype=torch.float32, device=dev)
        grads = torch.empty((3, n), dtype=torch.float32, device=dev)
        call("lb_d_loss", ptr(d_true), ptr(d_fake), ptr(d_aug), ptr(sums), n, n_global, self.penalty_gamma, ptr(out),
             ptr(grads[0]), ptr(grads[1]), ptr(grads[2]))
        torch.autograd.backward([d_true, d_fake, d_aug], [grads[0], grads[1], grads[2]])
        if update:
            self._reduce_and_step(self.d_opt)
        return out            # [hinge part (local share of the global mean), penalty, 0]

    def g_step(self, z, update=True):
        """main.py:160-171: generator forward through the frozen discriminator, loss, backward (+ Nadam when `update`)."""
        dev = z.device
        self.dis.requires_grad_(False)
        self.gen.zero_grad()
        if update:
            self._overlap_begin(self.g_opt, "g_step")
        d_fake = self.dis(self.gen(z)).view(-1)
        n = d_fake.numel()
        out = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty(n, dtype=torch.float32, device=dev)
        call("lb_g_loss", ptr(d_fake), n, float(n * dist.world_size()), ptr(out), ptr(grad))
        torch.autograd.backward([d_fake], [grad])
        if update:
            self._reduce_and_step(self.g_opt)
        self.dis.requires_grad_(True)
        return out

    def run_reference_schedule(self, batches, miniter=1, minibatches=1, diters=1, noise_fn=None):
        """The reference's loop body (main.py:137-172) over an iterable of (real, aug) batches, with ITS schedule:
        the discriminator gradient is recomputed (zero_grad + backward) on every batch but Nadam steps only when
        i % miniter == 0 (i counts from 1), and on every `diters`-th such step the generator runs `minibatches`
        zero_grad + backward passes on the SAME noise before one Nadam step -- so, exactly as in the reference, only
        the last of those gradients is applied (each pass still advances every spectral-norm u/v by one iteration).  Yields (i, d_out, g_out or None) per batch.
        `bench.py` / `step()` measure the miniter = minibatches = diters = 1 case."""
        g_out = None
        for i, (real, aug) in enumerate(batches, 1):
            z = noise_fn(real.shape[0]) if noise_fn is not None else torch.randn(
                (real.shape[0], CFG.INPUT_VECTOR_Z), device=real.device)
            d_out = self.d_step(real, aug, z, update=False)
            stepped = None
            if i % miniter == 0:
                self._reduce_and_step(self.d_opt)
                if (i // miniter) % diters == 0:
                    for _ in range(minibatches):
                        g_out = self.g_step(z, update=False)
                    self._reduce_and_step(self.g_opt)
                    stepped = g_out
            yield i, d_out, stepped

    def _eager_step(self, real, aug, z):
        d_out = self.d_step(real, aug, z)
        g_out = self.g_step(z)
        return d_out, g_out

    def step(self, real, aug, z):
        """One D update + one G update.  After `capture()` the whole step (about 3.7 k kernel launches, both
        backward passes, the collectives and both Nadam updates) is ONE cudaGraphLaunch: inputs are copied into
        the captured static buffers and the graph is replayed."""
        if self._graph is None:
            return self._eager_step(real, aug, z)
        if tuple(tuple(t.shape) for t in (real, aug, z)) != self._graph_shapes:
            raise ValueError(f"the captured graph is frozen to input shapes {self._graph_shapes}; "
                             "call release_graph() / capture() again for another batch shape")
        for dst, src in zip(self._static_in, (real, aug, z)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self._copy_stream is not None:                 # prefetch() may refill its staging buffers from here on
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        self._graph.replay()
        for opt in (self.d_opt, self.g_opt):          # host mirrors of the device-side step counters
            for a in opt.live_arenas():
                a["step"] += 1
                a["epoch"][0] += 1
        return self._static_out

    def prefetch(self, real, aug, z):
        """Start the host -> device copy of the NEXT batch on a side stream (pinned host tensors), so it overlaps the
        step that is running; `step()` with no arguments then trains on it.  This is the input side of the reference's
        loader loop (main.py:122-131) without its per-step synchronous `.to(device)`."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staged = None
        cs = self._copy_stream
        if self._stage_bufs is None or any(b.shape != t.shape for b, t in zip(self._stage_bufs, (real, aug, z))):
            dev = next(self.gen.parameters()).device
            self._stage_bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in (real, aug, z))
            # the caching allocator may hand back blocks that kernels still queued on the compute stream read:
            # order the first copy after everything enqueued so far
            cs.wait_stream(torch.cuda.current_stream())
            self._consumed = None
        if self._consumed is not None:                    # the previous batch has left the staging buffers
            cs.wait_event(self._consumed)
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage_bufs, (real, aug, z)):
                dst.copy_(src, non_blocking=True)
        self._staged = torch.cuda.Event()
        self._staged.record(cs)

    def step_prefetched(self):
        """Train on the batch handed to `prefetch()`."""
        if self._staged is None:
            raise RuntimeError("step_prefetched() without prefetch()")
        torch.cuda.current_stream().wait_event(self._staged)
        self._staged = None
        if self._graph is None:                           # eager: the step reads the staging buffers until it ends
            out = self._eager_step(*self._stage_bufs)
            self._consumed = torch.cuda.Event()
            self._consumed.record()
            return out
        return self.step(*self._stage_bufs)

    def capture(self, real, aug, z, warmup=3):
        """Capture the step into a CUDA graph (CUDA streams and graphs instead of a tracing compiler).  Everything
        step-dependent lives on the device (Nadam schedule, spectral-norm u/v, statistics), tensor maps and launch
        geometry are baked into the graph, so shapes are frozen to those of (real, aug, z).  Runs `warmup` eager
        steps first (lazy initialisation, allocator warm-up), which train the model like any other step."""
        from .conv_fn import invalidate_packs
        shapes = tuple(tuple(t.shape) for t in (real, aug, z))
        dev = next(self.gen.parameters()).device          # static inputs live on the model's device even for (pinned) host batches
        self._static_in = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev).copy_(t) for t in (real, aug, z))
        warmup = max(int(warmup), 1)                      # the first step switches D's u / v on (main.py:172): capture the steady state
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        invalidate_packs()                            # every replay starts with stale weight packs; capture the same
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._eager_step(*self._static_in)
        self._graph, self._static_out, self._graph_shapes = graph, out, shapes
        return self

    def release_graph(self):
        self._graph = self._static_in = self._static_out = None
