"""Nadam (libs/nadam.py:5-89) as ONE kernel over a flat parameter arena.

The reference loops over ~145 tensors in Python with ~8 ATen launches each.  Here the trainable
parameters of a group are re-homed into one flat fp32 buffer (param / grad / exp_avg / exp_avg_sq);
`step()` is a single lb_nadam_step launch, `zero_grad()` a single fill, and the data-parallel
all-reduce runs over contiguous buckets of the same buffer.  Backward kernels accumulate straight
into the grad arena (`param._lb_grad`, see ops._grad_sink).
"""
import torch
from torch.optim import Optimizer

from ._lib import call, ptr
from .conv_fn import PackEpoch, invalidate_packs  # noqa: F401  (invalidate_packs re-exported for manual weight edits)

_ALIGN = 4   # floats: keeps every view 16-byte aligned for the vectorised kernels


class Nadam(Optimizer):
    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, schedule_decay=4e-3):
        if weight_decay != 0:
            raise NotImplementedError("weight_decay is never used by the reference (utils.py:149)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, schedule_decay=schedule_decay)
        super().__init__(params, defaults)
        self._arenas = None
        if all(p.is_cuda for g in self.param_groups for p in g["params"]):
            self._attach()

    # -- arena ---------------------------------------------------------------------------------
    def _attach(self):
        """Two arenas per parameter group: the parameters trainable at construction, and the frozen ones (the
        spectral-norm weight_u / weight_v, requires_grad=False in the reference, spectral_norm.py:45-46).  The second
        arena only steps while its parameters require grad -- the reference's loop switches them on for the
        discriminator at the end of its first generator step (main.py:172) and its Nadam, whose parameter list holds
        them since construction (utils.py:149), creates their state (step counter, momentum schedule) at that point."""
        arenas = []
        for group in self.param_groups:
            arenas.append(self._make_arena([p for p in group["params"] if p.requires_grad]))
        for group in self.param_groups:
            arenas.append(self._make_arena([p for p in group["params"] if not p.requires_grad], late=True))
        self._arenas = arenas

    def _groups_arenas(self):
        n = len(self.param_groups)
        return list(zip(self.param_groups + self.param_groups, self._arenas, [False] * n + [True] * n))

    @staticmethod
    def _live(a):
        """A late arena takes part only while its parameters require grad (they switch together, Module.requires_grad_)."""
        return a is not None and (not a["late"] or a["params"][0].requires_grad)

    def _make_arena(self, train, late=False):
        if not train:
            return None
        if any(p.dtype != torch.float32 or not p.is_cuda for p in train):
            if late:
                return None
            raise RuntimeError("Nadam arena needs fp32 CUDA parameters (no CPU fallback)")
        offsets, total = [], 0
        for p in train:
            offsets.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        dev = train[0].device
        epoch = PackEpoch([0])         # bumped by step(): invalidates the bf16 weight packs of THIS arena only
        flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(train, offsets):
                n = p.numel()
                flat_p[off:off + n].copy_(p.data.reshape(-1))          # one-time setup copy
                gview = flat_g[off:off + n].view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                p.data = flat_p[off:off + n].view(p.shape)
                p._lb_grad = gview
                if not late:
                    p._lb_epoch = epoch
                    p.grad = gview
        return dict(param=flat_p, grad=flat_g, exp_avg=torch.zeros_like(flat_p), exp_avg_sq=torch.zeros_like(flat_p),
                    step=0, n=total, epoch=epoch, late=late, params=train, offsets=offsets,
                    sched=torch.tensor([0.0, 1.0], dtype=torch.float64, device=dev),   # {t, m_schedule}
                    hyper=torch.zeros(3, dtype=torch.float32, device=dev))

    def live_arenas(self):
        self._ensure()
        return [a for a in self._arenas if self._live(a)]

    def _ensure(self):
        if self._arenas is None:
            self._attach()

    @property
    def flat_grads(self):
        """Gradient buffers the data-parallel all-reduce must cover before step()."""
        self._ensure()
        return [a["grad"] for a in self._arenas if self._live(a)]

    @property
    def flat_params(self):
        self._ensure()
        return [a["param"] for a in self._arenas if a is not None and not a["late"]]

    def zero_grad(self, set_to_none=False):
        self._ensure()
        for a in self._arenas:
            if self._live(a):
                call("lb_fill", ptr(a["grad"]), a["n"], 0.0)
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad and p.grad is None and getattr(p, "_lb_grad", None) is not None:
                    p.grad = p._lb_grad

    # -- checkpointing (the reference saves no optimizer state, main.py:235-236; a resume here keeps the momenta) -----
    def state_dict(self):
        self._ensure()
        arenas = []
        for a in self._arenas:
            arenas.append(None if a is None else dict(exp_avg=a["exp_avg"].clone(), exp_avg_sq=a["exp_avg_sq"].clone(),
                                                      sched=a["sched"].clone(), step=a["step"], n=a["n"]))
        return dict(param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups], arenas=arenas)

    def load_state_dict(self, state):
        self._ensure()
        if len(state["arenas"]) != len(self._arenas):
            raise ValueError("optimizer state does not fit this parameter set")
        for a, sa in zip(self._arenas, state["arenas"]):
            if (a is None) != (sa is None) or (a is not None and a["n"] != sa["n"]):
                raise ValueError("optimizer state does not fit this parameter set")
            if a is None:
                continue
            with torch.no_grad():
                a["exp_avg"].copy_(sa["exp_avg"])
                a["exp_avg_sq"].copy_(sa["exp_avg_sq"])
                a["sched"].copy_(sa["sched"])
            a["step"] = sa["step"]
            a["epoch"][0] += 1
        for g, sg in zip(self.param_groups, state["param_groups"]):
            g.update(sg)

    # -- step (nadam.py:56-87) ------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._ensure()
        model = getattr(self, "_lb_model", None)
        if model is not None:
            model._finish_uv_grads()       # gradients of trainable spectral-norm v's (no-op unless a backward pass left some)
        for group, a, _late in self._groups_arenas():
            if not self._live(a):
                continue
            a["epoch"][0] += 1             # the raw-pointer update below does not bump torch's version counters
            beta1, beta2 = group["betas"]
            a["step"] += 1                 # host mirror only; the authoritative counter is a["sched"][0] on the device
            call("lb_nadam_schedule", ptr(a["sched"]), ptr(a["hyper"]), group["lr"], beta1, beta2, group["schedule_decay"])
            call("lb_nadam_step", ptr(a["param"]), ptr(a["grad"]), ptr(a["exp_avg"]), ptr(a["exp_avg_sq"]), a["n"],
                 beta1, beta2, group["eps"], ptr(a["hyper"]))
        return loss
