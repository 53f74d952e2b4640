"""torch.autograd.Function wrappers around the C-ABI kernels (include/locate_b200.h).

Conventions
 * activations are fp32 CUDA tensors, logical [B,C,H,W] with channels-last strides (physically
   [B][H][W][C]); vectors are [B,C].  Anything else is converted once at the boundary.
 * parameter gradients: every kernel accumulates (+=).  If the parameter carries an arena view
   (`param._lb_grad`, attached by locate_b200.optim.Nadam) the kernel adds straight into it and
   autograd receives None for that input; otherwise a fresh zero tensor is returned to autograd.
 * data-parallel hooks (global norm statistics) come from locate_b200.dist.

Each Function names the reference lines whose arithmetic the kernels replace.
"""
import math

import torch

from . import _lib, dist
from ._lib import call, ptr

_CL = torch.channels_last

# ---- optional per-kernel timing (bench.py roofline): CUDA events on the launching stream -----------
_TIMER = None


class KernelTimer:
    """`with KernelTimer() as t:` records (kernel family, algorithmic flops, bytes) + an event pair per
    launch of the GEMM-class kernels; `t.summary()` (after a synchronize) gives totals per family."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _TIMER
        _TIMER = self
        return self

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = None

    def summary(self, by_label=False):
        out = {}
        for name, flops, nbytes, e0, e1, label in self.records:
            d = out.setdefault((name, label) if by_label else name, dict(launches=0, flops=0.0, bytes=0.0, ms=0.0))
            d["launches"] += 1
            d["flops"] += flops
            d["bytes"] += nbytes
            d["ms"] += e0.elapsed_time(e1)
        return out


def _timed_call(family, flops, nbytes, name, *args):
    if _TIMER is None:
        return call(name, *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    g = next((a for a in args if hasattr(a, "in_c")), None)
    label = (f"{g.kh}x{g.kw}s{g.stride}m{g.mode} {g.in_c}->{g.out_c} in{g.in_h}x{g.in_w} out{g.out_h}x{g.out_w} b{g.batch}"
             if g is not None else "")
    _TIMER.records.append((family, flops, nbytes, e0, e1, label))


def _conv_work(spec, b, h, w_, oh, ow):
    """Algorithmic (flops, activation+weight bytes at fp32) of one conv GEMM (SURVEY.md section 8d:
    MAC x 2, padding taps counted; the transposed conv does kh*kw MACs per INPUT pixel)."""
    pixels = b * (h * w_ if spec.kind == "convT" else oh * ow)
    flops = 2.0 * pixels * spec.kh * spec.kw * spec.cin * spec.cout
    nbytes = 4.0 * (b * h * w_ * spec.cin + b * oh * ow * spec.cout + spec.kh * spec.kw * spec.cin * spec.cout)
    return flops, nbytes


F32, BF16 = 0, 1          # include/locate_b200.h: LB_F32 / LB_BF16


def _dt(t):
    """C-ABI storage code of an activation tensor."""
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype != torch.float32:
        raise TypeError(f"activations are fp32 or bf16, got {t.dtype}")
    return F32


def store_dtype(shape):
    """Storage type of an activation of this logical shape: in the tensor-core configuration every [B,C,H,W] map whose
    channel count TMA can address (C % 8 == 0) lives in HBM as bf16 -- half the bytes of every elementwise pass and
    directly the GEMM operand; narrow maps (RGB images, logits) and [B,C] vectors (style chain) stay fp32."""
    from .config import CFG
    if CFG.PRECISION == "bf16" and len(shape) == 4 and shape[1] % 8 == 0:
        return torch.bfloat16
    return torch.float32


def _cl_strides(shape):
    b, c, h, w = shape
    return (h * w * c, 1, w * c, c)


def _cast(t, dtype):
    """Same layout, other storage type (fp32 -> bf16 through lb_cast_bf16; the reverse only happens at API edges)."""
    if t.dtype == dtype:
        return t
    if dtype == torch.bfloat16 and t.dtype == torch.float32 and t.is_cuda and t.numel() > 0 and (
            t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=_CL))) and t.data_ptr() % 16 == 0:
        out = torch.empty_strided(t.shape, t.stride(), dtype=torch.bfloat16, device=t.device)
        call("lb_cast_bf16", ptr(t), ptr(out), t.numel())
        return out
    return t.to(dtype)


def _as_act(t):
    """Bring a tensor to the library's activation form: CUDA, storage type per store_dtype(), channels-last (4-D) or
    contiguous (other ranks).  Anything else is converted once at the boundary."""
    if getattr(t, "_lb_unwritten", False):
        raise RuntimeError("this norm output exists only as RootTanh(y) / RootTanh'(y) (emit='act'): its values were never written")
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    if t.dim() == 4:
        want = store_dtype(t.shape)
        if t.is_contiguous(memory_format=_CL):
            if t.stride() != _cl_strides(t.shape):          # size-1 dims leave strides ambiguous: same memory, canonical strides
                t = t.as_strided(t.shape, _cl_strides(t.shape))
            return _cast(t, want)
        if t.dtype != torch.float32:
            return _cast(t.contiguous(memory_format=_CL), want)
        return _to_channels_last(t, want)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _to_channels_last(t, dtype=torch.float32):
    """fp32 NCHW-contiguous -> channels-last (fp32 or bf16 storage) through lb_nchw_to_nhwc (no torch copy kernel)."""
    if not t.is_contiguous():
        t = t.contiguous()
    b, c, h, w = t.shape
    out = torch.empty_strided((b, c, h, w), (h * w * c, 1, w * c, c), dtype=dtype, device=t.device)
    call("lb_nchw_to_nhwc", ptr(t), ptr(out), b, c, h * w, BF16 if dtype == torch.bfloat16 else F32)
    return out


def to_nchw(t):
    """channels-last (either storage type) -> fp32 NCHW-contiguous copy (model boundary, e.g. image dumps)."""
    t = _as_act(t)
    b, c, h, w = t.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=t.device)
    call("lb_nhwc_to_nchw", ptr(t), ptr(out), b, c, h * w, _dt(t))
    return out


def _new_act(shape, like, dtype=None):
    if dtype is None:
        dtype = store_dtype(shape)
    if len(shape) == 4:
        return torch.empty_strided(tuple(shape), _cl_strides(shape), dtype=dtype, device=like.device)
    return torch.empty(shape, dtype=dtype, device=like.device)


def _bpc(t):
    """(batch, pixels, channels) of an activation."""
    if t.dim() == 4:
        return t.shape[0], t.shape[2] * t.shape[3], t.shape[1]
    if t.dim() == 3:           # [B, C, L] handled by callers via reshape to 4-D
        raise ValueError("3-D activations must be viewed as [B,C,L,1] first")
    return t.shape[0], 1, t.shape[1]


_STAT_WORK = {}


def _stat_work(device):
    """Scratch of the ordered grid reductions (lb_norm_stats / lb_gate_fwd_stats): zero-filled once, the kernels leave it
    zeroed.  One per device: every kernel of this library runs on torch's current stream, in order."""
    key = (device.type, device.index)
    w = _STAT_WORK.get(key)
    if w is None:
        w = torch.zeros(_lib.lib().lb_stat_work_doubles(), dtype=torch.float64, device=device)
        _STAT_WORK[key] = w
    return w


def _grad_sink(param):
    """(buffer the kernels accumulate into, value to hand back to autograd)."""
    arena = getattr(param, "_lb_grad", None)
    if arena is not None:
        if param.grad is None:                      # a plain nn.Module.zero_grad(set_to_none=True) dropped it: start from zero
            call("lb_fill", ptr(arena), arena.numel(), 0.0)
            param.grad = arena
        return arena, None
    fresh = torch.zeros_like(param, memory_format=torch.contiguous_format)
    return fresh, fresh


def _match(g, ref):
    """Bring an incoming gradient to the layout and storage type of `ref` (same logical shape)."""
    if g.shape != ref.shape:
        g = g.expand(tuple(ref.shape))
    if g.stride() != tuple(ref.stride()) or g.dtype != ref.dtype:
        if g.dim() == 4:
            if not g.is_contiguous(memory_format=_CL):
                g = _to_channels_last(g.float(), ref.dtype)     # NCHW gradients only arrive at the API boundary
            g = _cast(g, ref.dtype)
        else:
            g = _cast(g.contiguous(), ref.dtype)
        if g.stride() != tuple(ref.stride()):   # size-1 dims make strides ambiguous; data is identical
            g = g.as_strided(tuple(ref.shape), tuple(ref.stride()))
    return g


# ------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------
class RootTanhFn(torch.autograd.Function):
    """libs/activation.py:9-36."""

    @staticmethod
    def forward(ctx, x, growth):
        x = _as_act(x)
        y = torch.empty_like(x)
        call("lb_roottanh_fwd", ptr(x), ptr(y), x.numel(), growth, _dt(x))
        ctx.save_for_backward(x)
        ctx.growth = growth
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _match(g, x)
        dx = torch.empty_like(x)
        call("lb_roottanh_bwd", ptr(x), ptr(g), ptr(dx), x.numel(), ctx.growth, _dt(x))
        return dx, None


class TanhFn(torch.autograd.Function):
    """libs/models.py:66."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        y = torch.empty_like(x)
        call("lb_tanh_fwd", ptr(x), ptr(y), x.numel(), _dt(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _match(g, y)
        dx = torch.empty_like(y)
        call("lb_tanh_bwd", ptr(y), ptr(g), ptr(dx), y.numel(), _dt(y))
        return dx


class HingeFn(torch.autograd.Function):
    """libs/utils.py:133-134."""

    @staticmethod
    def forward(ctx, x):
        x = x.float().contiguous()
        y = torch.empty_like(x)
        call("lb_hinge_fwd", ptr(x), ptr(y), x.numel())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.float().contiguous()
        dx = torch.empty_like(x)
        call("lb_hinge_bwd", ptr(x), ptr(g), ptr(dx), x.numel())
        return dx


# ------------------------------------------------------------------------------------------
# whole-tensor norm
# ------------------------------------------------------------------------------------------
class GradMailbox:
    """A gated residual block out = (gamma * L(norm(x)) + 1) * x reads its input twice (merge.py:46-62 with block.py's
    Norm-wrapped layer modules), so x receives two gradients: the gate's (gamma y + 1) g and the one that comes back through
    the norm.  Autograd would add them with a kernel of its own; instead the gate's backward leaves its share here and the
    norm's input-gradient kernel, which always runs later (its output feeds the gate's y), adds it while it writes dx
    (`add` of lb_norm_bwd_apply).  Armed by the norm's forward only when its backward will produce dx."""
    __slots__ = ("armed", "dx")

    def __init__(self):
        self.armed = False
        self.dx = None


class WholeNormFn(torch.autograd.Function):
    """libs/inplace_norm.py:7-45 (MeanSubMulDivAdd + x.std() folded into one op).

    gain: [1,C,1,1] parameter or [B,C,1,1] style tensor; bias: [1,C,1,1]."""

    @staticmethod
    def forward(ctx, x, gain, bias, emit=None, mailbox=None):
        """emit == "act": the module that follows starts with RootTanh -> conv (conv.py:22-24) and nothing else reads the
        norm output.  In the tensor-core configuration the pass then writes RootTanh(y), that convolution's GEMM
        operand (`_lb_act16`), and -- when a backward pass will follow -- RootTanh'(y) (`_lb_dact16`), the factor its
        input gradient is multiplied by, INSTEAD of y: the returned tensor is a placeholder (`_lb_unwritten`)."""
        x = _as_act(x)
        dt = _dt(x)
        b, p, c = _bpc(x)
        per_sample = gain.shape[0] != 1          # [B,C,1,1] style gain (B == 1 degenerates to the shared form)
        if gain.numel() != (b if per_sample else 1) * c or bias.numel() != c:
            raise ValueError(f"norm: gain {tuple(gain.shape)} / bias {tuple(bias.shape)} do not fit {tuple(x.shape)}")
        gain_c = gain.float().contiguous()
        ready = getattr(x, "_lb_sums", None)
        if ready is not None:
            sums = ready.clone()                  # the producer (gate kernel) already reduced them; clone: all-reduced in place
        else:
            sums = torch.empty(2, dtype=torch.float64, device=x.device)
            call("lb_norm_stats", ptr(x), x.numel(), ptr(sums), ptr(_stat_work(x.device)), dt)
        n_total = float(x.numel()) * dist.all_reduce_sum_(sums)
        stats = torch.empty(4, dtype=torch.float32, device=x.device)
        call("lb_norm_finalize", ptr(sums), n_total, ptr(stats))
        y = torch.empty_like(x)
        gbs = c if per_sample else 0
        from .config import CFG
        if emit == "act" and dt == BF16 and c % 8 == 0 and x.dim() == 4 and CFG.ROOTTANH_GROWTH == 4:
            act = torch.empty_like(x)
            dact = torch.empty_like(x) if any(ctx.needs_input_grad) else None
            call("lb_norm_apply_ex", ptr(x), ptr(stats), ptr(gain_c), gbs, ptr(bias), None, ptr(act), ptr(dact), b, p, c, dt)
            y._lb_act16, y._lb_dact16, y._lb_unwritten = act, dact, True
        else:
            call("lb_norm_apply", ptr(x), ptr(stats), ptr(gain_c), gbs, ptr(bias), ptr(y), b, p, c, dt)
        ctx.save_for_backward(x, gain_c, stats)
        ctx.per_sample = per_sample
        ctx.gain_param, ctx.bias_param = gain, bias
        ctx.mailbox = mailbox if (mailbox is not None and ctx.needs_input_grad[0]) else None
        if ctx.mailbox is not None:
            mailbox.armed = True
        return y

    @staticmethod
    def backward(ctx, g):
        x, gain, stats = ctx.saved_tensors
        g = _match(g, x)
        dt = _dt(x)
        b, p, c = _bpc(x)
        part = torch.zeros((2, b, c), dtype=torch.float32, device=x.device)
        call("lb_norm_bwd_reduce", ptr(x), ptr(g), ptr(stats), ptr(part[0]), ptr(part[1]), b, p, c, dt)
        need_gain, need_bias = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if ctx.per_sample:
            dgain = torch.zeros_like(gain)           # the kernel always writes the per-sample slots
            dgain_ret = dgain.view(ctx.gain_param.shape) if need_gain else None
        elif need_gain:
            dgain, dgain_ret = _grad_sink(ctx.gain_param)
        else:
            dgain, dgain_ret = None, None
        if need_bias:
            dbias, dbias_ret = _grad_sink(ctx.bias_param)
        else:
            dbias, dbias_ret = None, None
        sc = torch.empty(2, dtype=torch.float64, device=x.device)
        call("lb_norm_bwd_finalize", ptr(part[0]), ptr(part[1]), ptr(gain), c if ctx.per_sample else 0, ptr(stats), b, c,
             ptr(dgain), ptr(dbias), ptr(sc))
        if need_gain and not ctx.per_sample:
            dist.grad_written(ctx.gain_param)
        if need_bias:
            dist.grad_written(ctx.bias_param)
        dist.all_reduce_sum_(sc)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            add = None
            if ctx.mailbox is not None and ctx.mailbox.dx is not None:
                add, ctx.mailbox.dx = ctx.mailbox.dx, None          # the gate's share of dL/dx (GradMailbox)
                if add.shape != x.shape or add.stride() != x.stride() or add.dtype != x.dtype:
                    raise RuntimeError("gradient mailbox: the gate and the norm of a residual block saw different inputs")
            call("lb_norm_bwd_apply", ptr(x), ptr(g), ptr(stats), ptr(gain), c if ctx.per_sample else 0, ptr(sc), ptr(add), ptr(dx),
                 b, p, c, dt)
        return dx, dgain_ret, dbias_ret, None, None


# ------------------------------------------------------------------------------------------
# gated residual
# ------------------------------------------------------------------------------------------
class GateFn(torch.autograd.Function):
    """libs/merge.py:19-39.  y is full-shape, or a [B,C,1,1] gate broadcast over pixels."""

    @staticmethod
    def forward(ctx, x, y, gamma, strict_reference, mailbox=None):
        ctx.mailbox = mailbox
        x = _as_act(x)
        dt = _dt(x)
        b, p, c = _bpc(x)
        bcast = y.shape != x.shape
        y_in_dtype = y.dtype
        if bcast:
            if tuple(y.shape[:2]) != (b, c) or y.numel() != b * c:
                raise ValueError(f"gate: cannot broadcast {tuple(y.shape)} over {tuple(x.shape)}")
            y = _cast(y.contiguous() if y.dim() != 4 else _as_act(y), x.dtype)
        else:
            y = _match(y, x)
        out = torch.empty_like(x)
        if x.dim() == 4 and c % 4 == 0 and x.numel() < (1 << 32):
            # every gate output is normalised next (block.py:46-51): leave its (sum, sum^2) for WholeNormFn
            sums = torch.empty(2, dtype=torch.float64, device=x.device)
            call("lb_gate_fwd_stats", ptr(x), ptr(y), ptr(gamma), ptr(out), ptr(sums), ptr(_stat_work(x.device)), b, p, c,
                 int(bcast), dt)
            out._lb_sums = sums
        else:
            call("lb_gate_fwd", ptr(x), ptr(y), ptr(gamma), ptr(out), b, p, c, int(bcast), dt)
        ctx.save_for_backward(x, y, gamma)
        ctx.bcast, ctx.strict, ctx.gamma_param, ctx.y_dtype = bcast, strict_reference, gamma, y_in_dtype
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, gamma = ctx.saved_tensors
        g = _match(g, x)
        dt = _dt(x)
        b, p, c = _bpc(x)
        dx = torch.empty_like(x)
        if ctx.bcast:
            dy, dyb = None, torch.zeros((b, c), dtype=torch.float32, device=x.device)
        else:
            dy, dyb = torch.empty_like(y), None
        if ctx.needs_input_grad[2]:
            dgamma, dgamma_ret = _grad_sink(ctx.gamma_param)
        else:
            dgamma, dgamma_ret = None, None
        call("lb_gate_bwd", ptr(x), ptr(y), ptr(gamma), ptr(g), ptr(dx), ptr(dy), ptr(dyb), ptr(dgamma), b, p, c, int(ctx.bcast),
             int(ctx.strict), dt)
        if ctx.needs_input_grad[2]:
            dist.grad_written(ctx.gamma_param)
        if ctx.bcast:
            dy = _cast(dyb, ctx.y_dtype).view(y.shape)
        if ctx.mailbox is not None and ctx.mailbox.armed and ctx.needs_input_grad[0]:
            ctx.mailbox.dx, dx = dx, None            # added by the norm's input-gradient kernel (GradMailbox)
        return dx, dy, dgamma_ret, None, None


# ------------------------------------------------------------------------------------------
# softmax
# ------------------------------------------------------------------------------------------
def _softmax_work(b, p, c, dev):
    n = _lib.lib().lb_softmax_pixels_work_floats(b, p, c)
    return (torch.empty(n, dtype=torch.float32, device=dev), n) if n else (None, 0)


class SoftmaxPixelsFn(torch.autograd.Function):
    """Softmax over HW for every (b,c) (attention.py:47 on the [B,F,HW] view)."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, p, c = _bpc(x)
        y = torch.empty_like(x)
        work, n = _softmax_work(b, p, c, x.device)
        call("lb_softmax_pixels_fwd", ptr(x), ptr(y), b, p, c, ptr(work), n, _dt(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _match(g, y)
        b, p, c = _bpc(y)
        dx = torch.empty_like(y)
        work, n = _softmax_work(b, p, c, y.device)
        call("lb_softmax_pixels_bwd", ptr(y), ptr(g), ptr(dx), b, p, c, ptr(work), n, _dt(y))
        return dx


class SoftmaxChannelsFn(torch.autograd.Function):
    """Softmax(dim=1) (attention.py:35).  Pixel count is 1 on the reference's path ([B,F,1,1])."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, p, c = _bpc(x)
        y = torch.empty_like(x)
        call("lb_softmax_rows_fwd", ptr(x), ptr(y), b * p, c, _dt(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _match(g, y)
        b, p, c = _bpc(y)
        dx = torch.empty_like(y)
        call("lb_softmax_rows_bwd", ptr(y), ptr(g), ptr(dx), b * p, c, _dt(y))
        return dx


# ------------------------------------------------------------------------------------------
# skip-path resampling
# ------------------------------------------------------------------------------------------
class FeaturePoolFn(torch.autograd.Function):
    """libs/scale.py:12-16."""

    @staticmethod
    def forward(ctx, x, c_out):
        x = _as_act(x)
        b, c, h, w = x.shape
        y = _new_act((b, c_out, h, w), x, x.dtype)
        call("lb_featpool_fwd", ptr(x), ptr(y), b, h, w, c, c_out, _dt(x))
        ctx.dims = (b, c, h, w, c_out)
        return y

    @staticmethod
    def backward(ctx, g):
        b, c, h, w, c_out = ctx.dims
        g = _as_act(g)
        dx = _new_act((b, c, h, w), g, g.dtype)
        call("lb_featpool_bwd", ptr(g), ptr(dx), b, h, w, c, c_out, _dt(g))
        return dx, None


class Upsample2xFn(torch.autograd.Function):
    """nn.Upsample(mode='bilinear', scale_factor=2, align_corners=False) (scale.py:37-38)."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, c, h, w = x.shape
        y = _new_act((b, c, 2 * h, 2 * w), x, x.dtype)
        call("lb_upsample2x_fwd", ptr(x), ptr(y), b, h, w, c, _dt(x))
        ctx.dims = (b, c, h, w)
        return y

    @staticmethod
    def backward(ctx, g):
        b, c, h, w = ctx.dims
        g = _as_act(g)
        dx = _new_act((b, c, h, w), g, g.dtype)
        call("lb_upsample2x_bwd", ptr(g), ptr(dx), b, h, w, c, _dt(g))
        return dx


class AvgPool2Fn(torch.autograd.Function):
    """nn.AvgPool2d(2, 2) (scale.py:40)."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, c, h, w = x.shape
        y = _new_act((b, c, h // 2, w // 2), x, x.dtype)
        call("lb_avgpool2_fwd", ptr(x), ptr(y), b, h, w, c, _dt(x))
        ctx.dims = (b, c, h, w)
        return y

    @staticmethod
    def backward(ctx, g):
        b, c, h, w = ctx.dims
        g = _as_act(g)
        dx = _new_act((b, c, h, w), g, g.dtype)
        call("lb_avgpool2_bwd", ptr(g), ptr(dx), b, h, w, c, _dt(g))
        return dx


def _esz(t):
    return 2 if t.dtype == torch.bfloat16 else 4


class CatFn(torch.autograd.Function):
    """torch.cat([a, b], dim=1) (style chain block.py:123; CatModule merge.py:15) via lb_copy_rows:
    channels are the contiguous axis of both [B,C] vectors and channels-last [B,C,H,W] maps."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _as_act(a), _as_act(b)
        ca, cb = a.shape[1], b.shape[1]
        rows = a.numel() // ca
        out = _new_act((a.shape[0], ca + cb, *a.shape[2:]), a)
        call("lb_copy_rows", ptr(a), ca, ptr(out), ca + cb, rows, ca, 0, _dt(a), _dt(out))
        call("lb_copy_rows", ptr(b), cb, out.data_ptr() + _esz(out) * ca, ca + cb, rows, cb, 0, _dt(b), _dt(out))
        ctx.meta = (ca, cb, rows, tuple(a.shape), tuple(b.shape), a.dtype, b.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        ca, cb, rows, sa, sb, da_t, db_t = ctx.meta
        g = _as_act(g)
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _new_act(sa, g, da_t)
            call("lb_copy_rows", ptr(g), ca + cb, ptr(da), ca, rows, ca, 0, _dt(g), _dt(da))
        if ctx.needs_input_grad[1]:
            db = _new_act(sb, g, db_t)
            call("lb_copy_rows", g.data_ptr() + _esz(g) * ca, ca + cb, ptr(db), cb, rows, cb, 0, _dt(g), _dt(db))
        return da, db


# functional aliases -------------------------------------------------------------------------
def roottanh(x, growth=4):
    return RootTanhFn.apply(x, growth)


def whole_norm(x, gain, bias, emit=None, mailbox=None):
    return WholeNormFn.apply(x, gain, bias, emit, mailbox)


def gate(x, y, gamma, strict_reference=True, mailbox=None):
    return GateFn.apply(x, y, gamma, strict_reference, mailbox)


# the convolution family lives in conv_fn.py (imports helpers from this module, hence the late import)
from .conv_fn import ConvSpec, SNConvFn, activated_pair, invalidate_packs, power_iterate, sn_conv  # noqa: E402,F401
