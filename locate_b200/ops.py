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


def _as_act(t):
    """fp32, CUDA, channels-last (4-D) or contiguous (other ranks)."""
    if getattr(t, "_lb_unwritten", False):
        raise RuntimeError("this tensor's fp32 values were never written (only its bf16 GEMM operand exists)")
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() == 4:
        return t if t.is_contiguous(memory_format=_CL) else _to_channels_last(t)
    return t.contiguous()


def _to_channels_last(t):
    """NCHW-contiguous -> channels-last through lb_nchw_to_nhwc (no torch copy kernel)."""
    if not t.is_contiguous():
        t = t.contiguous()
    b, c, h, w = t.shape
    out = torch.empty_strided((b, c, h, w), (h * w * c, 1, w * c, c), dtype=t.dtype, device=t.device)
    call("lb_nchw_to_nhwc", ptr(t), ptr(out), b, c, h * w)
    return out


def to_nchw(t):
    """channels-last -> NCHW-contiguous copy (model boundary, e.g. image dumps)."""
    t = _as_act(t)
    b, c, h, w = t.shape
    out = torch.empty((b, c, h, w), dtype=t.dtype, device=t.device)
    call("lb_nhwc_to_nchw", ptr(t), ptr(out), b, c, h * w)
    return out


def _new_act(shape, like):
    if len(shape) == 4:
        b, c, h, w = shape
        return torch.empty_strided((b, c, h, w), (h * w * c, 1, w * c, c), dtype=torch.float32, device=like.device)
    return torch.empty(shape, dtype=torch.float32, device=like.device)


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


# ------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------
class RootTanhFn(torch.autograd.Function):
    """libs/activation.py:9-36."""

    @staticmethod
    def forward(ctx, x, growth):
        x = _as_act(x)
        y = torch.empty_like(x)
        call("lb_roottanh_fwd", ptr(x), ptr(y), x.numel(), growth)
        ctx.save_for_backward(x)
        ctx.growth = growth
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _match(g, x)
        dx = torch.empty_like(x)
        call("lb_roottanh_bwd", ptr(x), ptr(g), ptr(dx), x.numel(), ctx.growth)
        return dx, None


class TanhFn(torch.autograd.Function):
    """libs/models.py:66."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        y = torch.empty_like(x)
        call("lb_tanh_fwd", ptr(x), ptr(y), x.numel())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _match(g, y)
        dx = torch.empty_like(y)
        call("lb_tanh_bwd", ptr(y), ptr(g), ptr(dx), y.numel())
        return dx


class HingeFn(torch.autograd.Function):
    """libs/utils.py:133-134."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        call("lb_hinge_fwd", ptr(x), ptr(y), x.numel())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(x)
        call("lb_hinge_bwd", ptr(x), ptr(g), ptr(dx), x.numel())
        return dx


def _match(g, ref):
    """Bring an incoming gradient to the layout of `ref` (same logical shape)."""
    if g.shape != ref.shape:
        g = g.expand_as(ref)
    if g.stride() != ref.stride() or g.dtype != torch.float32:
        g = _as_act(g.float())
        if g.stride() != ref.stride():          # size-1 dims make strides ambiguous; data is identical
            g = g.as_strided(ref.shape, ref.stride())
    return g


# ------------------------------------------------------------------------------------------
# whole-tensor norm
# ------------------------------------------------------------------------------------------
class WholeNormFn(torch.autograd.Function):
    """libs/inplace_norm.py:7-45 (MeanSubMulDivAdd + x.std() folded into one op).

    gain: [1,C,1,1] parameter or [B,C,1,1] style tensor; bias: [1,C,1,1]."""

    @staticmethod
    def forward(ctx, x, gain, bias, emit=None):
        """emit: None | "act" | "plain" -- also produce the bf16 operand of the conv that consumes the result
        (RootTanh applied first for "act"), attached to the output as `_lb_act16` / `_lb_plain16`.  With "plain" nothing
        but that conv reads the result, so the fp32 tensor is allocated but NOT written (`_lb_unwritten`)."""
        x = _as_act(x)
        b, p, c = _bpc(x)
        per_sample = gain.shape[0] != 1          # [B,C,1,1] style gain (B == 1 degenerates to the shared form)
        if gain.numel() != (b if per_sample else 1) * c or bias.numel() != c:
            raise ValueError(f"norm: gain {tuple(gain.shape)} / bias {tuple(bias.shape)} do not fit {tuple(x.shape)}")
        gain_c = gain.contiguous()
        ready = getattr(x, "_lb_sums", None)
        if ready is not None:
            sums = ready.clone()                  # the producer (gate kernel) already reduced them; clone: all-reduced in place
        else:
            sums = torch.empty(2, dtype=torch.float64, device=x.device)
            call("lb_norm_stats", ptr(x), x.numel(), ptr(sums), ptr(_stat_work(x.device)))
        n_total = float(x.numel()) * dist.all_reduce_sum_(sums)
        stats = torch.empty(4, dtype=torch.float32, device=x.device)
        call("lb_norm_finalize", ptr(sums), n_total, ptr(stats))
        y = torch.empty_like(x)
        from .config import CFG
        if emit is not None and CFG.PRECISION == "bf16" and c % 8 == 0 and x.dim() == 4:
            y16 = torch.empty_strided(x.shape, x.stride(), dtype=torch.bfloat16, device=x.device)
            lazy = emit == "plain" and c >= 32     # wide enough that the consumer is always a tensor-core GEMM
            call("lb_norm_apply_ex", ptr(x), ptr(stats), ptr(gain_c), c if per_sample else 0, ptr(bias), None if lazy else ptr(y),
                 ptr(y16), 1 if emit == "act" else 0, b, p, c)
            if emit == "act":
                y._lb_act16 = y16
            else:
                y._lb_plain16 = y16
                y._lb_unwritten = True
        else:
            call("lb_norm_apply", ptr(x), ptr(stats), ptr(gain_c), c if per_sample else 0, ptr(bias), ptr(y), b, p, c)
        ctx.save_for_backward(x, gain_c, stats)
        ctx.per_sample = per_sample
        ctx.gain_param, ctx.bias_param = gain, bias
        return y

    @staticmethod
    def backward(ctx, g):
        x, gain, stats = ctx.saved_tensors
        g = _match(g, x)
        b, p, c = _bpc(x)
        part = torch.zeros((2, b, c), dtype=torch.float32, device=x.device)
        call("lb_norm_bwd_reduce", ptr(x), ptr(g), ptr(stats), ptr(part[0]), ptr(part[1]), b, p, c)
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
        dist.all_reduce_sum_(sc)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("lb_norm_bwd_apply", ptr(x), ptr(g), ptr(stats), ptr(gain), c if ctx.per_sample else 0, ptr(sc), ptr(dx), b, p, c)
        return dx, dgain_ret, dbias_ret, None


# ------------------------------------------------------------------------------------------
# gated residual
# ------------------------------------------------------------------------------------------
class GateFn(torch.autograd.Function):
    """libs/merge.py:19-39.  y is full-shape, or a [B,C,1,1] gate broadcast over pixels."""

    @staticmethod
    def forward(ctx, x, y, gamma, strict_reference):
        x = _as_act(x)
        b, p, c = _bpc(x)
        bcast = y.shape != x.shape
        if bcast:
            if tuple(y.shape[:2]) != (b, c) or y.numel() != b * c:
                raise ValueError(f"gate: cannot broadcast {tuple(y.shape)} over {tuple(x.shape)}")
            y = y.contiguous()
        else:
            y = _match(y, x)
        out = torch.empty_like(x)
        if x.dim() == 4 and c % 4 == 0 and x.numel() < (1 << 32):
            # every gate output is normalised next (block.py:46-51): leave its (sum, sum^2) for WholeNormFn
            sums = torch.empty(2, dtype=torch.float64, device=x.device)
            call("lb_gate_fwd_stats", ptr(x), ptr(y), ptr(gamma), ptr(out), ptr(sums), ptr(_stat_work(x.device)), b, p, c, int(bcast))
            out._lb_sums = sums
        else:
            call("lb_gate_fwd", ptr(x), ptr(y), ptr(gamma), ptr(out), b, p, c, int(bcast))
        ctx.save_for_backward(x, y, gamma)
        ctx.bcast, ctx.strict, ctx.gamma_param = bcast, strict_reference, gamma
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, gamma = ctx.saved_tensors
        g = _match(g, x)
        b, p, c = _bpc(x)
        dx = torch.empty_like(x)
        dy = torch.zeros_like(y) if ctx.bcast else torch.empty_like(y)
        if ctx.needs_input_grad[2]:
            dgamma, dgamma_ret = _grad_sink(ctx.gamma_param)
        else:
            dgamma, dgamma_ret = None, None
        call("lb_gate_bwd", ptr(x), ptr(y), ptr(gamma), ptr(g), ptr(dx), ptr(dy), ptr(dgamma), b, p, c, int(ctx.bcast),
             int(ctx.strict))
        return dx, dy, dgamma_ret, None


# ------------------------------------------------------------------------------------------
# softmax
# ------------------------------------------------------------------------------------------
class SoftmaxPixelsFn(torch.autograd.Function):
    """Softmax over HW for every (b,c) (attention.py:47 on the [B,F,HW] view)."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, p, c = _bpc(x)
        y = torch.empty_like(x)
        call("lb_softmax_pixels_fwd", ptr(x), ptr(y), b, p, c)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _match(g, y)
        b, p, c = _bpc(y)
        dx = torch.empty_like(y)
        call("lb_softmax_pixels_bwd", ptr(y), ptr(g), ptr(dx), b, p, c)
        return dx


class SoftmaxChannelsFn(torch.autograd.Function):
    """Softmax(dim=1) (attention.py:35).  Pixel count is 1 on the reference's path ([B,F,1,1])."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, p, c = _bpc(x)
        y = torch.empty_like(x)
        call("lb_softmax_rows_fwd", ptr(x), ptr(y), b * p, c)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _match(g, y)
        b, p, c = _bpc(y)
        dx = torch.empty_like(y)
        call("lb_softmax_rows_bwd", ptr(y), ptr(g), ptr(dx), b * p, c)
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
        y = _new_act((b, c_out, h, w), x)
        call("lb_featpool_fwd", ptr(x), ptr(y), b, h, w, c, c_out)
        ctx.dims = (b, c, h, w, c_out)
        return y

    @staticmethod
    def backward(ctx, g):
        b, c, h, w, c_out = ctx.dims
        g = _as_act(g)
        dx = _new_act((b, c, h, w), g)
        call("lb_featpool_bwd", ptr(g), ptr(dx), b, h, w, c, c_out)
        return dx, None


class Upsample2xFn(torch.autograd.Function):
    """nn.Upsample(mode='bilinear', scale_factor=2, align_corners=False) (scale.py:37-38)."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, c, h, w = x.shape
        y = _new_act((b, c, 2 * h, 2 * w), x)
        call("lb_upsample2x_fwd", ptr(x), ptr(y), b, h, w, c)
        ctx.dims = (b, c, h, w)
        return y

    @staticmethod
    def backward(ctx, g):
        b, c, h, w = ctx.dims
        g = _as_act(g)
        dx = _new_act((b, c, h, w), g)
        call("lb_upsample2x_bwd", ptr(g), ptr(dx), b, h, w, c)
        return dx


class AvgPool2Fn(torch.autograd.Function):
    """nn.AvgPool2d(2, 2) (scale.py:40)."""

    @staticmethod
    def forward(ctx, x):
        x = _as_act(x)
        b, c, h, w = x.shape
        y = _new_act((b, c, h // 2, w // 2), x)
        call("lb_avgpool2_fwd", ptr(x), ptr(y), b, h, w, c)
        ctx.dims = (b, c, h, w)
        return y

    @staticmethod
    def backward(ctx, g):
        b, c, h, w = ctx.dims
        g = _as_act(g)
        dx = _new_act((b, c, h, w), g)
        call("lb_avgpool2_bwd", ptr(g), ptr(dx), b, h, w, c)
        return dx


class CatFn(torch.autograd.Function):
    """torch.cat([a, b], dim=1) (style chain block.py:123; CatModule merge.py:15) via lb_copy_rows:
    channels are the contiguous axis of both [B,C] vectors and channels-last [B,C,H,W] maps."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _as_act(a), _as_act(b)
        ca, cb = a.shape[1], b.shape[1]
        rows = a.numel() // ca
        out = _new_act((a.shape[0], ca + cb, *a.shape[2:]), a)
        call("lb_copy_rows", ptr(a), ca, ptr(out), ca + cb, rows, ca, 0)
        call("lb_copy_rows", ptr(b), cb, out.data_ptr() + 4 * ca, ca + cb, rows, cb, 0)
        ctx.meta = (ca, cb, rows, tuple(a.shape), tuple(b.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        ca, cb, rows, sa, sb = ctx.meta
        g = _as_act(g)
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _new_act(sa, g)
            call("lb_copy_rows", ptr(g), ca + cb, ptr(da), ca, rows, ca, 0)
        if ctx.needs_input_grad[1]:
            db = _new_act(sb, g)
            call("lb_copy_rows", g.data_ptr() + 4 * ca, ca + cb, ptr(db), cb, rows, cb, 0)
        return da, db


# functional aliases -------------------------------------------------------------------------
def roottanh(x, growth=4):
    return RootTanhFn.apply(x, growth)


def whole_norm(x, gain, bias, emit=None):
    return WholeNormFn.apply(x, gain, bias, emit)


def gate(x, y, gamma, strict_reference=True):
    return GateFn.apply(x, y, gamma, strict_reference)


# the convolution family lives in conv_fn.py (imports helpers from this module, hence the late import)
from .conv_fn import ConvSpec, SNConvFn, activated_pair, invalidate_packs, power_iterate, sn_conv  # noqa: E402,F401
