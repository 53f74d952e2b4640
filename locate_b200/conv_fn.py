"""The spectral-normed convolution family as one autograd Function (fwd, dgrad, wgrad + the power iteration).

Two kernel paths, chosen per layer:
  * CFG.PRECISION == "bf16" and the geometry is covered by the tensor-core kernels (stride 1|2, <= 64 taps per phase):
    tcgen05 implicit GEMM (lb_conv_tc_gemm_ws / lb_conv_tc_gemm_ex / lb_wgrad_tc), bf16 operands, fp32 accumulate.
    Wide activations ([B,C,H,W], C % 8 == 0) are STORED as bf16 (ops.store_dtype), so a GEMM reads its operand where the
    producer left it and writes the next one from its epilogue: no cast passes, RootTanh / RootTanh' fused.
  * otherwise: the fp32 SIMT gather-GEMM (lb_conv_gemm / lb_conv_wgrad), everything fp32.
`pre_act` fuses the RootTanh that precedes every conv of ActivatedBaseConv (conv.py:22-24) into this Function.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import ConvGeom, call, ptr
from .config import CFG
from . import dist
from .ops import BF16, F32, _as_act, _cast, _conv_work, _dt, _esz, _grad_sink, _match, _new_act, _timed_call, store_dtype

# weight packs are reused while the weights are unchanged (3 discriminator passes per D step);
# optim.Nadam.step / load_state_dict bump the epoch, torch's version counter covers in-place torch ops.
_PACK_EPOCH = [0]


def invalidate_packs():
    _PACK_EPOCH[0] += 1


class ConvSpec:
    """Static description of one spectral-normed linear map.

    kind: 'conv' (weight [Cout,Cin,kh,kw]), 'convT' (weight [Cin,Cout,kh,kw]), 'linear'
    (weight [out,in]), 'conv1d' (weight [Cout,Cin,1] applied per pixel)."""

    def __init__(self, kind, cin, cout, kh=1, kw=1, stride=1, pad=0, groups=1):
        self.kind, self.cin, self.cout = kind, cin, cout
        self.kh, self.kw, self.stride, self.pad = kh, kw, stride, pad
        self.groups = groups           # 1, or SEPARABLE's two grouped forms: depthwise (groups = cin = cout) / full-extent

    def out_hw(self, h, w):
        if self.kind == "convT":
            return (h - 1) * self.stride - 2 * self.pad + self.kh, (w - 1) * self.stride - 2 * self.pad + self.kw
        return (h + 2 * self.pad - self.kh) // self.stride + 1, (w + 2 * self.pad - self.kw) // self.stride + 1

    @property
    def taps(self):
        return self.kh * self.kw

    @property
    def sn_shape(self):
        """(height, width) of the matrix view spectral norm iterates on (spectral_norm.py:26)."""
        if self.groups > 1:            # weight (cout, cin / groups, kh, kw) -- for both Conv and ConvTranspose when cin = cout
            return self.cout, (self.cin // self.groups) * self.taps
        if self.kind == "convT":
            return self.cin, self.cout * self.taps
        return self.cout, self.cin * self.taps

    # (w_sk, w_sn, w_sty, w_stx) of W(tap,k,n) in the master layout
    def strides_fwd(self):            # k = cin, n = cout
        t = self.taps
        return (self.cout * t, t, self.kw, 1) if self.kind == "convT" else (t, self.cin * t, self.kw, 1)

    def strides_dgrad(self):          # k = cout, n = cin
        t = self.taps
        return (t, self.cout * t, self.kw, 1) if self.kind == "convT" else (self.cin * t, t, self.kw, 1)


def _geom(batch, ih, iw, ic, oh, ow, oc, spec, mode, ld_in, ld_out, strides):
    g = ConvGeom()
    g.batch, g.in_h, g.in_w, g.in_c = batch, ih, iw, ic
    g.out_h, g.out_w, g.out_c = oh, ow, oc
    g.kh, g.kw, g.stride, g.pad, g.mode = spec.kh, spec.kw, spec.stride, spec.pad, mode
    g.ld_in, g.ld_out = ld_in, ld_out
    g.w_sk, g.w_sn, g.w_sty, g.w_stx = strides
    return g


def power_iterate(w_bar, u, v, spec):
    """One in-place power iteration (spectral_norm.py:21-32); returns the [sigma, 1/sigma] buffer."""
    height, width = spec.sn_shape
    sigma = torch.empty(2, dtype=torch.float32, device=w_bar.device)
    work = torch.empty(_lib.lib().lb_sn_power_iter_work_floats(height, width), dtype=torch.float32, device=w_bar.device)
    call("lb_sn_power_iter", ptr(w_bar), height, width, ptr(u), ptr(v), ptr(sigma), ptr(work))
    return sigma


class PackEpoch(list):
    """[counter] of one optimizer arena (optim.Nadam bumps it in step()); carries the arena's pack plan."""
    plan = None


class _PackPlan:
    """Every bf16 weight pack of one optimizer arena, re-packed by ONE launch (lb_conv_tc_pack_batched) when the
    optimizer has moved the weights: entries are registered the first time a (weight, direction) is packed, the
    device tables are built outside CUDA-graph capture, and the pack buffers are reused step after step."""

    def __init__(self):
        self.entries = []              # [weakref(w_bar), tag, ent] with ent = [key, pk, geom copy, w data_ptr in the table]
        self.tables = None

    def register(self, w_bar, tag, ent):
        self.entries.append((weakref.ref(w_bar), tag, ent))
        self.tables = None

    def _build(self, dev):
        lib = _lib.lib()
        rec_bytes, chunk = lib.lb_pack_rec_bytes(), lib.lb_pack_chunk_items()
        live = [(r, t, e) for r, t, e in self.entries if r() is not None]
        self.entries = live
        buf = ctypes.create_string_buffer(rec_bytes * len(live))
        chunks = []
        for i, (ref, _tag, ent) in enumerate(live):
            w = ref()
            items = lib.lb_pack_rec_fill(ptr(w), ptr(ent[1]), ctypes.byref(ent[2]), ctypes.addressof(buf) + i * rec_bytes)
            if items < 0:
                raise _lib.LocateLibraryError("lb_pack_rec_fill failed")
            ent[3] = w.data_ptr()
            chunks.extend((i, first) for first in range(0, items, chunk))
        self.tables = (torch.from_numpy(np.frombuffer(buf.raw, dtype=np.uint8).copy()).to(dev),
                       torch.tensor(chunks, dtype=torch.int32).to(dev), len(chunks))

    def repack_all(self, epoch, dev):
        """False when the batched launch cannot run now (tables to (re)build while a CUDA graph is being captured)."""
        stale_table = self.tables is None or any(r() is None or r().data_ptr() != e[3] for r, _t, e in self.entries)
        if stale_table:
            if torch.cuda.is_current_stream_capturing():
                return False
            self._build(dev)
        if not self.entries:
            return False
        recs, chunks, n_chunks = self.tables
        call("lb_conv_tc_pack_batched", ptr(recs), ptr(chunks), n_chunks)
        for ref, _tag, ent in self.entries:
            w = ref()
            ent[0] = (_PACK_EPOCH[0], epoch[0], w._version, w.data_ptr())
        return True


def _packed_weight(w_bar, g, tag):
    """bf16 [tap][n][k] copy of the master weight for geometry g (cached per weight version)."""
    epoch = getattr(w_bar, "_lb_epoch", _PACK_EPOCH)       # per-optimizer counter when the weight lives in a Nadam arena
    key = (_PACK_EPOCH[0], epoch[0], w_bar._version, w_bar.data_ptr())
    cache = getattr(w_bar, "_lb_pack", None)
    if cache is None:
        cache = w_bar._lb_pack = {}
    ent = cache.get(tag)
    if ent is not None and ent[0] == key:
        return ent[1]
    plan = None
    if isinstance(epoch, PackEpoch):
        plan = epoch.plan
        if plan is None:
            plan = epoch.plan = _PackPlan()
    if ent is None:
        n = _lib.lib().lb_conv_tc_packed_elems(ctypes.byref(g))
        ent = cache[tag] = [None, torch.empty(n, dtype=torch.bfloat16, device=w_bar.device), ConvGeom.from_buffer_copy(g), 0]
        if plan is not None:
            plan.register(w_bar, tag, ent)
    elif plan is not None and plan.repack_all(epoch, w_bar.device) and ent[0] == key:
        return ent[1]                  # the whole arena was re-packed by one launch
    call("lb_conv_tc_pack", ptr(w_bar), ptr(ent[1]), g)
    ent[0] = key
    return ent[1]


def _uv_extra(pre_sigma, u):
    """(s_fwd, du, cacc) for lb_sn_weight_grad when this layer's u / v are trainable -- the reference's loop switches
    them on for the discriminator (main.py:172), see oracle._LiveSigma -- else three NULLs."""
    uv = getattr(pre_sigma, "_lb_uv", None)
    if uv is None or not CFG.STRICT_REFERENCE or not u.requires_grad:
        return None
    du = getattr(u, "_lb_grad", None)
    if du is None:
        return None
    return uv[0], du, uv[1], uv[2]


def _sn_weight_grad(dwn, w_bar, u, v, sigma, spec, packed_taps, extra, dev, dot=None):
    """grad(W_bar) += dwn/sigma - (sum dwn*W)/sigma^2 u v^T into the gradient sink; returns what autograd gets.
    `dot`: 2 doubles already holding sum dwn*W (left by lb_wgrad_tc's reduction pass)."""
    from .ops import _stat_work
    grad_w, dw_ret = _grad_sink(w_bar)
    have_dot = dot is not None
    if not have_dot:
        dot = torch.empty(2, dtype=torch.float64, device=dev)
    height, width = spec.sn_shape
    s_fwd = du = cacc = None
    if extra is not None:
        s_fwd, du, cacc, batch = extra
        if u.grad is None:
            u.grad, v.grad = u._lb_grad, v._lb_grad
        batch.uv_pending = True
    call("lb_sn_weight_grad", ptr(dwn), None if have_dot else ptr(w_bar), ptr(u.data), ptr(v.data), ptr(sigma), ptr(grad_w), height,
         width, packed_taps, ptr(dot), ptr(_stat_work(dev)), ptr(s_fwd), ptr(du), ptr(cacc))
    dist.grad_written(w_bar)
    return dw_ret


def _tc_gemm(fl, nbytes, a_ptr, pk, alpha_ptr, bias, out_ptr, g, dev, out_dtype=F32):
    """lb_conv_tc_gemm_ws with the split-K workspace the geometry asks for (weight-bound layers; 0 bytes otherwise).
    a_ptr / out_ptr are raw device addresses (operands are often channel slices of wider rows)."""
    need = _lib.lib().lb_conv_tc_workspace_bytes(ctypes.byref(g))
    work = torch.empty(need // 4, dtype=torch.float32, device=dev) if need else None
    _timed_call("conv_tc", fl, nbytes, "lb_conv_tc_gemm_ws", a_ptr, ptr(pk), alpha_ptr, ptr(bias), out_ptr, g, ptr(work), need,
                out_dtype)


class _Like:
    """shape / stride / dtype of a tensor that is gone (what ops._match needs to lay a gradient out)."""

    def __init__(self, shape, stride, dtype):
        self.shape, self._stride, self.dtype = torch.Size(shape), tuple(stride), dtype

    def stride(self):
        return self._stride


def _bf16_rows(t, cols, cols_p, growth=0, offset=0, ld=None):
    """bf16 [rows, cols_p] copy (rows padded to 16 bytes for TMA) of `cols` columns of the fp32 rows of `t` starting at
    column `offset` -- operands whose channel count is not a multiple of 8 (RGB images, logits, vectors of the style
    chain), optionally through RootTanh."""
    ld = ld if ld is not None else t.shape[1]
    rows = t.numel() // ld
    if t.dtype == torch.bfloat16:                 # an unaligned channel slice of bf16 rows (no activation on this path)
        if growth:
            raise ValueError("RootTanh on a bf16 slice copy")
        out = torch.zeros((rows, cols_p), dtype=torch.bfloat16, device=t.device)
        call("lb_copy_rows", t.data_ptr() + 2 * offset, ld, ptr(out), cols_p, rows, cols, 0, BF16, BF16)
        return out
    out = torch.empty((rows, cols_p), dtype=torch.bfloat16, device=t.device)
    call("lb_cast_bf16_rows", t.data_ptr() + 4 * offset, ld, ptr(out), cols_p, rows, cols, growth)
    return out


class SNConvFn(torch.autograd.Function):
    """y = conv(act?(x), W_bar)/sigma (+bias) with the power iteration run inside, i.e. SpectralNorm.forward
    (spectral_norm.py:57-59) around Conv2d / ConvTranspose2d / Conv1d(k=1) / Linear (conv.py:14-20,
    attention.py:26-34,44-46, scale.py:25-34, linear.py:10), optionally preceded by RootTanh (conv.py:23-24).

    If `cat_input` the result is cat([x, y], channels) (CatModule with an identity residual, merge.py:10-16):
    the GEMM writes straight into the channel slice of the wider output."""

    @staticmethod
    def forward(ctx, x, w_bar, u, v, bias, spec, cat_input, pre_act, pre_sigma):
        ready_act = getattr(x, "_lb_act16", None) if pre_act else None
        ready_dact = getattr(x, "_lb_dact16", None) if pre_act else None
        if getattr(x, "_lb_unwritten", False):       # a norm output that exists only as RootTanh(y) / RootTanh'(y)
            if ready_act is None:
                raise RuntimeError("norm output was emitted for an activated conv, but this conv does not start with RootTanh")
            x_vals = None
        else:
            x = _as_act(x)
            x_vals = x
        is_vec = x.dim() == 2
        if is_vec:
            b, h, w_, cin = x.shape[0], 1, 1, x.shape[1]
        else:
            b, cin, h, w_ = x.shape
        if cin != spec.cin:
            raise ValueError(f"expected {spec.cin} input channels, got {cin}")
        if cat_input and pre_act:
            raise ValueError("cat_input and pre_act are exclusive")
        # one power iteration per forward call (done here, or already done for the whole model by sn_batch)
        sigma = pre_sigma if pre_sigma is not None else power_iterate(w_bar, u.data, v.data, spec)
        oh, ow = spec.out_hw(h, w_)
        if cat_input and (oh, ow) != (h, w_):
            raise ValueError("cat_input needs a size-preserving conv")
        ctot = spec.cout + (cin if cat_input else 0)
        out = _new_act((b, ctot) if is_vec else (b, ctot, oh, ow), x)
        esz_o = _esz(out)
        off = cin if cat_input else 0                 # first column of the conv result inside the output rows
        t = spec.taps
        mode = 1 if spec.kind == "convT" else 0
        lib = _lib.lib()
        x16, out16 = x.dtype == torch.bfloat16, out.dtype == torch.bfloat16

        def geoms(ld_x, ld_dy):
            gf = _geom(b, h, w_, cin, oh, ow, spec.cout, spec, mode, ld_x, ctot, spec.strides_fwd())
            gd = _geom(b, oh, ow, spec.cout, h, w_, cin, spec, 1 - mode, ld_dy, cin, spec.strides_dgrad())
            if spec.kind == "convT":      # dense = x (cin), gathered = dy (cout): dw[ci][co][tap]
                gw = _geom(b, oh, ow, spec.cout, h, w_, cin, spec, 0, ld_dy, ld_x, (t, spec.cout * t, spec.kw, 1))
            else:                         # dense = dy (cout), gathered = x (cin): dw[co][ci][tap]
                gw = _geom(b, h, w_, cin, oh, ow, spec.cout, spec, 0, ld_x, ld_dy, (t, cin * t, spec.kw, 1))
            return gf, gd, gw

        # tensor-core operands are bf16 rows of >= 16 bytes (TMA stride rule): a wide bf16 activation is its own operand,
        # anything else is copied into rows padded to 8 channels
        cin_p, cout_p = (cin + 7) // 8 * 8, (spec.cout + 7) // 8 * 8
        # the gradient of a bf16 output is its own operand too when the conv's slice of the rows is 16-byte aligned
        gy_direct = out16 and off % 8 == 0 and spec.cout % 8 == 0
        ld_x = cin if x16 else cin_p
        ld_gy = ctot if gy_direct else cout_p
        tc = fused_cat = False
        g32 = geoms(cin, spec.cout)
        # tiny channel count on one side (D stem, G's last 1x1): direct kernels, activation fused
        # (a wide -> 3 layer such as G's last 1x1 is TMA-friendly on its input side and measured faster on the tensor cores)
        small = CFG.SMALL_KERNELS and cin <= 4 and not x16 and lib.lb_conv_small_supported(ctypes.byref(g32[0])) == 1
        if CFG.PRECISION == "bf16" and not small:
            g_fwd, g_dgrad, g_wgrad = geoms(ld_x, ld_gy)
            tc = (lib.lb_conv_tc_supported(ctypes.byref(g_fwd)) == 1 and lib.lb_conv_tc_supported(ctypes.byref(g_dgrad)) == 1
                  and lib.lb_wgrad_tc_supported(ctypes.byref(g_wgrad)) == 1)
        if not tc:
            if (x16 or out16) and not small:
                raise RuntimeError("bf16-stored activations reached a layer the tensor-core kernels do not cover")
            if x_vals is None:
                raise RuntimeError("placeholder norm output reached a layer outside the tensor-core path")
            g_fwd, g_dgrad, g_wgrad = geoms(cin, spec.cout)
        fl, by = _conv_work(spec, b, h, w_, oh, ow)
        n = x.numel()
        if tc:
            growth = CFG.ROOTTANH_GROWTH if pre_act else 0
            if x16:
                if not pre_act:
                    a = x
                elif ready_act is not None:
                    a = ready_act
                else:
                    a = torch.empty_like(x)
                    call("lb_roottanh_fwd", ptr(x), ptr(a), n, growth, BF16)
            else:
                a = _bf16_rows(x, cin, cin_p, growth)
            pk = _packed_weight(w_bar, g_fwd, "fwd")
            _tc_gemm(fl, _tc_bytes(g_fwd), ptr(a), pk, sigma.data_ptr() + 4, bias, out.data_ptr() + off * esz_o, g_fwd, x.device,
                     BF16 if out16 else F32)
        elif small:
            pointwise = spec.taps == 1 and spec.stride == 1 and spec.pad == 0
            small_growth = CFG.ROOTTANH_GROWTH if pre_act else 0
            if pre_act and not pointwise:         # every input pixel feeds up to 25 taps: activate it once, not per tap
                a = torch.empty_like(x)
                call("lb_roottanh_fwd", ptr(x), ptr(a), n, CFG.ROOTTANH_GROWTH, F32)
                small_growth = 0
            else:
                a = x                             # RootTanh is applied on load; nothing is materialised
            fused_cat = bool(cat_input and pointwise and cin <= 8)     # the D stem's [x | conv(x)] row in one pass
            _timed_call("conv_small", fl, by, "lb_conv_small", ptr(a), ptr(w_bar), sigma.data_ptr() + 4, ptr(bias),
                        out.data_ptr() + (0 if fused_cat else off * esz_o), g_fwd, small_growth, None, 0, 0, 1 if fused_cat else 0,
                        _dt(out))
        else:
            if pre_act:
                a = torch.empty_like(x)
                call("lb_roottanh_fwd", ptr(x), ptr(a), n, CFG.ROOTTANH_GROWTH, F32)
            else:
                a = x
            _timed_call("conv_gemm", fl, by, "lb_conv_gemm", ptr(a), ptr(w_bar), sigma.data_ptr() + 4, ptr(bias),
                        out.data_ptr() + off * 4, g_fwd)
        if cat_input and not fused_cat:
            call("lb_copy_rows", ptr(x), cin, ptr(out), ctot, b * h * w_, cin, 0, _dt(x), _dt(out))
        # RootTanh'(x) for the backward: the factor itself when the producer left it, else x to evaluate it from
        ctx.dact_ready = bool(pre_act and ready_dact is not None)
        ctx.save_for_backward((ready_dact if ctx.dact_ready else x_vals) if pre_act else None, a, w_bar, sigma)
        ctx.u, ctx.v = u, v                      # LIVE u/v: the reference's backward reads them at backward time
        ctx.uv_extra = _uv_extra(pre_sigma, u)
        ctx.bias_param, ctx.w_param = bias, w_bar
        ctx.meta = (spec, cat_input, pre_act, tc, (b, h, w_, cin, oh, ow, ctot), g_dgrad, g_wgrad, (fl, by))
        ctx.dtypes = (x.dtype, out.dtype, tuple(x.shape), gy_direct, ld_gy)
        ctx.out_like = (tuple(out.shape), tuple(out.stride()))
        ctx.raw_a = bool(small and pre_act and a is x)   # `a` is the pre-activation: backward applies RootTanh where it needs it
        return out

    @staticmethod
    def backward(ctx, gout):
        x, a, w_bar, sigma = ctx.saved_tensors
        spec, cat_input, pre_act, tc, (b, h, w_, cin, oh, ow, ctot), g_dgrad, g_wgrad, (fl, by) = ctx.meta
        x_dtype, out_dtype, x_shape, gy_direct, ld_gy = ctx.dtypes
        out_shape, out_stride = ctx.out_like
        gout = _match(gout, _Like(out_shape, out_stride, out_dtype))
        esz_g = _esz(gout)
        off = cin if cat_input else 0
        rows = b * oh * ow
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = dw_ret = dbias_ret = None
        dact_done = False
        t = spec.taps
        dev = gout.device
        x16 = x_dtype == torch.bfloat16
        lib = _lib.lib()
        if tc:
            if gy_direct:
                gy_ptr = gout.data_ptr() + off * 2     # the conv's slice of the bf16 gradient rows, row stride ctot
                gy_keep = gout
            else:
                gy_keep = _bf16_rows(gout, spec.cout, ld_gy, 0, off, ctot)
                gy_ptr = gy_keep.data_ptr()
            if need_dx:
                dx = _new_act(x_shape, gout, x_dtype)
                pk = _packed_weight(w_bar, g_dgrad, "dgrad")
                if x16:
                    aux_ok = pre_act and lib.lb_conv_tc_ex_supported(ctypes.byref(g_dgrad), 0, cin, cin, BF16) == 1
                    if aux_ok:      # dx = dgrad(gy) * RootTanh'(x) in the GEMM epilogue (x holds the factor itself if dact_ready)
                        _timed_call("conv_tc", fl, _tc_bytes(g_dgrad), "lb_conv_tc_gemm_ex", gy_ptr, ptr(pk), sigma.data_ptr() + 4, None,
                                    None, ptr(dx), None, cin, ptr(x), cin, BF16, EX_AUX_IS_FACTOR if ctx.dact_ready else 0, g_dgrad)
                        dact_done = True
                    else:
                        _tc_gemm(fl, _tc_bytes(g_dgrad), gy_ptr, pk, sigma.data_ptr() + 4, None, ptr(dx), g_dgrad, dev, BF16)
                else:               # narrow / vector input: fp32 rows of cin
                    _tc_gemm(fl, _tc_bytes(g_dgrad), gy_ptr, pk, sigma.data_ptr() + 4, None, ptr(dx), g_dgrad, dev, F32)
            if need_dw:
                ga_ptr, de_ptr = (gy_ptr, ptr(a)) if spec.kind == "convT" else (ptr(a), gy_ptr)
                dwn, dot = _wgrad_tc(ga_ptr, de_ptr, g_wgrad, w_bar, fl, dev)
                dw_ret = _sn_weight_grad(dwn, ctx.w_param, ctx.u, ctx.v, sigma, spec, 0, ctx.uv_extra, dev, dot)
        else:
            growth = CFG.ROOTTANH_GROWTH
            gy_ptr = gout.data_ptr() + off * esz_g    # gradient of the conv output slice, row stride ctot
            g_dgrad.ld_in = ctot
            small_in = CFG.SMALL_KERNELS and cin <= 4 and not x16
            if need_dx:
                dx = torch.empty_like(a)
                if small_in and lib.lb_conv_small_supported(ctypes.byref(g_dgrad)) == 1:
                    _timed_call("conv_small", fl, by, "lb_conv_small", gy_ptr, ptr(w_bar), sigma.data_ptr() + 4, None, ptr(dx),
                                g_dgrad, 0, ptr(x) if pre_act else None, cin, growth if pre_act else 0, 0, _dt(gout))
                    dact_done = True
                else:
                    _timed_call("conv_gemm", fl, by, "lb_conv_gemm", gy_ptr, ptr(w_bar), sigma.data_ptr() + 4, None, ptr(dx), g_dgrad)
            if need_dw:
                dwn = torch.zeros_like(w_bar, memory_format=torch.contiguous_format)
                if spec.kind == "convT":
                    g_wgrad.ld_in = ctot
                else:
                    g_wgrad.ld_out = ctot
                small_w = small_in and lib.lb_conv_small_wgrad_supported(ctypes.byref(g_wgrad)) == 1
                fuse_act = ctx.raw_a and small_w and spec.kind != "convT"     # RootTanh on the gathered operand's load
                if ctx.raw_a and not fuse_act:
                    act = torch.empty_like(a)
                    call("lb_roottanh_fwd", ptr(a), ptr(act), a.numel(), growth, F32)
                    a = act
                ga_ptr, de_ptr = (gy_ptr, ptr(a)) if spec.kind == "convT" else (ptr(a), gy_ptr)
                if small_w:
                    # the wide operand (lb_conv_small_wgrad): dense for a 1x1 layer with <= 4 gathered channels, else gathered
                    ga_dt, de_dt = (_dt(gout), _dt(a)) if spec.kind == "convT" else (_dt(a), _dt(gout))
                    pointwise = spec.taps == 1 and spec.stride == 1 and spec.pad == 0
                    wide = de_dt if (pointwise and g_wgrad.in_c <= 4) else ga_dt
                    _timed_call("conv_small_wgrad", fl, by, "lb_conv_small_wgrad", ga_ptr, de_ptr, ptr(dwn), g_wgrad,
                                growth if fuse_act else 0, wide)
                else:
                    _timed_call("conv_wgrad", fl, by, "lb_conv_wgrad", ga_ptr, de_ptr, ptr(dwn), g_wgrad)
                dw_ret = _sn_weight_grad(dwn, ctx.w_param, ctx.u, ctx.v, sigma, spec, 0, ctx.uv_extra, dev)
        if need_dx:
            if pre_act and not dact_done and ctx.dact_ready:
                call("lb_mul", ptr(x), ptr(dx), ptr(dx), dx.numel(), _dt(dx))                                  # dx *= RootTanh'(x), in place
            elif pre_act and not dact_done:
                call("lb_roottanh_bwd", ptr(x), ptr(dx), ptr(dx), dx.numel(), CFG.ROOTTANH_GROWTH, _dt(dx))   # in place
            if cat_input:
                call("lb_copy_rows", ptr(gout), ctot, ptr(dx), cin, b * h * w_, cin, 1, _dt(gout), _dt(dx))
        if ctx.bias_param is not None and ctx.needs_input_grad[4]:
            dbias, dbias_ret = _grad_sink(ctx.bias_param)
            call("lb_colsum", gout.data_ptr() + off * esz_g, rows, spec.cout, ctot, ptr(dbias), _dt(gout))
            dist.grad_written(ctx.bias_param)
        return dx, dw_ret, None, None, dbias_ret, None, None, None, None


def _tc_bytes(g, *_unused):
    """ALGORITHMIC bytes of one tensor-core GEMM launch: bf16 activations in + out and the bf16
    weight -- whatever extra copies (fp32 outputs, pre-activations for a fused RootTanh') the launch really moves."""
    rows_in = g.batch * g.in_h * g.in_w
    rows_out = g.batch * g.out_h * g.out_w
    return 2.0 * rows_in * g.in_c + 2.0 * g.kh * g.kw * g.in_c * g.out_c + 2.0 * rows_out * g.out_c


def _wgrad_bytes(g):
    """ALGORITHMIC bytes of one weight-gradient launch: both bf16 operands once + the fp32 gradient."""
    return (2.0 * g.batch * g.in_h * g.in_w * g.in_c + 2.0 * g.batch * g.out_h * g.out_w * g.out_c
            + 4.0 * g.kh * g.kw * g.in_c * g.out_c)


EX_AUX_IS_FACTOR, EX_OUT16_IS_DACT = 1, 2      # include/locate_b200.h LB_EX_*


def _ex_ok(g, out32, ld16, ld_aux, aux_dtype=F32):
    return _lib.lib().lb_conv_tc_ex_supported(ctypes.byref(g), int(out32), ld16, ld_aux, aux_dtype) == 1


_WGRAD_WORK = {}


def _wgrad_tc(gathered_ptr, dense_ptr, g_wgrad, w_bar, fl, dev):
    """Raw dW (fp32, master layout) of one conv on the tensor cores and sum dW * W_bar: lb_wgrad_tc with its split-K
    workspace (one buffer per device, grown to the largest layer; every layer overwrites it, stream order keeps the
    uses apart)."""
    from .ops import _stat_work
    need = _lib.lib().lb_wgrad_tc_workspace_floats(ctypes.byref(g_wgrad))
    work = _WGRAD_WORK.get(dev)
    if work is None or work.numel() < need:
        work = _WGRAD_WORK[dev] = torch.empty(max(need, 1 << 22), dtype=torch.float32, device=dev)
    dwn = torch.empty(w_bar.numel(), dtype=torch.float32, device=dev)
    dot = torch.empty(2, dtype=torch.float64, device=dev)
    _timed_call("wgrad_tc", fl, _wgrad_bytes(g_wgrad), "lb_wgrad_tc", gathered_ptr, dense_ptr, ptr(dwn), g_wgrad, ptr(work),
                work.numel(), ptr(w_bar), ptr(dot), ptr(_stat_work(dev)))
    return dwn, dot


def _sn_wgrad_tc(ctx_w, u, v, sigma, gathered_ptr, dense_ptr, g_wgrad, spec, fl, by, dev, extra=None):
    """dW of one spectral-normed conv on the tensor cores + the sigma correction, accumulated into the grad sink."""
    dwn, dot = _wgrad_tc(gathered_ptr, dense_ptr, g_wgrad, ctx_w, fl, dev)
    return _sn_weight_grad(dwn, ctx_w, u, v, sigma, spec, 0, extra, dev, dot)


class ActivatedPairFn(torch.autograd.Function):
    """ActivatedBaseConv.forward (conv.py:22-24) as ONE autograd node on the persistent tensor-core kernel, every
    activation stored as bf16:

        y1 = conv_1(RootTanh(conv_0(RootTanh(x))))         (both convs spectral-normed, no bias)

    forward : conv_0's epilogue writes y0 (bf16, kept for the activation backward) AND RootTanh(y0), the operand of
              conv_1 -- the activation never makes its own pass over memory; conv_1 writes y1 as bf16 (or fp32 rows when
              it has fewer than 8 channels: the generator's RGB output);
    backward: the incoming bf16 gradient is conv_1's operand as it stands; conv_1's input-gradient GEMM multiplies by
              RootTanh'(y0) in its epilogue and emits the bf16 operand of conv_0's two gradient GEMMs; conv_0's
              input-gradient GEMM multiplies by RootTanh'(x) the same way and writes dx as bf16.
    `x` (bf16) may carry `_lb_act16` = RootTanh(x) left by the kernel that produced it (norm apply)."""

    @staticmethod
    def forward(ctx, x, w0, u0, v0, w1, u1, v1, spec0, spec1, sigma0, sigma1, geoms, pre_act0):
        act16 = getattr(x, "_lb_act16", None) if pre_act0 else None
        dact_x = getattr(x, "_lb_dact16", None) if pre_act0 else None
        need_bwd = any(ctx.needs_input_grad)
        x_vals = None
        if not getattr(x, "_lb_unwritten", False):
            x_vals = x = _as_act(x)                  # bf16 (cin % 8 == 0, checked by activated_pair)
        elif act16 is None or (need_bwd and dact_x is None):
            raise RuntimeError("placeholder norm output without its RootTanh / RootTanh' companions")
        b, cin, h, w_ = x.shape
        oh, ow = spec0.out_hw(h, w_)
        mid, cout = spec0.cout, spec1.cout
        (gf0, gd0, gw0), (gf1, gd1, gw1) = geoms
        if not pre_act0:
            act16 = x
        elif act16 is None:
            act16 = torch.empty_like(x)
            call("lb_roottanh_fwd", ptr(x), ptr(act16), x.numel(), CFG.ROOTTANH_GROWTH, BF16)
        a0 = _new_act((b, mid, oh, ow), x, torch.bfloat16)
        # with a backward pass to come conv_0 also stores RootTanh'(y0) (not y0): conv_1's input gradient multiplies by it
        dact0 = torch.empty_like(a0) if need_bwd else None
        fl0, by0 = _conv_work(spec0, b, h, w_, oh, ow)
        fl1, by1 = _conv_work(spec1, b, oh, ow, oh, ow)
        _timed_call("conv_tc", fl0, _tc_bytes(gf0), "lb_conv_tc_gemm_ex", ptr(act16), ptr(_packed_weight(w0, gf0, "fwd")),
                    sigma0.data_ptr() + 4, None, None, ptr(dact0), ptr(a0), mid, None, 0, F32, EX_OUT16_IS_DACT if need_bwd else 0, gf0)
        y1 = _new_act((b, cout, oh, ow), x)
        # lb_conv_tc_gemm_ws picks the persistent kernel itself and keeps direct stores for rows that TMA cannot address
        # (cout = 3: G's last layer, fp32 rows)
        _tc_gemm(fl1, _tc_bytes(gf1), ptr(a0), _packed_weight(w1, gf1, "fwd"), sigma1.data_ptr() + 4, None, ptr(y1), gf1, x.device,
                 _dt(y1))
        # RootTanh'(x): the factor left by the norm kernel, else x itself (the epilogue then evaluates it)
        ctx.x_is_factor = bool(pre_act0 and dact_x is not None)
        ctx.save_for_backward((dact_x if ctx.x_is_factor else x_vals) if pre_act0 else None, act16, dact0, a0, w0, w1, sigma0, sigma1)
        ctx.uv = (u0, v0, u1, v1)                # LIVE u/v (see SNConvFn)
        ctx.uv_extra = (_uv_extra(sigma0, u0), _uv_extra(sigma1, u1))
        ctx.meta = (spec0, spec1, geoms, (b, cin, h, w_, oh, ow, mid, cout), (fl0, by0, fl1, by1), pre_act0)
        ctx.out_like = _Like(y1.shape, y1.stride(), y1.dtype)
        return y1

    @staticmethod
    def backward(ctx, gout):
        x, act16, dact0, a0, w0, w1, sigma0, sigma1 = ctx.saved_tensors
        u0, v0, u1, v1 = ctx.uv
        spec0, spec1, geoms, (b, cin, h, w_, oh, ow, mid, cout), (fl0, by0, fl1, by1), pre_act0 = ctx.meta
        (gf0, gd0, gw0), (gf1, gd1, gw1) = geoms
        gout = _match(gout, ctx.out_like)
        dev = gout.device
        need_dx, need_dw0, need_dw1 = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[4]
        cout_p = (cout + 7) // 8 * 8               # bf16 rows of >= 16 bytes for TMA
        g1 = gout if gout.dtype == torch.bfloat16 else _bf16_rows(gout, cout, cout_p, 0, 0, cout)
        dx = dw0 = dw1 = None
        if need_dw1:
            ga, de = (g1, a0) if spec1.kind == "convT" else (a0, g1)
            dw1 = _sn_wgrad_tc(w1, u1, v1, sigma1, ptr(ga), ptr(de), gw1, spec1, fl1, by1, dev, ctx.uv_extra[1])
        if need_dx or need_dw0:
            d0 = torch.empty_like(a0)            # bf16( dL/dy0 ) = bf16( dgrad_1(g1) * RootTanh'(y0) ), the factor stored by the forward
            _timed_call("conv_tc", fl1, _tc_bytes(gd1), "lb_conv_tc_gemm_ex", ptr(g1), ptr(_packed_weight(w1, gd1, "dgrad")),
                        sigma1.data_ptr() + 4, None, None, ptr(d0), None, mid, ptr(dact0), mid, BF16, EX_AUX_IS_FACTOR, gd1)
            if need_dw0:
                ga, de = (d0, act16) if spec0.kind == "convT" else (act16, d0)
                dw0 = _sn_wgrad_tc(w0, u0, v0, sigma0, ptr(ga), ptr(de), gw0, spec0, fl0, by0, dev, ctx.uv_extra[0])
            if need_dx:
                dx = _new_act((b, cin, h, w_), gout, torch.bfloat16)
                _timed_call("conv_tc", fl0, _tc_bytes(gd0), "lb_conv_tc_gemm_ex", ptr(d0), ptr(_packed_weight(w0, gd0, "dgrad")),
                            sigma0.data_ptr() + 4, None, None, ptr(dx), None, cin, ptr(x) if pre_act0 else None,
                            cin if pre_act0 else 0, BF16, EX_AUX_IS_FACTOR if ctx.x_is_factor else 0, gd0)
        return (dx, dw0, None, None, dw1) + (None,) * 8


def activated_pair(x, sn0, sn1, pre_act0=True):
    """conv_1(RootTanh(conv_0(RootTanh(x)))) for two SpectralNorm wrappers (layers.ActivatedBaseConv); returns None when
    the fused tensor-core path does not cover the layer (odd channel counts, fp32 mode, weight-bound split-K shapes)."""
    if CFG.PRECISION != "bf16" or x.dim() != 4 or CFG.ROOTTANH_GROWTH != 4:
        return None
    spec0, spec1 = sn0.spec, sn1.spec
    m0, m1 = sn0.module, sn1.module
    if getattr(m0, "bias", None) is not None or getattr(m1, "bias", None) is not None:
        return None
    if spec0.groups > 1 or spec1.groups > 1:          # SEPARABLE: depthwise conv_0 runs on its own direct kernel
        return None
    b, cin, h, w_ = x.shape
    mid, cout = spec0.cout, spec1.cout
    if cin % 8 or mid % 8 or cin != spec0.cin or mid != spec1.cin:
        return None
    cout_p = (cout + 7) // 8 * 8
    oh, ow = spec0.out_hw(h, w_)
    lib = _lib.lib()

    def geoms(spec, ih, iw, ci, co, o_h, o_w, ld_co):
        """ld_co: row stride of the bf16 gradient w.r.t. this conv's output (co padded to 8)."""
        mode = 1 if spec.kind == "convT" else 0
        t = spec.taps
        gf = _geom(b, ih, iw, ci, o_h, o_w, co, spec, mode, ci, co, spec.strides_fwd())
        gd = _geom(b, o_h, o_w, co, ih, iw, ci, spec, 1 - mode, ld_co, ci, spec.strides_dgrad())
        if spec.kind == "convT":
            gw = _geom(b, o_h, o_w, co, ih, iw, ci, spec, 0, ld_co, ci, (t, co * t, spec.kw, 1))
        else:
            gw = _geom(b, ih, iw, ci, o_h, o_w, co, spec, 0, ci, ld_co, (t, ci * t, spec.kw, 1))
        return gf, gd, gw

    g0, g1 = geoms(spec0, h, w_, cin, mid, oh, ow, mid), geoms(spec1, oh, ow, mid, cout, oh, ow, cout_p)
    ok = (_ex_ok(g0[0], False, mid, 0) and lib.lb_conv_tc_supported(ctypes.byref(g1[0])) == 1 and _ex_ok(g1[1], False, mid, mid, BF16)
          and _ex_ok(g0[1], False, cin, cin if pre_act0 else 0, BF16)
          and lib.lb_wgrad_tc_supported(ctypes.byref(g0[2])) == 1 and lib.lb_wgrad_tc_supported(ctypes.byref(g1[2])) == 1)
    if not ok:
        return None
    sig = []
    for sn in (sn0, sn1):
        m = sn.module
        for _ in range(sn.power_iterations - 1):
            power_iterate(m.weight_bar, m.weight_u.data, m.weight_v.data, sn.spec)
        pre, sn._pre_sigma = sn._pre_sigma, None
        sig.append(pre if pre is not None else power_iterate(m.weight_bar, m.weight_u.data, m.weight_v.data, sn.spec))
    return ActivatedPairFn.apply(x, m0.weight_bar, m0.weight_u, m0.weight_v, m1.weight_bar, m1.weight_u, m1.weight_v,
                                 spec0, spec1, sig[0], sig[1], (g0, g1), pre_act0)


class GroupedSNConvFn(torch.autograd.Function):
    """The grouped spectral-normed convolutions of SEPARABLE = True (config.py:53) on direct HBM-bound kernels:
    depthwise k x k Conv2d / ConvTranspose2d (conv.py:17, groups = channels) optionally preceded by RootTanh
    (conv.py:23), and feature attention's full-extent grouped conv (attention.py:15-21, kernel = the whole map)."""

    @staticmethod
    def forward(ctx, x, w_bar, u, v, spec, pre_act, pre_sigma):
        x = _as_act(x)
        b, c, h, w_ = x.shape
        if c != spec.cin:
            raise ValueError(f"expected {spec.cin} input channels, got {c}")
        sigma = pre_sigma if pre_sigma is not None else power_iterate(w_bar, u.data, v.data, spec)
        dt = _dt(x)
        a = x
        if pre_act:
            a = torch.empty_like(x)
            call("lb_roottanh_fwd", ptr(x), ptr(a), x.numel(), CFG.ROOTTANH_GROWTH, dt)
        full = spec.groups != spec.cin                    # full-extent grouped conv: [B,F,S,S] -> [B,F/r,1,1]
        if full:
            if (spec.kh, spec.kw) != (h, w_) or spec.stride != 1 or spec.pad != 0 or spec.kind != "conv" or pre_act:
                raise NotImplementedError("grouped conv other than depthwise / full-extent (attention.py:15-21)")
            r = spec.cin // spec.groups
            if spec.cout != spec.groups:
                raise NotImplementedError("full-extent grouped conv with more than one output per group")
            out = _new_act((b, spec.cout, 1, 1), x, x.dtype)
            call("lb_gfull_fwd", ptr(a), ptr(w_bar), sigma.data_ptr() + 4, ptr(out), b, h * w_, c, r, dt)
            out = _cast(out, store_dtype(out.shape))      # F/r may fall off the bf16 storage rule (C % 8)
        else:
            if spec.cout != spec.cin:
                raise NotImplementedError("depthwise conv with a channel multiplier (FEATURE_MULTIPLIER != 1)")
            oh, ow = spec.out_hw(h, w_)
            out = _new_act((b, c, oh, ow), x, x.dtype)
            call("lb_dw_conv", ptr(a), ptr(w_bar), sigma.data_ptr() + 4, ptr(out), b, h, w_, oh, ow, c, spec.kh, spec.kw, spec.stride,
                 spec.pad, 1 if spec.kind == "convT" else 0, dt)
        ctx.save_for_backward(x if pre_act else None, a, w_bar, sigma)
        ctx.u, ctx.v = u, v
        ctx.uv_extra = _uv_extra(pre_sigma, u)
        ctx.w_param = w_bar
        ctx.meta = (spec, pre_act, full, tuple(x.shape), _Like(out.shape, out.stride(), out.dtype))
        return out

    @staticmethod
    def backward(ctx, gout):
        x, a, w_bar, sigma = ctx.saved_tensors
        spec, pre_act, full, (b, c, h, w_), out_like = ctx.meta
        gout = _match(gout, out_like)
        if gout.dtype != a.dtype:
            gout = gout.to(a.dtype)
        dt = _dt(gout)
        dev = gout.device
        dx = dw_ret = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(a)
            if full:
                call("lb_gfull_dgrad", ptr(gout), ptr(w_bar), sigma.data_ptr() + 4, ptr(dx), b, h * w_, c, spec.cin // spec.groups, dt)
            else:
                oh, ow = out_like.shape[2], out_like.shape[3]
                call("lb_dw_conv", ptr(gout), ptr(w_bar), sigma.data_ptr() + 4, ptr(dx), b, oh, ow, h, w_, c, spec.kh, spec.kw,
                     spec.stride, spec.pad, 0 if spec.kind == "convT" else 1, dt)
            if pre_act:
                call("lb_roottanh_bwd", ptr(x), ptr(dx), ptr(dx), dx.numel(), CFG.ROOTTANH_GROWTH, dt)
        if ctx.needs_input_grad[1]:
            dwn = torch.zeros_like(w_bar, memory_format=torch.contiguous_format)
            if full:
                call("lb_gfull_wgrad", ptr(a), ptr(gout), ptr(dwn), b, h * w_, c, spec.cin // spec.groups, dt)
            elif spec.kind == "convT":    # gathered = dy (the larger map), dense = x
                oh, ow = out_like.shape[2], out_like.shape[3]
                call("lb_dw_wgrad", ptr(gout), ptr(a), ptr(dwn), b, oh, ow, h, w_, c, spec.kh, spec.kw, spec.stride, spec.pad, dt)
            else:
                oh, ow = out_like.shape[2], out_like.shape[3]
                call("lb_dw_wgrad", ptr(a), ptr(gout), ptr(dwn), b, h, w_, oh, ow, c, spec.kh, spec.kw, spec.stride, spec.pad, dt)
            dw_ret = _sn_weight_grad(dwn, ctx.w_param, ctx.u, ctx.v, sigma, spec, 0, ctx.uv_extra, dev)
        return dx, dw_ret, None, None, None, None, None


def sn_conv(x, w_bar, u, v, bias, spec, cat_input=False, pre_act=False, pre_sigma=None):
    if spec.groups > 1:
        if bias is not None or cat_input:
            raise NotImplementedError("grouped convolutions carry no bias / concat on the reference's path")
        return GroupedSNConvFn.apply(x, w_bar, u, v, spec, pre_act, pre_sigma)
    return SNConvFn.apply(x, w_bar, u, v, bias, spec, cat_input, pre_act, pre_sigma)
