"""Direct kernels for tiny-channel layers (lb_conv_small / lb_conv_small_wgrad) against the fp32 SIMT gather-GEMM
(lb_conv_gemm / lb_conv_wgrad) with the fused RootTanh / RootTanh' applied by torch on the reference side.  The narrow
side (<= 4 channels) is always fp32; the wide side is stored as fp32 or bf16 (`wide`): the bf16 cases feed both kernels
the same bf16-rounded operand, so only the output rounding differs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

from locate_b200 import _lib                     # noqa: E402
from locate_b200._lib import ConvGeom, call, ptr  # noqa: E402

DEV = "cuda:0"


def geom(b, ih, iw, ic, oh, ow, oc, kh, kw, s, p, mode, ld_in, ld_out, strides):
    g = ConvGeom()
    g.batch, g.in_h, g.in_w, g.in_c, g.out_h, g.out_w, g.out_c = b, ih, iw, ic, oh, ow, oc
    g.kh, g.kw, g.stride, g.pad, g.mode, g.ld_in, g.ld_out = kh, kw, s, p, mode, ld_in, ld_out
    g.w_sk, g.w_sn, g.w_sty, g.w_stx = strides
    return g


def roottanh(x):
    return (x * x + 1) ** 0.25 * torch.tanh(x)


def roottanh_grad(x):
    q = x * x + 1
    th = torch.tanh(x)
    return (2 * q * (1 - th * th) + x * th) * q ** 0.25 / (2 * q)


CASES = {
    # name: (kind, b, h, w, cin, cout, k, s, p, direction)
    "stem_1x1_3_29": ("conv", 3, 16, 16, 3, 29, 1, 1, 0, "fwd"),
    "stem_5x5s2": ("conv", 3, 16, 16, 3, 3, 5, 2, 2, "fwd"),
    "stem_5x5s2_odd": ("conv", 2, 13, 11, 3, 3, 5, 2, 2, "fwd"),
    "stem_1x1_3_32": ("conv", 3, 8, 8, 3, 32, 1, 1, 0, "fwd"),
    "last_1x1_48_3": ("conv", 3, 16, 16, 48, 3, 1, 1, 0, "fwd"),
    "stem_5x5s2_dgrad": ("conv", 3, 16, 16, 3, 3, 5, 2, 2, "dgrad"),
    "stem_1x1_3_32_dgrad": ("conv", 3, 8, 8, 3, 32, 1, 1, 0, "dgrad"),
    "last_1x1_48_3_dgrad": ("conv", 3, 16, 16, 48, 3, 1, 1, 0, "dgrad"),
    "convT_4x4s2_3": ("convT", 2, 8, 8, 3, 3, 4, 2, 1, "fwd"),
    "convT_4x4s2_3_dgrad": ("convT", 2, 8, 8, 3, 3, 4, 2, 1, "dgrad"),
}


F32, BF16 = 0, 1


@pytest.mark.parametrize("wide", [F32, BF16])
@pytest.mark.parametrize("fuse", ["plain", "act_in", "dact_out", "factor_out"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_small_matches_simt(name, fuse, wide):
    kind, b, h, w, cin, cout, k, s, p, direction = CASES[name]
    gen = torch.Generator().manual_seed(hash(name) % 1000)
    t = k * k
    if kind == "conv":
        wt = torch.randn((cout, cin, k, k), generator=gen)
        oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        fwd, dgr, mode_f, mode_d = (t, cin * t, k, 1), (cin * t, t, k, 1), 0, 1
    else:
        wt = torch.randn((cin, cout, k, k), generator=gen)
        oh, ow = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
        fwd, dgr, mode_f, mode_d = (cout * t, t, k, 1), (t, cout * t, k, 1), 1, 0
    wt = wt.to(DEV)
    pad_out = 5                              # write into a channel slice of a wider tensor
    if direction == "fwd":
        src = torch.randn((b, h, w, cin), generator=gen).to(DEV)
        g = geom(b, h, w, cin, oh, ow, cout, k, k, s, p, mode_f, cin, cout + pad_out, fwd)
        out_shape = (b, oh, ow, cout + pad_out)
    else:
        src = torch.randn((b, oh, ow, cout), generator=gen).to(DEV)
        g = geom(b, oh, ow, cout, h, w, cin, k, k, s, p, mode_d, cout, cin + pad_out, dgr)
        out_shape = (b, h, w, cin + pad_out)
    n = g.out_c
    assert _lib.lib().lb_conv_small_supported(ctypes.byref(g)) == 1
    in_wide, out_wide = g.in_c > 4, g.in_c <= 4            # which side carries the `wide` storage type
    if wide == BF16 and g.in_c <= 4 and g.out_c <= 4:
        pytest.skip("both sides narrow: an all-fp32 layer")
    src_k = src.bfloat16() if (wide == BF16 and in_wide) else src
    src = src_k.float()
    alpha = torch.tensor([0.37], device=DEV)
    bias = torch.randn(n, generator=gen).to(DEV)
    ref = torch.full(out_shape, -7.0, device=DEV)
    got = torch.full(out_shape, -7.0, device=DEV, dtype=torch.bfloat16 if (wide == BF16 and out_wide) else torch.float32)
    src_ref = roottanh(src) if fuse == "act_in" else src
    call("lb_conv_gemm", ptr(src_ref.contiguous()), ptr(wt), ptr(alpha), ptr(bias), ptr(ref), ctypes.byref(g))
    xpre = None
    if fuse == "dact_out":
        xpre = 2.0 * torch.randn(out_shape[:3] + (n,), generator=gen).to(DEV)
        if wide == BF16 and out_wide:
            xpre = xpre.bfloat16()
        ref[..., :n] *= roottanh_grad(xpre.float())
    if fuse == "factor_out":                 # xpre holds the derivative itself (stored by the forward pass): growth_out = -1
        xpre = torch.randn(out_shape[:3] + (n,), generator=gen).to(DEV)
        if wide == BF16 and out_wide:
            xpre = xpre.bfloat16()
        ref[..., :n] *= xpre.float()
    call("lb_conv_small", ptr(src_k), ptr(wt), ptr(alpha), ptr(bias), ptr(got), ctypes.byref(g), 4 if fuse == "act_in" else 0,
         ptr(xpre), n, {"dact_out": 4, "factor_out": -1}.get(fuse, 0), 0, wide)
    torch.cuda.synchronize()
    got = got.float()
    assert torch.equal(got[..., n:], ref[..., n:]), "wrote outside its channel slice"
    err = (got - ref)[..., :n].abs().max().item()
    scale = ref[..., :n].abs().max().item()
    tol = 8e-3 if got.dtype != ref.dtype or (wide == BF16 and out_wide) else 1e-4
    assert err <= tol * scale + 1e-5, f"{name}/{fuse}: max err {err:.3e} vs scale {scale:.3e}"


WG_CASES = {
    # name: (kind, b, h, w, cin, cout, k, s, p)   h,w = layer INPUT size
    "wg_stem_1x1_3_29": ("conv", 3, 16, 16, 3, 29, 1, 1, 0),
    "wg_stem_5x5s2": ("conv", 3, 16, 16, 3, 3, 5, 2, 2),
    "wg_stem_5x5s2_odd": ("conv", 2, 13, 11, 3, 3, 5, 2, 2),
    "wg_stem_1x1_3_32": ("conv", 5, 8, 8, 3, 32, 1, 1, 0),
    "wg_last_1x1_48_3": ("conv", 3, 16, 16, 48, 3, 1, 1, 0),
    "wg_convT_4x4s2_3": ("convT", 2, 8, 8, 3, 3, 4, 2, 1),
}


@pytest.mark.parametrize("wide", [F32, BF16])
@pytest.mark.parametrize("act", [0, 4])
@pytest.mark.parametrize("name", sorted(WG_CASES))
def test_small_wgrad_matches_simt(name, act, wide):
    kind, b, h, w, cin, cout, k, s, p = WG_CASES[name]
    gen = torch.Generator().manual_seed(hash(name) % 1000)
    t = k * k
    if kind == "conv":
        oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        x = torch.randn((b, h, w, cin), generator=gen).to(DEV)
        dy = torch.randn((b, oh, ow, cout), generator=gen).to(DEV)
        gathered, dense = x, dy
        g = geom(b, h, w, cin, oh, ow, cout, k, k, s, p, 0, cin, cout, (t, cin * t, k, 1))
        shape = (cout, cin, k, k)
    else:
        oh, ow = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
        x = torch.randn((b, h, w, cin), generator=gen).to(DEV)
        dy = torch.randn((b, oh, ow, cout), generator=gen).to(DEV)
        gathered, dense = dy, x
        g = geom(b, oh, ow, cout, h, w, cin, k, k, s, p, 0, cout, cin, (t, cout * t, k, 1))
        shape = (cin, cout, k, k)
    assert _lib.lib().lb_conv_small_wgrad_supported(ctypes.byref(g)) == 1
    dense_wide = k == 1 and s == 1 and p == 0 and g.in_c <= 4     # else the gathered operand is the wide one
    gath_k, dense_k = gathered, dense
    if wide == BF16:
        if g.in_c <= 4 and g.out_c <= 4:
            pytest.skip("both operands narrow: all fp32")
        if dense_wide:
            dense_k = dense.bfloat16(); dense = dense_k.float()
        else:
            gath_k = gathered.bfloat16(); gathered = gath_k.float()
    ref = torch.zeros(shape, device=DEV)
    got = torch.zeros(shape, device=DEV)
    gath_ref = (roottanh(gathered) if act else gathered).contiguous()
    call("lb_conv_wgrad", ptr(gath_ref), ptr(dense), ptr(ref), ctypes.byref(g))
    call("lb_conv_small_wgrad", ptr(gath_k), ptr(dense_k), ptr(got), ctypes.byref(g), act, wide)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-4 * scale + 1e-5, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("b,h,w,cin,cout", [(3, 16, 16, 3, 29), (5, 9, 7, 3, 29), (2, 16, 16, 4, 12)])
def test_small_fused_cat(b, h, w, cin, cout):
    """CatModule(identity, conv 1x1) (merge.py:10-16, scale.py:31-34): rows [x | conv(x)] written in one pass."""
    gen = torch.Generator().manual_seed(11)
    wt = torch.randn((cout, cin, 1, 1), generator=gen).to(DEV)
    x = torch.randn((b, h, w, cin), generator=gen).to(DEV)
    bias = torch.randn(cout, generator=gen).to(DEV)
    alpha = torch.tensor([0.6], device=DEV)
    ctot = cin + cout
    g = geom(b, h, w, cin, h, w, cout, 1, 1, 1, 0, 0, cin, ctot, (1, cin, 1, 1))
    ref = torch.full((b, h, w, ctot), -7.0, device=DEV)
    got = torch.full((b, h, w, ctot), -7.0, device=DEV)
    call("lb_conv_gemm", ptr(x), ptr(wt), ptr(alpha), ptr(bias), ref.data_ptr() + 4 * cin, ctypes.byref(g))
    ref[..., :cin] = x
    call("lb_conv_small", ptr(x), ptr(wt), ptr(alpha), ptr(bias), ptr(got), ctypes.byref(g), 0, None, 0, 0, 1, F32)
    torch.cuda.synchronize()
    assert torch.equal(got[..., :cin], x)
    err = (got - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-5
    # the same rows stored as bf16 (the tensor-core configuration: the concat output is a wide, bf16 tensor)
    got16 = torch.full((b, h, w, ctot), -7.0, device=DEV, dtype=torch.bfloat16)
    call("lb_conv_small", ptr(x), ptr(wt), ptr(alpha), ptr(bias), ptr(got16), ctypes.byref(g), 0, None, 0, 0, 1, BF16)
    torch.cuda.synchronize()
    assert (got16.float() - ref).abs().max().item() <= 8e-3 * ref.abs().max().item() + 1e-5


def test_small_wgrad_concat_slice():
    """Weight gradient of the D stem's concat conv: the dense operand is the channel slice [3:32] of a 32-wide gradient
    (rows 16-byte aligned, slice start not) -- the aligned-row float4 path with a column shift."""
    gen = torch.Generator().manual_seed(21)
    b, h, w, cin, cout = 5, 16, 16, 3, 29
    x = torch.randn((b, h, w, cin), generator=gen).to(DEV)
    gfull = torch.randn((b, h, w, cin + cout), generator=gen).to(DEV)
    g = geom(b, h, w, cin, h, w, cout, 1, 1, 1, 0, 0, cin, cin + cout, (1, cin, 1, 1))
    ref = torch.zeros((cout, cin, 1, 1), device=DEV)
    got = torch.zeros((cout, cin, 1, 1), device=DEV)
    dense_ptr = gfull.data_ptr() + 4 * cin
    call("lb_conv_wgrad", ptr(x), dense_ptr, ptr(ref), ctypes.byref(g))
    call("lb_conv_small_wgrad", ptr(x), dense_ptr, ptr(got), ctypes.byref(g), 0, F32)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    assert err <= 2e-4 * ref.abs().max().item() + 1e-5
    # bf16 gradient rows: the slice starts 6 bytes into 64-byte rows (4-element loads with a 3-element shift)
    g16 = gfull.bfloat16()
    ref16 = torch.zeros((cout, cin, 1, 1), device=DEV)
    got16 = torch.zeros((cout, cin, 1, 1), device=DEV)
    gf = g16.float()
    call("lb_conv_wgrad", ptr(x), gf.data_ptr() + 4 * cin, ptr(ref16), ctypes.byref(g))
    call("lb_conv_small_wgrad", ptr(x), g16.data_ptr() + 2 * cin, ptr(got16), ctypes.byref(g), 0, BF16)
    torch.cuda.synchronize()
    assert (got16 - ref16).abs().max().item() <= 2e-4 * ref16.abs().max().item() + 1e-5
