"""tcgen05 / TMA implicit-GEMM (lb_conv_tc_gemm) against the fp32 SIMT gather-GEMM (lb_conv_gemm) on the SAME
bf16-rounded operands: any difference beyond fp32 accumulation order is a layout / descriptor bug."""
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


CASES = {
    # name: (kind, b, h, w, cin, cout, k, s, p, direction)
    "1x1_64": ("conv", 2, 16, 16, 64, 64, 1, 1, 0, "fwd"),
    "1x1_96_48": ("conv", 3, 16, 16, 96, 48, 1, 1, 0, "fwd"),
    "1x1_wide": ("conv", 5, 4, 4, 1536, 768, 1, 1, 0, "fwd"),
    "3x3_48": ("conv", 2, 16, 16, 48, 48, 3, 1, 1, "fwd"),
    "3x3_odd": ("conv", 2, 9, 9, 40, 24, 3, 1, 1, "fwd"),
    "5x5s2": ("conv", 2, 32, 32, 64, 64, 5, 2, 2, "fwd"),
    "5x5s2_c32": ("conv", 3, 64, 64, 32, 32, 5, 2, 2, "fwd"),
    "5x5s2_dgrad": ("conv", 2, 32, 32, 64, 64, 5, 2, 2, "dgrad"),
    "3x3_dgrad": ("conv", 2, 16, 16, 48, 48, 3, 1, 1, "dgrad"),
    "convT4": ("convT", 2, 8, 8, 96, 96, 4, 2, 1, "fwd"),
    "convT4_big": ("convT", 4, 4, 4, 1536, 1536, 4, 2, 1, "fwd"),
    "convT4_b2x2": ("convT", 40, 2, 2, 128, 128, 4, 2, 1, "fwd"),
    "convT4_dgrad": ("convT", 2, 8, 8, 96, 96, 4, 2, 1, "dgrad"),
    "convT1x1": ("convT", 3, 8, 8, 192, 96, 1, 1, 0, "fwd"),
    "5x5s2_tiny": ("conv", 6, 2, 2, 1024, 1024, 5, 2, 2, "fwd"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_tc_matches_simt(name):
    kind, b, h, w, cin, cout, k, s, p, direction = CASES[name]
    gen = torch.Generator().manual_seed(hash(name) % 1000)
    t = k * k
    if kind == "conv":
        wt = torch.randn((cout, cin, k, k), generator=gen)
        oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        fwd = (t, cin * t, k, 1)          # k=cin, n=cout
        dgr = (cin * t, t, k, 1)          # k=cout, n=cin
        mode_f, mode_d = 0, 1
    else:
        wt = torch.randn((cin, cout, k, k), generator=gen)
        oh, ow = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
        fwd = (cout * t, t, k, 1)
        dgr = (t, cout * t, k, 1)
        mode_f, mode_d = 1, 0
    wt = (wt / (cin * t) ** 0.5).bfloat16().float().to(DEV)
    pad_out = 8                              # write into a channel slice of a wider tensor
    if direction == "fwd":
        src = torch.randn((b, h, w, cin), generator=gen).bfloat16().to(DEV)
        g = geom(b, h, w, cin, oh, ow, cout, k, k, s, p, mode_f, cin, cout + pad_out, fwd)
        out_shape = (b, oh, ow, cout + pad_out)
    else:
        src = torch.randn((b, oh, ow, cout), generator=gen).bfloat16().to(DEV)
        g = geom(b, oh, ow, cout, h, w, cin, k, k, s, p, mode_d, cout, cin + pad_out, dgr)
        out_shape = (b, h, w, cin + pad_out)
    alpha = torch.tensor([0.37], device=DEV)
    bias = torch.randn(g.out_c, generator=gen).to(DEV)
    ref = torch.full(out_shape, -7.0, device=DEV)
    got = torch.full(out_shape, -7.0, device=DEV)
    srcf = src.float().contiguous()
    call("lb_conv_gemm", ptr(srcf), ptr(wt), ptr(alpha), ptr(bias), ptr(ref), ctypes.byref(g))
    assert _lib.lib().lb_conv_tc_supported(ctypes.byref(g)) == 1
    packed = torch.empty(_lib.lib().lb_conv_tc_packed_elems(ctypes.byref(g)), dtype=torch.bfloat16, device=DEV)
    call("lb_conv_tc_pack", ptr(wt), ptr(packed), ctypes.byref(g))
    call("lb_conv_tc_gemm", ptr(src), ptr(packed), ptr(alpha), ptr(bias), ptr(got), ctypes.byref(g))
    torch.cuda.synchronize()
    assert torch.equal(got[..., g.out_c:], ref[..., g.out_c:]), "wrote outside its channel slice"
    err = (got - ref).abs().max().item()
    scale = ref[..., :g.out_c].abs().max().item()
    assert err <= 2e-4 * scale + 1e-5, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


WG_CASES = {
    # name: (kind, b, h, w, cin, cout, k, s, p)   h,w = layer INPUT size
    "wg_1x1": ("conv", 2, 16, 16, 64, 64, 1, 1, 0),
    "wg_1x1_96_48": ("conv", 3, 16, 16, 96, 48, 1, 1, 0),
    "wg_3x3": ("conv", 2, 16, 16, 48, 48, 3, 1, 1),
    "wg_5x5s2": ("conv", 2, 32, 32, 64, 64, 5, 2, 2),
    "wg_5x5s2_c32": ("conv", 3, 64, 64, 32, 32, 5, 2, 2),
    "wg_convT4": ("convT", 2, 8, 8, 96, 96, 4, 2, 1),
    "wg_convT4_wide": ("convT", 4, 4, 4, 384, 384, 4, 2, 1),
    "wg_convT1x1": ("convT", 3, 8, 8, 192, 96, 1, 1, 0),
    "wg_3x3_odd": ("conv", 2, 9, 9, 40, 24, 3, 1, 1),
    "wg_5x5s2_tiny": ("conv", 6, 2, 2, 256, 320, 5, 2, 2),
    # halo-tile main loop (dense map >= 8 wide, several taps): tap groups per parity view, ragged tiles, both channel boxes
    "wg_halo_3x3_b5": ("conv", 5, 24, 40, 48, 48, 3, 1, 1),
    "wg_halo_3x3_wide": ("conv", 2, 16, 16, 160, 272, 3, 1, 1),
    "wg_halo_5x5s2_odd": ("conv", 3, 21, 19, 40, 72, 5, 2, 2),
    "wg_halo_5x5s1": ("conv", 2, 12, 12, 16, 32, 5, 1, 2),
    "wg_halo_convT4_192": ("convT", 3, 16, 16, 192, 192, 4, 2, 1),
    "wg_halo_convT4_b9": ("convT", 9, 8, 8, 96, 136, 4, 2, 1),
    "wg_halo_shallow": ("conv", 7, 2, 16, 24, 24, 3, 1, 1),
}


@pytest.mark.parametrize("name", sorted(WG_CASES))
def test_tc_wgrad_matches_simt(name):
    kind, b, h, w, cin, cout, k, s, p = WG_CASES[name]
    gen = torch.Generator().manual_seed(hash(name) % 1000)
    t = k * k
    if kind == "conv":
        oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        x = torch.randn((b, h, w, cin), generator=gen).bfloat16().to(DEV)
        dy = torch.randn((b, oh, ow, cout), generator=gen).bfloat16().to(DEV)
        gathered, dense = x, dy
        g = geom(b, h, w, cin, oh, ow, cout, k, k, s, p, 0, cin, cout, (t, cin * t, k, 1))
        shape, d0, d1 = (cout, cin, k, k), cout, cin
    else:
        oh, ow = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
        x = torch.randn((b, h, w, cin), generator=gen).bfloat16().to(DEV)
        dy = torch.randn((b, oh, ow, cout), generator=gen).bfloat16().to(DEV)
        gathered, dense = dy, x
        g = geom(b, oh, ow, cout, h, w, cin, k, k, s, p, 0, cout, cin, (t, cout * t, k, 1))
        shape, d0, d1 = (cin, cout, k, k), cin, cout
    ref = torch.zeros(shape, device=DEV)
    gath_f, dense_f = gathered.float().contiguous(), dense.float().contiguous()      # keep alive across the async launch
    call("lb_conv_wgrad", ptr(gath_f), ptr(dense_f), ptr(ref), ctypes.byref(g))
    assert _lib.lib().lb_wgrad_tc_supported(ctypes.byref(g)) == 1
    need = _lib.lib().lb_wgrad_tc_workspace_floats(ctypes.byref(g))
    assert need >= ref.numel() and need % ref.numel() == 0
    got = torch.full(shape, float("nan"), device=DEV)                 # overwritten, no zero-fill contract
    work = torch.full((need + 16,), float("nan"), device=DEV)
    wbar = torch.randn(shape, generator=gen).to(DEV)
    dot = torch.zeros(2, dtype=torch.float64, device=DEV)
    stat_work = torch.zeros(_lib.lib().lb_stat_work_doubles(), dtype=torch.float64, device=DEV)
    call("lb_wgrad_tc", ptr(gathered), ptr(dense), ptr(got), ctypes.byref(g), ptr(work), need, ptr(wbar), ptr(dot), ptr(stat_work))
    torch.cuda.synchronize()
    want_dot = (ref.double() * wbar.double()).sum().item()
    assert abs(dot[0].item() - want_dot) <= 3e-4 * (ref.double() * wbar.double()).abs().sum().item() + 1e-6
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 3e-4 * scale + 1e-5, f"{name}: max err {err:.3e} vs scale {scale:.3e}"
    assert torch.isnan(work[need:]).all()                             # stays inside the workspace it asked for
    # ordered split-K: bit-identical on a second run
    again = torch.empty(shape, device=DEV)
    dot2 = torch.zeros(2, dtype=torch.float64, device=DEV)
    call("lb_wgrad_tc", ptr(gathered), ptr(dense), ptr(again), ctypes.byref(g), ptr(work), need, ptr(wbar), ptr(dot2), ptr(stat_work))
    assert torch.equal(got, again) and dot[0].item() == dot2[0].item()
    call("lb_wgrad_tc", ptr(gathered), ptr(dense), ptr(again), ctypes.byref(g), ptr(work), need, None, None, None)
    assert torch.equal(got, again)
    dwp = got.reshape(d0, d1, t).permute(2, 0, 1).contiguous()        # the tap-major form lb_sn_weight_grad also accepts
    # the packed read path of the spectral-norm epilogue, and the "dot already computed" path
    u = torch.randn(d0, generator=gen).to(DEV)
    v = torch.randn(d1 * t, generator=gen).to(DEV)
    sigma = torch.tensor([2.0, 0.5], device=DEV)
    ga, gb = torch.zeros(shape, device=DEV), torch.zeros(shape, device=DEV)
    work = torch.empty(2, dtype=torch.float64, device=DEV)
    stat_work = torch.zeros(_lib.lib().lb_stat_work_doubles(), dtype=torch.float64, device=DEV)
    call("lb_sn_weight_grad", ptr(ref), ptr(wbar), ptr(u), ptr(v), ptr(sigma), ptr(ga), d0, d1 * t, 0, ptr(work), ptr(stat_work),
         None, None, None)
    call("lb_sn_weight_grad", ptr(dwp), ptr(wbar), ptr(u), ptr(v), ptr(sigma), ptr(gb), d0, d1 * t, t, ptr(work), ptr(stat_work),
         None, None, None)
    torch.cuda.synchronize()
    assert (ga - gb).abs().max().item() <= 3e-4 * ga.abs().max().item() + 1e-5
    gc = torch.zeros(shape, device=DEV)
    call("lb_sn_weight_grad", ptr(got), None, ptr(u), ptr(v), ptr(sigma), ptr(gc), d0, d1 * t, 0, ptr(dot), ptr(stat_work),
         None, None, None)
    torch.cuda.synchronize()
    assert (ga - gc).abs().max().item() <= 3e-4 * ga.abs().max().item() + 1e-5


# ---- persistent kernel with the fused epilogue (lb_conv_tc_gemm_ex) against lb_conv_tc_gemm + torch elementwise ----
def _roottanh(x):
    return (x * x + 1) ** 0.25 * torch.tanh(x)


def _roottanh_grad(x):
    q = x * x + 1
    th = torch.tanh(x)
    return (2 * q * (1 - th * th) + x * th) * q ** 0.25 / (2 * q)


EX_CASES = dict(CASES)
EX_CASES.update({
    "convT4_c192": ("convT", 5, 16, 16, 192, 192, 4, 2, 1, "fwd"),       # block_n = 192, many tiles per CTA
    "convT4_many": ("convT", 40, 32, 32, 96, 96, 4, 2, 1, "fwd"),        # > 148 tiles: persistent loop + TMEM ping-pong
    "1x1_many": ("conv", 24, 64, 64, 96, 48, 1, 1, 0, "fwd"),
    "1x1_c384": ("conv", 9, 16, 16, 384, 384, 1, 1, 0, "fwd"),           # two channel tiles
    "3x3_edge": ("conv", 3, 20, 12, 48, 40, 3, 1, 1, "fwd"),             # tiles overhang every dimension
    "linear": ("conv", 37, 1, 1, 256, 1536, 1, 1, 0, "fwd"),
    "1x1_c352": ("conv", 4, 2, 2, 32, 352, 1, 1, 0, "fwd"),              # channel tile not a multiple of the 32-column slab
    "1x1_c416": ("convT", 4, 1, 1, 192, 416, 1, 1, 0, "fwd"),
    "featattn_dgrad": ("conv", 6, 16, 16, 64, 16, 16, 1, 0, "dgrad"),     # full-extent kernel: 2 live taps of 16 per tile
    # halo mode (one shared-memory tile per source view serves every tap shift; 8 x 16 pixel tiles)
    "halo_convT4_64": ("convT", 3, 64, 64, 96, 96, 4, 2, 1, "fwd"),       # 4 output phases x 4 taps, two channel chunks
    "halo_3x3_c48": ("conv", 4, 64, 48, 48, 48, 3, 1, 1, "fwd"),          # 9 taps of one view, partial channel chunk
    "halo_3x3_dgrad": ("conv", 12, 40, 24, 48, 48, 3, 1, 1, "dgrad"),     # mode 1 stride 1, tiles overhang both axes
    "halo_5x5s2_c32": ("conv", 10, 64, 64, 32, 32, 5, 2, 2, "fwd"),       # 4 parity views: 9 / 6 / 6 / 4 taps
    "halo_5x5s2_dgrad": ("conv", 3, 64, 64, 32, 32, 5, 2, 2, "dgrad"),    # mode 1 stride 2: 4 phases, up to 9 taps each
    "halo_convT4_dgrad": ("convT", 40, 16, 16, 192, 192, 4, 2, 1, "dgrad"),  # mode 0 stride 2, 4 taps per view, 3 chunks
    "halo_odd": ("conv", 26, 21, 19, 40, 24, 3, 1, 1, "fwd"),
})
EX_MODES = ["o32", "o16", "o16act", "o16both", "o16dact", "both_aux", "o16_aux", "o16_aux16", "o16_auxfac16"]


@pytest.mark.parametrize("mode", EX_MODES)
@pytest.mark.parametrize("name", sorted(EX_CASES))
def test_tc_ex_matches_v1(name, mode):
    kind, b, h, w, cin, cout, k, s, p, direction = EX_CASES[name]
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
    wt = (wt / (cin * t) ** 0.5).bfloat16().float().to(DEV)
    pad_out = 8
    if direction == "fwd":
        src = torch.randn((b, h, w, cin), generator=gen).bfloat16().to(DEV)
        g = geom(b, h, w, cin, oh, ow, cout, k, k, s, p, mode_f, cin, cout + pad_out, fwd)
        out_shape = (b, oh, ow, cout + pad_out)
    else:
        src = torch.randn((b, oh, ow, cout), generator=gen).bfloat16().to(DEV)
        g = geom(b, oh, ow, cout, h, w, cin, k, k, s, p, mode_d, cout, cin + pad_out, dgr)
        out_shape = (b, h, w, cin + pad_out)
    n = g.out_c
    alpha = torch.tensor([1.7], device=DEV)
    bias = torch.randn(n, generator=gen).to(DEV)
    ref = torch.full(out_shape, -7.0, device=DEV)
    packed = torch.empty(_lib.lib().lb_conv_tc_packed_elems(ctypes.byref(g)), dtype=torch.bfloat16, device=DEV)
    call("lb_conv_tc_pack", ptr(wt), ptr(packed), ctypes.byref(g))
    # reference: the fp32 SIMT gather-GEMM on the same bf16-rounded operands (an independent kernel: lb_conv_tc_gemm would
    # route halo-eligible shapes through the persistent kernel under test)
    call("lb_conv_gemm", ptr(src.float().contiguous()), ptr(wt), ptr(alpha), ptr(bias), ptr(ref), ctypes.byref(g))
    ref = ref[..., :n]

    want32 = mode in ("o32", "both_aux")
    want16 = mode in ("o16", "o16both", "o16dact", "both_aux", "o16_aux", "o16_aux16", "o16_auxfac16")   # bf16 of the accumulator
    want16a = mode in ("o16act", "o16both", "o16dact")                             # bf16 of RootTanh(accumulator)
    dact = mode == "o16dact"                                                       # out16 holds RootTanh'(accumulator) instead
    use_aux = "aux" in mode
    aux16 = mode.endswith("16") and use_aux
    auxfac = "auxfac" in mode                                                      # aux is the factor itself
    flags = (1 if auxfac else 0) | (2 if dact else 0)
    ld16 = (n + 7) // 8 * 8 + 8
    ld_aux = ((n + 7) // 8 * 8 + 8) if aux16 else ((n + 3) // 4 * 4 + 4)
    got32 = torch.full(out_shape, -7.0, device=DEV) if want32 else None
    got16 = torch.full(out_shape[:3] + (ld16,), -7.0, device=DEV, dtype=torch.bfloat16) if want16 else None
    got16a = torch.full(out_shape[:3] + (ld16,), -7.0, device=DEV, dtype=torch.bfloat16) if want16a else None
    aux = None
    if use_aux:
        aux = torch.full(out_shape[:3] + (ld_aux,), float("nan"), device=DEV)
        aux[..., :n] = 2.0 * torch.randn(out_shape[:3] + (n,), generator=gen).to(DEV)
        if aux16:
            aux = aux.bfloat16()
    any16 = want16 or want16a
    if _lib.lib().lb_conv_tc_ex_supported(ctypes.byref(g), int(want32), ld16 if any16 else 0, ld_aux if use_aux else 0,
                                          1 if aux16 else 0) != 1:
        pytest.skip("weight-bound shape: stays on the split-K path of k_conv_tc")
    call("lb_conv_tc_gemm_ex", ptr(src), ptr(packed), ptr(alpha), ptr(bias), ptr(got32), ptr(got16), ptr(got16a),
         ld16 if any16 else 0, ptr(aux), ld_aux if use_aux else 0, 1 if aux16 else 0, flags, ctypes.byref(g))
    torch.cuda.synchronize()
    exp = ref
    if use_aux:
        exp = ref * (aux[..., :n].float() if auxfac else _roottanh_grad(aux[..., :n].float()))
    scale = exp.abs().max().item()
    if want32:
        assert torch.all(got32[..., n:] == -7.0), "fp32 store outside its channel slice"
        err = (got32[..., :n] - exp).abs().max().item()
        assert err <= 2e-4 * scale + 1e-5, f"{name}/{mode}: fp32 max err {err:.3e} vs scale {scale:.3e}"
    if want16:
        assert torch.all(got16[..., n:].float() == -7.0), "bf16 store outside its channel slice"
        exp16 = _roottanh_grad(exp) if dact else exp
        err = (got16[..., :n].float() - exp16).abs().max().item()
        assert err <= 6e-3 * exp16.abs().max().item() + 1e-5, f"{name}/{mode}: bf16 max err {err:.3e}"
    if want16a:
        assert torch.all(got16a[..., n:].float() == -7.0), "bf16 (activated) store outside its channel slice"
        # with a stored pre-activation the function value belongs to the STORED (bf16-rounded) argument
        arg = got16[..., :n].float() if (want16 and not dact) else exp
        exp16 = _roottanh(arg)
        err = (got16a[..., :n].float() - exp16).abs().max().item()
        assert err <= 6e-3 * exp16.abs().max().item() + 1e-5, f"{name}/{mode}: activated bf16 max err {err:.3e}"


def test_tc_ex_resident_weights():
    """LB_TC2_RESIDENT=1: the weights of one output phase stay in shared memory (opt-in variant of the persistent kernel).
    The library reads its debugging switches once per process, so the variant runs in a child process."""
    import os
    import subprocess
    import sys
    if os.environ.get("LB_TC2_RESIDENT"):
        pytest.skip("already inside the resident-weights child run")
    env = dict(os.environ, LB_TC2_RESIDENT="1")
    sel = "test_tc_ex_matches_v1 and (convT4 or 3x3_48 or 1x1_many or 5x5s2_c32) and (both_aux or o16both)"
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k", sel,
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout
