"""micro-benchmark of the tiny-channel kernels at the stem / last-layer shapes. usage: small_micro.py [batch]"""
import ctypes, sys, torch
sys.path.insert(0, '.')
from locate_b200 import _lib
from locate_b200._lib import ConvGeom, call, ptr
DEV = 'cuda:0'
def geom(b, ih, iw, ic, oh, ow, oc, kh, kw, s, p, mode, ld_in, ld_out, strides):
    g = ConvGeom()
    g.batch, g.in_h, g.in_w, g.in_c, g.out_h, g.out_w, g.out_c = b, ih, iw, ic, oh, ow, oc
    g.kh, g.kw, g.stride, g.pad, g.mode, g.ld_in, g.ld_out = kh, kw, s, p, mode, ld_in, ld_out
    g.w_sk, g.w_sn, g.w_sty, g.w_stx = strides
    return g
b = int(sys.argv[1]) if len(sys.argv) > 1 else 192
reps = 3
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# (name, h, cin, cout, k, s, p, ld_out_extra, growth_in, dact)
FWD = [("1x1 3->29 cat", 128, 3, 29, 1, 1, 0, 3, 0, 0), ("1x1 48->3 act", 128, 48, 3, 1, 1, 0, 0, 4, 0),
       ("5x5s2 3->3 act", 128, 3, 3, 5, 2, 2, 0, 4, 0), ("1x1 3->32 act", 64, 3, 32, 1, 1, 0, 0, 4, 0)]
for name, h, cin, cout, k, s, p, extra, gin, dact in FWD:
    t = k * k
    oh = (h + 2 * p - k) // s + 1
    wt = torch.randn((cout, cin, k, k), device=DEV)
    x = torch.randn((b, h, h, cin), device=DEV)
    out = torch.empty((b, oh, oh, cout + extra), device=DEV)
    g = geom(b, h, h, cin, oh, oh, cout, k, k, s, p, 0, cin, cout + extra, (t, cin * t, k, 1))
    ms = timeit(lambda: call("lb_conv_small", ptr(x), ptr(wt), None, None, out.data_ptr() + 4 * extra, g, gin, None, 0, 0, 0))
    gb = (x.numel() + b * oh * oh * cout) * 4 / 1e9
    print(f"fwd   {name:16s} {ms*1e3:8.1f} us  {gb/ms*1e3:7.0f} GB/s", flush=True)
    # input gradient (mode 1) with RootTanh' of the input fused
    dy = torch.randn((b, oh, oh, cout), device=DEV)
    dx = torch.empty((b, h, h, cin), device=DEV)
    gd = geom(b, oh, oh, cout, h, h, cin, k, k, s, p, 1, cout, cin, (cin * t, t, k, 1))
    ms = timeit(lambda: call("lb_conv_small", ptr(dy), ptr(wt), None, None, ptr(dx), gd, 0, ptr(x), cin, 4, 0))
    gb = (2 * x.numel() + dy.numel()) * 4 / 1e9
    print(f"dgrad {name:16s} {ms*1e3:8.1f} us  {gb/ms*1e3:7.0f} GB/s", flush=True)
    gw = geom(b, h, h, cin, oh, oh, cout, k, k, s, p, 0, cin, cout, (t, cin * t, k, 1))
    if _lib.lib().lb_conv_small_wgrad_supported(ctypes.byref(gw)) == 1:
        dw = torch.zeros_like(wt)
        ms = timeit(lambda: call("lb_conv_small_wgrad", ptr(x), ptr(dy), ptr(dw), gw, gin))
        gb = (x.numel() + dy.numel()) * 4 / 1e9
        print(f"wgrad {name:16s} {ms*1e3:8.1f} us  {gb/ms*1e3:7.0f} GB/s", flush=True)
