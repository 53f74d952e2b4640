"""micro-benchmark of one conv geometry on the tcgen05 kernel (forward GEMM only), CUDA-event timed."""
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
CASES = {
  "convT96": dict(kind="convT", b=96, h=64, cin=96, cout=96, k=4, s=2, p=1),
  "convT192": dict(kind="convT", b=96, h=32, cin=192, cout=192, k=4, s=2, p=1),
  "convT1536": dict(kind="convT", b=96, h=4, cin=1536, cout=1536, k=4, s=2, p=1),
  "c3x3_48": dict(kind="conv", b=96, h=128, cin=48, cout=48, k=3, s=1, p=1),
  "c1x1_96": dict(kind="conv", b=96, h=64, cin=96, cout=96, k=1, s=1, p=0),
  "c1x1_48_96": dict(kind="conv", b=96, h=128, cin=48, cout=96, k=1, s=1, p=0),
}
names = sys.argv[1:] or list(CASES)
reps = 5
for name in names:
    c = CASES[name]; k, s, p, t = c["k"], c["s"], c["p"], c["k"] ** 2
    b, h, cin, cout = c["b"], c["h"], c["cin"], c["cout"]
    if c["kind"] == "conv":
        oh = (h + 2 * p - k) // s + 1
        wt = torch.randn((cout, cin, k, k), device=DEV); strides = (t, cin * t, k, 1); mode = 0
        flops = 2.0 * b * oh * oh * t * cin * cout
    else:
        oh = (h - 1) * s - 2 * p + k
        wt = torch.randn((cin, cout, k, k), device=DEV); strides = (cout * t, t, k, 1); mode = 1
        flops = 2.0 * b * h * h * t * cin * cout
    x = torch.randn((b, h, h, cin), device=DEV).bfloat16()
    out = torch.empty((b, oh, oh, cout), device=DEV)
    g = geom(b, h, h, cin, oh, oh, cout, k, k, s, p, mode, cin, cout, strides)
    packed = torch.empty(_lib.lib().lb_conv_tc_packed_elems(ctypes.byref(g)), dtype=torch.bfloat16, device=DEV)
    call("lb_conv_tc_pack", ptr(wt), ptr(packed), g)
    for _ in range(2): call("lb_conv_tc_gemm", ptr(x), ptr(packed), None, None, ptr(out), g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call("lb_conv_tc_gemm", ptr(x), ptr(packed), None, None, ptr(out), g)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    io = (x.numel() * 2 + out.numel() * 4 + packed.numel() * 2) / 1e9
    print(f"{name:12s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TF/s  io {io/ms*1e3:7.1f} GB/s  ({io*1e3:.0f} MB)")
