"""micro-benchmark: k_conv_tc (v1) vs the persistent fused-epilogue kernel (lb_conv_tc_gemm_ex), forward GEMMs of the
generator/discriminator layer shapes, CUDA-event timed.  usage: conv_micro2.py [batch] [case ...]"""
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
  "convT96": dict(kind="convT", h=64, cin=96, cout=96, k=4, s=2, p=1),
  "convT192": dict(kind="convT", h=32, cin=192, cout=192, k=4, s=2, p=1),
  "convT384": dict(kind="convT", h=16, cin=384, cout=384, k=4, s=2, p=1),
  "convT768": dict(kind="convT", h=8, cin=768, cout=768, k=4, s=2, p=1),
  "convT1536": dict(kind="convT", h=4, cin=1536, cout=1536, k=4, s=2, p=1),
  "c3x3_48": dict(kind="conv", h=128, cin=48, cout=48, k=3, s=1, p=1),
  "c1x1_96_48": dict(kind="conv", h=128, cin=96, cout=48, k=1, s=1, p=0),
  "c1x1_192_96": dict(kind="conv", h=64, cin=192, cout=96, k=1, s=1, p=0),
  "c1x1_384_192": dict(kind="conv", h=32, cin=384, cout=192, k=1, s=1, p=0),
  "c1x1_96": dict(kind="conv", h=64, cin=96, cout=96, k=1, s=1, p=0),
  "c5x5s2_32": dict(kind="conv", h=64, cin=32, cout=32, k=5, s=2, p=2),
  "c5x5s2_64": dict(kind="conv", h=32, cin=64, cout=64, k=5, s=2, p=2),
  "c5x5s2_128": dict(kind="conv", h=16, cin=128, cout=128, k=5, s=2, p=2),
}
args = sys.argv[1:]
b = int(args[0]) if args and args[0].isdigit() else 96
names = [a for a in args if not a.isdigit()] or list(CASES)
reps = 5
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"batch {b}")
for name in names:
    c = CASES[name]; k, s, p, t = c["k"], c["s"], c["p"], c["k"] ** 2
    h, cin, cout = c["h"], c["cin"], c["cout"]
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
    out16 = torch.empty((b, oh, oh, cout), device=DEV, dtype=torch.bfloat16)
    aux = torch.randn((b, oh, oh, cout), device=DEV)
    g = geom(b, h, h, cin, oh, oh, cout, k, k, s, p, mode, cin, cout, strides)
    packed = torch.empty(_lib.lib().lb_conv_tc_packed_elems(ctypes.byref(g)), dtype=torch.bfloat16, device=DEV)
    call("lb_conv_tc_pack", ptr(wt), ptr(packed), g)
    if _lib.lib().lb_conv_tc_ex_supported(ctypes.byref(g), cout, cout) != 1:
        print(name, "not covered by the persistent kernel"); continue
    ms1 = timeit(lambda: call("lb_conv_tc_gemm", ptr(x), ptr(packed), None, None, ptr(out), g))
    ms2 = timeit(lambda: call("lb_conv_tc_gemm_ex", ptr(x), ptr(packed), None, None, ptr(out), None, 0, 0, None, 0, g))
    ms3 = timeit(lambda: call("lb_conv_tc_gemm_ex", ptr(x), ptr(packed), None, None, ptr(out), ptr(out16), cout, 1, None, 0, g))
    ms4 = timeit(lambda: call("lb_conv_tc_gemm_ex", ptr(x), ptr(packed), None, None, None, ptr(out16), cout, 0, ptr(aux), cout, g))
    ms5 = timeit(lambda: call("lb_conv_tc_gemm_ex", ptr(x), ptr(packed), None, None, None, ptr(out16), cout, 0, None, 0, g))
    # weight gradient of the same layer
    dy16 = torch.randn((b, oh, oh, cout), device=DEV).bfloat16()
    if c["kind"] == "conv":
        gw = geom(b, h, h, cin, oh, oh, cout, k, k, s, p, 0, cin, cout, (t, cin * t, k, 1)); ga, de = x, dy16
    else:
        gw = geom(b, oh, oh, cout, h, h, cin, k, k, s, p, 0, cout, cin, (t, cout * t, k, 1)); ga, de = dy16, x
    dwp = torch.zeros(t * cin * cout, device=DEV)
    msw = timeit(lambda: call("lb_wgrad_tc", ptr(ga), ptr(de), ptr(dwp), gw))
    io32 = (x.numel() * 2 + out.numel() * 4) / 1e9
    tf = lambda ms: flops / ms / 1e9
    print(f"{name:13s} v1 {ms1*1e3:8.1f}us {tf(ms1):7.1f}TF | ex32 {ms2*1e3:8.1f}us {tf(ms2):7.1f}TF {io32/ms2*1e3:6.0f}GB/s | "
          f"ex32+16act {ms3*1e3:8.1f}us | ex16+aux {ms4*1e3:8.1f}us | ex16 {ms5*1e3:8.1f}us {tf(ms5):7.1f}TF | wgrad {msw*1e3:8.1f}us {tf(msw):7.1f}TF", flush=True)
