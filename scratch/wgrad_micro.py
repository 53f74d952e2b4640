"""Weight-gradient GEMM micro-benchmark over the layer shapes of the 128x128 step (batch 512 by default).
usage: [LB_WGRAD_HALO=0] python scratch/wgrad_micro.py [batch]   -- run once per setting, the switch is read once."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from locate_b200 import _lib
from locate_b200._lib import ConvGeom, call, ptr

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
SHAPES = [  # kind, h (layer input), cin, cout, k, s, p
    ("conv", 128, 48, 48, 3, 1, 1), ("convT", 64, 96, 96, 4, 2, 1), ("conv", 64, 32, 32, 5, 2, 2),
    ("convT", 32, 192, 192, 4, 2, 1), ("convT", 16, 384, 384, 4, 2, 1), ("convT", 8, 768, 768, 4, 2, 1),
    ("convT", 4, 1536, 1536, 4, 2, 1), ("conv", 32, 64, 64, 5, 2, 2), ("conv", 16, 128, 128, 5, 2, 2),
    ("conv", 8, 256, 256, 5, 2, 2), ("conv", 4, 512, 512, 5, 2, 2), ("conv", 128, 48, 96, 1, 1, 0),
    ("conv", 64, 96, 96, 1, 1, 0), ("conv", 8, 256, 256, 1, 1, 0),
]


def geom(b, ih, iw, ic, oh, ow, oc, kh, kw, s, p, mode, ld_in, ld_out, strides):
    g = ConvGeom()
    g.batch, g.in_h, g.in_w, g.in_c, g.out_h, g.out_w, g.out_c = b, ih, iw, ic, oh, ow, oc
    g.kh, g.kw, g.stride, g.pad, g.mode, g.ld_in, g.ld_out = kh, kw, s, p, mode, ld_in, ld_out
    g.w_sk, g.w_sn, g.w_sty, g.w_stx = strides
    return g


def timeit(fn, n=10):
    """GPU time of one call, replayed from a CUDA graph (host launch gaps excluded), L2 flushed before every replay."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    fn(); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


if len(sys.argv) > 2 and sys.argv[2] == "weights":      # tiny batch: the time is the partial stores + k_wgrad_reduce (+ dot)
    SHAPES = [("convT", 4, 1536, 1536, 4, 2, 1), ("conv", 4, 1024, 1024, 5, 2, 2), ("conv", 2, 1024, 1024, 3, 1, 1),
              ("convT", 8, 768, 768, 4, 2, 1), ("conv", 8, 512, 512, 5, 2, 2), ("conv", 8, 768, 1536, 1, 1, 0)]
stat_work = torch.zeros(_lib.lib().lb_stat_work_doubles(), dtype=torch.float64, device=DEV)
dot = torch.zeros(2, dtype=torch.float64, device=DEV)
for kind, h, cin, cout, k, s, p in SHAPES:
    t = k * k
    if kind == "conv":
        oh = (h + 2 * p - k) // s + 1
        x = torch.randn((B, h, h, cin), device=DEV).bfloat16()
        dy = torch.randn((B, oh, oh, cout), device=DEV).bfloat16()
        g = geom(B, h, h, cin, oh, oh, cout, k, k, s, p, 0, cin, cout, (t, cin * t, k, 1)); ga, de = x, dy
        px = B * oh * oh
    else:
        oh = (h - 1) * s - 2 * p + k
        x = torch.randn((B, h, h, cin), device=DEV).bfloat16()
        dy = torch.randn((B, oh, oh, cout), device=DEV).bfloat16()
        g = geom(B, oh, oh, cout, h, h, cin, k, k, s, p, 0, cout, cin, (t, cout * t, k, 1)); ga, de = dy, x
        px = B * h * h
    need = _lib.lib().lb_wgrad_tc_workspace_floats(ctypes.byref(g))
    work = torch.empty(need, device=DEV)
    dwn = torch.empty(t * cin * cout, device=DEV)
    wbar = torch.randn(t * cin * cout, device=DEV)
    ms = timeit(lambda: call("lb_wgrad_tc", ptr(ga), ptr(de), ptr(dwn), ctypes.byref(g), ptr(work), need, ptr(wbar), ptr(dot), ptr(stat_work)))
    flops = 2.0 * px * t * cin * cout
    byts = (x.numel() + dy.numel()) * 2 + dwn.numel() * 4
    print(f"{kind:5s} {k}x{k}s{s} {cin:4d}->{cout:4d} in{h:3d}  {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TF/s  {byts/ms/1e6:7.1f} GB/s  "
          f"splits {need // dwn.numel()}")
