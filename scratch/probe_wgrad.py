import ctypes, sys, torch
sys.path.insert(0, '.')
from locate_b200 import _lib
from locate_b200._lib import ConvGeom, call, ptr
DEV='cuda:0'
def geom(b, ih, iw, ic, oh, ow, oc, kh, kw, s, p, mode, ld_in, ld_out, strides):
    g = ConvGeom()
    g.batch, g.in_h, g.in_w, g.in_c, g.out_h, g.out_w, g.out_c = b, ih, iw, ic, oh, ow, oc
    g.kh, g.kw, g.stride, g.pad, g.mode, g.ld_in, g.ld_out = kh, kw, s, p, mode, ld_in, ld_out
    g.w_sk, g.w_sn, g.w_sty, g.w_stx = strides
    return g
b,h,w,cin,cout=1,8,8,128,64
g = geom(b,h,w,cin,h,w,cout,1,1,1,0,0,cin,cout,(1,cin,1,1))
for (p0,c0,p1,n0) in [(0,0,0,0),(0,1,0,0),(0,8,0,0),(0,0,0,1),(0,0,0,8),(1,0,1,0),(8,0,8,0),(9,3,9,5),(17,70,17,33),(63,127,63,63),(5,64,5,0), (0,0,1,0)]:
    x = torch.zeros((b,h*w,cin), dtype=torch.bfloat16, device=DEV); dy = torch.zeros((b,h*w,cout), dtype=torch.bfloat16, device=DEV)
    x[0,p0,c0]=1; dy[0,p1,n0]=1
    dwp = torch.zeros((1,cout,cin), device=DEV)
    call("lb_wgrad_tc", ptr(x), ptr(dy), ptr(dwp), ctypes.byref(g)); torch.cuda.synchronize()
    nz = dwp[0].nonzero().tolist()
    print(f"x@(pix{p0},ch{c0}) dy@(pix{p1},ch{n0}) -> nonzero (n,m): {nz[:6]} vals {[round(dwp[0][a][b_].item(),3) for a,b_ in nz[:6]]}")
