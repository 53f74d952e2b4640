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
gen = torch.Generator().manual_seed(0)
for (b,h,w,cin,cout) in [(1,8,8,128,64),(1,8,8,64,64),(1,8,8,128,128),(2,8,8,128,64),(1,16,16,128,64),(2,16,16,64,64),(1,8,8,96,48),(1,8,8,256,64),(1,8,8,128,256),(1,8,8,128,320)]:
    g = geom(b,h,w,cin,h,w,cout,1,1,1,0,0,cin,cout,(1,cin,1,1))
    x = torch.randn((b,h*w,cin), generator=gen).bfloat16().to(DEV); dy = torch.randn((b,h*w,cout), generator=gen).bfloat16().to(DEV)
    dwp = torch.zeros((1,cout,cin), device=DEV)
    call("lb_wgrad_tc", ptr(x), ptr(dy), ptr(dwp), ctypes.byref(g)); torch.cuda.synchronize()
    ref = torch.einsum('bpn,bpm->nm', dy.float(), x.float())
    err = (dwp[0]-ref).abs()
    print((b,h,w,cin,cout), 'max err', err.max().item(), 'scale', ref.abs().max().item(), 'bad rows(n)', (err.max(dim=1).values>0.1).sum().item(), 'bad cols(m)', (err.max(dim=0).values>0.1).sum().item())
