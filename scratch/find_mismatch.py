"""Run one D/G step and, for every plain tensor-core GEMM, compare k_conv_tc (v1) with the persistent kernel on the same operands."""
import ctypes, os, sys
os.environ["LB_TC_V1_ONLY"] = "1"
sys.path.insert(0, '.')
import torch
import locate_b200 as L
from locate_b200 import ops, conv_fn, _lib
from locate_b200._lib import call, ptr
cudart = ctypes.CDLL("libcudart.so.12")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
L.configure(IMAGE_SIZE=S)
dev = 'cuda:0'
torch.manual_seed(999)
gen, g_opt = L.get_model(L.Generator(), L.CFG.GLR, dev)
dis, d_opt = L.get_model(L.Discriminator(), L.CFG.DLR, dev)
tr = L.GanTrainer(gen, dis, g_opt, d_opt)
seen = {}
orig = conv_fn._timed_call
def hooked(family, fl, by, name, *args):
    orig(family, fl, by, name, *args)
    if name != "lb_conv_tc_gemm":
        return
    a, pk, alpha, bias, out, g = args
    key = (g.kh, g.kw, g.stride, g.mode, g.in_c, g.out_c, g.in_h, g.in_w, g.out_h, g.out_w, g.batch, g.ld_in, g.ld_out, bias is not None, out % 16)
    if key in seen:
        return
    if _lib.lib().lb_conv_tc_ex_supported(ctypes.byref(g), 0, 0) != 1 or g.ld_out % 4 or out % 16:
        seen[key] = "n/a"; return
    rows = g.batch * g.out_h * g.out_w
    ref = torch.empty(rows * g.ld_out, device=dev)
    tmp = torch.full((rows * g.ld_out,), 0.0, device=dev)
    torch.cuda.synchronize()
    nbytes = (rows * g.ld_out - (g.ld_out - g.out_c)) * 4
    cudart.cudaMemcpy(ctypes.c_void_p(ref.data_ptr()), ctypes.c_void_p(out), ctypes.c_size_t(nbytes), 3)
    call("lb_conv_tc_gemm_ex", a, pk, alpha, bias, ptr(tmp), None, 0, 0, None, 0, g)
    torch.cuda.synchronize()
    r = ref.view(rows, g.ld_out)[:, :g.out_c]; t = tmp.view(rows, g.ld_out)[:, :g.out_c]
    err = (r - t).abs().max().item(); sc = r.abs().max().item()
    seen[key] = (err, sc)
    flag = "BAD" if err > 1e-3 * sc + 1e-6 else "ok"
    print(f"{flag} err {err:.3e} scale {sc:.3e}  {g.kh}x{g.kw}s{g.stride}m{g.mode} {g.in_c}->{g.out_c} in{g.in_h}x{g.in_w} out{g.out_h}x{g.out_w} b{g.batch} ld_in {g.ld_in} ld_out {g.ld_out} bias {bias is not None}", flush=True)
conv_fn._timed_call = hooked
g = torch.Generator().manual_seed(0)
real = torch.randn((B, 3, S, S), generator=g).clamp_(-1, 1).to(dev); aug = real.clone(); z = torch.randn((B, S), generator=g).to(dev)
tr.step(real, aug, z)
torch.cuda.synchronize()
print("done", len(seen))
