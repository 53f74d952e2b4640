"""CPU dry run of the Python plumbing: every C-ABI call is replaced by an arity / ctypes-conversion check (no kernel
runs, tensors hold garbage), so signature drift between include/locate_b200.h, _lib.py and the autograd glue shows up
without a GPU.  usage: python scratch/dry_run.py [res] [separable]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import locate_b200 as L
from locate_b200 import _lib, conv_fn, ops, optim, sn_batch, train

calls = {}


def fake_ptr(t):
    return None if t is None else (t.data_ptr() if t.numel() else 16)


def fake_call(name, *args):
    argtypes, _ = _lib._SIGNATURES[name]
    if len(args) + 1 != len(argtypes):
        raise TypeError(f"{name}: {len(args)} args + stream, signature has {len(argtypes)}")
    for i, (a, t) in enumerate(zip(args, argtypes)):
        try:
            if isinstance(a, ctypes.Structure):
                a = ctypes.byref(a)
            t.from_param(a)
        except Exception as exc:
            raise TypeError(f"{name}: argument {i} = {a!r} does not convert to {t}: {exc}")
    calls[name] = calls.get(name, 0) + 1


for mod in (_lib, ops, conv_fn, optim, sn_batch, train):
    if hasattr(mod, "call"):
        mod.call = fake_call
    if hasattr(mod, "ptr"):
        mod.ptr = fake_ptr
# Nadam's arena wants CUDA parameters: pretend
optim.Nadam._attach_orig = optim.Nadam._attach

res = int(sys.argv[1]) if len(sys.argv) > 1 else 32
separable = len(sys.argv) > 2 and sys.argv[2] == "separable"
for precision in ("bf16", "fp32"):
    L.config.reset()
    L.configure(IMAGE_SIZE=res, PRECISION=precision, BASE_FEATURE_FACTOR=2 if res <= 32 else 8, SEPARABLE=separable)
    torch.manual_seed(0)
    gen, dis = L.Generator(), L.Discriminator()
    g_opt = L.Nadam(gen.parameters(), lr=1e-3)
    d_opt = L.Nadam(dis.parameters(), lr=1e-3)
    g_opt._arenas = d_opt._arenas = []          # CPU: no arena
    tr = L.GanTrainer(gen, dis, g_opt, d_opt)
    b = 4
    real, aug, z = torch.randn(b, 3, res, res), torch.randn(b, 3, res, res), torch.randn(b, L.CFG.INPUT_VECTOR_Z)
    for step in range(2):
        tr.step(real, aug, z)
    print(precision, "ok", sum(calls.values()), "calls")
print(sorted(calls.items()))
