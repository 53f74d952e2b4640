"""In-tree build of the C-ABI library: nvcc (sm_100a) -> locate_b200/csrc/liblocate_b200.so.

No JIT cache: the .so sits next to the sources so it travels to the GPU box with the repo snapshot.
`python -m locate_b200.build [--force]`.
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIB = os.path.join(CSRC, "liblocate_b200.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math" if False else "-DLB_PRECISE_MATH",   # precise division/sqrt: fp32 parity with the oracle
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-I", INCLUDE, "-I", CSRC,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link liblocate_b200.so.  Returns the path."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")) + [__file__]
    srcs = sources()
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    todo = [(s, o) for s, o in zip(srcs, objs) if force or _stale(o, [s] + headers)]

    def compile_one(pair):
        src, obj = pair
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return res.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:
        logs = list(pool.map(compile_one, todo))
    if verbose:
        print("\n".join(logs))
    if todo or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
