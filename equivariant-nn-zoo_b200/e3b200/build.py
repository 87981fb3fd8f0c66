"""Builds libe3b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG), "csrc")
LIB_DIR = os.path.join(os.path.dirname(PKG), "lib")
LIB_PATH = os.path.join(LIB_DIR, "libe3b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
SOURCES = ["e3b200.cu", "tp_fast.cu", "tp_fast_p1.cu", "tp_fast_p2.cu", "tp_fast_p3.cu", "gemm_tf32x3.cu", "wgrad_tf32x3.cu"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    gen = os.path.join(CSRC, "gen_tp.py")
    gen_out = [os.path.join(CSRC, "tp_generated.cuh"), os.path.join(CSRC, "cg_tables.cuh")]
    gen_deps = [gen, os.path.join(PKG, "cg.py"), os.path.join(PKG, "plan.py"), os.path.join(PKG, "irreps.py")]
    if force or any(_stale(o, gen_deps) for o in gen_out):
        subprocess.check_call([sys.executable, gen])
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(PKG)), "include", "e3b200.h"))
    sources = [os.path.join(CSRC, src) for src in SOURCES]
    if not force and not _stale(LIB_PATH, sources + headers):
        return LIB_PATH
    # all translation units in parallel (~45 s); the objects are removed after the link: with -lineinfo they are 35 MB that
    # would travel with every snapshot of the tree, and staleness is judged on the library alone
    objs, procs = [], []
    for s in sources:
        o = os.path.join(LIB_DIR, os.path.basename(s).replace(".cu", ".o"))
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd)))
    failed = [cmd for cmd, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed: " + " ".join(failed[0]))
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs)
    for o in objs:
        os.remove(o)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
