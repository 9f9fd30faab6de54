"""Build libkuq.so (the CUDA extension) in-tree with nvcc for sm_100a.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libkuq.so")
SOURCES = ["kuq_kernels.cu", "kuq_api.cu", "kuq_dbbuild.cu", "kuq_microbench.cu", "kuq_layout_exp.cu", "kuq_clades.cu"]
HEADERS = ["kuq_kernels.cuh", "kuq_minimizer.cuh", os.path.join(ROOT, "include", "kuq.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-bindir", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "shared"]


def nvcc():
    for c in (os.environ.get("KUQ_NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libkuq.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


CLASSIFY = os.path.join(PKG, "bin", "classify")


def build_classify(force: bool = False) -> str:
    """The drop-in `classify` executable (host C++ over libkuq.so) and, from the same source with -DEXACT_COUNTING
    like the reference's Makefile, `classifyExact` (what `krakenuniq --exact` runs)."""
    src = os.path.join(CSRC, "classify_main.cpp")
    build(force=force)
    exact = CLASSIFY + "Exact"
    newest = max(os.path.getmtime(src), os.path.getmtime(LIB))
    if not force and all(os.path.exists(x) and os.path.getmtime(x) >= newest for x in (CLASSIFY, exact)):
        return CLASSIFY
    os.makedirs(os.path.dirname(CLASSIFY), exist_ok=True)
    for out, defs in ((CLASSIFY, []), (exact, ["-DEXACT_COUNTING"])):
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-fopenmp"] + defs + [src, "-o", out, "-L" + os.path.dirname(LIB),
               "-lkuq", "-lz", "-ldl", "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building " + os.path.basename(out))
    return CLASSIFY


def build_dbtools(force: bool = False):
    """The drop-in `db_sort` and `set_lcas` executables (host C++ over libkuq.so)."""
    src = os.path.join(CSRC, "dbbuild_main.cpp")
    build(force=force)
    outs = [(os.path.join(PKG, "bin", "db_sort"), []), (os.path.join(PKG, "bin", "set_lcas"), ["-DTOOL_SET_LCAS"])]
    newest = max(os.path.getmtime(src), os.path.getmtime(LIB))
    if not force and all(os.path.exists(o) and os.path.getmtime(o) >= newest for o, _ in outs):
        return [o for o, _ in outs]
    os.makedirs(os.path.join(PKG, "bin"), exist_ok=True)
    for out, defs in outs:
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall"] + defs + [src, "-o", out, "-L" + os.path.dirname(LIB), "-lkuq",
               "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building " + os.path.basename(out))
    return [o for o, _ in outs]


if __name__ == "__main__":
    build_dbtools(force="--force" in sys.argv)
    build_classify(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
