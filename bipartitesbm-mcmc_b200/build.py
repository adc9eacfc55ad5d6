"""Build recipe for the native parts: libbisbm.so (CUDA, sm_100a) and bin/mcmc (C++ CLI).

nvcc cross-compiles without a GPU; the outputs are built IN-TREE so they travel to the GPU
box with the repo snapshot (they are git-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbisbm.so")
CLI = os.path.join(ROOT, "bin", "mcmc")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    out = [os.path.join(ROOT, "include", "bisbm.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".cc", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def build_lib(force=False, verbose=False):
    srcs = _sources()
    if not force and not _newer(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "capi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


def build_cli(force=False):
    src = os.path.join(CSRC, "mcmc_main.cc")
    if not os.path.exists(src):
        return None
    if not force and not _newer(CLI, [src, os.path.join(ROOT, "include", "bisbm.h"), LIB]):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", CLI, src,
           "-L", HERE, "-lbisbm", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../bipartitesbm-mcmc_b200"]
    subprocess.check_call(cmd)
    return CLI


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
