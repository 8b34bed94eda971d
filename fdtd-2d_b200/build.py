"""Build recipe for libfdtd2d.so (hand-written CUDA for sm_100a, C ABI in include/fdtd2d.h).

The library is built IN-TREE (fdtd-2d_b200/libfdtd2d.so) so that it travels to the GPU box with the
repository snapshot.  nvcc cross-compiles for sm_100a without a GPU.
  -fmad=false : no FMA contraction -- results must be bit-identical to the reference's numpy
                arithmetic; the kernels also use explicit *_rn intrinsics.
  -lineinfo   : so ncu's source page maps back to the .cu/.cuh files.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libfdtd2d.so")
SOURCES = ["api.cu", "resident_x2.cu"]  # two translation units, compiled side by side (each takes 1-1.5 minutes)
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(PKG_DIR, "build")  # (git-ignored)


def _newest_source_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libfdtd2d.so if it is missing or older than its sources; return its path."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *extra, *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        jobs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, job in jobs:
        if job.wait() != 0:
            for _, other in jobs:
                if other.poll() is None:
                    other.kill()
            raise subprocess.CalledProcessError(job.returncode, cmd)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH, *objs]
    if verbose:
        print(" ".join(link), file=sys.stderr)
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
