"""Builds libLRNDE.so (nvcc, sm_100a) and liblrnde_hostcheck.so (g++, CPU) in-tree.

    python localregneuralde.jl_b200/build.py [--force]

The CUDA library is the product; the host-check library only re-exports the scalar
step-size logic of csrc/lrnde_controller.h so the CPU test-suite can pin it bit-for-bit
against the oracle without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libLRNDE.so")
HOSTCHECK = os.path.join(HERE, "liblrnde_hostcheck.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources():
    out = []
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                out.append(os.path.join(d, f))
    out.append(os.path.abspath(__file__))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if force or not _newer(LIB, srcs):
        cus = [s for s in srcs if s.endswith(".cu")]
        cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *cus, "-lcuda"]
        if os.environ.get("LRNDE_TRACE"):
            cmd.insert(1, "-DLRNDE_UMMA_TRACE")
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.run(cmd, check=True, cwd=CSRC)
    if force or not _newer(HOSTCHECK, srcs):
        cpps = [s for s in srcs if s.endswith(".cpp")]
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", HOSTCHECK, *cpps]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
