"""Build libsphb.so (the CUDA product library) in-tree for sm_100a.

`python -m sphugo_b200.build` or `sphugo_b200.build.build()`.  nvcc cross-compiles without a GPU.
The reference is a GOAMD64=v1 build (no FMA contraction): everything that decides a neighbour set
or moves a particle (d^2, drift, kick, wrap, reflections) uses explicit __dmul_rn/__dadd_rn and is bit-identical;
the 32-term density / force sums may contract to FMA (differences ~1e-16, tolerance 1e-12).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsphb.so")
SOURCES = ["sphb.cu"]
DEPS = ["sphb.cu", "sphb_kernels.cuh", "sphb_reuse.cuh", "sphb_ring.cuh", "sphb_slab.inc", "sphb_ring.inc", os.path.join("..", "..", "include", "sphb.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550", "-ldl",
]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
