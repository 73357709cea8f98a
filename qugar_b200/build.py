"""Builds libqugar_b200.so (CUDA engine + C ABI) in-tree for sm_100a with nvcc.

-fmad=false: no implicit FMA contraction -- the cut-cell path fixes its fp64 operation order
(see csrc/qg_math.cuh).  -lineinfo keeps ncu's source view usable.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libqugar_b200.so")
SOURCES = ["qg_engine.cu"]
HEADERS = ["qg_math.cuh", "qg_funcs.cuh", "qg_quad.cuh", "qg_pipeline.cuh", "qg_device_mem.cuh", "qg_kernels_pipeline.cuh",
           "qg_kernels_domain.cuh", "qg_kernels_rules.cuh", "gauss_table.inc",
           os.path.join("..", "..", "include", "qugar_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-mavx", "-shared",
    "--expt-relaxed-constexpr",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libqugar_b200.so")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
