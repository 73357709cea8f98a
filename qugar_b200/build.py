"""Builds libqugar_b200.so (CUDA engine + C ABI) in-tree for sm_100a with nvcc.

Translation units, compiled in parallel and linked into one shared library:
  csrc/qg_engine.cu                host objects, pass scheduling, domain / rule kernels without a level-set kind, C ABI
  csrc/qg_kind.cu  x 7             the pipeline kernels (CTL, K0, K1, K2, K3) instantiated for one level-set kind each:
                                   sphere, Bezier (2-D / 3-D), Schoen gyroid, Schwarz diamond, and the run-time-kind units
                                   (2-D / 3-D) for the rest
-fmad=false: no implicit FMA contraction -- the cut-cell path fixes its fp64 operation order (see csrc/qg_math.cuh).
-lineinfo keeps ncu's source view usable.  Objects go to qugar_b200/_build/ (git-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
OUT = os.path.join(HERE, "libqugar_b200.so")
# headers every unit includes / headers only qg_engine.cu includes (a change there rebuilds the engine unit alone)
HEADERS_COMMON = ["qg_math.cuh", "qg_funcs.cuh", "qg_quad.cuh", "qg_pipeline.cuh", "qg_launch_cfg.cuh", "qg_kernels_kind.cuh",
                  "qg_reparam.cuh", "qg_kernels_reparam.cuh"]
HEADERS_ENGINE = ["qg_device_mem.cuh", "qg_kernels_pipeline.cuh", "qg_kernels_domain.cuh", "qg_kernels_rules.cuh", "qg_coeffs.cuh",
                  "gauss_table.inc", os.path.join("..", "..", "include", "qugar_b200.h")]
HEADERS = HEADERS_COMMON + HEADERS_ENGINE
# (source, object name, extra defines)
def _kind_unit(kind, tag, dim=None):
    defs = [f"-DQG_KIND={kind}", f"-DQG_KIND_TAG={tag}"]
    if dim is not None:
        defs.append(f"-DQG_ONLY_DIM={dim}")
    if kind < 0:
        defs.append("-DQG_BEZIER_GENERIC_ONLY")
    return ("qg_kind.cu", f"kind_{tag}.o", defs)


# slowest first; the run-time-kind and the Bezier unit are split by dimension
UNITS = [_kind_unit(-1, "gend3", 3), ("qg_engine.cu", "engine.o", ["-DQG_BEZIER_GENERIC_ONLY"]), _kind_unit(-1, "gend2", 2),
         _kind_unit(2, "k2d3", 3), _kind_unit(2, "k2d2", 2), _kind_unit(0, "k0"), _kind_unit(10, "k10"), _kind_unit(14, "k14")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-mavx",
    "--expt-relaxed-constexpr",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in HEADERS + ["qg_engine.cu", "qg_kind.cu"]] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in deps)


def _unit_stale(unit) -> bool:
    src, obj, _ = unit
    o = os.path.join(OBJ, obj)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    hdrs = HEADERS if src == "qg_engine.cu" else HEADERS_COMMON
    deps = [os.path.join(CSRC, f) for f in hdrs + [src]] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    todo = [u for u in UNITS if force or _unit_stale(u)]

    def compile_unit(unit):
        src, obj, defs = unit
        cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o",
                                                                                    os.path.join(OBJ, obj)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return unit, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_unit, todo))
    failed = False
    for unit, rc, out in results:
        if verbose or rc != 0:
            sys.stderr.write(f"--- {unit[0]} {' '.join(unit[2])}\n{out}")
        failed = failed or rc != 0
    if failed:
        raise RuntimeError("nvcc failed building libqugar_b200.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + [os.path.join(OBJ, u[1]) for u in UNITS]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libqugar_b200.so")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
