#!/usr/bin/env python
"""Throughput of the reparameterization (not BASELINE.json's metric -- bench.py measures that).

    python tools/bench_reparam.py [--n 256] [--order 4] [--steps 5] [--cpu-cells 20000]

One step = create_reparameterization_levelset / create_reparameterization over the cut cells of the gyroid (2,2,2) grid,
merge off / on.  Reports, as one JSON line per case: device-resident time (CUDA events on the engine stream around qg_reparam,
mesh left in HBM), end-to-end time (qg_reparam + qg_mesh_view: the mesh downloaded to host arrays) and the oracle port on one
host thread over a sample of the same cut cells (test infrastructure used as the CPU baseline, labelled "port").
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu-cells", type=int, default=20000)
    a = ap.parse_args()
    from qugar_b200 import cpp
    lib = cpp._load()
    vp = C.c_void_p
    lib.qg_timer_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.qg_timer_start.argtypes = [vp]
    lib.qg_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    lib.qg_mesh_device_view.argtypes = [vp] + [C.POINTER(C.c_int64)] * 3 + [C.POINTER(vp)] * 3
    per = [2., 2., 2.]
    breaks = [np.linspace(0, 1, a.n + 1)] * 3
    dom = cpp.create_unfitted_impl_domain(cpp.create_Schoen(per), cpp.create_cart_grid(breaks))
    cut = dom.get_cut_cells()
    timer = vp()
    cpp._check(lib.qg_timer_create(0, C.byref(timer)))
    eps = float(np.finfo(np.float64).eps)

    import oracle_lib as O
    import test_oracle_reparam as T
    od = None
    for levelset in (True, False):
        for merge in (False, True):
            mode = 1 if levelset else 0
            dev_ms, e2e_ms = [], []
            sizes = None
            for it in range(a.warmup + a.steps):
                h = vp()
                cpp._check(lib.qg_timer_start(timer))
                cpp._check(lib.qg_reparam(dom._h, cpp._ptr(cut), len(cut), a.order, mode, 1 if merge else 0, 1.0e4 * eps, C.byref(h)))
                ms = C.c_double()
                cpp._check(lib.qg_timer_stop(timer, C.byref(ms)))
                npts, ncells, nw = C.c_int64(), C.c_int64(), C.c_int64()
                lib.qg_mesh_device_view(h, C.byref(npts), C.byref(ncells), C.byref(nw), None, None, None)
                sizes = (npts.value, ncells.value, nw.value)
                lib.qg_mesh_free(h)
                t0 = time.perf_counter()
                m = (cpp.create_reparameterization_levelset if levelset else cpp.create_reparameterization)(dom, cut, a.order, merge)
                t1 = time.perf_counter()
                del m
                if it >= a.warmup:
                    dev_ms.append(ms.value)
                    e2e_ms.append((t1 - t0) * 1e3)
            # CPU: the oracle port, one thread, a contiguous sample of the cut cells
            if od is None:
                od = O.OracleDomain(3, O.K_SCHOEN, per, breaks, libm=False, nthreads=os.cpu_count() or 1)
            sample = cut[len(cut) // 3: len(cut) // 3 + a.cpu_cells]
            olib = T._lib(False)
            t0 = time.perf_counter()
            hh = olib.orc_reparam_domain(od.h, np.ascontiguousarray(sample), len(sample), a.order, int(levelset), int(merge), 1)
            cpu_s = time.perf_counter() - t0
            olib.orc_reparam_free(hh)
            d = float(np.median(dev_ms))
            e = float(np.median(e2e_ms))
            print(json.dumps({
                "metric": "reparameterized cut cells/s", "config": {"workload": "gyroid (2,2,2) %d^3, order %d, %s, merge %s" % (
                    a.n, a.order, "level set" if levelset else "volume (cut cells only)", "on" if merge else "off")},
                "cut_cells": int(len(cut)), "points": sizes[0], "cells": sizes[1], "wire_edges": sizes[2],
                "device_ms": round(d, 3), "value": round(len(cut) / d * 1e3, 1), "e2e_ms": round(e, 3),
                "e2e_value": round(len(cut) / e * 1e3, 1),
                "cpu_baseline": {"kind": "port", "cores": 1, "sample": "%d cut cells" % len(sample),
                                 "value": round(len(sample) / cpu_s, 1)},
                "steps": a.steps, "warmup": a.warmup, "dtype": "f64"}), flush=True)


if __name__ == "__main__":
    main()
