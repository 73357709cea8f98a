"""Wall and device time of the device-resident step on z-slabs of the 256^3 gyroid grid (what a rank of an N-GPU run does).
usage: python tools/prof_slab.py [layers ...]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench

eng = bench.Engine()
cpp = eng.cpp
GRID, NPTS = 256, 8
full = np.linspace(0.0, 1.0, GRID + 1)
func = cpp.create_Schoen([2.0, 2.0, 2.0])
grid = cpp.create_cart_grid([full] * 3)
for layers in [int(v) for v in sys.argv[1:]] or [256, 128, 64, 32]:
    ts = []
    for it in range(6):
        t0 = time.perf_counter()
        ncut, n1, n2, ms1, ms2 = eng.device_step(func, grid, 0, 0, layers, NPTS)
        ts.append(time.perf_counter() - t0)
    t = min(ts[2:]) * 1e3
    print(f"layers {layers:4d}: cut {ncut:7d}  step {t:7.2f} ms  (interior {ms1[7]:.2f} boundary {ms2[7]:.2f} -> domain+rest {t - ms1[7] - ms2[7]:.2f})"
          f"  passes int {[round(v, 2) for v in ms1[:7]]} bnd {[round(v, 2) for v in ms2[:7]]}")
