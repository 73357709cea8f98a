"""Host-API interior rule (compute + download) timing, for the transport path.  usage: python tools/prof_e2e.py [reps]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench

eng = bench.Engine()
cpp, L = eng.cpp, eng.lib
full = np.linspace(0, 1, 257)
func = cpp.create_Schoen([2.0, 2.0, 2.0])
grid = cpp.create_cart_grid([full] * 3)
dom = cpp.create_unfitted_impl_domain(func, grid)
cut = dom.get_cut_cells()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t0 = time.perf_counter()
    q = cpp.create_quadrature(dom, cut, 8)
    t1 = time.perf_counter()
    chk = float(q.points[0, 0] + q.points[-1, 2] + q.weights[-1])
    print(f"create_quadrature: {1e3 * (t1 - t0):.1f} ms  ({len(q.weights)} nodes, check {chk:.6f})")
    del q
