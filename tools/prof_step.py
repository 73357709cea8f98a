"""Device-resident step (domain + interior rule + boundary rule) on a gyroid grid of a chosen size, with the
engine's own per-pass CUDA-event times.  Used under ncu (smaller grids keep the replay passes cheap).
usage: python tools/prof_step.py [grid=256] [npts=8] [reps=3] [periods=2,2,2 | bezier3 | sphere | sphere_bzr]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench

grid_n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
what = sys.argv[4] if len(sys.argv) > 4 else "2,2,2"
eng = bench.Engine()
cpp = eng.cpp
full = np.linspace(0.0, 1.0, grid_n + 1)
if what == "bezier3":      # BASELINE.json configs[3]: degree-3 BezierTP, coefficients 2U-1, rng 20260101
    func = cpp.create_bezier([4, 4, 4], 2.0 * np.random.default_rng(20260101).random(64) - 1.0)
elif what == "sphere":     # configs[1]
    func = cpp.create_sphere(0.4, np.array([0.5, 0.5, 0.5]), False)
elif what == "sphere_bzr":
    func = cpp.create_sphere(0.4, np.array([0.5, 0.5, 0.5]), True)
else:
    func = cpp.create_Schoen([float(v) for v in what.split(",")])
grid = cpp.create_cart_grid([full] * 3)
names = ["ctl", "k0", "k1", "k2", "k3", "scans", "other", "total"]
for it in range(reps):
    ncut, n1, n2, ms1, ms2 = eng.device_step(func, grid, 0, 0, grid_n, npts)
    print(f"step {it}: cut {ncut} interior {n1} boundary {n2}")
    print("  interior:", " ".join(f"{n}={v:.2f}" for n, v in zip(names, ms1)))
    print("  boundary:", " ".join(f"{n}={v:.2f}" for n, v in zip(names, ms2)))
print(f"lines: interior {eng.lines_interior} boundary {eng.lines_boundary} nodes: interior {n1} boundary {n2}")
