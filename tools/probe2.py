"""Per-call wall times of the bench's device-resident step and of the host (e2e) calls."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench

eng = bench.Engine()
cpp, L = eng.cpp, eng.lib
full = np.linspace(0, 1, bench.GRID + 1)
func = cpp.create_Schoen(bench.PERIODS)
grid = cpp.create_cart_grid([full] * 3)


def tic():
    return time.perf_counter()


for it in range(4):
    t0 = tic()
    dom, dc, r1, r2 = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    cpp._check(L.qg_domain_create_slab(func._h, grid._h, 0, 0, 256, C.byref(dom)))
    t1 = tic()
    cpp._check(L.qg_domain_cut_cells_device(dom, C.byref(dc)))
    t2 = tic()
    cpp._check(L.qg_quad_cells_device(dom, dc, 8, C.byref(r1)))
    t3 = tic()
    cpp._check(L.qg_quad_unf_bdry_device(dom, dc, 8, C.byref(r2)))
    t4 = tic()
    L.qg_rule_free(r1)
    L.qg_rule_free(r2)
    L.qg_dcells_free(dc)
    L.qg_domain_free(dom)
    t5 = tic()
    print("dev step: domain %.1f cutlist %.1f cells %.1f surf %.1f free %.1f ms" % tuple(1e3 * v for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))

for it in range(3):
    t0 = tic()
    hdom = C.c_void_p()
    cpp._check(L.qg_domain_create_slab(func._h, grid._h, 0, 0, 256, C.byref(hdom)))
    t1 = tic()
    cnt = (C.c_int64 * 4)()
    L.qg_domain_counts(hdom, cnt)
    cut = np.empty(cnt[2], np.int64)
    cpp._check(L.qg_domain_cells(hdom, 0, cut.ctypes.data_as(C.c_void_p)))
    t2 = tic()
    dc, r1 = C.c_void_p(), C.c_void_p()
    cpp._check(L.qg_domain_upload_cells(hdom, cut.ctypes.data_as(C.c_void_p), len(cut), C.byref(dc)))
    cpp._check(L.qg_quad_cells_device(hdom, dc, 8, C.byref(r1)))
    t3 = tic()
    cpp._check(L.qg_rule_download(r1))
    t4 = tic()
    L.qg_rule_free(r1)
    L.qg_dcells_free(dc)
    L.qg_domain_free(hdom)
    t5 = tic()
    print("host: domain %.1f cells->host %.1f compute %.1f download %.1f (%.1f GB/s) free %.1f ms" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 674035200 * 32 / (t4 - t3) / 1e9, 1e3 * (t5 - t4)))
