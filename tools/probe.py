"""Quick timing probe of the device-resident path (not the bench)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from qugar_b200 import cpp

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
per = [float(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2., 2., 2.]
lib = cpp._load()
breaks = [np.linspace(0, 1, ng + 1)] * 3
f = cpp.create_Schoen(per)
g = cpp.create_cart_grid(breaks)
for it in range(2):
    t = time.time()
    dom = cpp.create_unfitted_impl_domain(f, g)
    print("domain create %.3f s" % (time.time() - t), dom._counts)
cut = dom.get_cut_cells()
dc = C.c_void_p()
cpp._check(lib.qg_domain_upload_cells(dom._h, cut.ctypes.data_as(C.c_void_p), len(cut), C.byref(dc)))
names = ["ctl", "k0", "k1", "k2", "k3", "scans", "other", "total"]
for surf in (False, True):
    for it in range(3):
        h = C.c_void_p()
        t = time.time()
        fn = lib.qg_quad_unf_bdry_device if surf else lib.qg_quad_cells_device
        cpp._check(fn(dom._h, dc, p, C.byref(h)))
        wall = time.time() - t
        ms = (C.c_double * 8)()
        lib.qg_rule_pass_ms(h, ms)
        ne, npts = C.c_int64(), C.c_int64()
        lib.qg_rule_device_view(h, C.byref(ne), C.byref(npts), None, None, None, None)
        print("surf" if surf else "cell", "wall %.1f ms" % (wall * 1e3), "npts", npts.value, "pts/cell %.1f" % (npts.value / max(len(cut), 1)),
              " ".join("%s=%.2f" % (n, v) for n, v in zip(names, ms)))
        lib.qg_rule_free(h)
