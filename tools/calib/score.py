"""Score an oracle knob setting against the reference's golden totals and centroids
(cpp/test/test_impl_general_quad_0.cpp:153-306)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
import oracle_lib as O

GOLD2 = (213, 60, 0)
GOLD3 = (795837, 177421, 4626)
GOLD3_CEN_INT = np.array([0.5000000310737452, 0.4999998788332324, 0.4999999483123341])
GOLD3_CEN_BND = np.array([0.4999957334273454, 0.4999960466067558, 0.4999963189195731])
GOLD3_CEN_FAC = np.array([0.50000000000001399, 0.50000000000003864])


def cen(r):
    # quadrature_test_utils.hpp:36-53 (sequential accumulation)
    w = r.weights
    return (r.points * w[:, None]).sum(0) / w.sum()


def score(libm=False, do3d=True, nthreads=8, do2d=True):
    out = {}
    if do2d:
        d = O.OracleDomain(2, O.K_SCHOEN, [1., 1., 0.], [O.cpp_breaks(4)] * 2, libm=libm)
        cut = d.cut_cells
        q = d.quad_cells(cut, 3); b = d.quad_unf_bdry(cut, 3, True, True)
        c, f = d.facets_where(O.is_cut_facet)
        r = d.quad_facets(c, f, 3, True)
        out["c2"] = (len(d.full_cells), len(d.empty_cells), len(cut))
        out["t2"] = (len(q.weights), len(b.weights), len(r.weights))
        out["per2"] = (q.n_per.tolist(), b.n_per.tolist())
    if do3d:
        d = O.OracleDomain(3, O.K_SCHOEN, [2., 2., 2.], [O.cpp_breaks(16)] * 3, libm=libm, nthreads=nthreads)
        cut = d.cut_cells
        q = d.quad_cells(cut, 3, nthreads=nthreads); b = d.quad_unf_bdry(cut, 3, True, True, nthreads=nthreads)
        c, f = d.facets_where(O.is_cut_facet)
        r = d.quad_facets(c, f, 3, True, nthreads=nthreads)
        out["c3"] = (len(d.full_cells), len(d.empty_cells), len(cut))
        out["t3"] = (len(q.weights), len(b.weights), len(r.weights))
        out["d3"] = tuple(a - g for a, g in zip(out["t3"], GOLD3))
        out["cen_int"] = (cen(q) - 0.5) * 1e7      # gold: +0.31 -1.21 -0.52
        out["cen_bnd"] = (cen(b) - 0.5) * 1e6      # gold: -4.27 -3.95 -3.68
        out["cen_fac"] = (cen(r) - 0.5) * 1e14 if len(r.weights) else None  # gold: 1.4 3.9
        out["rules3"] = (q, b, r, d)
    return out


def show(tag, r):
    print(tag, r.get("t2"), r.get("per2"), r.get("c3"), r.get("t3"), r.get("d3"),
          np.round(r["cen_int"], 3) if "cen_int" in r else "", np.round(r["cen_bnd"], 3) if "cen_bnd" in r else "",
          np.round(r["cen_fac"], 2) if r.get("cen_fac") is not None else "", flush=True)


if __name__ == "__main__":
    print("gold cen_int*1e7", (GOLD3_CEN_INT - .5) * 1e7, "cen_bnd*1e6", (GOLD3_CEN_BND - .5) * 1e6, "fac*1e14",
          (GOLD3_CEN_FAC - .5) * 1e14)
    lib = O.load(True)
    for ts in (1, 0):
        lib.orc_set_tol_scale(ts)
        show("tol_scale=%d" % ts, score(libm=True))
