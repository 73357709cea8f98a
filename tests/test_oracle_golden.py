"""The oracle against every golden / known-answer value the reference's own tests hold for the general
(non-Bezier) path.  Sources are cited per test (paths relative to /root/reference).

Pinned exactly: every cell and facet classification count of cpp/test/test_impl_general_quad_0.cpp and
test_impl_domain_0.cpp, and the 144 facet points of the Schwarz-Diamond case.
Pinned to tolerance: analytic volumes / areas / centroids (test_impl_general_quad_1.cpp, _2.cpp), and the golden
centroids of the 3-D Schoen rules at the reference's own tolerance (1e-6).
NOT reproduced: the five hard-coded point totals of the Schoen cases (213 / 60, 795837 / 177421 / 4626).  The tests
`test_schoen_*_reference_point_totals` assert the REFERENCE's numbers and are expected to fail (xfail, strict) with
the measured distance; tools/calibrate_oracle.py sweeps every from-memory choice of the restatement against them
(profiles/r2_oracle_calibration.tsv).  The reference excludes exactly those checks from its CI
because they drift across platforms (.github/workflows/build_and_test.yml:62-71).
"""
import math

import numpy as np
import pytest

import oracle_lib as O


def _facet_counts(d):
    fs = d.facet_status
    n_f = 2 * d.dim
    n_rec = len(d.facet_cells)
    full_cells_plain = len(d.full_cells) - int(np.isin(d.full_cells, d.facet_cells).sum())
    empty_cells_plain = len(d.empty_cells)
    return dict(cut=int(O.is_cut_facet(fs).sum()), full=int((fs == O.F_FULL).sum()) + n_f * full_cells_plain,
                empty=int(O.is_empty_facet(fs).sum()) + n_f * empty_cells_plain, full_unf=int((fs == O.F_FULL_UNF).sum()),
                unf=int(O.has_unf_bdry(fs).sum()), n_rec=n_rec)


@pytest.mark.parametrize("libm", [False, True], ids=["det_sincos", "glibc_sincos"])
def test_schwarz_diamond_2d_counts(libm):
    # cpp/test/test_impl_general_quad_0.cpp:61-150
    d = O.OracleDomain(2, O.K_SCHWARZ_D, [1., 1., 0.], [O.cpp_breaks(12)] * 2, libm=libm)
    assert (len(d.full_cells), len(d.empty_cells), len(d.cut_cells)) == (72, 72, 0)
    fc = _facet_counts(d)
    assert (fc["cut"], fc["full"], fc["empty"], fc["full_unf"], fc["unf"]) == (0, 240, 288, 48, 48)
    q = d.quad_cells(d.cut_cells, 3)
    assert len(q.weights) == 0
    b = d.quad_unf_bdry(d.cut_cells, 3, True, True)
    assert len(b.weights) == 0
    c, f = d.facets_where(lambda s: s == O.F_FULL_UNF)
    r = d.quad_facets(c, f, 3, True)
    assert len(r.weights) == 144
    cen = (r.points[:, 0] * r.weights).sum() / r.weights.sum()
    assert abs(cen - 0.5) < 1e-6
    c, f = d.facets_where(O.has_unf_bdry)
    r = d.quad_facets(c, f, 3, True)
    assert len(r.weights) == 144


@pytest.mark.parametrize("libm", [False, True], ids=["det_sincos", "glibc_sincos"])
def test_schoen_2d_counts(libm):
    # cpp/test/test_impl_general_quad_0.cpp:153-233
    d = O.OracleDomain(2, O.K_SCHOEN, [1., 1., 0.], [O.cpp_breaks(4)] * 2, libm=libm)
    assert (len(d.full_cells), len(d.empty_cells), len(d.cut_cells)) == (4, 4, 8)
    fc = _facet_counts(d)
    assert (fc["cut"], fc["full"], fc["empty"], fc["full_unf"], fc["unf"]) == (8, 28, 28, 0, 0)
    cut = d.cut_cells
    q = d.quad_cells(cut, 3)
    cen = (q.points * q.weights[:, None]).sum(0) / q.weights.sum()
    # points are in cell coordinates; centroid of the cell-local points as the reference computes it
    assert np.allclose(cen, [0.5, 0.5], atol=1e-6)
    b = d.quad_unf_bdry(cut, 3, True, True)
    cen = (b.points * b.weights[:, None]).sum(0) / b.weights.sum()
    assert np.allclose(cen, [0.5, 0.5], atol=1e-6)
    c, f = d.facets_where(O.is_cut_facet)
    r = d.quad_facets(c, f, 3, True)
    assert len(r.weights) == 0            # golden: 0 points (no cut facet is exterior)


@pytest.mark.xfail(strict=True, reason="oracle gives 180 / 48: every cell bisects once less than Algoim does (parity unpinned)")
def test_schoen_2d_reference_point_totals():
    # cpp/test/test_impl_general_quad_0.cpp:205,217 -- the reference's numbers, not the oracle's
    d = O.OracleDomain(2, O.K_SCHOEN, [1., 1., 0.], [O.cpp_breaks(4)] * 2, libm=True)
    q = d.quad_cells(d.cut_cells, 3)
    b = d.quad_unf_bdry(d.cut_cells, 3, True, True)
    assert (len(q.weights), len(b.weights)) == (213, 60)


def test_schoen_2d_point_totals_distance_to_reference():
    # how far the restatement is from the golden totals (regression guard for the calibration)
    d = O.OracleDomain(2, O.K_SCHOEN, [1., 1., 0.], [O.cpp_breaks(4)] * 2, libm=True)
    q = d.quad_cells(d.cut_cells, 3)
    b = d.quad_unf_bdry(d.cut_cells, 3, True, True)
    assert abs(len(q.weights) - 213) <= 33 and abs(len(b.weights) - 60) <= 12


@pytest.mark.parametrize("libm", [False, True], ids=["det_sincos", "glibc_sincos"])
def test_schoen_3d_counts(libm):
    # cpp/test/test_impl_general_quad_0.cpp:236-306
    d = O.OracleDomain(3, O.K_SCHOEN, [2., 2., 2.], [O.cpp_breaks(16)] * 3, libm=libm, nthreads=4)
    assert (len(d.full_cells), len(d.empty_cells), len(d.cut_cells)) == (896, 896, 2304)
    fc = _facet_counts(d)
    assert (fc["cut"], fc["full"], fc["empty"]) == (8448, 8064, 8064)
    cut = d.cut_cells
    q = d.quad_cells(cut, 3, nthreads=4)
    cen = (q.points * q.weights[:, None]).sum(0) / q.weights.sum()
    assert np.allclose(cen, [0.5000000310737452, 0.4999998788332324, 0.4999999483123341], atol=1e-6)
    b = d.quad_unf_bdry(cut, 3, True, True, nthreads=4)
    cen = (b.points * b.weights[:, None]).sum(0) / b.weights.sum()
    assert np.allclose(cen, [0.4999957334273454, 0.4999960466067558, 0.4999963189195731], atol=1e-6)
    c, f = d.facets_where(O.is_cut_facet)
    r = d.quad_facets(c, f, 3, True, nthreads=4)
    cen = (r.points * r.weights[:, None]).sum(0) / r.weights.sum()
    assert np.allclose(cen, [0.50000000000001399, 0.50000000000003864], atol=1e-6)
    # golden totals (795837, 177421, 4626): distance of the restatement, see the xfail test below
    assert abs(len(q.weights) - 795837) / 795837 < 0.002      # 797184: +0.17 %
    assert abs(len(b.weights) - 177421) / 177421 < 0.012      # 175488: -1.09 %
    assert abs(len(r.weights) - 4626) / 4626 < 0.17
    # exact volume fraction of the gyroid on whole periods is 1/2
    vol = (len(d.full_cells) + q.weights.sum()) / 16 ** 3
    assert abs(vol - 0.5) < 1e-6


@pytest.mark.xfail(strict=True, reason="oracle gives 797184 / 175488 / 3888 (+0.17 % / -1.1 % / -16 %): parity unpinned")
def test_schoen_3d_reference_point_totals():
    # cpp/test/test_impl_general_quad_0.cpp:278,290,300 -- the reference's numbers, not the oracle's
    d = O.OracleDomain(3, O.K_SCHOEN, [2., 2., 2.], [O.cpp_breaks(16)] * 3, libm=True, nthreads=4)
    cut = d.cut_cells
    q = d.quad_cells(cut, 3, nthreads=4)
    b = d.quad_unf_bdry(cut, 3, True, True, nthreads=4)
    c, f = d.facets_where(O.is_cut_facet)
    r = d.quad_facets(c, f, 3, True, nthreads=4)
    assert (len(q.weights), len(b.weights), len(r.weights)) == (795837, 177421, 4626)


def test_axis_aligned_planes_have_no_cut_cells():
    # cpp/test/test_impl_domain_0.cpp:62-139 (Plane<2> case): interface on grid lines => no cut cells
    d = O.OracleDomain(2, O.K_PLANE, [0.5, 0.0, 1.0, 0.0], [O.cpp_breaks(4), O.cpp_breaks(5)])
    assert len(d.cut_cells) == 0
    ids = np.arange(20)
    assert np.array_equal(d.full_cells, ids[ids % 4 < 2])
    d = O.OracleDomain(2, O.K_PLANE, [0.0, 0.5, 0.0, 1.0], [O.cpp_breaks(5), O.cpp_breaks(4)])
    assert len(d.cut_cells) == 0
    assert np.array_equal(d.full_cells, np.arange(10))


def _volume_area(d, n):
    cut = d.cut_cells
    q = d.quad_cells(cut, n)
    b = d.quad_unf_bdry(cut, n, True, True)
    nc = np.array(d.n_cells_dir)
    lens = np.array([br[1] - br[0] for br in d.breaks])   # uniform grids here
    cellvol = float(np.prod(lens))
    vol = (len(d.full_cells) + q.weights.sum()) * cellvol
    # quadrature_test_utils.hpp:150-178: area = sum w * |DF^-T n| * cell volume
    area = (b.weights * np.linalg.norm(b.normals / lens, axis=1)).sum() * cellvol
    return vol, area, q, nc


def test_sphere_general_volume_area():
    # cpp/test/test_impl_general_quad_1.cpp: sphere r=0.4 on 4x5x6, n=7, tol 1e-6 (3-D); disk 4x5, tol 1e-8
    d = O.OracleDomain(3, O.K_SPHERE, [0.4, .5, .5, .5], [O.cpp_breaks(4), O.cpp_breaks(5), O.cpp_breaks(6)])
    vol, area, _, _ = _volume_area(d, 7)
    assert abs(vol - 4 / 3 * math.pi * 0.4 ** 3) < 1e-6
    assert abs(area - 4 * math.pi * 0.16) < 1e-6
    d = O.OracleDomain(2, O.K_SPHERE, [0.4, .5, .5], [O.cpp_breaks(4), O.cpp_breaks(5)])
    vol, area, _, _ = _volume_area(d, 7)
    assert abs(vol - math.pi * 0.16) < 1e-8
    assert abs(area - 2 * math.pi * 0.4) < 1e-8


def _centroid(d, q):
    # quadrature_test_utils.hpp:67-104 (compute_centroid): full cells by mid point, cut cells by the rule
    lens = np.array([br[1] - br[0] for br in d.breaks])
    n = np.array(d.n_cells_dir)

    def lo_of(ids):
        t = np.zeros((len(ids), d.dim))
        c = ids.copy()
        for k in range(d.dim):
            t[:, k] = c % n[k]
            c = c // n[k]
        return t * lens + np.array([br[0] for br in d.breaks])

    cellvol = float(np.prod(lens))
    full = d.full_cells
    vol = len(full) * cellvol
    cen = (lo_of(full) + 0.5 * lens).sum(0) * cellvol
    cut = d.cut_cells
    lo = np.repeat(lo_of(cut), q.n_per, axis=0)
    x = lo + q.points * lens
    vol += q.weights.sum() * cellvol
    cen = cen + (x * q.weights[:, None]).sum(0) * cellvol
    return cen / vol


def test_plane_general_volume_area_centroid():
    # cpp/test/test_impl_general_quad_2.cpp:36-90: plane through (0.5,..) with normals (2,1[,0]) etc., n = 2,
    # V = 0.5, A = sqrt(1.25), centroid (13/48, 5/12[, 0.5]); tolerance 10 eps in the reference
    tol = 1e-14
    d = O.OracleDomain(2, O.K_PLANE, [0.5, 0.5, 2.0, 1.0], [O.cpp_breaks(2), O.cpp_breaks(2)])
    vol, area, q, _ = _volume_area(d, 2)
    assert abs(vol - 0.5) < tol and abs(area - math.sqrt(1.25)) < tol
    assert np.allclose(_centroid(d, q), [13 / 48, 5 / 12], atol=tol)
    perms = [([2., 1., 0.], [13 / 48, 5 / 12, 0.5]), ([0., 2., 1.], [0.5, 13 / 48, 5 / 12]), ([1., 0., 2.], [5 / 12, 0.5, 13 / 48])]
    for nrm, cen in perms:
        d = O.OracleDomain(3, O.K_PLANE, [0.5, 0.5, 0.5, *nrm], [O.cpp_breaks(4), O.cpp_breaks(5), O.cpp_breaks(6)])
        vol, area, q, _ = _volume_area(d, 2)
        assert abs(vol - 0.5) < tol and abs(area - math.sqrt(1.25)) < tol
        assert np.allclose(_centroid(d, q), cen, atol=tol)


def _check_vol_centroid_area(d, n, vol_t, cen_t, area_t, tol):
    """test_volume_and_centroid (cpp/test/quadrature_test_utils.hpp:180-235): absolute tolerance on all three."""
    vol, area, q, _ = _volume_area(d, n)
    assert abs(vol - vol_t) < tol, (vol, vol_t)
    assert abs(area - area_t) < tol, (area, area_t)
    assert np.allclose(_centroid(d, q), cen_t, rtol=0, atol=tol)


def test_annulus_and_torus_general_quad():
    # cpp/test/test_impl_general_quad_3.cpp:36-61: annulus R0 = 0.10, R1 = 0.40 at (0.48, 0.55), grid 16x17, n = 7, tol 1e-6
    r0, r1 = 0.25 - 0.15, 0.25 + 0.15
    d = O.OracleDomain(2, O.K_ANNULUS, [r0, r1, 0.48, 0.55], [O.cpp_breaks(16), O.cpp_breaks(17)])
    _check_vol_centroid_area(d, 7, math.pi * (r1 * r1 - r0 * r0), [0.48, 0.55], 2 * math.pi * (r0 + r1), 1e-6)
    # :64-91: torus R = 0.35, r = 0.08 at (0.45, 0.51, 0.49), axis (1, 1.2, 0.9), grid 10x12x14, n = 7, tol 1e-3
    ax = np.array([1.0, 1.2, 0.9])
    ax = ax / np.sqrt((ax * ax).sum())
    d = O.OracleDomain(3, O.K_TORUS, [0.35, 0.08, 0.45, 0.51, 0.49, *ax], [O.cpp_breaks(10), O.cpp_breaks(12), O.cpp_breaks(14)])
    _check_vol_centroid_area(d, 7, 2 * math.pi ** 2 * 0.08 ** 2 * 0.35, [0.45, 0.51, 0.49], 4 * math.pi ** 2 * 0.08 * 0.35, 1e-3)


def test_cylinder_general_quad():
    # cpp/test/test_impl_general_quad_4.cpp:36-62: r = 0.45 through (0.5,0.5,0.5) along each axis, grid 4x5x6, n = 7, tol 1e-6
    for k in range(3):
        ax = [0.0, 0.0, 0.0]
        ax[k] = 1.0
        d = O.OracleDomain(3, O.K_CYLINDER, [0.45, 0.5, 0.5, 0.5, *ax], [O.cpp_breaks(4), O.cpp_breaks(5), O.cpp_breaks(6)])
        _check_vol_centroid_area(d, 7, math.pi * 0.45 ** 2, [0.5, 0.5, 0.5], 2 * math.pi * 0.45, 1e-6)


def test_ellipse_and_ellipsoid_general_quad():
    # cpp/test/test_impl_general_quad_5.cpp:36-61: ellipse (0.25, 0.15) in RefSystem((0.45,0.55), x-axis (1,0.5)), grid 9x10,
    # n = 7, tol 1e-6, perimeter 1.2763499392751274 ("computed numerically" upstream)
    from qugar_b200 import cpp
    rs = cpp.create_ref_system([0.45, 0.55], [1.0, 0.5])
    d = O.OracleDomain(2, O.K_ELLIPSOID, [0.25, 0.15, *rs.origin, *rs.basis.reshape(-1)], [O.cpp_breaks(9), O.cpp_breaks(10)])
    _check_vol_centroid_area(d, 7, math.pi * 0.25 * 0.15, [0.45, 0.55], 1.2763499392751274, 1e-6)
    # :64-93: ellipsoid (0.3, 0.4, 0.35) in RefSystem((0.45,0.55,0.49), (1,0.5,1), (0,1,0)), grid 9x10x11, area 1.535246912774376
    rs = cpp.create_ref_system([0.45, 0.55, 0.49], [1.0, 0.5, 1.0], [0.0, 1.0, 0.0])
    d = O.OracleDomain(3, O.K_ELLIPSOID, [0.3, 0.4, 0.35, *rs.origin, *rs.basis.reshape(-1)],
                       [O.cpp_breaks(9), O.cpp_breaks(10), O.cpp_breaks(11)])
    _check_vol_centroid_area(d, 7, 4 / 3 * math.pi * 0.3 * 0.4 * 0.35, [0.45, 0.55, 0.49], 1.535246912774376, 1e-6)


def test_gyroid_volume_converges_to_half():
    # exact by symmetry; checks that no root is lost on fine cells far from the origin
    for n_cells, p in ((24, 5), (40, 6)):
        d = O.OracleDomain(3, O.K_SCHOEN, [2., 2., 2.], [np.linspace(0, 1, n_cells + 1)] * 3, nthreads=4)
        q = d.quad_cells(d.cut_cells, p, nthreads=4)
        vol = (len(d.full_cells) + q.weights.sum()) / n_cells ** 3
        assert abs(vol - 0.5) < 1e-12


def test_det_sincos_accuracy():
    import mpmath as mp
    lib = O.load(False)
    import ctypes as C
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-50, 50, 400), rng.uniform(-1e5, 1e5, 200), np.arange(-8, 9) * (math.pi / 4)])
    mp.mp.dps = 40
    worst = 0.0
    for x in xs:
        s, c = C.c_double(), C.c_double()
        lib.orc_sincos(float(x), C.byref(s), C.byref(c))
        es, ec = mp.sin(mp.mpf(float(x))), mp.cos(mp.mpf(float(x)))
        for got, ex in ((s.value, es), (c.value, ec)):
            ulp = math.ulp(float(ex)) if float(ex) != 0 else 5e-324
            worst = max(worst, abs(float(mp.mpf(got) - ex)) / ulp)
    assert worst < 1.5, worst


def test_gauss_table_matches_reference_checksum():
    import hashlib
    import struct
    import os
    vals = []
    for line in open(os.path.join(O.ORACLE_DIR, "gauss_table.inc")):
        if line.startswith("//"):
            continue
        a, b = line.strip().rstrip(",").split(", ")
        vals += [float.fromhex(a), float.fromhex(b)]
    digest = hashlib.sha256(struct.pack("<%dd" % len(vals), *vals)).hexdigest()
    want = open(os.path.join(os.path.dirname(__file__), "golden", "gauss_table.sha256")).read().split()[0]
    assert digest == want
    # product copy identical
    prod = open(os.path.join(O.ROOT, "qugar_b200", "csrc", "gauss_table.inc")).read()
    assert prod == open(os.path.join(O.ORACLE_DIR, "gauss_table.inc")).read()


@pytest.mark.parametrize("dim", [2, 3])
def test_bezier_rescale_domain_and_extract_facet(dim):
    """cpp/test/test_impl_bezier_tp_3.cpp:46-127: a random Bezier of order 3..5 rescaled to a random box in [-2,2]^dim has,
    at the rescaled points, the values of the original (tol 1e-10); so do its facets."""
    rng = np.random.default_rng(1234 + dim)
    for _ in range(5):
        order = rng.integers(3, 6, dim)
        coefs = rng.random(int(np.prod(order))) * 2.0 - 1.0
        p0, p1 = rng.uniform(-2, 2, dim), rng.uniform(-2, 2, dim)
        lo, hi = np.minimum(p0, p1), np.maximum(p0, p1)
        new, sg = O.bezier_rescale(order, coefs, lo, hi)
        assert sg == (1 if (new > 0).all() else -1 if (new < 0).all() else 0)
        par = np.concatenate([order.astype(np.float64), coefs])
        par_new = np.concatenate([order.astype(np.float64), new])
        pts = lo + rng.random((5, dim)) * (hi - lo)
        scaled = (pts - lo) / (hi - lo)
        v, _ = O.func_eval(dim, O.K_BEZIER, par, pts)
        vs, _ = O.func_eval(dim, O.K_BEZIER, par_new, scaled)
        assert np.abs(v - vs).max() < 1e-10
        # facets: the coefficient slice at index 0 / order-1 of the constant direction (bezier_tp.cpp:350-376)
        c = new.reshape(order[::-1])
        for lf in range(2 * dim):
            cd, side = lf // 2, lf % 2
            sl = np.take(c, 0 if side == 0 else order[cd] - 1, axis=dim - 1 - cd)
            sp = scaled.copy()
            sp[:, cd] = float(side)
            ev = pts.copy()
            ev[:, cd] = lo[cd] if side == 0 else hi[cd]
            v, _ = O.func_eval(dim, O.K_BEZIER, par, ev)
            vs, _ = O.func_eval(dim, O.K_BEZIER, par_new, sp)
            assert np.abs(v - vs).max() < 1e-10
            if dim == 3:
                fo = np.delete(order, cd)
                fpar = np.concatenate([fo.astype(np.float64), np.ascontiguousarray(sl).reshape(-1)])
                vf, _ = O.func_eval(2, O.K_BEZIER, fpar, np.delete(sp, cd, axis=1))
                assert np.abs(v - vf).max() < 1e-10


def test_calibration_knobs_move_the_golden_totals():
    """The from-memory choices the calibration sweeps (tools/calibrate_oracle.py) are live: the sin/cos remainder constant
    changes the 2-D totals the way profiles/r2_oracle_calibration.tsv records, and the defaults are restored afterwards."""
    import ctypes as C
    lib = O.load(True)
    lib.orc_set_knob.argtypes = [C.c_char_p, C.c_double]

    def totals():
        d = O.OracleDomain(2, O.K_SCHOEN, [1., 1., 0.], [O.cpp_breaks(4)] * 2, libm=True)
        return len(d.quad_cells(d.cut_cells, 3).weights), len(d.quad_unf_bdry(d.cut_cells, 3, True, True).weights)

    assert totals() == (180, 48)
    try:
        assert lib.orc_set_knob(b"trig_C", 1.5) == 0
        assert totals() == (288, 60)              # the boundary total of the reference, the interior one overshoots
        assert lib.orc_set_knob(b"rf_cap_mode", 2.0) == 0
        assert totals() == (252, 48)
        assert lib.orc_set_knob(b"no_such_knob", 1.0) == -1
    finally:
        lib.orc_set_knob(b"trig_C", 1.0)
        lib.orc_set_knob(b"rf_cap_mode", 0.0)
    assert totals() == (180, 48)


def test_bezier_variants_of_the_primitives_are_the_reference_polynomials():
    """CylinderBzr / EllipsoidBzr / AnnulusBzr / TorusBzr (primitive_funcs_lib.cpp:162-197, 277-311, 418-470, 570-716): the
    Bernstein form built by qugar_b200.cpp evaluates (through the oracle's de Casteljau) to the polynomial the reference
    tabulates -- |y|^2 - (a.y)^2 - r^2, sum (v_d.y)^2 / s_d^2 - 1, (|y|^2 - r^2)(|y|^2 - R^2), (|y|^2 + R^2 - r^2)^2 -
    4 R^2 (|y|^2 - (a.y)^2) -- and has the sign of the closed-form level set."""
    from qugar_b200 import cpp
    rng = np.random.default_rng(0)
    p3, p2 = rng.random((300, 3)), rng.random((300, 2))

    def bez(f, pts):
        return O.func_eval(f.dim, O.K_BEZIER, np.concatenate([np.asarray(f.order, float), f.coefs]), pts)[0]

    o, a = np.array([.5, .45, .5]), np.array([.6, 0., .8])
    y = p3 - o
    assert np.abs(bez(cpp.create_cylinder(0.3, o, a, True), p3) - ((y * y).sum(1) - (y @ a) ** 2 - 0.09)).max() < 1e-13
    rs = cpp.create_ref_system(np.array([.5, .5, .5]), np.array([.8, .6, 0.]), np.array([-.6, .8, 0.]))
    y = p3 - rs.origin
    ref = sum((y @ rs.basis[d]) ** 2 / s ** 2 for d, s in enumerate([.4, .3, .2])) - 1.0
    assert np.abs(bez(cpp.create_ellipsoid([.4, .3, .2], rs, True), p3) - ref).max() < 1e-13
    c = np.array([.5, .5])
    y2 = ((p2 - c) ** 2).sum(1)
    f = cpp.create_annulus(0.2, 0.42, c, True)
    assert np.abs(bez(f, p2) - (y2 - .04) * (y2 - .42 ** 2)).max() < 1e-13
    closed = O.func_eval(2, O.K_ANNULUS, [0.2, 0.42, .5, .5], p2)[0]
    assert (np.sign(bez(f, p2)) == np.sign(closed)).all()
    c, a = np.array([.5, .5, .5]), np.array([0., .6, .8])
    y = p3 - c
    y2 = (y * y).sum(1)
    f = cpp.create_torus(0.3, 0.12, c, a, True)
    assert np.abs(bez(f, p3) - ((y2 + .09 - .0144) ** 2 - 4 * .09 * (y2 - (y @ a) ** 2))).max() < 1e-13
    closed = O.func_eval(3, O.K_TORUS, [0.3, 0.12, .5, .5, .5, 0., .6, .8], p3)[0]
    assert (np.sign(bez(f, p3)) == np.sign(closed)).all()
