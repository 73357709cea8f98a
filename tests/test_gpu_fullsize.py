"""BASELINE.json configurations at FULL size on the GPU, checked through size-independent properties (the oracle needs
minutes at these sizes): analytic volumes / areas, phi <-> -phi complement, per-cell bounds, and the divergence theorem
    3 |Omega| = int_Gamma x.n dS + int_{boundary of the box, inside Omega} x.n dS
which ties the three rule types together (interior rule, unfitted-boundary rule + normals, exterior facet rules),
like the reference's acceptance demo python/demo/demo_div_thm_no_fenicsx.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cpp():
    from qugar_b200 import cpp
    if cpp.device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu tests must run the CUDA path")
    return cpp


def _cell_origins(dom, cells, n, dim):
    idx = np.stack([(cells // (n ** d)) % n for d in range(dim)], axis=1)
    return idx / float(n)


def _div_theorem(cpp, func, n, p, dim, vol_exact=None, area_exact=None, tol=1e-9):
    """Returns (volume, area, 3V - flux) on [0,1]^dim with n cells per direction."""
    grid = cpp.create_cart_grid([np.linspace(0.0, 1.0, n + 1)] * dim)
    dom = cpp.create_unfitted_impl_domain(func, grid)
    h = 1.0 / n
    cut = dom.get_cut_cells()
    full = dom.get_full_cells()
    q = cpp.create_quadrature(dom, cut, p)
    assert q.n_pts_per_entity.sum() == len(q.weights) and (q.weights > 0).all()
    assert q.points.min() >= 0.0 and q.points.max() <= 1.0
    per_cell = np.add.reduceat(q.weights, np.concatenate([[0], np.cumsum(q.n_pts_per_entity)[:-1]]))
    assert (per_cell > 0).all() and (per_cell < 1.0 + 1e-12).all()          # a cut cell is neither empty nor over-full
    vol = (len(full) + q.weights.sum()) * h ** dim
    # unfitted boundary: x.n dS
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, False, False)
    assert np.allclose(np.sqrt((b.normals ** 2).sum(axis=1)), 1.0, rtol=0, atol=1e-12)
    org = np.repeat(_cell_origins(dom, cut, n, dim), b.n_pts_per_entity, axis=0)
    x = org + b.points * h
    ds = b.weights * h ** (dim - 1)
    area = ds.sum()
    flux = ((x * b.normals).sum(axis=1) * ds).sum()
    # exterior facets of the box that lie (partly) in Omega: x.n = 1 on the max sides, 0 on the min sides
    for d in range(dim):
        lf = 2 * d + 1
        cells = grid.get_boundary_cells(lf)
        fq = cpp.create_facets_quadrature_exterior_integral(dom, cells, np.full(len(cells), lf, np.int32), p)
        flux += fq.weights.sum() * h ** (dim - 1)
    res = dim * vol - flux
    assert abs(res) < tol, (vol, area, res)
    if vol_exact is not None:
        assert abs(vol - vol_exact) < 1e-9 * max(1.0, vol_exact), (vol, vol_exact)
    if area_exact is not None:
        assert abs(area - area_exact) < 1e-7 * area_exact, (area, area_exact)
    return vol, area, res


def test_C1_disk_64x64():
    cpp = _cpp()
    r = 0.4
    for use_bzr in (False, True):
        f = cpp.create_disk(r, np.array([0.5, 0.5]), use_bzr)
        vol, area, _ = _div_theorem(cpp, f, 64, 4, 2)
        assert abs(vol - np.pi * r * r) < 1e-8 and abs(area - 2 * np.pi * r) < 1e-7


def test_C2_sphere_128():
    cpp = _cpp()
    r = 0.4
    for use_bzr in (False, True):
        f = cpp.create_sphere(r, np.array([0.5, 0.5, 0.5]), use_bzr)
        vol, area, _ = _div_theorem(cpp, f, 128, 5, 3)
        assert abs(vol - 4.0 / 3.0 * np.pi * r ** 3) < 1e-9 and abs(area - 4 * np.pi * r * r) < 1e-7


def test_C3_gyroid_256():
    cpp = _cpp()
    for per in ([2.0, 2.0, 2.0], [1.9, 2.1, 2.3]):
        f = cpp.create_Schoen(per)
        # the non-aligned gyroid cuts the box faces; the flux balance then carries the quadrature error of the
        # curved facet / surface rules (measured 5e-9).  The aligned one passes through grid vertices: cells there run into
        # the 16-level bisection cap, whose midpoint test reads the default PsiCode slots (sign -1 since round 2, DESIGN.md
        # section 2) and drops boxes of 2^-16 of a cell: measured closure 3.7e-8, volume 0.5 - 1.7e-10 (with "unconstrained"
        # slots the dropped pieces cancelled by symmetry: 1e-9 / 1e-13)
        vol, area, _ = _div_theorem(cpp, f, 256, 8, 3, tol=1e-7)
        if per[0] == 2.0:
            assert abs(vol - 0.5) < 1e-9       # phi(x + (1/4,1/4,1/4)-type symmetry: half of the unit cube
        # complement: {phi < 0} and {-phi < 0} tile the cube
        vol_c, area_c, _ = _div_theorem(cpp, cpp.create_negative(f), 256, 8, 3, tol=1e-7)
        assert abs(vol + vol_c - 1.0) < 1e-9 and abs(area - area_c) < 1e-7 * area


def test_C4_bezier_deg3_256():
    cpp = _cpp()
    coefs = 2.0 * np.random.default_rng(20260101).random(64) - 1.0
    f = cpp.create_bezier([4, 4, 4], coefs)
    vol, area, _ = _div_theorem(cpp, f, 256, 8, 3, tol=1e-7)
    vol_c, area_c, _ = _div_theorem(cpp, cpp.create_negative(f), 256, 8, 3, tol=1e-7)
    assert abs(vol + vol_c - 1.0) < 1e-12 and abs(area - area_c) < 1e-9 * area


def test_C5_schwarz_diamond_512():
    cpp = _cpp()
    f = cpp.create_SchwarzDiamond([2.0, 2.0, 2.0])
    vol, area, _ = _div_theorem(cpp, f, 512, 6, 3, tol=1e-7)
    assert abs(vol - 0.5) < 1e-9
