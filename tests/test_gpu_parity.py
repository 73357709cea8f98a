"""GPU parity: the CUDA path, called through the C ABI (qugar_b200.cpp -> libqugar_b200.so), against the
CPU oracle on the same inputs.  Bit-exact: classification, per-cell counts, points, weights, normals."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _cpp():
    from qugar_b200 import cpp
    if cpp.device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu tests must run the CUDA path")
    return cpp


CASES = [
    # name, dim, oracle kind, params, factory, grid n, pts/dir
    ("schoen3d_aligned", 3, O.K_SCHOEN, [2., 2., 2.], "create_Schoen", 16, 3),
    ("schoen3d", 3, O.K_SCHOEN, [1.9, 2.1, 2.3], "create_Schoen", 24, 4),
    ("schoen2d", 2, O.K_SCHOEN, [1., 1., 0.], "create_Schoen", 4, 3),
    ("schwarzd3d", 3, O.K_SCHWARZ_D, [1.3, 0.7, 1.1], "create_SchwarzDiamond", 20, 2),
    ("schwarzd2d_aligned", 2, O.K_SCHWARZ_D, [1., 1., 0.], "create_SchwarzDiamond", 12, 3),
    ("schwarzp3d", 3, O.K_SCHWARZ_P, [1.0, 1.5, 0.8], "create_SchwarzPrimitive", 14, 5),
    ("iwp3d", 3, O.K_SCHOEN_IWP, [1.1, 0.9, 1.0], "create_SchoenIWP", 12, 3),
    # power-of-two frequencies: the compile-time same-argument instantiation (FuncKSame) of K0 / K1 / K2
    ("iwp3d_pow2", 3, O.K_SCHOEN_IWP, [2.0, 1.0, 0.5], "create_SchoenIWP", 10, 3),
    ("schoen3d_pow2", 3, O.K_SCHOEN, [1.0, 4.0, 2.0], "create_Schoen", 14, 4),
    ("schoen2d_pow2", 2, O.K_SCHOEN, [2.0, 1.0, 0.25], "create_Schoen", 20, 5),
    ("schoen3d_mixed", 3, O.K_SCHOEN, [2.0, 1.7, 1.0], "create_Schoen", 12, 3),
    ("frd3d", 3, O.K_SCHOEN_FRD, [0.9, 1.0, 1.1], "create_SchoenFRD", 12, 3),
    ("fks3d", 3, O.K_FISCHER_KOCH_S, [1.0, 0.8, 1.2], "create_FischerKochS", 12, 3),
    ("sphere3d", 3, O.K_SPHERE, [0.4, .5, .5, .5], "create_sphere", 12, 5),
    ("disk2d", 2, O.K_SPHERE, [0.4, .5, .5], "create_disk", 64, 4),
]


def _make_func(cpp, factory, dim, params):
    if factory in ("create_sphere", "create_disk"):
        return getattr(cpp, factory)(params[0], np.array(params[1:]), False)
    if dim == 2:
        return getattr(cpp, factory)(params[:2], params[2])
    return getattr(cpp, factory)(params)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_domain_and_rules_match_oracle(case):
    cpp = _cpp()
    name, dim, kind, params, factory, n, p = case
    breaks = [O.cpp_breaks(n)] * dim
    dom = cpp.create_unfitted_impl_domain(_make_func(cpp, factory, dim, params), cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, kind, params, breaks)
    # classification: bit-exact
    assert np.array_equal(od.status, dom._cell_status())
    rc, rs = dom._facet_records()
    assert np.array_equal(od.facet_cells, rc)
    assert np.array_equal(od.facet_status, rs)
    cut = dom.get_cut_cells()
    assert np.array_equal(cut, od.cut_cells)
    # interior rule
    q = cpp.create_quadrature(dom, cut, p)
    ro = od.quad_cells(cut, p)
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points, q.points)
    assert np.array_equal(ro.weights, q.weights)
    # unfitted-boundary rule (python callers pass include_facet_unf_bdry=False, exclude_ext_bdry=False)
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, False, False)
    rb = od.quad_unf_bdry(cut, p, False, False)
    assert np.array_equal(rb.n_per, b.n_pts_per_entity)
    assert np.array_equal(rb.points, b.points)
    assert np.array_equal(rb.weights, b.weights)
    assert np.array_equal(rb.normals, b.normals)


def test_mixed_cell_list_full_empty_cut():
    cpp = _cpp()
    n, p = 10, 3
    breaks = [np.linspace(0, 1, n + 1)] * 3
    per = [1.2, 0.9, 1.4]
    dom = cpp.create_unfitted_impl_domain(cpp.create_Schoen(per), cpp.create_cart_grid(breaks))
    od = O.OracleDomain(3, O.K_SCHOEN, per, breaks)
    rng = np.random.default_rng(1)
    cells = rng.permutation(n**3)[:400].astype(np.int64)
    q = cpp.create_quadrature(dom, cells, p)
    ro = od.quad_cells(cells, p)
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points, q.points)
    assert np.array_equal(ro.weights, q.weights)
    assert np.array_equal(q.cells, cells)


def test_func_eval_matches_oracle():
    cpp = _cpp()
    rng = np.random.default_rng(0)
    pts = rng.random((1000, 3))
    f = cpp.create_Schoen([1.9, 2.1, 2.3])
    v, g = O.func_eval(3, O.K_SCHOEN, [1.9, 2.1, 2.3], pts)
    assert np.array_equal(f.eval(pts), v)
    assert np.array_equal(f.grad(pts), g)
    fn = cpp.create_negative(f)
    assert np.array_equal(fn.eval(pts), -v)


def test_empty_cell_list_and_bad_ids():
    cpp = _cpp()
    breaks = [np.linspace(0, 1, 5)] * 2
    dom = cpp.create_unfitted_impl_domain(cpp.create_disk(0.3, np.array([0.5, 0.5]), False), cpp.create_cart_grid(breaks))
    q = cpp.create_quadrature(dom, np.zeros(0, np.int64), 3)
    assert len(q.weights) == 0 and len(q.n_pts_per_entity) == 0
    with pytest.raises(ValueError):
        cpp.create_quadrature(dom, np.array([999], np.int64), 3)
    with pytest.raises(ValueError):
        cpp.create_quadrature(dom, np.array([0], np.int64), 0)


FACET_CASES = [c for c in CASES if c[0] in ("schoen3d_aligned", "schoen3d", "schoen2d", "schwarzd2d_aligned", "sphere3d", "disk2d")]


def _domain_pair(case):
    cpp = _cpp()
    name, dim, kind, params, factory, n, p = case
    breaks = [O.cpp_breaks(n)] * dim
    dom = cpp.create_unfitted_impl_domain(_make_func(cpp, factory, dim, params), cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, kind, params, breaks)
    return cpp, dom, od, dim, n, p


def _interesting_cells(od, dim, n):
    """Cells with facet records (cut, full with unfitted boundary) plus a sample of plain full / empty / boundary cells."""
    rng = np.random.default_rng(7)
    other = rng.permutation(n**dim)[:60].astype(np.int64)
    return np.unique(np.concatenate([od.facet_cells, other, np.arange(min(n, 8), dtype=np.int64)]))


@pytest.mark.parametrize("case", FACET_CASES, ids=[c[0] for c in FACET_CASES])
@pytest.mark.parametrize("exterior", [False, True])
def test_facet_rules_match_oracle(case, exterior):
    cpp, dom, od, dim, n, p = _domain_pair(case)
    base = _interesting_cells(od, dim, n)
    nf = 2 * dim
    cells = np.repeat(base, nf)
    facets = np.tile(np.arange(nf, dtype=np.int32), len(base))
    make = cpp.create_facets_quadrature_exterior_integral if exterior else cpp.create_facets_quadrature_interior_integral
    q = make(dom, cells, facets, p)
    ro = od.quad_facets(cells, facets, p, exterior)
    assert q.points.shape[1] == dim - 1
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points.reshape(-1, dim - 1), q.points)
    assert np.array_equal(ro.weights, q.weights)
    assert np.array_equal(q.cells, cells) and np.array_equal(q.facets, facets)
    if not exterior:
        assert q.n_pts_per_entity.sum() > 0


@pytest.mark.parametrize("case", FACET_CASES, ids=[c[0] for c in FACET_CASES])
@pytest.mark.parametrize("exclude_ext", [False, True])
def test_unf_bdry_rule_with_facet_parts(case, exclude_ext):
    cpp, dom, od, dim, n, p = _domain_pair(case)
    cells = _interesting_cells(od, dim, n)
    b = cpp.create_unfitted_bound_quadrature(dom, cells, p, True, exclude_ext)
    rb = od.quad_unf_bdry(cells, p, True, exclude_ext)
    assert np.array_equal(rb.n_per, b.n_pts_per_entity)
    assert np.array_equal(rb.points.reshape(-1, dim), b.points)
    assert np.array_equal(rb.weights, b.weights)
    assert np.array_equal(rb.normals.reshape(-1, dim), b.normals)


def test_facet_rule_bad_arguments():
    cpp = _cpp()
    breaks = [np.linspace(0, 1, 5)] * 2
    dom = cpp.create_unfitted_impl_domain(cpp.create_disk(0.3, np.array([0.5, 0.5]), False), cpp.create_cart_grid(breaks))
    with pytest.raises(ValueError):
        cpp.create_facets_quadrature_interior_integral(dom, np.array([0], np.int64), np.array([4], np.int32), 2)
    with pytest.raises(ValueError):
        cpp.create_facets_quadrature_exterior_integral(dom, np.array([99], np.int64), np.array([0], np.int32), 2)
    q = cpp.create_facets_quadrature_exterior_integral(dom, np.zeros(0, np.int64), np.zeros(0, np.int32), 2)
    assert len(q.weights) == 0


# ---- Bezier level sets (BezierTP de Casteljau evaluation, integrated with the general algorithm) -------------
def _bezier_case(name):
    from test_emu_parity import bezier_params
    cpp = _cpp()
    if name == "sphere_bzr":
        return 3, cpp.create_sphere(0.4, np.array([.5, .5, .5])), bezier_params(3, name), 10, 4      # use_bzr defaults to True
    if name == "disk_bzr":
        return 2, cpp.create_disk(0.4, np.array([.5, .5]), True), bezier_params(2, name), 32, 4
    par = bezier_params(3, name)
    return 3, cpp.create_bezier([4, 4, 4], par[3:]), par, 8, 3


@pytest.mark.parametrize("name", ["sphere_bzr", "disk_bzr", "random_deg3"])
def test_bezier_domain_and_rules_match_oracle(name):
    cpp = _cpp()
    dim, func, params, n, p = _bezier_case(name)
    breaks = [O.cpp_breaks(n)] * dim
    dom = cpp.create_unfitted_impl_domain(func, cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, O.K_BEZIER, params, breaks)
    assert np.array_equal(od.status, dom._cell_status())
    rc, rs = dom._facet_records()
    assert np.array_equal(od.facet_cells, rc) and np.array_equal(od.facet_status, rs)
    cut = dom.get_cut_cells()
    assert len(cut) > 0
    q = cpp.create_quadrature(dom, cut, p)
    ro = od.quad_cells(cut, p)
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points, q.points) and np.array_equal(ro.weights, q.weights)
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, False, False)
    rb = od.quad_unf_bdry(cut, p, False, False)
    assert np.array_equal(rb.n_per, b.n_pts_per_entity)
    assert np.array_equal(rb.points, b.points) and np.array_equal(rb.weights, b.weights)
    assert np.array_equal(rb.normals, b.normals)
    # function values / gradients
    pts = np.random.default_rng(5).random((500, dim))
    v, g = O.func_eval(dim, O.K_BEZIER, params, pts)
    assert np.array_equal(func.eval(pts), v) and np.array_equal(func.grad(pts), g)


def test_bezier_sphere_volume_and_area():
    """cpp/test/test_impl_poly_quad_2.cpp:56-73 (sphere r=0.4 in [0,1]^3, Bezier level set): volume and area."""
    cpp = _cpp()
    n, p, r = 16, 5, 0.4
    breaks = [np.linspace(0, 1, n + 1)] * 3
    f = cpp.create_sphere(r, np.array([.5, .5, .5]))
    assert type(f).__name__ == "SphereBzr"
    pts = np.random.default_rng(2).random((100, 3))
    assert np.allclose(f.eval(pts), ((pts - 0.5) ** 2).sum(axis=1) - r * r, rtol=0, atol=1e-14)
    assert np.allclose(f.grad(pts), 2 * (pts - 0.5), rtol=0, atol=1e-13)
    dom = cpp.create_unfitted_impl_domain(f, cpp.create_cart_grid(breaks))
    cut = dom.get_cut_cells()
    cell_vol = 1.0 / n ** 3
    q = cpp.create_quadrature(dom, cut, p)
    vol = (len(dom.get_full_cells()) + q.weights.sum()) * cell_vol
    assert abs(vol - 4.0 / 3.0 * np.pi * r ** 3) < 1e-6 * vol
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, False, False)
    # physical area = w * |cell| * |DF^-T n| ; uniform grid: |DF^-T n| = n (quadrature_test_utils.hpp:150-178)
    area = b.weights.sum() * cell_vol * n
    assert abs(area - 4.0 * np.pi * r ** 2) < 1e-6 * area


# ---- remaining primitives (general variants) -----------------------------------------------------------------
def _prim_cases():
    cpp = _cpp()
    rs3 = cpp.create_ref_system([.5, .5, .5], [0.8, 0.6, 0.0], [-0.6, 0.8, 0.0])
    rs2 = cpp.create_ref_system([.5, .5], [0.8, 0.6])
    return [
        ("cylinder", 3, O.K_CYLINDER, cpp.create_cylinder(0.3, [.5, .45, .5], [0.6, 0.0, 0.8], False), 10, 3),
        ("ellipsoid", 3, O.K_ELLIPSOID, cpp.create_ellipsoid([0.4, 0.3, 0.2], rs3, False), 10, 3),
        ("ellipse", 2, O.K_ELLIPSOID, cpp.create_ellipse([0.4, 0.25], rs2, False), 16, 4),
        ("annulus", 2, O.K_ANNULUS, cpp.create_annulus(0.2, 0.42, [.5, .5], False), 16, 3),
        ("torus", 3, O.K_TORUS, cpp.create_torus(0.3, 0.12, [.5, .5, .5], [0.0, 0.6, 0.8], False), 12, 3),
        ("iwp_generic_instantiation", 3, O.K_SCHOEN_IWP, cpp.create_SchoenIWP([1.1, 0.9, 1.0]), 10, 3),
        ("dim_linear3d", 3, O.K_DIM_LINEAR, cpp.create_dim_linear([-0.3, 0.2, 0.1, 0.5, -0.4, 0.3, -0.1, 0.6]), 6, 3),
        ("dim_linear2d", 2, O.K_DIM_LINEAR, cpp.create_dim_linear([-0.3, 0.2, 0.4, -0.5]), 8, 4),
    ]


@pytest.mark.parametrize("idx", range(8))
def test_primitives_match_oracle(idx):
    cpp = _cpp()
    name, dim, kind, func, n, p = _prim_cases()[idx]
    params = func._params
    breaks = [O.cpp_breaks(n)] * dim
    dom = cpp.create_unfitted_impl_domain(func, cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, kind, params, breaks)
    assert np.array_equal(od.status, dom._cell_status()), name
    cut = dom.get_cut_cells()
    assert len(cut) > 0
    q = cpp.create_quadrature(dom, cut, p)
    ro = od.quad_cells(cut, p)
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points, q.points) and np.array_equal(ro.weights, q.weights)
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, True, False)
    rb = od.quad_unf_bdry(cut, p, True, False)
    assert np.array_equal(rb.n_per, b.n_pts_per_entity)
    assert np.array_equal(rb.points, b.points) and np.array_equal(rb.weights, b.weights)
    assert np.array_equal(rb.normals, b.normals)
    pts = np.random.default_rng(11).random((300, dim))
    v, g = O.func_eval(dim, kind, params, pts)
    assert np.array_equal(func.eval(pts), v) and np.array_equal(func.grad(pts), g)


def test_primitive_volumes():
    cpp = _cpp()
    n, p = 16, 5
    cases = [
        (cpp.create_ellipsoid([0.4, 0.3, 0.2], cpp.create_ref_system([.5, .5, .5], [0.8, 0.6, 0.0], [-0.6, 0.8, 0.0]), False),
         4.0 / 3.0 * np.pi * 0.4 * 0.3 * 0.2, 1e-6),
        (cpp.create_torus(0.3, 0.12, [.5, .5, .5], [0.0, 0.6, 0.8], False), 2.0 * np.pi ** 2 * 0.3 * 0.12 ** 2, 1e-6),
        (cpp.create_annulus(0.2, 0.42, [.5, .5], False), np.pi * (0.42 ** 2 - 0.2 ** 2), 1e-6),
    ]
    # the polynomial (Bzr) twins -- the reference's default (cpp/test/test_impl_poly_quad_3..5.cpp: tol 1e-6 / 1e-3)
    cases += [
        (cpp.create_ellipsoid([0.4, 0.3, 0.2], cpp.create_ref_system([.5, .5, .5], [0.8, 0.6, 0.0], [-0.6, 0.8, 0.0]), True),
         4.0 / 3.0 * np.pi * 0.4 * 0.3 * 0.2, 1e-6),
        (cpp.create_torus(0.3, 0.12, [.5, .5, .5], [0.0, 0.6, 0.8], True), 2.0 * np.pi ** 2 * 0.3 * 0.12 ** 2, 1e-5),
        (cpp.create_annulus(0.2, 0.42, [.5, .5], True), np.pi * (0.42 ** 2 - 0.2 ** 2), 1e-6),
        (cpp.create_cylinder(0.3, [.5, .5, .5], [0.0, 0.0, 1.0], True), np.pi * 0.09, 1e-6),
    ]
    for f, exact, tol in cases:
        dim = f.dim
        dom = cpp.create_unfitted_impl_domain(f, cpp.create_cart_grid([np.linspace(0, 1, n + 1)] * dim))
        q = cpp.create_quadrature(dom, dom.get_cut_cells(), p)
        vol = (len(dom.get_full_cells()) + q.weights.sum()) / n ** dim
        assert abs(vol - exact) < tol * exact, (type(f).__name__, vol, exact)


# ---- composite functions: affinely transformed, added, subtracted ------------------------------------------------
@pytest.mark.parametrize("name,dim,n,p", [("transformed_sphere", 3, 10, 3), ("schoen_minus_constant", 3, 8, 3),
                                          ("disk_plus_transformed_disk", 2, 16, 3), ("rotated_square", 2, 12, 3),
                                          ("rotated_cube", 3, 8, 2), ("nested_2d", 2, 16, 3), ("nested_3d", 3, 8, 3)])
def test_composite_functions_match_oracle(name, dim, n, p):
    from test_emu_parity import composite_func
    cpp = _cpp()
    func = composite_func(name)
    breaks = [O.cpp_breaks(n)] * dim
    dom = cpp.create_unfitted_impl_domain(func, cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, O.K_COMPOSITE, func._params, breaks)
    assert np.array_equal(od.status, dom._cell_status())
    cut = dom.get_cut_cells()
    assert len(cut) > 0
    q = cpp.create_quadrature(dom, cut, p)
    ro = od.quad_cells(cut, p)
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points, q.points) and np.array_equal(ro.weights, q.weights)
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, True, False)
    rb = od.quad_unf_bdry(cut, p, True, False)
    assert np.array_equal(rb.n_per, b.n_pts_per_entity)
    assert np.array_equal(rb.points, b.points) and np.array_equal(rb.weights, b.weights) and np.array_equal(rb.normals, b.normals)
    pts = np.random.default_rng(3).random((200, dim))
    v, g = O.func_eval(dim, O.K_COMPOSITE, func._params, pts)
    assert np.array_equal(func.eval(pts), v) and np.array_equal(func.grad(pts), g)
    # the negative of a composite flips the whole expression
    neg = cpp.create_negative(func)
    assert np.array_equal(neg.eval(pts), -v)


def test_transformed_sphere_is_the_ellipsoid():
    from test_emu_parity import composite_func
    cpp = _cpp()
    n, p = 16, 5
    f = composite_func("transformed_sphere")
    dom = cpp.create_unfitted_impl_domain(f, cpp.create_cart_grid([np.linspace(0, 1, n + 1)] * 3))
    q = cpp.create_quadrature(dom, dom.get_cut_cells(), p)
    vol = (len(dom.get_full_cells()) + q.weights.sum()) / n ** 3
    exact = 4.0 / 3.0 * np.pi * 0.35 * 0.2 * 0.28
    assert abs(vol - exact) < 1e-6 * exact
    # arbitrary nesting is a postfix program; only its size is bounded (6 leaves, stack depth 4)
    g = cpp.create_functions_addition(f, cpp.create_functions_addition(f, f))
    pts = np.random.default_rng(5).random((50, 3))
    assert np.allclose(g.eval(pts), 3.0 * f.eval(pts), rtol=1e-14, atol=1e-14)
    with pytest.raises(NotImplementedError):
        h = g
        for _ in range(3):
            h = cpp.create_functions_addition(h, g)


def test_square_volume_and_perimeter():
    """create_square (Square<dim>, impl_funcs_lib.cpp:44-98): the level set is only piecewise smooth and its Interval
    evaluation follows the branch of the box centre, so cells holding an edge of the box are integrated approximately
    (the oracle gives the same numbers: area to 2e-4, perimeter to 5 % on this grid); faces through cell interiors are exact."""
    cpp = _cpp()
    n, p = 24, 4
    f = cpp.create_square(cpp.create_affine_transformation([.5, .45], [2., 1.], 0.3, 0.2))
    dom = cpp.create_unfitted_impl_domain(f, cpp.create_cart_grid([np.linspace(0, 1, n + 1)] * 2))
    cut = dom.get_cut_cells()
    q = cpp.create_quadrature(dom, cut, p)
    area = (len(dom.get_full_cells()) + q.weights.sum()) / n ** 2
    assert abs(area - 4 * 0.3 * 0.2) < 1e-3 * area
    b = cpp.create_unfitted_bound_quadrature(dom, cut, p, False, False)
    assert abs(b.weights.sum() / n - 4 * (0.3 + 0.2)) < 0.1 * 2.0
    # axis-aligned plain box [-1,1]^3 clipped by the grid [0,2]^3: the octant [0,1]^3, planar faces through grid planes
    g = cpp.create_square_3D()
    pts = np.random.default_rng(5).uniform(-2, 2, (500, 3))
    assert np.array_equal(g.eval(pts), np.abs(pts).max(axis=1) - 1.0)
    gr = g.grad(pts)
    ind = np.abs(pts).argmax(axis=1)
    assert np.array_equal(gr[np.arange(500), ind], np.sign(pts[np.arange(500), ind])) and np.abs(gr).sum() == 500
    dom = cpp.create_unfitted_impl_domain(g, cpp.create_cart_grid([np.linspace(0.1, 1.9, 10)] * 3))
    q = cpp.create_quadrature(dom, dom.get_cut_cells(), 3)
    vol = (len(dom.get_full_cells()) + q.weights.sum()) * 0.2 ** 3
    assert abs(vol - 0.9 ** 3) < 1e-12


# ---- edge cases ------------------------------------------------------------------------------------------------------
def test_nonuniform_grid_negative_function_and_duplicates():
    """Graded (non-uniform) breaks, the negated level set, a cell list with repeats in arbitrary order."""
    cpp = _cpp()
    rng = np.random.default_rng(9)
    breaks = []
    for n in (9, 7, 11):
        b = np.sort(np.concatenate([[0.0, 1.0], rng.random(n - 1)]))
        breaks.append(b)
    per = [1.3, 0.8, 1.1]
    f = cpp.create_negative(cpp.create_SchwarzDiamond(per))
    dom = cpp.create_unfitted_impl_domain(f, cpp.create_cart_grid(breaks))
    od = O.OracleDomain(3, O.K_SCHWARZ_D, per, breaks, negative=True)
    assert np.array_equal(od.status, dom._cell_status())
    rc, rs = dom._facet_records()
    assert np.array_equal(od.facet_cells, rc) and np.array_equal(od.facet_status, rs)
    cut = dom.get_cut_cells()
    cells = np.concatenate([cut[::-1], cut[:5], cut[:5]])
    for p in (1, 4):
        q = cpp.create_quadrature(dom, cells, p)
        ro = od.quad_cells(cells, p)
        assert np.array_equal(ro.n_per, q.n_pts_per_entity)
        assert np.array_equal(ro.points, q.points) and np.array_equal(ro.weights, q.weights)
        b = cpp.create_unfitted_bound_quadrature(dom, cells, p, True, True)
        rb = od.quad_unf_bdry(cells, p, True, True)
        assert np.array_equal(rb.n_per, b.n_pts_per_entity)
        assert np.array_equal(rb.points, b.points) and np.array_equal(rb.weights, b.weights) and np.array_equal(rb.normals, b.normals)


@pytest.mark.parametrize("p", [16, 33, 100])
def test_many_points_per_direction(p):
    """Large rules: the K3 tile holds fewer groups than threads (p = 100 is the table's maximum, quadrature.hpp:47)."""
    cpp = _cpp()
    n = 3
    breaks = [O.cpp_breaks(n)] * 3
    per = [0.6, 0.5, 0.4]
    dom = cpp.create_unfitted_impl_domain(cpp.create_Schoen(per), cpp.create_cart_grid(breaks))
    od = O.OracleDomain(3, O.K_SCHOEN, per, breaks)
    cut = dom.get_cut_cells()[:3 if p == 100 else 8]
    assert len(cut) > 0
    q = cpp.create_quadrature(dom, cut, p)
    ro = od.quad_cells(cut, p)
    assert np.array_equal(ro.n_per, q.n_pts_per_entity)
    assert np.array_equal(ro.points, q.points) and np.array_equal(ro.weights, q.weights)
    with pytest.raises(ValueError):
        cpp.create_quadrature(dom, cut, 101)


def test_domain_without_cut_cells():
    cpp = _cpp()
    grid = cpp.create_cart_grid([np.linspace(0, 1, 9)] * 3)
    inside = cpp.create_unfitted_impl_domain(cpp.create_sphere(5.0, np.array([.5, .5, .5]), False), grid)     # cube inside the sphere
    assert len(inside.get_cut_cells()) == 0 and len(inside.get_full_cells()) == 512 and len(inside.get_empty_cells()) == 0
    outside = cpp.create_unfitted_impl_domain(cpp.create_sphere(0.2, np.array([3., 3., 3.]), False), grid)
    assert len(outside.get_empty_cells()) == 512
    for dom in (inside, outside):
        cells = np.arange(10, dtype=np.int64)
        q = cpp.create_quadrature(dom, cells, 3)
        assert q.n_pts_per_entity.tolist() == ([27] * 10 if dom is inside else [0] * 10)
        b = cpp.create_unfitted_bound_quadrature(dom, cells, 3, True, True)
        assert b.n_pts_per_entity.sum() == 0 and b.normals.shape == (0, 3)
        assert not dom.has_facets_with_unf_bdry


def test_device_resident_rules_are_zero_copy_views():
    """create_quadrature_device / create_unfitted_bound_quadrature_device: the arrays stay in HBM and are wrapped by torch
    through __cuda_array_interface__; their contents equal the host rules."""
    import torch
    cpp = _cpp()
    n, p = 12, 4
    dom = cpp.create_unfitted_impl_domain(cpp.create_Schoen([1.9, 2.1, 2.3]), cpp.create_cart_grid([np.linspace(0, 1, n + 1)] * 3))
    cut = dom.get_cut_cells()
    hq = cpp.create_quadrature(dom, cut, p)
    dq = cpp.create_quadrature_device(dom, cut, p)
    pts = torch.as_tensor(dq.points, device="cuda")
    wts = torch.as_tensor(dq.weights, device="cuda")
    npe = torch.as_tensor(dq.n_pts_per_entity, device="cuda")
    assert pts.data_ptr() == dq.points.__cuda_array_interface__["data"][0]          # no copy
    assert pts.shape == (len(hq.weights), 3) and dq.normals is None
    assert np.array_equal(pts.cpu().numpy(), hq.points) and np.array_equal(wts.cpu().numpy(), hq.weights)
    assert np.array_equal(npe.cpu().numpy(), hq.n_pts_per_entity)
    hb = cpp.create_unfitted_bound_quadrature(dom, cut, p, False, False)
    db = cpp.create_unfitted_bound_quadrature_device(dom, cut, p)
    assert np.array_equal(torch.as_tensor(db.normals, device="cuda").cpu().numpy(), hb.normals)
    assert np.array_equal(torch.as_tensor(db.weights, device="cuda").cpu().numpy(), hb.weights)
    # a consumer on the GPU: the cut volume without any transfer of nodes
    vol = float(wts.sum()) / n ** 3 + len(dom.get_full_cells()) / n ** 3
    assert abs(vol - (hq.weights.sum() + len(dom.get_full_cells())) / n ** 3) < 1e-14
