"""CPU tests of the host-side logic of the Python mirror (no compute calls): monomial -> Bernstein conversion, reference
systems, factory argument handling, index iteration -- the parts of qugar_b200/cpp.py that restate reference host code."""
from math import comb

import numpy as np
import pytest


def _bernstein_eval(order, coefs, x):
    dim = len(order)
    v = 0.0
    for flat in range(int(np.prod(order))):
        idx, r = [], flat
        for d in range(dim):
            idx.append(r % order[d])
            r //= order[d]
        b = coefs[flat]
        for d in range(dim):
            n = order[d] - 1
            b *= comb(n, idx[d]) * x[d] ** idx[d] * (1.0 - x[d]) ** (n - idx[d])
        v += b
    return v


def test_tensor_index_iteration_is_direction0_fastest():
    from qugar_b200 import cpp
    assert list(cpp._tensor_indices([0, 0], [2, 3])) == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]
    assert list(cpp._tensor_indices([1, 2], [3, 3])) == [(1, 2), (2, 2)]
    assert list(cpp._tensor_indices([0], [0])) == []


def test_monomials_to_bezier_reproduces_the_polynomial():
    """MonomialsTP::transform_coefs_to_Bezier (monomials_tp.cpp:295-331)."""
    from qugar_b200 import cpp
    rng = np.random.default_rng(0)
    for order in ([3, 3], [2, 4], [3, 3, 3], [4, 2, 3]):
        mon = rng.standard_normal(int(np.prod(order)))
        bz = cpp._monomials_to_bezier(order, mon)
        for x in rng.random((20, len(order))):
            ref, flat = 0.0, 0
            for tid in cpp._tensor_indices([0] * len(order), order):
                ref += mon[flat] * np.prod([x[d] ** tid[d] for d in range(len(order))])
                flat += 1
            assert abs(_bernstein_eval(order, bz, x) - ref) < 1e-12


def test_sphere_and_plane_bezier_coefficients():
    from qugar_b200 import cpp
    f = cpp.create_sphere(0.4, np.array([0.5, 0.25, 0.75]))          # the reference's default: SphereBzr
    assert type(f).__name__ == "SphereBzr" and f.order == (3, 3, 3)
    rng = np.random.default_rng(1)
    for x in rng.random((50, 3)):
        assert abs(_bernstein_eval([3, 3, 3], f.coefs, x) - (((x - [0.5, 0.25, 0.75]) ** 2).sum() - 0.16)) < 1e-14
    d = cpp.create_disk(0.3, np.array([0.2, 0.6]), True)
    for x in rng.random((50, 2)):
        assert abs(_bernstein_eval([3, 3], d.coefs, x) - (((x - [0.2, 0.6]) ** 2).sum() - 0.09)) < 1e-14
    pl = cpp.create_plane([0.1, 0.2, 0.3], [3.0, 0.0, 4.0])          # PlaneBzr normalises the normal
    for x in rng.random((50, 3)):
        assert abs(_bernstein_eval([2, 2, 2], pl.coefs, x) - ((x - [0.1, 0.2, 0.3]) * [0.6, 0.0, 0.8]).sum()) < 1e-14
    ln = cpp.create_line([0.5, 0.5], [0.0, 2.0], True)
    assert abs(_bernstein_eval([2, 2], ln.coefs, [0.3, 0.9]) - 0.4) < 1e-15
    # bool in the place of the first optional argument means use_bzr (nanobind overloads create_sphere(radius, use_bzr))
    assert type(cpp.create_sphere(0.3, False)).__name__ == "Sphere"
    with pytest.raises(ValueError):
        cpp.create_sphere(-1.0)
    with pytest.raises(ValueError):
        cpp.create_bezier([3, 3], np.zeros(8))


def test_reference_systems_are_orthonormal():
    """create_orthonormal_system (ref_system.cpp:28-138)."""
    from qugar_b200 import cpp
    rs = cpp.create_ref_system([0.1, 0.2], [3.0, 4.0])
    assert np.allclose(rs.basis, [[0.6, 0.8], [-0.8, 0.6]])
    for args in (([0, 0, 0], [1.0, 2.0, 2.0], [0.0, 1.0, 0.0]), ([1, 1, 1], [0.0, 0.0, 5.0]), ([0, 0, 0], [1.0, 1e-3, 0.0])):
        b = cpp.create_ref_system(*args).basis
        assert np.allclose(b @ b.T, np.eye(3), atol=1e-14)
        assert np.linalg.det(b) > 0.99                      # right-handed
    # (origin, axis_x, axis_y): x along axis_x, z = x cross y
    b = cpp.create_ref_system([0, 0, 0], [2.0, 0.0, 0.0], [1.0, 1.0, 0.0]).basis
    assert np.allclose(b, np.eye(3))
    # (origin, axis) in 3-D: the axis is the z-axis
    b = cpp.create_ref_system([0, 0, 0], [0.0, 0.0, 2.0]).basis
    assert np.allclose(b[2], [0, 0, 1])
    assert np.allclose(cpp.create_ref_system([0.5, 0.5, 0.5]).basis, np.eye(3))
    with pytest.raises(ValueError):
        cpp.create_ref_system([0, 0, 0], [1.0, 0.0, 0.0], [2.0, 0.0, 0.0])


def test_primitive_factories_validate_and_store_parameters():
    from qugar_b200 import cpp
    c = cpp.create_cylinder(0.2, [0.5, 0.5, 0.5], [0.0, 3.0, 4.0], False)
    assert np.allclose(c.axis, [0.0, 0.6, 0.8]) and c.radius == 0.2
    t = cpp.create_torus(0.3, 0.1, [0.5, 0.5, 0.5], False)
    assert np.allclose(t.axis, [0, 0, 1])
    e = cpp.create_ellipse([0.4, 0.2], False)
    assert np.allclose(e.ref_system.basis, np.eye(2)) and len(e._params) == 8
    with pytest.raises(ValueError):
        cpp.create_annulus(0.4, 0.2, [0.5, 0.5], False)     # inner >= outer
    with pytest.raises(ValueError):
        cpp.create_torus(0.1, 0.3, [0.5, 0.5, 0.5], False)  # minor >= major
    with pytest.raises(ValueError):
        cpp.create_ellipsoid([0.4, -0.1, 0.2], False)
    k = cpp.create_constant_3D(-2.5)                          # ConstantBzr: degree-0 Bernstein polynomial
    assert k.order == (1, 1, 1) and k.coefs[0] == -2.5
    n = cpp.create_negative(cpp.create_negative(c))
    assert not n._negative and np.array_equal(n._params, c._params)


def test_oracle_bezier_evaluation_matches_the_bernstein_definition():
    """cpp/test/test_impl_bezier_tp_0.cpp ("Testing Beziers evaluation"): de Casteljau value and gradient against the
    Bernstein-basis definition, random orders 1..5 per direction (the kernels are then checked bit for bit against
    this oracle in test_emu_parity / test_gpu_parity)."""
    import oracle_lib as O
    rng = np.random.default_rng(42)
    for dim in (2, 3):
        for _ in range(6):
            order = [int(v) for v in rng.integers(1, 6, dim)]
            coefs = rng.standard_normal(int(np.prod(order)))
            params = np.concatenate([np.asarray(order, float), coefs])
            pts = rng.random((15, dim))
            v, g = O.func_eval(dim, O.K_BEZIER, params, pts)
            for x, vx, gx in zip(pts, v, g):
                assert abs(vx - _bernstein_eval(order, coefs, x)) < 1e-12
                for d in range(dim):
                    h = 1e-6
                    xp, xm = x.copy(), x.copy()
                    xp[d] += h
                    xm[d] -= h
                    fd = (_bernstein_eval(order, coefs, xp) - _bernstein_eval(order, coefs, xm)) / (2 * h)
                    assert abs(gx[d] - fd) < 1e-6 * max(1.0, abs(fd))


def test_oracle_square_is_the_max_norm_box():
    """Square<dim> (impl_funcs_lib.cpp:44-98): value max_d |x_d| - 1 and gradient sgn(x_ind) e_ind on doubles; an
    axis-aligned box whose faces cross cell interiors is integrated exactly (planar faces, corners on cell edges); the
    affine placement goes through the same composite leaf as TransformedFunction."""
    import oracle_lib as O
    from qugar_b200 import cpp
    rng = np.random.default_rng(5)
    for dim in (2, 3):
        pts = rng.uniform(-2, 2, (400, dim))
        v, g = O.func_eval(dim, O.K_SQUARE, [], pts)
        ind = np.abs(pts).argmax(axis=1)
        assert np.array_equal(v, np.abs(pts).max(axis=1) - 1.0)
        assert np.array_equal(g[np.arange(400), ind], np.sign(pts[np.arange(400), ind])) and np.abs(g).sum() == 400
    d = O.OracleDomain(3, O.K_SQUARE, [], [np.linspace(0.1, 1.9, 10)] * 3)
    r = d.quad_cells(d.cut_cells, 3)
    assert len(d.cut_cells) == 61
    assert abs((len(d.full_cells) + r.weights.sum()) * 0.2 ** 3 - 0.9 ** 3) < 1e-14
    b = d.quad_unf_bdry(d.cut_cells, 3, False, False)
    # three faces of 0.9 x 0.9 (uniform grid: area = w h^2); a cell holding an edge or the corner of the box sees only the face
    # of the branch taken at its centre (ties: direction 0 first), so part of the other faces is not integrated there
    area = b.weights.sum() * 0.2 ** 2
    assert 0.85 * 3 * 0.9 ** 2 < area < 3 * 0.9 ** 2
    assert set(map(tuple, np.unique(b.normals, axis=0))) == {(1., 0., 0.), (0., 1., 0.), (0., 0., 1.)}
    # a rotated box inside the grid: half-sides 0.3 x 0.2
    f = cpp.create_square(cpp.create_affine_transformation([.5, .45], [2., 1.], 0.3, 0.2))
    assert type(f).__name__ == "TransformedFunction_2D" and f.dim == 2
    n = 24
    d = O.OracleDomain(2, O.K_COMPOSITE, f._params, [np.linspace(0, 1, n + 1)] * 2)
    r = d.quad_cells(d.cut_cells, 4)
    assert abs((len(d.full_cells) + r.weights.sum()) / n ** 2 - 0.24) < 1e-3 * 0.24
    with pytest.raises(ValueError):
        cpp.create_square()
    assert cpp.create_square_3D().dim == 3 and cpp.create_square_2D().dim == 2


def test_rule_array_views_keep_their_owner_alive():
    """The reference binds rule arrays with reference_internal (python/nanobind_wrappers/cut_quad.cpp:49-78) and its
    own consumer drops the container at once (python/qugar/mesh/unfitted_domain.py:559-566): a view must keep the
    owner of the buffer alive, through slices and reshapes too."""
    import ctypes as C
    import gc
    import weakref
    from qugar_b200 import cpp

    class Owner:
        pass

    backing = np.arange(12, dtype=np.float64)
    own = Owner()
    alive = weakref.ref(own)
    a = cpp._owned_view(backing.ctypes.data, C.c_double, (4, 3), np.float64, own)
    assert not a.flags.writeable and a.shape == (4, 3) and a[1, 2] == 5.0
    col = a[:, 1]
    del own, a
    gc.collect()
    assert alive() is not None, "a slice of the view must still own the rule"
    del col
    gc.collect()
    assert alive() is None
    # empty views never touch the address
    assert cpp._owned_view(0, C.c_double, (0, 3), np.float64, None).shape == (0, 3)


def test_custom_coeffs_restatement_layout():
    """The NumPy restatement of custom_coefficients.py (tests/ref_custom_coeffs.py, the oracle of the device packer):
    an FFCx-style reader that follows the offsets finds every entity's block (custom_coefficients.py:370-449, 730-775)."""
    from ref_custom_coeffs import pack_custom_coeffs_ref
    rng = np.random.default_rng(3)
    for dtype in (np.float64, np.float32):
        n_ent, n_old = 9, 5
        old = rng.random((n_ent, n_old)).astype(dtype)
        custom = np.array([1, 4, 6], np.int64)
        empty = np.array([0, 8], np.int64)
        n_per1 = np.array([2, 0, 3], np.int32)
        n_per2 = np.array([1, 2, 0], np.int32)
        q1 = (n_per1, rng.random((5, 3)), rng.random(5), None)
        q2 = (n_per2, rng.random((3, 3)), rng.random(3), rng.random((3, 3)))
        new = pack_custom_coeffs_ref(old, custom, empty, [q1, q2])
        n_extra = 1 if dtype == np.float64 else 2
        assert new.shape[1] == n_old + n_extra and new.dtype == dtype
        assert np.array_equal(new[:n_ent, :n_old], old)
        rel = new[:n_ent, -n_extra:].copy().view(np.intp).reshape(-1)
        assert list(rel[[0, 8]]) == [0, 0] and list(rel[[2, 3, 5, 7]]) == [-1] * 4
        flat = new.ravel()
        starts1 = np.concatenate([[0], np.cumsum(n_per1)])
        starts2 = np.concatenate([[0], np.cumsum(n_per2)])
        for k, row in enumerate(custom):
            pos = row * new.shape[1] + rel[row]                 # (const T*)w + offset, w = the entity's row
            n = int(flat[pos]); pos += 1
            assert n == n_per1[k]
            assert np.array_equal(flat[pos:pos + 3 * n], q1[1][starts1[k]:starts1[k] + n].ravel().astype(dtype)); pos += 3 * n
            assert np.array_equal(flat[pos:pos + n], q1[2][starts1[k]:starts1[k] + n].astype(dtype)); pos += n
            n = int(flat[pos]); pos += 1
            assert n == n_per2[k]
            assert np.array_equal(flat[pos:pos + 3 * n], q2[1][starts2[k]:starts2[k] + n].ravel().astype(dtype)); pos += 3 * n
            assert np.array_equal(flat[pos:pos + n], q2[2][starts2[k]:starts2[k] + n].astype(dtype)); pos += n
            assert np.array_equal(flat[pos:pos + 3 * n], q2[3][starts2[k]:starts2[k] + n].ravel().astype(dtype))


def test_reference_demo_names_resolve_under_the_shim():
    """Every `qugar.cpp.<name>` the reference's FEniCSx-free demo touches (python/demo/demo_div_thm_no_fenicsx.py) exists
    in the shim package (qugar_b200/shim/qugar), and so do the methods / properties it calls on grid, domain and rule
    objects.  Runs only where the reference tree is present (this container); the flow itself is restated for the GPU box in
    tests/test_gpu_round2.py::test_divergence_theorem_demo_flow."""
    import ast
    import os
    import sys
    demo = "/root/reference/python/demo/demo_div_thm_no_fenicsx.py"
    if not os.path.exists(demo):
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "qugar_b200", "shim"))
    try:
        import qugar
        import qugar.cpp as shim
    finally:
        sys.path.pop(0)
    assert qugar.cpp is shim
    names = set()
    for node in ast.walk(ast.parse(open(demo).read())):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Attribute) and node.value.attr == "cpp" \
                and isinstance(node.value.value, ast.Name) and node.value.value.id == "qugar":
            names.add(node.attr)
    assert {"create_sphere", "create_cart_grid", "create_unfitted_impl_domain", "create_quadrature",
            "create_unfitted_bound_quadrature", "create_facets_quadrature_exterior_integral", "create_Gauss_quad_01"} <= names
    missing = sorted(n for n in names if not hasattr(shim, n))
    assert not missing, missing
    # object-level API the demo uses
    from qugar_b200 import cpp
    for cls, attrs in ((cpp.CartGridTP_3D, ["get_cell_domain", "get_boundary_cells"]),
                       (cpp.BoundBox_3D, ["as_array", "volume", "length"]),
                       (cpp.UnfittedImplDomain_3D, ["get_cut_cells", "get_full_cells", "get_full_facets", "get_cut_facets"]),
                       (cpp.CutCellsQuad_3D, ["cells", "n_pts_per_entity", "points", "weights"]),
                       (cpp.CutUnfBoundsQuad_3D, ["normals"]), (cpp.CutIsoBoundsQuad_2D, ["cells", "points", "weights"])):
        for a in attrs:
            assert hasattr(cls, a), (cls.__name__, a)


def test_gauss_quad_01_unequal_point_counts():
    """create_Gauss_quad_01(n_pts_dir[dim]) (quadrature.cpp:199-277): direction 0 fastest, one count per direction."""
    from qugar_b200 import cpp
    q = cpp.create_Gauss_quad_01([2, 3, 4])
    assert q.points.shape == (24, 3) and abs(q.weights.sum() - 1.0) < 1e-15
    x0 = cpp.create_Gauss_quad_01([2]); x1 = cpp.create_Gauss_quad_01([3]); x2 = cpp.create_Gauss_quad_01([4])
    k = 0
    for c in range(4):
        for b in range(3):
            for a in range(2):
                assert q.points[k, 0] == x0.points[a, 0] and q.points[k, 1] == x1.points[b, 0] and q.points[k, 2] == x2.points[c, 0]
                assert q.weights[k] == x0.weights[a] * x1.weights[b] * x2.weights[c]
                k += 1
    # equal counts: the library's own tensor rule and the composed one agree bit for bit
    e = cpp.create_Gauss_quad_01([3, 3])
    assert np.array_equal(e.points[:, 0], np.tile(x1.points[:, 0], 3)) and np.array_equal(e.weights, np.tile(x1.weights, 3) * np.repeat(x1.weights, 3))
    with pytest.raises(ValueError):
        cpp.create_Gauss_quad_01([2, 0])


def test_reparameterization_argument_forms():
    """create_reparameterization(unf_domain[, cells], n_pts_dir, merge_points): both overloads of reparam.cpp:101-151, positional or
    by keyword (argument parsing only: no device call)."""
    from qugar_b200 import cpp
    f = cpp._reparam_args
    c = np.arange(3)
    assert f((4, True), {}, "x") == (None, 4, True)
    got = f((c, 4, False), {}, "x")
    assert got[0] is c and got[1:] == (4, False)
    assert f((), dict(n_pts_dir=5, merge_points=True), "x") == (None, 5, True)
    got = f((c,), dict(n_pts_dir=5, merge_points=True), "x")
    assert got[0] is c and got[1:] == (5, True)
    assert f((4,), dict(merge_points=False), "x") == (None, 4, False)
    got = f((), dict(cells=c, n_pts_dir=3, merge_points=False), "x")
    assert got[0] is c and got[1:] == (3, False)
    import pytest
    with pytest.raises(TypeError):
        f((4,), {}, "x")
    for name in ("ReparamMesh_1_2", "ReparamMesh_2_2", "ReparamMesh_2_3", "ReparamMesh_3_3", "create_reparameterization",
                 "create_reparameterization_levelset"):
        assert hasattr(cpp, name)
