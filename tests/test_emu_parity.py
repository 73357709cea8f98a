"""CPU check of the kernel bodies: tests/emu runs the per-thread pass bodies of qugar_b200/csrc (the code the
CUDA kernels execute) thread by thread on the host and must reproduce the oracle bit for bit.  This is a
development aid for a container without a GPU; the gpu-marked tests are the parity tests proper."""
import numpy as np
import pytest

import emu_lib as E
import oracle_lib as O

CASES = [
    (3, O.K_SCHOEN, [2., 2., 2.], 16, 3),
    (2, O.K_SCHOEN, [1., 1., 0.], 4, 3),
    (3, O.K_SCHOEN, [1.9, 2.1, 2.3], 20, 4),
    (3, O.K_SPHERE, [0.4, .5, .5, .5], 12, 5),
    (2, O.K_SPHERE, [0.4, .5, .5], 16, 4),
    (3, O.K_SCHWARZ_D, [1.3, 0.7, 1.1], 16, 2),
    (3, O.K_SCHOEN_IWP, [1.1, 0.9, 1.0], 10, 3),
    (3, O.K_SCHOEN_IWP, [2.0, 1.0, 0.5], 8, 3),      # power-of-two periods: value- and gradient-style arguments coincide
    (3, O.K_SCHOEN, [1.0, 4.0, 1.7], 10, 3),
    (3, O.K_SCHOEN_FRD, [0.9, 1.0, 1.1], 10, 2),
    (3, O.K_FISCHER_KOCH_S, [1., .8, 1.2], 10, 3),
    (2, O.K_SCHWARZ_P, [1.5, 1.0, 0.1], 12, 6),
    (2, O.K_PLANE, [0.5, 0.5, 2.0, 1.0], 6, 2),
    (3, O.K_CYLINDER, [0.3, .5, .45, .5, 0.6, 0.0, 0.8], 10, 3),
    (3, O.K_ELLIPSOID, [0.4, 0.3, 0.2, .5, .5, .5, 0.8, 0.6, 0., -0.6, 0.8, 0., 0., 0., 1.], 10, 3),
    (2, O.K_ELLIPSOID, [0.4, 0.25, .5, .5, 0.8, 0.6, -0.6, 0.8], 16, 4),
    (2, O.K_ANNULUS, [0.2, 0.42, .5, .5], 16, 3),
    (3, O.K_TORUS, [0.3, 0.12, .5, .5, .5, 0.0, 0.6, 0.8], 12, 3),
    (3, O.K_DIM_LINEAR, [-0.3, 0.2, 0.1, 0.5, -0.4, 0.3, -0.1, 0.6], 6, 3),
    (2, O.K_DIM_LINEAR, [-0.3, 0.2, 0.4, -0.5], 8, 4),
    (3, O.K_COMPOSITE, "transformed_sphere", 10, 3),
    (3, O.K_COMPOSITE, "schoen_minus_constant", 8, 3),
    (2, O.K_COMPOSITE, "disk_plus_transformed_disk", 16, 3),
    (2, O.K_COMPOSITE, "rotated_square", 12, 3),
    (3, O.K_COMPOSITE, "rotated_cube", 8, 2),
    (2, O.K_COMPOSITE, "nested_2d", 16, 3),
    (3, O.K_COMPOSITE, "nested_3d", 8, 3),
    (3, O.K_BEZIER, "sphere_bzr", 8, 3),
    (2, O.K_BEZIER, "disk_bzr", 16, 4),
    (3, O.K_BEZIER, "random_deg3", 6, 3),
]


def composite_func(name):
    """The composite test level sets, built with the Python mirror (whose parameter encoding the C ABI, the oracle and the
    emulation harness all decode)."""
    from qugar_b200 import cpp
    if name == "transformed_sphere":       # an oblique ellipsoid: unit sphere under y = B^t S x + o
        t = cpp.create_affine_transformation([.5, .45, .5], [1., 1., 0.], [0., 1., 0.2], 0.35, 0.2, 0.28)
        return cpp.create_affinely_transformed_function(cpp.create_sphere(1.0, np.zeros(3), False), t)
    if name == "schoen_minus_constant":    # a thicker gyroid phase
        return cpp.create_functions_subtraction(cpp.create_Schoen([1.1, 0.9, 1.3]), cpp.create_constant_3D(0.3, False))
    if name == "disk_plus_transformed_disk":
        t = cpp.create_affine_transformation([.6, .55], [2., 1.], 0.5, 0.25)
        moved = cpp.create_affinely_transformed_function(cpp.create_negative(cpp.create_disk(1.0, np.zeros(2), False)), t)
        return cpp.create_functions_addition(cpp.create_disk(0.42, np.array([.45, .5]), False), moved)
    if name == "rotated_square":           # Square<2> under a rotation + anisotropic scaling: half-sides 0.3 x 0.2
        return cpp.create_square(cpp.create_affine_transformation([.5, .45], [2., 1.], 0.3, 0.2))
    if name == "rotated_cube":             # Square<3>: an oblique box with half-sides 0.3 x 0.2 x 0.25
        return cpp.create_square(cpp.create_affine_transformation([.5, .45, .5], [1., 1., 0.], [0., 1., 0.2], 0.3, 0.2, 0.25))
    if name == "nested_2d":
        # Subtract(Transformed(Add(disk, Negative(Transformed(disk))), T1), constant): three levels of nesting
        t0 = cpp.create_affine_transformation([.1, .05], [2., 1.], 0.5, 0.3)
        inner = cpp.create_functions_addition(
            cpp.create_disk(0.42, np.array([.0, .05]), False),
            cpp.create_negative(cpp.create_affinely_transformed_function(cpp.create_disk(1.0, np.zeros(2), False), t0)))
        t1 = cpp.create_affine_transformation([.5, .5], [1., .3], 0.9, 1.1)
        return cpp.create_functions_subtraction(cpp.create_affinely_transformed_function(inner, t1), cpp.create_constant_2D(0.02, False))
    if name == "nested_3d":
        # Add(Transformed(Subtract(Schoen, constant), T), Negative(Transformed(Negative(sphere), T2)))
        t = cpp.create_affine_transformation([.05, .0, .1], [1., .2, 0.], [0., 1., .1], 1.1, 0.9, 1.0)
        thick = cpp.create_affinely_transformed_function(
            cpp.create_functions_subtraction(cpp.create_Schoen([1.1, 0.9, 1.3]), cpp.create_constant_3D(0.3, False)), t)
        t2 = cpp.create_affine_transformation([.5, .5, .5], [1., 0., 0.], [0., 1., 0.], 0.45, 0.4, 0.35)
        ball = cpp.create_negative(cpp.create_affinely_transformed_function(
            cpp.create_negative(cpp.create_sphere(1.0, np.zeros(3), False)), t2))
        return cpp.create_functions_addition(thick, ball)
    raise KeyError(name)


def bezier_params(dim, name):
    """order[dim] + Bernstein coefficients (direction 0 fastest) of the Bezier test level sets."""
    from qugar_b200 import cpp
    if name in ("sphere_bzr", "disk_bzr"):
        order = [3] * dim
        mon = np.zeros(3 ** dim)
        c, r = [0.5] * dim, 0.4
        mon[0] = -r * r
        for d in range(dim):
            mon[0] += c[d] * c[d]
            mon[3 ** d] = -2.0 * c[d]
            mon[2 * 3 ** d] = 1.0
        coefs = cpp._monomials_to_bezier(order, mon)
    else:
        order = [4] * dim
        coefs = 2.0 * np.random.default_rng(20260101).random(4 ** dim) - 1.0
    return np.concatenate([np.asarray(order, np.float64), coefs])


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}d-k{c[1]}-n{c[3]}-p{c[4]}-{c[2] if isinstance(c[2], str) else ''}" for c in CASES])
def test_pipeline_bodies_match_oracle(case):
    dim, kind, params, ng, p = case
    if kind == O.K_BEZIER:
        params = bezier_params(dim, params)
    if kind == O.K_COMPOSITE:
        params = composite_func(params)._params
    br = [np.linspace(0, 1, ng + 1)] * dim
    d = O.OracleDomain(dim, kind, params, br)
    cut = d.cut_cells
    assert len(cut) > 0
    r = d.quad_cells(cut, p)
    e = E.run(dim, kind, params, br, cut, E.JOB_VOLUME, p, E.SINK_CELL)
    assert e["err"] == 0
    assert np.array_equal(r.n_per, e["n_per"])
    assert np.array_equal(r.points, e["points"])
    assert np.array_equal(r.weights, e["weights"])
    r = d.quad_unf_bdry(cut, p, False, False)
    e = E.run(dim, kind, params, br, cut, E.JOB_SURFACE, p, E.SINK_SURF)
    # the oracle purges facet-coincident nodes; on these grids none occur except in aligned cases
    if np.array_equal(r.n_per, e["n_per"]):
        assert np.array_equal(r.points, e["points"])
        assert np.array_equal(r.weights, e["weights"])
        assert np.array_equal(r.normals, e["normals"])
    else:
        assert (r.n_per <= e["n_per"]).all()


def test_facet_rules_match_oracle():
    dim, kind, params, ng, p = 3, O.K_SCHOEN, [1.3, 0.9, 1.1], 8, 3
    br = [np.linspace(0, 1, ng + 1)] * dim
    d = O.OracleDomain(dim, kind, params, br)
    c, f = d.facets_where(lambda s: s == O.F_CUT)
    # exterior cut facets: the oracle's exterior-integral rule on them is the plain facet rule
    n = np.array(d.n_cells_dir)
    t = np.stack([(c // int(np.prod(n[:k]))) % n[k] for k in range(dim)], axis=1)
    cd, side = f // 2, f % 2
    ext = t[np.arange(len(c)), cd] == np.where(side == 0, 0, n[cd] - 1)
    c, f = c[ext], f[ext]
    assert len(c) > 10
    r = d.quad_facets(c, f, p, True)
    e = E.run(dim, kind, params, br, c, f, p, E.SINK_FACET)
    assert np.array_equal(r.n_per, e["n_per"])
    assert np.array_equal(r.points, e["points"])
    assert np.array_equal(r.weights, e["weights"])


def test_det_sincos_same_bits_in_oracle_and_kernels():
    import ctypes as C
    lib = O.load(False)
    emu = E.load()
    rng = np.random.default_rng(3)
    for x in np.concatenate([rng.uniform(-100, 100, 2000), np.arange(-16, 17) * (np.pi / 8)]):
        s1, c1, s2, c2 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        lib.orc_sincos(float(x), C.byref(s1), C.byref(c1))
        emu.emu_sincos(C.c_double(float(x)), C.byref(s2), C.byref(c2))
        assert s1.value == s2.value and c1.value == c2.value


# ---- reparameterization bodies (qugar_b200/csrc/qg_reparam.cuh) on the CPU against the oracle, before the merge -------------
import test_oracle_reparam as TR


@pytest.mark.parametrize("case", range(len(TR.BOX_CASES)))
def test_emu_reparam_box(case):
    (dim, kind, params), lo, hi, levelset, facet, gold, cen = TR.BOX_CASES[case]
    mode = 1 if levelset else (2 + facet if facet >= 0 else 0)
    breaks = [np.array([lo[d], hi[d]], float) for d in range(dim)]
    for order in (4, 7):
        o = TR.reparam_box(dim, kind, params, lo, hi, order, levelset, facet, merge_tol=-1.0, libm=False)
        pts, conn, wires, err = E.reparam(dim, kind, params, breaks, [0], [], order, mode)
        assert err == 0
        assert pts.shape == o[0].shape and np.array_equal(pts, o[0])
        assert np.array_equal(conn, o[1]) and np.array_equal(wires, o[2])


@pytest.mark.parametrize("case", [
    (3, O.K_SCHOEN, [2., 2., 2.], 5, False), (3, O.K_SCHOEN, [2., 2., 2.], 5, True), (3, O.K_SCHWARZ_D, [2., 2., 2.], 4, True),
    (2, O.K_SCHOEN, [1., 1., 0.], 4, False), (2, O.K_SCHOEN, [1., 1., 0.], 4, True), (3, O.K_SPHERE, [0.4, .5, .5, .5], 4, False)])
def test_emu_reparam_grid(case):
    dim, kind, params, n, levelset = case
    breaks = [O.cpp_breaks(n)] * dim
    dom = O.OracleDomain(dim, kind, params, breaks, libm=False)
    cells = np.arange(n ** dim, dtype=np.int64)
    lib = TR._lib(False)
    o = TR.take_mesh(lib, lib.orc_reparam_domain(dom.h, cells, len(cells), 4, int(levelset), 0, 0), dim)
    pts, conn, wires, err = E.reparam(dim, kind, params, breaks, dom.cut_cells, [] if levelset else dom.full_cells, 4,
                                      1 if levelset else 0)
    assert err == 0
    assert pts.shape == o[0].shape and np.array_equal(pts, o[0])
    assert np.array_equal(conn, o[1]) and np.array_equal(wires, o[2])
