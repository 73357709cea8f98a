"""GPU parity of the reparameterization: the CUDA path through the C ABI (qugar_b200.cpp -> qg_reparam) against
the oracle (oracle/oracle_reparam.hpp) bit for bit, and against the golden sizes / centroids of the reference's own tests
(cpp/test/test_impl_general_reparam_{0,1}.cpp) directly."""
import numpy as np
import pytest

import oracle_lib as O
import test_oracle_reparam as T

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def _cpp():
    from qugar_b200 import cpp
    if cpp.device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu tests must run the CUDA path")
    return cpp


def _func(cpp, dim, kind, params):
    if kind == O.K_SPHERE:
        return (cpp.create_disk if dim == 2 else cpp.create_sphere)(params[0], np.array(params[1:]), False)
    if kind == O.K_PLANE:
        return (cpp.create_line if dim == 2 else cpp.create_plane)(np.array(params[:dim]), np.array(params[dim:]), False)
    fac = {O.K_SCHOEN: "create_Schoen", O.K_SCHWARZ_D: "create_SchwarzDiamond", O.K_FISCHER_KOCH_S: "create_FischerKochS",
           O.K_SCHOEN_IWP: "create_SchoenIWP"}[kind]
    return getattr(cpp, fac)(params[:2], params[2]) if dim == 2 else getattr(cpp, fac)(params)


def _same(mesh, oracle_mesh):
    pts, conn, wires = oracle_mesh[:3]
    assert mesh.points.shape == pts.shape
    assert np.array_equal(mesh.points, pts)
    assert np.array_equal(mesh.cells_conn.reshape(-1), conn)
    assert np.array_equal(mesh.wirebasket_conn.reshape(-1), wires)


@pytest.mark.parametrize("case", range(len(T.BOX_CASES)))
def test_reparam_box_matches_oracle_and_goldens(case):
    cpp = _cpp()
    (dim, kind, params), lo, hi, levelset, facet, gold, cen = T.BOX_CASES[case]
    f = _func(cpp, dim, kind, params)
    for order in (4, 3, 8):
        # before the merge: points in insertion order, connectivity, wirebasket -- bit for bit
        m = cpp.reparam_general(f, lo, hi, order, levelset, None if facet < 0 else facet)
        o = T.reparam_box(dim, kind, params, lo, hi, order, levelset, facet, merge_tol=-1.0, libm=False)
        _same(m, o)
        assert (m.dim, m.range, m.order) == (dim - 1 if (levelset or facet >= 0) else dim, dim, order)
        # merged (stable sorts on both sides)
        m = cpp.reparam_general(f, lo, hi, order, levelset, None if facet < 0 else facet, merge_tol=2.0e4 * EPS)
        o = T.reparam_box(dim, kind, params, lo, hi, order, levelset, facet, merge_tol=2.0e4 * EPS, stable=1, libm=False)
        _same(m, o)
        if order == 4:
            # the reference's golden sizes and centroid (test_impl_general_reparam_0.cpp)
            got, c = T.summary(m.points, m.cells_conn.reshape(-1), m.wirebasket_conn.reshape(-1))
            assert got[:3] == tuple(gold)[:3]
            if gold[0]:
                assert np.abs(c - np.asarray(cen)).max() <= 101 * EPS


GRIDS = [(c[0], c[1], c[2], c[3], c[4], c[5], c[6]) for c in T.GRID_CASES] + [
    ((3, O.K_SCHOEN, [2., 2., 2.]), (0, 0, 0), (1, 1, 1), 5, False, (275722, 470720, 18928),
     (0.499995160198210664, 0.49994071822141356, 0.500341156991491731)),
    ((3, O.K_SCHOEN, [2., 2., 2.]), (0, 0, 0), (1, 1, 1), 5, True, (41450, 61168, 11712),
     (0.502672595808896783, 0.504815611380007412, 0.503249213105162796)),
    ((3, O.K_SCHOEN, [1.9, 2.1, 2.3]), (0, 0, 0), (1, 1, 1), 6, False, None, None),
    ((3, O.K_SCHWARZ_D, [2., 2., 2.]), (0, 0, 0), (1, 1, 1), 4, True, None, None),
    ((2, O.K_SCHOEN, [1., 1., 0.]), (0, 0), (1, 1), 4, False, None, None),
    ((2, O.K_SCHOEN, [1., 1., 0.]), (0, 0), (1, 1), 4, True, None, None),
    ((3, O.K_FISCHER_KOCH_S, [2., 2., 2.]), (0, 0, 0), (1, 1, 1), 2, False, None, None),
]


@pytest.mark.parametrize("case", range(len(GRIDS)))
def test_reparam_grid_matches_oracle_and_goldens(case):
    cpp = _cpp()
    (dim, kind, params), lo, hi, n, levelset, gold, cen = GRIDS[case]
    order = 3 if kind == O.K_FISCHER_KOCH_S else 4
    breaks = [O.cpp_breaks(n, float(lo[d]), float(hi[d])) for d in range(dim)]
    dom = cpp.create_unfitted_impl_domain(_func(cpp, dim, kind, params), cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, kind, params, breaks, libm=False)
    assert np.array_equal(od.status, dom._cell_status())
    fac = cpp.create_reparameterization_levelset if levelset else cpp.create_reparameterization
    cells = np.arange(n ** dim, dtype=np.int64)
    lib = T._lib(False)
    for merge in (False, True):
        m = fac(dom, cells[::-1].copy(), order, merge)      # the cell list is sorted by the call (impl_reparam.cpp:66-67)
        o = T.take_mesh(lib, lib.orc_reparam_domain(od.h, cells, len(cells), order, int(levelset), int(merge), 1), dim)
        _same(m, o)
    m2 = fac(dom, order, True)                                # overload without cells = every cell
    assert np.array_equal(m2.points, m.points) and np.array_equal(m2.cells_conn, m.cells_conn)
    if gold is not None:
        got, c = T.summary(m.points, m.cells_conn.reshape(-1), m.wirebasket_conn.reshape(-1))
        if kind in (O.K_SPHERE, O.K_PLANE):
            assert got[:3] == tuple(gold)[:3]
            assert np.abs(c - np.asarray(cen)).max() <= 101 * EPS
        else:
            # trigonometric level sets: the reference calls glibc's sin / cos, the device its own < 1.5 ulp sincos; roots move by
            # an ulp and a handful of near-coincident points merge differently (275 717 instead of 275 722 points on the 5^3
            # gyroid).  The golden sizes are asserted exactly on the glibc build of the oracle (tests/test_oracle_reparam.py).
            assert got[1:3] == tuple(gold)[1:3]
            assert abs(got[0] - gold[0]) <= 1e-4 * gold[0]
            assert np.abs(c - np.asarray(cen)).max() <= 1e-4


def test_reparam_subset_of_cells_and_errors():
    cpp = _cpp()
    n = 6
    breaks = [O.cpp_breaks(n)] * 3
    per = [1.3, 0.9, 1.1]
    dom = cpp.create_unfitted_impl_domain(cpp.create_Schoen(per), cpp.create_cart_grid(breaks))
    od = O.OracleDomain(3, O.K_SCHOEN, per, breaks, libm=False)
    rng = np.random.default_rng(3)
    cells = rng.permutation(n ** 3)[:70].astype(np.int64)
    lib = T._lib(False)
    for levelset in (False, True):
        fac = cpp.create_reparameterization_levelset if levelset else cpp.create_reparameterization
        m = fac(dom, cells, 5, False)
        o = T.take_mesh(lib, lib.orc_reparam_domain(od.h, cells, len(cells), 5, int(levelset), 0, 1), 3)
        _same(m, o)
    with pytest.raises(ValueError):
        cpp.create_reparameterization(dom, np.array([n ** 3], np.int64), 4, False)
    with pytest.raises(ValueError):
        cpp.create_reparameterization(dom, cells, 1, False)


@pytest.mark.parametrize("name", ["sphere_bzr", "disk_bzr", "random_deg3"])
def test_reparam_bezier_level_sets_by_the_general_algorithm(name):
    """Bezier level sets (the reference's use_bzr=True default) are reparameterized by the general algorithm, with a warning
    (upstream: reparam_Bezier, not built): bit-identical to the oracle's reparam_general on the same polynomial."""
    from test_gpu_parity import _bezier_case
    cpp = _cpp()
    dim, func, params, n, _ = _bezier_case(name)
    n = min(n, 8)
    breaks = [O.cpp_breaks(n)] * dim
    dom = cpp.create_unfitted_impl_domain(func, cpp.create_cart_grid(breaks))
    od = O.OracleDomain(dim, O.K_BEZIER, params, breaks, libm=False)
    assert np.array_equal(od.status, dom._cell_status())
    cells = np.arange(n ** dim, dtype=np.int64)
    lib = T._lib(False)
    for levelset in (False, True):
        fac = cpp.create_reparameterization_levelset if levelset else cpp.create_reparameterization
        for merge in (False, True):
            with pytest.warns(UserWarning):
                m = fac(dom, cells, 4, merge)
            o = T.take_mesh(lib, lib.orc_reparam_domain(od.h, cells, len(cells), 4, int(levelset), int(merge), 1), dim)
            _same(m, o)


def test_reparam_mesh_device_views():
    """The mesh stays in HBM: device views (__cuda_array_interface__) without a download, equal to the host arrays."""
    import torch
    cpp = _cpp()
    n = 6
    dom = cpp.create_unfitted_impl_domain(cpp.create_Schoen([1.3, 0.9, 1.1]), cpp.create_cart_grid([O.cpp_breaks(n)] * 3))
    m = cpp.create_reparameterization_levelset(dom, 4, True)
    assert m._host is None and m.num_points > 0 and m.num_cells > 0
    dp = torch.as_tensor(m.device_points, device="cuda")
    dc = torch.as_tensor(m.device_cells_conn, device="cuda")
    dw = torch.as_tensor(m.device_wirebasket_conn, device="cuda")
    assert m._host is None                       # nothing was downloaded for the device views
    assert np.array_equal(dp.cpu().numpy(), m.points) and np.array_equal(dc.cpu().numpy(), m.cells_conn)
    assert np.array_equal(dw.cpu().numpy(), m.wirebasket_conn)
    assert dc.max().item() < m.num_points
