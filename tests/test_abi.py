"""The C-ABI library loads and exports every symbol include/qugar_b200.h declares; argument validation that
needs no GPU behaves; compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from qugar_b200 import build
    build.build()
    return C.CDLL(os.path.join(ROOT, "qugar_b200", "libqugar_b200.so"))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "qugar_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(qg_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_reference_for_each_entry_point():
    hdr = open(os.path.join(ROOT, "include", "qugar_b200.h")).read()
    for key in ("cut_quadrature.cpp", "impl_quadrature.cpp", "impl_unfitted_domain.cpp", "cut_quad.cpp", "unf_domain.cpp"):
        assert key in hdr


def test_validation_without_gpu(lib):
    lib.qg_last_error.restype = C.c_char_p
    g = C.c_void_p()
    b = np.array([0.0, 0.5, 0.25])   # not increasing
    arr = (C.POINTER(C.c_double) * 2)(b.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)))
    nb = (C.c_int64 * 2)(3, 3)
    assert lib.qg_grid_create(2, arr, nb, C.byref(g)) == -1
    assert b"increasing" in lib.qg_last_error()
    f = C.c_void_p()
    p = np.array([-1.0, 0.0, 0.0])
    assert lib.qg_func_create(0, 2, p.ctypes.data_as(C.POINTER(C.c_double)), 3, 0, C.byref(f)) == -1   # radius <= 0
    assert lib.qg_func_create(99, 2, p.ctypes.data_as(C.POINTER(C.c_double)), 3, 0, C.byref(f)) == -5
    pts = np.zeros((27, 3))
    wts = np.zeros(27)
    assert lib.qg_gauss_01(3, 3, pts.ctypes.data_as(C.c_void_p), wts.ctypes.data_as(C.c_void_p)) == 0
    assert abs(wts.sum() - 1.0) < 1e-15
    assert np.allclose(pts[1], [0.5, pts[0, 1], pts[0, 2]])   # direction 0 fastest


def test_reparam_entry_points_validate_without_gpu(lib):
    lib.qg_last_error.restype = C.c_char_p
    out = C.c_void_p()
    lib.qg_reparam.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_void_p)]
    assert lib.qg_reparam(None, None, 0, 4, 0, 0, 0.0, C.byref(out)) == -1          # null domain
    lib.qg_mesh_view.argtypes = [C.c_void_p] + [C.c_void_p] * 9
    assert lib.qg_mesh_view(None, *([None] * 9)) == -1
    lib.qg_mesh_free.argtypes = [C.c_void_p]
    lib.qg_mesh_free(None)                                                         # a no-op, like free(NULL)
    hdr = open(os.path.join(ROOT, "include", "qugar_b200.h")).read()
    for key in ("impl_reparam.cpp", "impl_reparam_general.cpp", "reparam_mesh.cpp", "reparam.cpp"):
        assert key in hdr


def test_python_mirror_has_reference_names():
    from qugar_b200 import cpp
    for name in ("create_cart_grid", "create_bound_box", "create_sphere", "create_disk", "create_plane", "create_line",
                 "create_Schoen", "create_SchoenIWP", "create_SchoenFRD", "create_FischerKochS", "create_SchwarzPrimitive",
                 "create_SchwarzDiamond", "create_negative", "create_unfitted_impl_domain", "create_quadrature",
                 "create_unfitted_bound_quadrature", "create_facets_quadrature_interior_integral",
                 "create_facets_quadrature_exterior_integral", "create_Gauss_quad_01", "__version__", "__git_commit_hash__"):
        assert hasattr(cpp, name), name
    g = cpp.create_cart_grid([np.linspace(0, 1, 5), np.linspace(0, 2, 3)])
    assert g.dim == 2 and list(g.num_cells_dir) == [4, 2]
    assert np.allclose(g.get_cell_domain(5).as_array(), [[0.25, 0.5], [1.0, 2.0]])
    assert np.array_equal(g.get_boundary_cells(1), [3, 7])
    q = cpp.create_Gauss_quad_01([2, 2])
    assert q.points.shape == (4, 2) and abs(q.weights.sum() - 1) < 1e-15
    f = cpp.create_sphere(0.3)   # the reference's default: the Bezier variant
    assert type(f).__name__ == "SphereBzr" and f.order == (3, 3, 3) and len(f.coefs) == 27
    t = cpp.create_torus(0.3, 0.1)       # the reference's default: TorusBzr, a degree-4 polynomial (order 5 per direction)
    assert type(t).__name__ == "TorusBzr" and t.order == (5, 5, 5) and t.major_radius == 0.3
    assert type(cpp.create_torus(0.3, 0.1, None, None, False)).__name__ == "Torus"
    assert type(cpp.create_cylinder(0.2, [0.5, 0.5, 0.5])).__name__ == "CylinderBzr"
    assert type(cpp.create_ellipse([0.3, 0.2])).__name__ == "EllipseBzr" and type(cpp.create_annulus(0.1, 0.3)).__name__ == "AnnulusBzr"
    for name in ("create_bezier", "create_constant_2D", "create_constant_3D", "create_cylinder", "create_ellipsoid", "create_ellipse",
                 "create_annulus", "create_torus", "create_ref_system"):
        assert hasattr(cpp, name), name


def test_div_rcp_exact():
    """div_rcp(a, b, 1/b) (csrc/qg_math.cuh, compiled for the host by the emulation harness with the same fma) must equal
    a / b bit for bit on the operand ranges K3 produces: coordinates in a cell over the cell extent, weights over volumes."""
    import ctypes as C
    import emu_lib as E
    lib = E.load()
    lib.emu_div_rcp_mismatches.restype = C.c_longlong
    rng = np.random.default_rng(0)
    n = 4_000_000
    for scale in ("cell", "pow2", "volume", "signed"):
        if scale == "cell":
            b = np.full(n, 1.0 / 256)
            a = rng.random(n) * b
        elif scale == "pow2":
            b = (rng.random(n) + 0.5) * 2.0 ** -rng.integers(0, 30, n)
            a = rng.random(n) * b
        elif scale == "volume":
            b = rng.random(n) * rng.random(n) * rng.random(n) * 1e-6 + 1e-300
            a = rng.random(n) * b * rng.random(n)
        else:
            b = rng.random(n) + 1e-3
            a = (rng.random(n) * 2 - 1) * b * 4
        bad = lib.emu_div_rcp_mismatches(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_longlong(n))
        assert bad == 0, scale


def test_compute_fails_loudly_without_device():
    from qugar_b200 import cpp
    if cpp.device_count() > 0:
        pytest.skip("a GPU is present")
    g = cpp.create_cart_grid([np.linspace(0, 1, 5)] * 2)
    f = cpp.create_disk(0.3, np.array([0.5, 0.5]), False)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cpp.create_unfitted_impl_domain(f, g)


def test_function_kind_enumerators_agree_everywhere():
    """The level-set kind travels as a plain int through the C ABI: include/qugar_b200.h (qg_func_kind), the engine
    (qg_funcs.cuh FuncKind), the oracle (oracle_funcs.hpp FuncKind), tests/oracle_lib.py and qugar_b200/cpp.py must agree."""
    import os
    import re

    import oracle_lib as O
    from qugar_b200 import cpp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def enum_of(path, prefix):
        text = open(os.path.join(root, path)).read()
        return {m.group(1): int(m.group(2)) for m in re.finditer(prefix + r"([A-Z_0-9]+)\s*=\s*(\d+)", text)}

    hdr = enum_of("include/qugar_b200.h", "QG_")
    eng = enum_of("qugar_b200/csrc/qg_funcs.cuh", "QG_FUNC_")
    orc = enum_of("oracle/oracle_funcs.hpp", r"\bK_")
    alias = {"SCHWARZ_DIAMOND": "SCHWARZ_D", "SCHWARZ_PRIMITIVE": "SCHWARZ_P"}
    kinds = {alias.get(k, k): v for k, v in hdr.items() if alias.get(k, k) in eng}
    assert len(kinds) == 17 and kinds["SQUARE"] == 16
    for name, v in kinds.items():
        assert eng[name] == v and orc[name] == v, name
        assert getattr(O, "K_" + name) == v, name
    for name in ("SPHERE", "PLANE", "BEZIER", "CONSTANT", "CYLINDER", "ELLIPSOID", "ANNULUS", "TORUS", "COMPOSITE", "DIM_LINEAR", "SQUARE"):
        assert getattr(cpp, "QG_" + name) == kinds[name], name
