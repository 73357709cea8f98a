"""ctypes access to the CPU oracle (oracle/_build/liboracle*.so).

Test infrastructure only: nothing under qugar_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# FuncKind (oracle/oracle_funcs.hpp)
K_SPHERE, K_PLANE, K_BEZIER, K_CONSTANT, K_CYLINDER, K_ELLIPSOID, K_ANNULUS, K_TORUS, K_COMPOSITE, K_DIM_LINEAR = range(10)
K_SQUARE = 16
K_SCHOEN, K_SCHOEN_IWP, K_SCHOEN_FRD, K_FISCHER_KOCH_S, K_SCHWARZ_D, K_SCHWARZ_P = 10, 11, 12, 13, 14, 15

CELL_CUT, CELL_FULL, CELL_EMPTY = 0, 1, 2
(F_CUT, F_FULL, F_EMPTY, F_CUT_UNF, F_CUT_EXT, F_CUT_UNF_EXT, F_FULL_UNF, F_FULL_EXT, F_UNF, F_EXT,
 F_UNF_EXT) = range(11)

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


_libs = {}


def load(libm: bool = False):
    key = "libm" if libm else "det"
    if key in _libs:
        return _libs[key]
    path = os.path.join(ORACLE_DIR, "_build", "liboracle_libm.so" if libm else "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    lib.orc_set_knobs.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.orc_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.orc_domain_create.restype = C.c_void_p
    lib.orc_domain_create.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f64p, _i64p]
    lib.orc_domain_create_mt.restype = C.c_void_p
    lib.orc_domain_create_mt.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f64p, _i64p, C.c_int]
    lib.orc_domain_free.argtypes = [C.c_void_p]
    lib.orc_domain_num_cells.restype = C.c_int64
    lib.orc_domain_num_cells.argtypes = [C.c_void_p]
    lib.orc_domain_num_facet_records.restype = C.c_int64
    lib.orc_domain_num_facet_records.argtypes = [C.c_void_p]
    lib.orc_domain_status.argtypes = [C.c_void_p, _u8p]
    lib.orc_domain_facet_records.argtypes = [C.c_void_p, _i64p, _u8p]
    lib.orc_quad_cells.restype = C.c_void_p
    lib.orc_quad_cells.argtypes = [C.c_void_p, _i64p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
    lib.orc_quad_unf_bdry.restype = C.c_void_p
    lib.orc_quad_unf_bdry.argtypes = [C.c_void_p, _i64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_quad_facets.restype = C.c_void_p
    lib.orc_quad_facets.argtypes = [C.c_void_p, _i64p, _i32p, C.c_int64, C.c_int, C.c_int, C.c_int]
    lib.orc_rule_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int)]
    lib.orc_rule_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_rule_free.argtypes = [C.c_void_p]
    lib.orc_func_eval.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f64p, C.c_int64, _f64p, C.c_void_p]
    lib.orc_func_eval_interval.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f64p, _f64p,
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _libs[key] = lib
    return lib


@dataclass
class Rule:
    n_per: np.ndarray
    points: np.ndarray
    weights: np.ndarray
    normals: np.ndarray | None


def _take_rule(lib, h) -> Rule:
    ne, npts, pdim, hn = C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
    lib.orc_rule_sizes(h, C.byref(ne), C.byref(npts), C.byref(pdim), C.byref(hn))
    n_per = np.zeros(ne.value, np.int32)
    pts = np.zeros((npts.value, pdim.value), np.float64)
    wts = np.zeros(npts.value, np.float64)
    nrm = np.zeros((npts.value, pdim.value), np.float64) if hn.value else None
    lib.orc_rule_copy(h, n_per.ctypes.data, pts.ctypes.data, wts.ctypes.data,
                      nrm.ctypes.data if nrm is not None else None)
    lib.orc_rule_free(h)
    return Rule(n_per, pts, wts, nrm)


def cpp_breaks(n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """CartGridTP(domain, n) breaks: x0 + dx*i (cart_grid_tp.cpp:62-69)."""
    dx = (hi - lo) / float(n)
    return np.array([lo + dx * float(i) for i in range(n + 1)], np.float64)


class OracleDomain:
    """UnfittedImplDomain restatement (oracle/oracle.cpp)."""

    def __init__(self, dim, kind, params, breaks, negative=False, libm=False, nthreads=1):
        self.lib = load(libm)
        self.dim = dim
        self.breaks = [np.ascontiguousarray(b, np.float64) for b in breaks]
        assert len(self.breaks) == dim
        self.n_cells_dir = [len(b) - 1 for b in self.breaks]
        p = np.ascontiguousarray(params, np.float64)
        allb = np.concatenate(self.breaks)
        nb = np.array([len(b) for b in self.breaks], np.int64)
        self.h = self.lib.orc_domain_create_mt(dim, kind, int(negative), p, len(p), allb, nb, int(nthreads))
        n = self.lib.orc_domain_num_cells(self.h)
        self.status = np.zeros(n, np.uint8)
        self.lib.orc_domain_status(self.h, self.status)
        nr = self.lib.orc_domain_num_facet_records(self.h)
        self.facet_cells = np.zeros(nr, np.int64)
        self.facet_status = np.zeros((nr, 2 * dim), np.uint8)
        if nr:
            self.lib.orc_domain_facet_records(self.h, self.facet_cells, self.facet_status.reshape(-1))

    def __del__(self):
        try:
            self.lib.orc_domain_free(self.h)
        except Exception:
            pass

    def cells(self, status):
        return np.nonzero(self.status == status)[0].astype(np.int64)

    @property
    def cut_cells(self):
        return self.cells(CELL_CUT)

    @property
    def full_cells(self):
        return self.cells(CELL_FULL)

    @property
    def empty_cells(self):
        return self.cells(CELL_EMPTY)

    def facets_where(self, pred):
        """(cells, local facets) over facet records for which pred(status) holds (sorted by cell)."""
        mask = pred(self.facet_status)
        idx = np.nonzero(mask)
        return self.facet_cells[idx[0]].astype(np.int64), idx[1].astype(np.int32)

    def quad_cells(self, cells, n, nthreads=1, stats=False):
        cells = np.ascontiguousarray(cells, np.int64)
        st = (C.c_longlong * 3)()
        h = self.lib.orc_quad_cells(self.h, cells, len(cells), n, nthreads, C.cast(st, C.c_void_p))
        r = _take_rule(self.lib, h)
        return (r, list(st)) if stats else r

    def quad_unf_bdry(self, cells, n, include_facet_unf_bdry=False, exclude_ext_bdry=False, nthreads=1):
        cells = np.ascontiguousarray(cells, np.int64)
        h = self.lib.orc_quad_unf_bdry(self.h, cells, len(cells), n, int(include_facet_unf_bdry),
                                       int(exclude_ext_bdry), nthreads)
        return _take_rule(self.lib, h)

    def quad_facets(self, cells, facets, n, exterior, nthreads=1):
        cells = np.ascontiguousarray(cells, np.int64)
        facets = np.ascontiguousarray(facets, np.int32)
        h = self.lib.orc_quad_facets(self.h, cells, facets, len(cells), n, int(exterior), nthreads)
        return _take_rule(self.lib, h)


def is_cut_facet(s):
    return (s == F_CUT) | (s == F_CUT_UNF) | (s == F_CUT_EXT) | (s == F_CUT_UNF_EXT)


def is_empty_facet(s):
    return (s == F_EMPTY) | (s == F_FULL_EXT) | (s == F_EXT)


def has_unf_bdry(s):
    return (s == F_CUT_UNF) | (s == F_CUT_UNF_EXT) | (s == F_FULL_UNF) | (s == F_UNF) | (s == F_UNF_EXT)


def func_eval(dim, kind, params, pts, negative=False, libm=False, grad=True):
    lib = load(libm)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, dim)
    p = np.ascontiguousarray(params, np.float64)
    val = np.zeros(len(pts))
    g = np.zeros((len(pts), dim)) if grad else None
    lib.orc_func_eval(dim, kind, int(negative), p, len(p), pts.reshape(-1), len(pts), val,
                      g.ctypes.data if g is not None else None)
    return val, g


def bezier_rescale(order, coefs, lo, hi):
    """orc_bezier_rescale: restatement of BezierTP::rescale_domain + sign (bezier_tp.cpp:403-431)."""
    lib = load(False)
    order = np.ascontiguousarray(order, np.int32)
    coefs = np.ascontiguousarray(coefs, np.float64)
    lo = np.ascontiguousarray(lo, np.float64)
    hi = np.ascontiguousarray(hi, np.float64)
    out = np.empty_like(coefs)
    lib.orc_bezier_rescale.restype = C.c_int
    lib.orc_bezier_rescale.argtypes = [C.c_int, _i32p, _f64p, _f64p, _f64p, _f64p]
    sg = lib.orc_bezier_rescale(len(order), order, coefs, lo, hi, out)
    return out, sg
