"""ctypes access to the CPU emulation harness of the kernel bodies (tests/emu/emu.cpp). Test aid only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
SINK_CELL, SINK_SURF, SINK_FACET, SINK_RAW = 0, 1, 2, 3
JOB_VOLUME, JOB_SURFACE = -1, -2

_lib = None


def build_emu():
    out = os.path.join(EMU_DIR, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libemu.so")
    src = os.path.join(EMU_DIR, "emu.cpp")
    csrc = os.path.join(HERE, "..", "qugar_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma",
                               "-x", "c++", src, "-o", so])
    return so


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_emu())
        _lib.emu_run.restype = C.c_void_p
    return _lib


def run(dim, kind, params, breaks, cells, types, p, mode, negative=False):
    lib = load()
    params = np.ascontiguousarray(params, np.float64)
    allb = np.concatenate([np.ascontiguousarray(b, np.float64) for b in breaks])
    nb = np.array([len(b) for b in breaks], np.int64)
    cells = np.ascontiguousarray(cells, np.int64)
    types = np.ascontiguousarray(np.broadcast_to(types, cells.shape), np.int32)
    h = lib.emu_run(C.c_int(dim), C.c_int(kind), C.c_int(int(negative)), params.ctypes.data_as(C.c_void_p),
                    C.c_int(len(params)), allb.ctypes.data_as(C.c_void_p), nb.ctypes.data_as(C.c_void_p),
                    cells.ctypes.data_as(C.c_void_p), types.ctypes.data_as(C.c_void_p), C.c_longlong(len(cells)),
                    C.c_int(p), C.c_int(mode))
    h = C.c_void_p(h)
    npts, pdim, err = C.c_longlong(), C.c_int(), C.c_int()
    stats = (C.c_longlong * 3)()
    lib.emu_sizes(h, C.byref(npts), C.byref(pdim), C.byref(err), stats)
    n_per = np.zeros(len(cells), np.int32)
    pts = np.zeros((npts.value, pdim.value))
    wts = np.zeros(npts.value)
    nrm = np.zeros((npts.value, dim))
    lib.emu_copy(h, n_per.ctypes.data_as(C.c_void_p), pts.ctypes.data_as(C.c_void_p),
                 wts.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p))
    lib.emu_free(h)
    return dict(n_per=n_per, points=pts, weights=wts, normals=nrm, err=err.value, stats=list(stats))


def reparam(dim, kind, params, breaks, cut_cells, full_cells, order, mode, negative=False):
    """The reparameterization bodies (qugar_b200/csrc/qg_reparam.cuh) run on the CPU, before the merge: points, connectivity,
    wirebasket connectivity, error bits.  mode: 0 volume, 1 level set, 2 + f face f."""
    lib = load()
    lib.emu_reparam.restype = C.c_void_p
    params = np.ascontiguousarray(params, np.float64)
    allb = np.concatenate([np.ascontiguousarray(b, np.float64) for b in breaks])
    nb = np.array([len(b) for b in breaks], np.int64)
    cut = np.ascontiguousarray(cut_cells, np.int64)
    full = np.ascontiguousarray(full_cells, np.int64)
    h = lib.emu_reparam(C.c_int(dim), C.c_int(kind), C.c_int(int(negative)), params.ctypes.data_as(C.c_void_p),
                        C.c_int(len(params)), allb.ctypes.data_as(C.c_void_p), nb.ctypes.data_as(C.c_void_p),
                        cut.ctypes.data_as(C.c_void_p), C.c_longlong(len(cut)), full.ctypes.data_as(C.c_void_p),
                        C.c_longlong(len(full)), C.c_int(order), C.c_int(mode))
    h = C.c_void_p(h)
    a, b, c, err = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_int()
    lib.emu_mesh_sizes(h, C.byref(a), C.byref(b), C.byref(c), C.byref(err))
    pts = np.zeros(a.value)
    conn = np.zeros(b.value, np.int64)
    wires = np.zeros(c.value, np.int64)
    lib.emu_mesh_copy(h, pts.ctypes.data_as(C.c_void_p), conn.ctypes.data_as(C.c_void_p), wires.ctypes.data_as(C.c_void_p))
    lib.emu_mesh_free(h)
    return pts.reshape(-1, dim), conn, wires, err.value
