"""Custom-coefficient packing on the GPU: the consumer of the quadrature rules in QUGaR's FEniCSx layer.

Mirrors, for cell and unfitted-boundary integrals with real coefficients, what
python/qugar/dolfinx/custom_coefficients.py does on the host after the rules have been created:
    _allocate_new_coeffs (:278-312), _compute_n_vals_per_custom_entity (:314-368), _compute_offsets (:370-449),
    _compute_new_vals (:730-775), _copy_vals (:451-512; the reference's own TODO: "this loop over the cells may be slow").
Here the rules never leave HBM (`cpp.create_quadrature_device`, `cpp.create_unfitted_bound_quadrature_device`); one CUDA
pass lays out `new_coeffs` and only that array crosses PCIe (half the bytes for float32 forms).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import cpp


class PackedCoeffs:
    """`new_coeffs` of custom_coefficients.py: shape (n_rows + n_extra_rows, n_old_cols + n_extra_cols)."""

    def __init__(self, handle):
        self._h = handle
        lib = cpp._load()
        nr, nc, es, d2h = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
        ph, pd = C.c_void_p(), C.c_void_p()
        cpp._check(lib.qg_coeffs_view(handle, C.byref(nr), C.byref(nc), C.byref(es), C.byref(ph), C.byref(pd), C.byref(d2h)))
        self.shape = (int(nr.value), int(nc.value))
        self.dtype = np.dtype(np.float64 if es.value == 8 else np.float32)
        self.d2h_bytes = int(d2h.value)
        self._keep = _CoeffsHandle(handle)
        self.host = None
        if ph.value:
            ctype = C.c_double if es.value == 8 else C.c_float
            self.host = cpp._owned_view(ph.value, ctype, self.shape, self.dtype, self._keep)
        self.device = cpp._DeviceArray(pd.value, self.shape, "<f8" if es.value == 8 else "<f4", self._keep)


class _CoeffsHandle:
    __slots__ = ("h", "__weakref__")

    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            cpp._load().qg_coeffs_free(self.h)
        except Exception:
            pass


def pack_custom_coeffs(old_coeffs, custom_entity_ids, empty_entity_ids, rules, to_host=True, device=None):
    """new_coeffs for one integral.

    old_coeffs: (n_entities, n_cols) float32 / float64 host array (the standard coefficients of the integral's entities);
    custom_entity_ids / empty_entity_ids: positions (rows) of the cut / empty entities, as _get_custom_entities_ids /
    _get_empty_entities_ids return them (:640-707); rules: one device-resident rule per quadrature of the integrand
    (cpp.DeviceRule), each with one entity per custom id, in custom_entity_ids order."""
    old = np.ascontiguousarray(old_coeffs)
    if old.ndim != 2 or old.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise NotImplementedError("pack_custom_coeffs: real float32 / float64 coefficient arrays of shape (n_entities, n_cols)")
    cust = np.ascontiguousarray(custom_entity_ids, np.int64).reshape(-1)
    emp = np.ascontiguousarray(empty_entity_ids, np.int64).reshape(-1)
    if not rules:
        raise ValueError("pack_custom_coeffs: at least one rule")
    lib = cpp._load()
    lib.qg_pack_custom_coeffs.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.qg_coeffs_view.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    lib.qg_coeffs_free.argtypes = [C.c_void_p]
    arr = (C.c_void_p * len(rules))(*[r._h for r in rules])
    if device is None:
        device = 0
    h = C.c_void_p()
    cpp._check(lib.qg_pack_custom_coeffs(int(device), old.dtype.itemsize, cpp._ptr(old), old.shape[0], old.shape[1], cpp._ptr(cust),
                                         len(cust), cpp._ptr(emp), len(emp), arr, len(rules), int(bool(to_host)), C.byref(h)))
    return PackedCoeffs(h)
