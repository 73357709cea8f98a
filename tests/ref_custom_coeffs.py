"""NumPy restatement of the reference's custom-coefficient packing -- TEST INFRASTRUCTURE (the oracle of SURVEY 8(f)3).

Follows python/qugar/dolfinx/custom_coefficients.py for cell / unfitted-boundary integrals with real coefficients:
  _get_n_extra_cols (:253-276), _allocate_new_coeffs (:278-312), _compute_n_vals_per_custom_entity (:314-368),
  _compute_offsets (:370-449), _compute_new_vals (:730-775), _copy_vals (:451-512).
`quads`: list of (n_pts_per_entity int32[nc], points f64[N,gdim], weights f64[N], normals f64[N,gdim] or None).
"""
import numpy as np


def pack_custom_coeffs_ref(old_coeffs, custom_entity_ids, empty_entity_ids, quads):
    coeffs_dtype = np.dtype(old_coeffs.dtype)
    offsets_dtype = np.dtype(np.intp)
    n_extra_cols = int(np.ceil(offsets_dtype.itemsize / coeffs_dtype.itemsize))
    n_custom = len(custom_entity_ids)
    # _compute_n_vals_per_custom_entity
    n_vals = np.zeros(n_custom, np.int32)
    for n_per, pts, _w, nrm in quads:
        gdim = pts.shape[1]
        per_pt = gdim + 1 + (gdim if nrm is not None else 0)
        assert n_per.size == n_custom
        n_vals += n_per * per_pt
        n_vals += 1
    # _allocate_new_coeffs
    n_cols = old_coeffs.shape[1] + n_extra_cols
    n_tot = np.sum(n_vals, dtype=offsets_dtype)
    n_rows = offsets_dtype.type(old_coeffs.shape[0])
    n_extra_rows = offsets_dtype.type(np.ceil(n_tot / n_cols))
    new = np.zeros((n_rows + n_extra_rows, n_cols), dtype=coeffs_dtype)
    new[: old_coeffs.shape[0], : old_coeffs.shape[1]] = old_coeffs
    # _compute_offsets
    n_entities = old_coeffs.shape[0]
    first_extra_pos = offsets_dtype.type(n_entities * n_cols)
    first_col_pos = np.arange(n_entities, dtype=offsets_dtype) * n_cols
    if n_custom == 0:
        offsets = np.empty(0, dtype=offsets_dtype)
    else:
        offsets = (np.concatenate(([0], np.cumsum(n_vals[:-1]))) + first_extra_pos).astype(offsets_dtype)
    rel = offsets - first_col_pos[custom_entity_ids]
    all_rel = np.full(n_entities, -1, dtype=rel.dtype)
    all_rel[custom_entity_ids] = rel
    all_rel[empty_entity_ids] = 0
    new[:n_entities, -n_extra_cols:] = all_rel.view(coeffs_dtype).reshape([-1, n_extra_cols])
    # _compute_new_vals + _copy_vals
    flat = new.ravel()
    offs = np.copy(offsets)
    for n_per, pts, w, nrm in quads:
        vals = [np.ascontiguousarray(pts).ravel(), w] + ([nrm.ravel()] if nrm is not None else [])
        n_per = n_per.astype(offsets_dtype)
        flat[offs] = n_per
        offs += 1
        per = np.array([v.size for v in vals], dtype=offsets_dtype)
        tot = np.sum(n_per)
        if tot > 0:
            per //= tot
        vo = np.zeros(len(vals), dtype=offsets_dtype)
        for c in range(n_custom):
            n = n_per[c]
            c0 = offs[c]
            for i, v in enumerate(vals):
                nv = n * per[i]
                flat[c0:c0 + nv] = v[vo[i]:vo[i] + nv]
                c0 += nv
                vo[i] += nv
            offs[c] = c0
    return new
