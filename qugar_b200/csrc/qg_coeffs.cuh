// qugar_b200 -- custom-coefficient packing on the device: the consumer side of the rules.
// The reference's FEniCSx layer smuggles every custom cell's quadrature into the DOLFINx coefficient array
// (python/qugar/dolfinx/custom_coefficients.py): per custom entity and per quadrature
//     [ n_pts | points (gdim * n_pts) | weights (n_pts) | normals (gdim * n_pts) if unfitted boundary ]      (:730-775)
// appended after the n_rows original rows, and per row a ptrdiff_t offset in the trailing column(s):
//     -1 standard (full) entity, 0 empty entity, else the position of the entity's block relative to the row  (:370-449)
// These kernels produce that array from device-resident rules; only the packed array crosses PCIe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qg {

constexpr int kMaxCoeffRules = 8;

struct CoeffRules
{
  int n_rules;
  int pdim[kMaxCoeffRules];              // point dimension
  int has_nrm[kMaxCoeffRules];
  const int *n_per[kMaxCoeffRules];      // [n_custom]
  const long long *ent_off[kMaxCoeffRules];  // [n_custom + 1] exclusive scan of n_per
  const double *pts[kMaxCoeffRules], *wts[kMaxCoeffRules], *nrm[kMaxCoeffRules];
};

// _compute_n_vals_per_custom_entity (custom_coefficients.py:332-368): real slots per custom entity
__global__ void coeff_counts_kernel(CoeffRules r, long long n_custom, int *n_vals)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_custom) return;
  int v = 0;
  for (int q = 0; q < r.n_rules; ++q) v += 1 + r.n_per[q][e] * (r.pdim[q] + 1 + (r.has_nrm[q] ? r.pdim[q] : 0));
  n_vals[e] = v;
}

// writes a ptrdiff_t into the trailing column(s) of row `row` (two 32-bit halves: rows of float32 coefficients are only
// 4-byte aligned; one intp spans two float32 cells, custom_coefficients.py:253-276, 426-449)
template<typename T> __device__ __forceinline__ void put_offset(T *coeffs, long long n_cols, int n_extra, long long row, long long v)
{
  unsigned *dst = reinterpret_cast<unsigned *>(coeffs + row * n_cols + (n_cols - n_extra));
  dst[0] = (unsigned)((unsigned long long)v & 0xffffffffull);
  dst[1] = (unsigned)((unsigned long long)v >> 32);
}

// _compute_offsets (:370-449): -1 for every row first ...
template<typename T> __global__ void coeff_offsets_default_kernel(T *coeffs, long long n_rows, long long n_cols, int n_extra)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows) put_offset(coeffs, n_cols, n_extra, i, -1);
}
// ... then the custom entities: absolute position of the block minus the first position of the entity's row ...
template<typename T>
__global__ void coeff_offsets_custom_kernel(T *coeffs, long long n_rows, long long n_cols, int n_extra, const long long *custom_ids,
  long long n_custom, const long long *val_off)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_custom) return;
  const long long row = custom_ids[e];
  put_offset(coeffs, n_cols, n_extra, row, n_rows * n_cols + val_off[e] - row * n_cols);
}
// ... and 0 for the empty entities, last (an id listed both ways ends up empty, as upstream)
template<typename T>
__global__ void coeff_offsets_empty_kernel(T *coeffs, long long n_cols, int n_extra, const long long *empty_ids, long long n_empty)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_empty) put_offset(coeffs, n_cols, n_extra, empty_ids[e], 0);
}

// _copy_vals (:451-512): one block per custom entity; every array segment is copied with consecutive threads on
// consecutive slots (coalesced reads of the rule, coalesced writes of the block), converted to the coefficient type.
template<typename T>
__global__ void __launch_bounds__(256) coeff_pack_kernel(CoeffRules r, T *coeffs, long long first_extra_pos, const long long *val_off,
  long long n_custom)
{
  const long long e = blockIdx.x;
  if (e >= n_custom) return;
  T *dst = coeffs + first_extra_pos + val_off[e];
  for (int q = 0; q < r.n_rules; ++q) {
    const int n = r.n_per[q][e];
    const long long o = r.ent_off[q][e];
    const int pd = r.pdim[q];
    if (threadIdx.x == 0) dst[0] = (T)n;   // new_coeffs_real[offsets] = n_pts_per_cell (numeric conversion)
    ++dst;
    const double *sp = r.pts[q] + o * pd;
    for (int i = threadIdx.x; i < n * pd; i += blockDim.x) dst[i] = (T)sp[i];
    dst += (long long)n * pd;
    const double *sw = r.wts[q] + o;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (T)sw[i];
    dst += n;
    if (r.has_nrm[q]) {
      const double *sn = r.nrm[q] + o * pd;
      for (int i = threadIdx.x; i < n * pd; i += blockDim.x) dst[i] = (T)sn[i];
      dst += (long long)n * pd;
    }
  }
}

}  // namespace qg
