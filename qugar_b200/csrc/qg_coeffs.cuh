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

// ---- BezierTP::rescale_domain / sign for many boxes at once (bezier_tp.cpp:403-431) ------------------------------------
// One thread per box: the global coefficient block (direction 0 fastest) is restricted to the box by de Casteljau
// subdivision, one direction after the other, and the sign of the restricted coefficients is tested -- the per-cell
// Bernstein coefficients the polynomial path starts from (impl_quadrature.cpp:165-168, impl_unfitted_domain.cpp:351-360).
constexpr int kMaxBezierCoefs = 216;  // order <= 6 per direction in 3-D
__global__ void __launch_bounds__(64) bezier_rescale_kernel(int dim, int o0, int o1, int o2, const double *coefs, const double *boxes,
  long long nbox, double *out, signed char *sign)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbox) return;
  const int order[3] = { o0, o1, dim > 2 ? o2 : 1 };
  const int n = o0 * o1 * order[2];
  double c[kMaxBezierCoefs];
  for (int k = 0; k < n; ++k) c[k] = coefs[k];
  int stride = 1;
  for (int d = 0; d < dim; ++d) {
    const int P = order[d];
    const double a = boxes[i * 2 * dim + d], b = boxes[i * 2 * dim + dim + d];
    const bool left_first = fabs(b) >= fabs(a - 1.0);
    const double t0 = left_first ? b : a;
    const double t1 = left_first ? a / b : (b - a) / (1.0 - a);
    for (int base = 0; base < n; ++base) {
      if ((base / stride) % P != 0) continue;
      double *l = c + base;
      for (int pass = 0; pass < 2; ++pass) {
        const double tau = pass == 0 ? t0 : t1;
        if ((pass == 0) == left_first) {  // one-sided subdivision keeping [0,tau]
          for (int r = 1; r < P; ++r)
            for (int j = P - 1; j >= r; --j) l[j * stride] = l[j * stride] * tau + l[(j - 1) * stride] * (1.0 - tau);
        } else {                          // keeping [tau,1]
          for (int r = 1; r < P; ++r)
            for (int j = 0; j < P - r; ++j) l[j * stride] = l[j * stride] * (1.0 - tau) + l[(j + 1) * stride] * tau;
        }
      }
    }
    stride *= P;
  }
  bool pos = true, neg = true;
  for (int k = 0; k < n; ++k) {
    out[i * n + k] = c[k];
    pos = pos && c[k] > 0.0;
    neg = neg && c[k] < 0.0;
  }
  sign[i] = pos ? 1 : (neg ? -1 : 0);
}

}  // namespace qg
