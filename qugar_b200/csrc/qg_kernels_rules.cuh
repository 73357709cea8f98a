// qugar_b200 -- kernels of the rule assembly: entity dispatch, Gauss fill, purge / compaction, facet rules, merge
// (part of the translation unit qg_engine.cu; included there, in this order)
#pragma once

namespace qg {

// ---- rule assembly ------------------------------------------------------------------------------------
enum EntKind : int { ENT_NONE = 0, ENT_JOB = 1, ENT_GAUSS = 2 };

// create_cell_quadrature dispatch (impl_quadrature.cpp:510-536): cut & recorded -> Algoim; full -> Gauss; else 0
__global__ void entity_kind_cells_kernel(long long ne, const long long *cells, long long ncells,
  const unsigned char *status, const int *rec_idx, int *kind, int *bad)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const long long c = cells[e];
  if (c < 0 || c >= ncells) {
    kind[e] = ENT_NONE;
    *bad = 1;
    return;
  }
  const unsigned char s = status[c];
  kind[e] = (s == ST_CUT && rec_idx[c] >= 0) ? ENT_JOB : (s == ST_FULL ? ENT_GAUSS : ENT_NONE);
}
// create_cell_unfitted_bound_quadrature dispatch (impl_quadrature.cpp:556-580)
__global__ void entity_kind_unf_kernel(long long ne, const long long *cells, long long ncells,
  const unsigned char *status, const int *rec_idx, int *kind, int *bad)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const long long c = cells[e];
  if (c < 0 || c >= ncells) {
    kind[e] = ENT_NONE;
    *bad = 1;
    return;
  }
  kind[e] = (status[c] == ST_CUT && rec_idx[c] >= 0) ? ENT_JOB : ENT_NONE;
}
__global__ void flag_kind_kernel(long long ne, const int *kind, int v, int *flag)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne) flag[e] = kind[e] == v ? 1 : 0;
}
__global__ void gather_jobs_kernel(long long ne, const long long *cells, const int *flag, const long long *off,
  long long *job_cell, int *job_ent)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || !flag[e]) return;
  job_cell[off[e]] = cells[e];
  job_ent[off[e]] = (int)e;
}
__global__ void entity_counts_kernel(long long ne, const int *kind, int gauss_n, int *cnt)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne) cnt[e] = kind[e] == ENT_GAUSS ? gauss_n : 0;
}
__global__ void job_counts_to_entities_kernel(long long njobs, const int *job_ent, const long long *job_pt_off,
  const int *job_keep_cnt, int *cnt)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  cnt[job_ent[j]] = job_keep_cnt ? job_keep_cnt[j] : (int)(job_pt_off[j + 1] - job_pt_off[j]);
}
__global__ void job_shift_kernel(long long njobs, const int *job_ent, const long long *job_pt_off,
  const long long *ent_off, long long *shift)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < njobs) shift[j] = ent_off[job_ent[j]] - job_pt_off[j];
}
// Quadrature<dim>::create_Gauss_01 (quadrature.cpp:199-277): direction 0 fastest, w = (w0*w1)*w2
__global__ void gauss_fill_kernel(long long ne, const int *kind, const long long *ent_off, int dim, int p,
  const double *gxw, double *pts, double *wts)
{
  // one warp per entity (most lists hold no full cell at all: the early exit is the common case)
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= ne || kind[e] != ENT_GAUSS) return;
  int tot = 1;
  for (int d = 0; d < dim; ++d) tot *= p;
  const long long base = ent_off[e];
  for (int t = lane; t < tot; t += 32) {
    int idx = t;
    double w = 1.0;
    for (int d = 0; d < dim; ++d) {
      const int i = idx % p;
      idx /= p;
      pts[(base + t) * dim + d] = gxw[2 * i];
      w = d == 0 ? gxw[2 * i + 1] : w * gxw[2 * i + 1];
    }
    wts[base + t] = w;
  }
}

// purge_facet_points, cell version (impl_quadrature.cpp:238-288): a surface node is dropped if it lies on a
// facet that has unfitted boundary, its normal is (+-) the facet normal and it is on the level set.
template<int N>
__global__ void purge_flags_surf_kernel(FuncDesc f, long long njobs, const long long *job_cell,
  const long long *job_pt_off, const double *job_box, const int *rec_idx, const unsigned char *rec_stat,
  const double *pts, const double *nrm, unsigned char *keep, int *job_keep_cnt)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const int r = rec_idx[job_cell[j]];
  const double tol = 1000.0 * kEps;
  const double *lo = job_box + 6 * j, *hi = lo + 3;
  int kept = 0;
  const long long q0 = job_pt_off[j], q1 = job_pt_off[j + 1];
  for (long long q = q0; q < q1; ++q) keep[q] = 1;
  // the reference erases facet by facet (ascending); a node erased once stays erased
  for (int lf = 0; lf < 2 * N; ++lf) {
    const unsigned char s = rec_stat[(size_t)r * 2 * N + lf];
    if (!facet_has_unf_bdry(s)) continue;
    const int cd = lf >> 1, side = lf & 1;
    const double fc = side == 0 ? 0.0 : 1.0;
    for (long long q = q0; q < q1; ++q) {
      if (!keep[q]) continue;
      if (!(fabs(pts[q * N + cd] - fc) <= tol)) continue;
      const double nc = nrm[q * N + cd];
      if ((1.0 - fabs(nc)) > tol) continue;
      double x[N];
      for (int d = 0; d < N; ++d) x[d] = pts[q * N + d] * (hi[d] - lo[d]) + lo[d];
      if (fabs(phi_val<N>(f, x)) <= tol) keep[q] = 0;
    }
  }
  for (long long q = q0; q < q1; ++q) kept += keep[q];
  job_keep_cnt[j] = kept;
}
__global__ void compact_points_kernel(long long njobs, const long long *job_pt_off, const unsigned char *keep,
  const long long *job_dst, int pdim, const double *pts, const double *wts, const double *nrm, double *opts,
  double *owts, double *onrm)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  long long o = job_dst[j];
  for (long long q = job_pt_off[j]; q < job_pt_off[j + 1]; ++q) {
    if (!keep[q]) continue;
    for (int d = 0; d < pdim; ++d) opts[o * pdim + d] = pts[q * pdim + d];
    owts[o] = wts[q];
    if (nrm)
      for (int d = 0; d < pdim; ++d) onrm[o * pdim + d] = nrm[q * pdim + d];
    ++o;
  }
}
__global__ void job_dst_kernel(long long njobs, const int *job_ent, const long long *ent_off, long long *dst)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < njobs) dst[j] = ent_off[job_ent[j]];
}

// ---- facet rules (create_facet_quadrature, impl_quadrature.cpp:595-649 and its drivers) -------------------
__host__ __device__ inline bool facet_is_cut_st(unsigned char s)
{
  return s == QG_FACET_CUT || s == QG_FACET_CUT_UNF_BDRY || s == QG_FACET_CUT_EXT_BDRY || s == QG_FACET_CUT_UNF_BDRY_EXT_BDRY;
}
enum FacetMode : int { FMODE_INTERIOR = 0, FMODE_EXTERIOR = 1, FMODE_UNF_BDRY = 2 };
enum PurgeBits : unsigned char { PURGE_UNF = 1, PURGE_UNF_EXT = 2, PURGE_CUT = 4 };

// One entity = (cell, local facet).  Decides between nothing / the (dim-1) Gauss rule / a cut-facet job, and the
// purge switches of the job: the dispatch of create_facets_quadrature_{interior,exterior}_integral
// (cut_quadrature.cpp:178-262), add_unf_bdry_quad (impl_quadrature.cpp:349-389) and create_facet_quadrature.
template<int N>
__global__ void facet_entity_kernel(GridDesc g, long long ne, const long long *cells, const int *facets, int facets_stride,
  int mode, int exclude_ext, long long ncells, const unsigned char *status, const int *rec_idx,
  const unsigned char *rec_stat, int *kind, unsigned char *pflags, int *bad)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  // facets == nullptr: the entities are all 2N facets of each cell (cell = e / 2N, facet = e % 2N)
  const long long c = facets ? cells[e] : cells[e / (2 * N)];
  const int lf = facets ? facets[e * facets_stride] : (int)(e % (2 * N));
  kind[e] = ENT_NONE;
  pflags[e] = 0;
  if (c < 0 || c >= ncells || lf < 0 || lf >= 2 * N) {
    *bad = 1;
    return;
  }
  const int r = rec_idx[c];
  const bool has_rec = r >= 0;
  const unsigned char st = has_rec ? rec_stat[(size_t)r * 2 * N + lf] : (unsigned char)QG_FACET_EMPTY;
  const bool f_full = has_rec ? st == QG_FACET_FULL : status[c] == ST_FULL;
  const bool f_full_unf = has_rec && st == QG_FACET_FULL_UNF_BDRY;
  const bool f_cut = has_rec && facet_is_cut_st(st);
  const bool f_unf = has_rec && facet_has_unf_bdry(st);
  // exterior facet: the cell touches the grid boundary on that side
  long long id = c;
  bool ext = false;
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const long long t = id % g.n[d];
    id /= g.n[d];
    if (d == (lf >> 1)) ext = (lf & 1) ? (t == g.n[d] - 1) : (t == 0);
  }
  bool run = false, remove_unf = false, remove_cut = false;
  if (mode == FMODE_INTERIOR) {
    run = !ext;
    remove_unf = true;
  } else if (mode == FMODE_EXTERIOR) {
    if (ext) {
      run = true;
    } else if (f_unf) {
      run = true;
      remove_cut = true;
    }
  } else {
    run = f_unf && !(exclude_ext && ext);
    remove_cut = true;
  }
  if (!run) return;
  if (f_full_unf || f_full) {
    if (!(f_full_unf && remove_unf)) kind[e] = ENT_GAUSS;
  } else if (f_cut || f_unf) {
    kind[e] = ENT_JOB;
    unsigned char pf = 0;
    if (remove_unf && f_unf) pf |= PURGE_UNF;
    if (f_unf) pf |= PURGE_UNF_EXT;
    if (remove_cut && f_cut) pf |= PURGE_CUT;
    pflags[e] = pf;
  }
}
template<int N>
__global__ void gather_facet_jobs_kernel(long long ne, const long long *cells, const int *facets, int facets_stride,
  const int *flag, const long long *off, long long *job_cell, int *job_ent, int *job_type)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || !flag[e]) return;
  job_cell[off[e]] = facets ? cells[e] : cells[e / (2 * N)];
  job_type[off[e]] = facets ? facets[e * facets_stride] : (int)(e % (2 * N));
  job_ent[off[e]] = (int)e;
}
// purge_facet_points, facet version (impl_quadrature.cpp:291-347)
template<int N>
__global__ void purge_flags_facet_kernel(FuncDesc f, long long njobs, const int *job_type, const int *job_ent,
  const unsigned char *pflags, const long long *job_pt_off, const double *job_box, const double *pts,
  unsigned char *keep, int *job_keep_cnt)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const unsigned char pf = pflags[job_ent[j]];
  const double tol = 1000.0 * kEps;
  const double *lo = job_box + 6 * j, *hi = lo + 3;
  const int lf = job_type[j], cd = lf >> 1, side = lf & 1;
  const double fc = side == 0 ? 0.0 : 1.0;
  int kept = 0;
  for (long long q = job_pt_off[j]; q < job_pt_off[j + 1]; ++q) {
    bool erase = false;
    if (pf) {
      double x[N];
      int k = 0;
      for (int d = 0; d < N; ++d) {
        const double p01 = (d == cd) ? fc : pts[q * (N - 1) + k];
        if (d != cd) ++k;
        x[d] = p01 * (hi[d] - lo[d]) + lo[d];
      }
      if (fabs(phi_val<N>(f, x)) <= tol) {
        double g[N];
        phi_grad<N>(f, x, g);
        double gc = g[0];
        for (int d = 1; d < N; ++d)
          if (d == cd) gc = g[d];
        const bool outside = side == 0 ? (gc < -tol) : (tol < gc);
        erase = (outside && (pf & PURGE_UNF)) || (!outside && (pf & PURGE_UNF_EXT));
      } else if (pf & PURGE_CUT) {
        erase = true;
      }
    }
    keep[q] = erase ? 0 : 1;
    kept += erase ? 0 : 1;
  }
  job_keep_cnt[j] = kept;
}
// add_unf_bdry_quad (impl_quadrature.cpp:349-389): per cell, the surface nodes followed by the nodes of its facets
// with unfitted boundary, lifted to the cell (coordinate fc on the facet direction, normal -+ e_cd)
template<int N>
__global__ void merge_counts_kernel(long long ne, const int *n_surf, const int *n_facet, int *n_out)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int n = n_surf[e];
  for (int lf = 0; lf < 2 * N; ++lf) n += n_facet[e * 2 * N + lf];
  n_out[e] = n;
}
template<int N>
__global__ void merge_rules_kernel(long long ne, const long long *off_out, const long long *off_surf, const int *n_surf,
  const double *sp, const double *sw, const double *sn, const long long *off_facet, const int *n_facet, const double *fp,
  const double *fw, double *op, double *ow, double *on)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  long long o = off_out[e];
  for (long long q = off_surf[e]; q < off_surf[e] + n_surf[e]; ++q, ++o) {
    for (int d = 0; d < N; ++d) {
      op[o * N + d] = sp[q * N + d];
      on[o * N + d] = sn[q * N + d];
    }
    ow[o] = sw[q];
  }
  for (int lf = 0; lf < 2 * N; ++lf) {
    const long long fe = e * 2 * N + lf;
    const int cd = lf >> 1, side = lf & 1;
    for (long long q = off_facet[fe]; q < off_facet[fe] + n_facet[fe]; ++q, ++o) {
      int k = 0;
      for (int d = 0; d < N; ++d) {
        op[o * N + d] = (d == cd) ? (side == 0 ? 0.0 : 1.0) : fp[q * (N - 1) + k];
        if (d != cd) ++k;
        on[o * N + d] = (d == cd) ? (side == 0 ? -1.0 : 1.0) : 0.0;
      }
      ow[o] = fw[q];
    }
  }
}

}  // namespace qg
