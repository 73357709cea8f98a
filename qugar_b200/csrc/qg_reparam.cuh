// qugar_b200 -- reparameterization of cut cells by a general level set: per-thread bodies (host/device inline).
//
// The reference (/root/reference/cpp/qugar/impl_reparam_general.cpp:70-875, ImplicitGeneralReparam) runs the same
// recursion as the quadrature (prune / height direction / bisection) and, at every leaf of that recursion, builds
// Lagrange cells of order^pd points between consecutive roots of the level set: first "reference intervals" at interval
// midpoints (:275-305), then, point by point, "similar" intervals that keep the root count of the reference ones
// (:423-551), points inserted into cells keyed by the interval indices (:314-329, 587-650).
// Here the leaves are the chain items the CTL pass already emits (qg_quad.cuh); one thread per chain item runs that leaf as a
// flat state machine over at most three stages with fixed-size local arrays (no maps, no recursion on the heap), once to count
// its cells and points and once to write them at the offsets of an exclusive scan -- so points and cells come out in the
// reference's insertion order.  Orientation (impl_reparam_mesh.cpp:183-278), wirebasket (:100-178, reparam_mesh.cpp:502-605) and
// full cells (reparam_mesh.cpp:352-390) are one thread per cell / per (cell, edge).
#pragma once
#include "qg_quad.cuh"
#include <cmath>
#include <vector>

// the large helpers of the leaf walk are called from many sites of a three-stage template recursion: kept out of line on the
// device (the fully inlined kernel took 20 minutes to compile)
#if defined(__CUDACC__)
#define QG_RP_FN __host__ __device__ __noinline__
#else
#define QG_RP_FN inline
#endif

namespace qg {

enum { ERR_RP_CAP = 8 };
constexpr int kRpR0 = kMaxRoots + 2;  // roots of the base line incl. end points
constexpr int kRpR = 6;               // roots per monotone line: end points + one per psi (<= 2) + slack
constexpr int kRpIv0 = 12;            // intervals of the base line that may carry cells
constexpr int kRpIv1 = 3;
constexpr int kRpCells = 48;          // cells per leaf
constexpr int kRpMaxOrder = 32;

template<int R> struct RpIv
{
  double r[R + 1];
  signed char fid[R + 1];
  int n;
  unsigned active;
};

// Tolerance().update(10 eps x extent): the LARGER of 10 eps and 10 eps x extent (impl_reparam_general.cpp:152-157, tolerance.cpp:47-55)
QG_HD double rp_tol(const ChainItem &it, int k)
{
  double ext = it.hi[0] - it.lo[0];
  for (int d = 1; d < 3; ++d)
    if (d == k) ext = it.hi[d] - it.lo[d];
  const double t = kNearEps * ext;
  return t > kNearEps ? t : kNearEps;
}
QG_HD bool rp_eq(double tol, double a, double b) { return fabs(a - b) <= tol; }

template<int R> QG_HD void rp_add(RpIv<R> &iv, double r, int f, int *err)
{
  if (iv.n < R) {
    iv.r[iv.n] = r;
    iv.fid[iv.n] = (signed char)f;
    ++iv.n;
  } else {
    *err |= ERR_RP_CAP;
  }
}

// RootsIntervals::adjust_roots (impl_utils.cpp:143-206); the sort is stable (std::sort is an insertion sort at these sizes)
template<int R> QG_HD void rp_adjust(RpIv<R> &iv, double tol, double x0, double x1)
{
  if (iv.n == 0) return;
  for (int i = 1; i < iv.n; ++i) {
    const double v = iv.r[i];
    const signed char fv = iv.fid[i];
    int j = i - 1;
    while (j >= 0 && iv.r[j] > v) {
      iv.r[j + 1] = iv.r[j];
      iv.fid[j + 1] = iv.fid[j];
      --j;
    }
    iv.r[j + 1] = v;
    iv.fid[j + 1] = fv;
  }
  double ro[R + 1];
  signed char fo[R + 1];
  const int no = iv.n;
  for (int i = 0; i < no; ++i) {
    ro[i] = iv.r[i];
    fo[i] = iv.fid[i];
  }
  int n = 0;
  iv.r[n] = x0;
  iv.fid[n] = -1;
  n = 1;
  bool has0 = false, has1 = false;
  for (int i = 0; i < no; ++i) {
    double root = ro[i];
    const int f = fo[i];
    if (rp_eq(tol, root, x0)) {
      if (f == -1) {
        has0 = true;
        continue;
      }
      root = x0;
    } else if (rp_eq(tol, root, x1)) {
      if (f == -1) {
        has1 = true;
        continue;
      }
      root = x1;
    } else if (0 < i && rp_eq(tol, root, ro[i - 1])) {
      root = ro[i - 1];
    }
    iv.r[n] = root;  // n <= no <= R here: the array holds R + 1 entries
    iv.fid[n] = (signed char)f;
    ++n;
  }
  if (!has0) {
    for (int i = 1; i < n; ++i) {
      iv.r[i - 1] = iv.r[i];
      iv.fid[i - 1] = iv.fid[i];
    }
    --n;
  }
  if (has1) {
    iv.r[n] = x1;
    iv.fid[n] = -1;
    ++n;
  }
  iv.n = n;
}

template<int N, class F = FuncDesc> struct RpCtx
{
  F f;
  ChainItem it;
  int order;
  const double *nodes;  // the order nodes of [0,1] (equispaced below order 7, modified Chebyshev nodes from there)
  int top;              // top stage of the chain
  bool surf;            // the top stage is the level set itself
  int fixdir;           // facet job: its constant direction, else -1
  RpIv<kRpR0> r0;
  RpIv<kRpR> r1[kRpIv0];
  RpIv<kRpR> r2[kRpIv0 * kRpIv1];
  int keys[kRpCells];
  int ncells;
  long long npts;
  bool emit;
  double *pts;        // [npoints][N]
  long long *conn;    // [ncells][npc]
  long long pt_base, cell_base;
  int npc;
  int err;
};

template<int N> QG_HD double rp_lo(const ChainItem &it, int k)
{
  double v = it.lo[0];
#pragma unroll
  for (int d = 1; d < N; ++d)
    if (d == k) v = it.lo[d];
  return v;
}
template<int N> QG_HD double rp_hi(const ChainItem &it, int k)
{
  double v = it.hi[0];
#pragma unroll
  for (int d = 1; d < N; ++d)
    if (d == k) v = it.hi[d];
  return v;
}
template<int N> QG_HD void rp_setk(double *x, int k, double v)
{
#pragma unroll
  for (int d = 0; d < N; ++d)
    if (d == k) x[d] = v;
}

// compute_roots (impl_reparam_general.cpp:183-199): psi i of stage s at x; appends to out, returns the new count
template<int N, class F> QG_RP_FN int rp_roots(RpCtx<N, F> &c, int s, int i, const double *x_in, double *out, int nr)
{
  const ChainItem &it = c.it;
  const int k = it.dim[s];
  double x[N];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = x_in[d];
  set_nonfree<N>(it, nonfree_mask(it, s, N), psi_sides(it.psi[s][i]), x);
  LineEval<N, F, -1> le;
  le.init(c.f, x, k);
  const double xmin = rp_lo<N>(it, k), xmax = rp_hi<N>(it, k);
  if (s == 0) return root_find_general<N>(le, xmin, xmax, out, nr, &c.err);
  double r;
  if (newton_bisection<N>(le, xmin, xmax, line_tol(xmin, xmax), r)) {
    if (nr < kMaxRoots)
      out[nr++] = r;
    else
      c.err |= ERR_ROOT_CAP;
  }
  return nr;
}

template<int N, class F> QG_HD int rp_npsi(const RpCtx<N, F> &c, int s) { return (c.surf && s == c.top) ? 1 : c.it.npsi[s]; }

// check_active_interval (:210-226)
template<int N, class F> QG_RP_FN bool rp_check_active(RpCtx<N, F> &c, int s, double *x, double x0, double x1)
{
  const ChainItem &it = c.it;
  const int k = it.dim[s];
  if (rp_eq(rp_tol(it, k), x1, x0)) return false;
  bool okay = true;
  rp_setk<N>(x, k, 0.5 * (x0 + x1));
  const int mask = nonfree_mask(it, s, N);
  const int npsi = rp_npsi<N>(c, s);
  for (int i = 0; i < npsi && okay; ++i) {
    set_nonfree<N>(it, mask, psi_sides(it.psi[s][i]), x);
    const int sg = psi_sign(it.psi[s][i]);
    okay = okay && (phi_val<N>(c.f, x) > 0.0 ? (sg >= 0) : (sg <= 0));
  }
  return okay;
}

// compute_all_intervals (:232-264)
template<int N, int R, class F> QG_RP_FN void rp_all_intervals(RpCtx<N, F> &c, int s, const double *x_in, RpIv<R> &iv)
{
  const ChainItem &it = c.it;
  const int k = it.dim[s];
  const double x0 = rp_lo<N>(it, k), x1 = rp_hi<N>(it, k);
  double x[N];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = x_in[d];
  iv.n = 0;
  iv.active = 0;
  rp_add(iv, x0, -1, &c.err);
  rp_add(iv, x1, -1, &c.err);
  const int npsi = rp_npsi<N>(c, s);
  for (int i = 0; i < npsi; ++i) {
    double rt[kMaxRoots];
    const int nr = rp_roots<N>(c, s, i, x, rt, 0);
    for (int j = 0; j < nr; ++j) rp_add(iv, rt[j], i, &c.err);
  }
  rp_adjust(iv, rp_tol(it, k), x0, x1);
  if (!(c.surf && s == c.top)) {
    for (int i = 0; i + 1 < iv.n; ++i)
      if (rp_check_active<N>(c, s, x, iv.r[i], iv.r[i + 1])) iv.active |= 1u << i;
  }
}

// compute_similar_roots (:423-501), including the (sic) root picked by get_ref_intervals_roots (:399-411): the root stored
// at index psi_id, once per root of that psi.  refpoint: the point the reference intervals were computed at.
template<int N, int R, class F>
QG_RP_FN int rp_similar_roots(RpCtx<N, F> &c, int s, const RpIv<R> &ref, const double *refpoint, int i, const double *x_in, double *out)
{
  const ChainItem &it = c.it;
  const int k = it.dim[s];
  int nref = 0;
  for (int j = 0; j < ref.n; ++j)
    if (ref.fid[j] == i) ++nref;
  if (nref == 0) return 0;
  const double refval = ref.r[i];
  double point[N];
#pragma unroll
  for (int d = 0; d < N; ++d) point[d] = x_in[d];
  double roots[kMaxRoots];
  int nr = rp_roots<N>(c, s, i, point, roots, 0);
  const double tol = rp_tol(it, k);
  const double x0 = rp_lo<N>(it, k), x1 = rp_hi<N>(it, k);
  {
    int m = 0;
    for (int j = 0; j < nr; ++j)
      if (!(rp_eq(tol, roots[j], x0) || rp_eq(tol, roots[j], x1))) roots[m++] = roots[j];
    nr = m;
  }
  bool use_ref = false;
  if (nr == nref) {
    // as computed
  } else if (nr > nref) {
    use_ref = true;
  } else {
    set_nonfree<N>(it, nonfree_mask(it, s, N), psi_sides(it.psi[s][i]), point);
    rp_setk<N>(point, k, x0);
    const double v0 = phi_val<N>(c.f, point);
    rp_setk<N>(point, k, x1);
    const double v1 = phi_val<N>(c.f, point);
    bool r0 = false, r1 = false;
    double scale = 1.0;
    for (int t = 0; t < 4; ++t) {
      const double nt = tol * scale;  // tol * pow(10, t)
      scale *= 10.0;
      r0 = fabs(v0) <= nt;
      r1 = fabs(v1) <= nt;
      if (r0 || r1) break;
    }
    if (r0 != r1) {
      for (int j = nr; j < nref && j < kMaxRoots; ++j) roots[j] = r0 ? x0 : x1;
      nr = nref;
    } else if (r0 && r1) {
      if (nr + 2 == nref) {
        roots[nr++] = x0;
        roots[nr++] = x1;
      } else if (nr + 1 == nref) {
        double xx[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xx[d] = refpoint[d];
        set_nonfree<N>(it, nonfree_mask(it, s, N), psi_sides(it.psi[s][i]), xx);
        rp_setk<N>(xx, k, 0.5 * (x0 + refval));
        const bool sign = phi_val<N>(c.f, xx) > 0;
        for (int a = 1; a < nr; ++a) {  // std::sort
          const double v = roots[a];
          int b = a - 1;
          while (b >= 0 && roots[b] > v) {
            roots[b + 1] = roots[b];
            --b;
          }
          roots[b + 1] = v;
        }
        const double xmid = nr == 0 ? x1 : roots[0];
        rp_setk<N>(point, k, 0.5 * (x0 + xmid));
        const bool new_sign = phi_val<N>(c.f, point) > 0;
        roots[nr++] = (sign == new_sign) ? x1 : x0;
      }
    }
    if (nr != nref) use_ref = true;
  }
  if (use_ref) {
    for (int j = 0; j < nref && j < kMaxRoots; ++j) roots[j] = refval;
    nr = nref;
  }
  for (int j = 0; j < nr; ++j) out[j] = roots[j];
  return nr;
}

// compute_similar_intervals (:513-551)
template<int N, int R, class F>
QG_RP_FN void rp_similar_intervals(RpCtx<N, F> &c, int s, const RpIv<R> &ref, const double *refpoint, const double *x, RpIv<R> &iv)
{
  const bool surf = c.surf && s == c.top;
  iv.n = 0;
  iv.active = 0;
  if (ref.n == 2) {
    if (!surf) iv = ref;
    return;
  }
  const int npsi = rp_npsi<N>(c, s);
  for (int i = 0; i < npsi; ++i) {
    double rt[kMaxRoots];
    const int nr = rp_similar_roots<N>(c, s, ref, refpoint, i, x, rt);
    for (int j = 0; j < nr; ++j) rp_add(iv, rt[j], i, &c.err);
  }
  const int k = c.it.dim[s];
  const double x0 = rp_lo<N>(c.it, k), x1 = rp_hi<N>(c.it, k);
  if (!surf) {
    rp_add(iv, x0, -1, &c.err);
    rp_add(iv, x1, -1, &c.err);
    iv.active = ref.active;
  }
  rp_adjust(iv, rp_tol(c.it, k), x0, x1);
}

// insert_point_reparam (:314-329) + ReparamMesh::insert_cell_point (reparam_mesh.cpp:489-498)
template<int N, class F> QG_HD void rp_insert(RpCtx<N, F> &c, const double *x, const int *idx, const int *pt, int skip_dir)
{
  const int key = idx[0] | (idx[1] << 8) | (idx[2] << 16);
  int cell = -1;
  for (int i = 0; i < c.ncells; ++i)
    if (c.keys[i] == key) cell = i;
  if (cell < 0) {
    if (c.ncells < kRpCells) {
      cell = c.ncells;
      c.keys[c.ncells++] = key;
    } else {
      c.err |= ERR_RP_CAP;
      return;
    }
  }
  if (c.emit) {
    int flat = 0, stride = 1;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (d == skip_dir) continue;
      flat += pt[d] * stride;
      stride *= c.order;
    }
    const long long gp = c.pt_base + c.npts;
#pragma unroll
    for (int d = 0; d < N; ++d) c.pts[gp * N + d] = x[d];
    c.conn[(c.cell_base + cell) * c.npc + flat] = gp;
  }
  ++c.npts;
}

template<int N, class F> QG_HD RpIv<kRpR> *rp_ref_slot(RpCtx<N, F> &c, int s, const int *idx)
{
  if (s == 1) return &c.r1[idx[0]];
  return &c.r2[idx[0] * kRpIv1 + idx[1]];
}

// the point the reference intervals of stage s were computed at: lower stages at the midpoints of their reference intervals
// (or of the box for tensor-product stages)
template<int N, class F> QG_HD void rp_refpoint(RpCtx<N, F> &c, int s, const int *idx, double *x)
{
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = 0.0;
  for (int t = 0; t < s; ++t) {
    const int k = c.it.dim[t];
    double v;
    if (c.it.kind[t] == ST_PRESET) {
      v = (rp_lo<N>(c.it, k) + rp_hi<N>(c.it, k)) * 0.5;
    } else if (t == 0) {
      v = 0.5 * (c.r0.r[idx[0]] + c.r0.r[idx[0] + 1]);
    } else {
      const RpIv<kRpR> &r = c.r1[idx[0]];
      v = 0.5 * (r.r[idx[1]] + r.r[idx[1] + 1]);
    }
    rp_setk<N>(x, k, v);
  }
}

// compute_ref_intervals (:270-305), stage S upwards
template<int N, int S, class F> QG_HD void rp_ref(RpCtx<N, F> &c, const double *x_in, const int *idx_in)
{
  double x[N];
  int idx[3];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = x_in[d];
  for (int d = 0; d < 3; ++d) idx[d] = idx_in[d];
  const int k = c.it.dim[S];
  unsigned active;
  int n;
  if constexpr (S == 0) {
    rp_all_intervals<N>(c, 0, x, c.r0);
    active = c.r0.active;
    n = c.r0.n;
  } else {
    RpIv<kRpR> *iv = rp_ref_slot<N>(c, S, idx);
    rp_all_intervals<N>(c, S, x, *iv);
    active = iv->active;
    n = iv->n;
  }
  if (S == c.top) return;
  if constexpr (S < 2) {
    for (int i = 0; i + 1 < n; ++i) {
      if (!(active & (1u << i))) continue;
      if (i >= (S == 0 ? kRpIv0 : kRpIv1)) {
        c.err |= ERR_RP_CAP;
        continue;
      }
      idx[S] = i;
      double a, b;
      if constexpr (S == 0) {
        a = c.r0.r[i];
        b = c.r0.r[i + 1];
      } else {
        const RpIv<kRpR> *iv = rp_ref_slot<N>(c, S, idx_in);
        a = iv->r[i];
        b = iv->r[i + 1];
      }
      rp_setk<N>(x, k, 0.5 * (a + b));
      rp_ref<N, S + 1>(c, x, idx);
    }
  }
}

// process (:587-650), stage S upwards
template<int N, int S, class F> QG_HD void rp_process(RpCtx<N, F> &c, const double *x_in, const int *idx_in, const int *pt_in)
{
  double x[N];
  int idx[3], pt[3];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = x_in[d];
  for (int d = 0; d < 3; ++d) {
    idx[d] = idx_in[d];
    pt[d] = pt_in[d];
  }
  const int k = c.it.dim[S];
  const bool surf = c.surf && S == c.top;
  // this point's intervals
  double rr[kRpR0 + 1];
  unsigned active;
  int n;
  if constexpr (S == 0) {
    n = c.r0.n;
    active = c.r0.active;
    for (int i = 0; i < n; ++i) rr[i] = c.r0.r[i];
  } else {
    const RpIv<kRpR> *ref = rp_ref_slot<N>(c, S, idx);
    double refpoint[N];
    rp_refpoint<N>(c, S, idx, refpoint);
    RpIv<kRpR> iv;
    rp_similar_intervals<N>(c, S, *ref, refpoint, x, iv);
    n = iv.n;
    active = iv.active;
    for (int i = 0; i < n; ++i) rr[i] = iv.r[i];
  }
  if (surf) {
    for (int i = 0; i < n; ++i) {
      rp_setk<N>(x, k, rr[i]);
      idx[S] = 0;
      rp_insert<N>(c, x, idx, pt, k);
    }
    return;
  }
  for (int i = 0; i + 1 < n; ++i) {
    if (!(active & (1u << i))) continue;
    if (S != c.top && i >= (S == 0 ? kRpIv0 : kRpIv1)) continue;  // flagged by rp_ref
    const double x0 = rr[i], x1 = rr[i + 1];
    idx[S] = i;
    for (int j = 0; j < c.order; ++j) {
      rp_setk<N>(x, k, x0 + (x1 - x0) * c.nodes[j]);
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (d == k) pt[d] = j;
      if (S == c.top) {
        if (c.fixdir >= 0) {
          set_nonfree<N>(c.it, nonfree_mask(c.it, S, N), psi_sides(c.it.psi[S][0]), x);
          rp_insert<N>(c, x, idx, pt, c.fixdir);
        } else {
          rp_insert<N>(c, x, idx, pt, -1);
        }
      } else {
        if constexpr (S < 2) rp_process<N, S + 1>(c, x, idx, pt);
      }
    }
  }
}

// One chain item = one leaf of the reference's recursion.  Returns through c.ncells / c.npts; with c.emit writes points / cells.
template<int N, class F> QG_HD void rp_leaf(RpCtx<N, F> &c)
{
  const ChainItem &it = c.it;
  c.ncells = 0;
  c.npts = 0;
  int top = 0;
  for (int s = 0; s < 3; ++s)
    if (it.kind[s] != ST_PASS) top = s;
  c.top = top;
  c.surf = it.kind[top] == ST_SURF;
  int m = 0;  // leading tensor-product stages
  while (m <= top && it.kind[m] == ST_PRESET) ++m;
  const int zero[3] = { 0, 0, 0 };
  if (m > top) {
    // tensor_product_reparam_full (:331-362): the whole box (cell) or the whole face
    double x[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = 0.0;
    if (c.fixdir >= 0) set_nonfree<N>(it, 1 << c.fixdir, psi_sides(it.psi[top][0]), x);
    int npt = 1;
    for (int s = 0; s <= top; ++s) npt *= c.order;
    for (int fl = 0; fl < npt; ++fl) {
      int pt[3] = { 0, 0, 0 };
      int r = fl;
      for (int s = 0; s <= top; ++s) {  // stages hold the free directions in ascending order
        const int d = it.dim[s];
        const int t = r % c.order;
        r /= c.order;
#pragma unroll
        for (int e = 0; e < 3; ++e)
          if (e == d) pt[e] = t;
        const double lo = rp_lo<N>(it, d), hi = rp_hi<N>(it, d);
        rp_setk<N>(x, d, lo + (hi - lo) * c.nodes[t]);
      }
      rp_insert<N>(c, x, zero, pt, c.fixdir);
    }
    return;
  }
  if (m == 0) {
    double x[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = 0.0;
    rp_ref<N, 0>(c, x, zero);
    rp_process<N, 0>(c, x, zero, zero);
    return;
  }
  // tensor_product_reparam_base (:365-391): stages 0..m-1 are tensor-product directions of the box
  double xm[N];
#pragma unroll
  for (int d = 0; d < N; ++d) xm[d] = 0.0;
  for (int s = 0; s < m; ++s) {
    const int d = it.dim[s];
    rp_setk<N>(xm, d, (rp_lo<N>(it, d) + rp_hi<N>(it, d)) * 0.5);
  }
  if (m == 1)
    rp_ref<N, 1>(c, xm, zero);
  else
    rp_ref<N, 2>(c, xm, zero);
  int npt = 1;
  for (int s = 0; s < m; ++s) npt *= c.order;
  for (int fl = 0; fl < npt; ++fl) {
    double x[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = 0.0;
    int pt[3] = { 0, 0, 0 };
    int r = fl;
    for (int s = 0; s < m; ++s) {
      const int d = it.dim[s];
      const int t = r % c.order;
      r /= c.order;
#pragma unroll
      for (int e = 0; e < 3; ++e)
        if (e == d) pt[e] = t;
      const double lo = rp_lo<N>(it, d), hi = rp_hi<N>(it, d);
      rp_setk<N>(x, d, lo + (hi - lo) * c.nodes[t]);
    }
    if (m == 1)
      rp_process<N, 1>(c, x, zero, pt);
    else
      rp_process<N, 2>(c, x, zero, pt);
  }
}

// ---- host tables ---------------------------------------------------------------------------------------------
constexpr int kChebyshevOrder = 7;  // reparam_mesh.hpp:53

// algoim::bernstein::modifiedChebyshevNode (not in tree): the Chebyshev roots rescaled so that the end nodes are 0 and 1
inline double modified_chebyshev_node(int i, int n)
{
  if (n == 1) return 0.5;
  const double tn = 3.141592653589793238462643383279502884 / (2 * n);
  return 0.5 - 0.5 * std::cos((2 * i + 1) * tn) / std::cos(tn);
}
inline std::vector<double> reparam_nodes(int order)
{
  std::vector<double> v((size_t)order);
  for (int i = 0; i < order; ++i)
    v[(size_t)i] = order >= kChebyshevOrder ? modified_chebyshev_node(i, order) : double(i) / double(order - 1);
  return v;
}
// Lagrange basis and first derivatives at the midpoint of [0,1]^pd on (sic) the modified Chebyshev nodes, flat index with
// direction 0 fastest (lagrange_tp_utils.cpp:38-150 as called from impl_reparam_mesh.cpp:188-193, 236-243: chebyshev = false)
inline void lagrange_mid(int pd, int order, std::vector<double> &basis, std::vector<double> &ders)
{
  const double x = 0.5;
  std::vector<double> c((size_t)order), v, d;
  for (int i = 0; i < order; ++i) c[(size_t)i] = modified_chebyshev_node(i, order);
  for (int i = 0; i < order; ++i) {
    double prod = 1.0;
    for (int j = 0; j < order; ++j)
      if (i != j) prod *= (x - c[(size_t)j]) / (c[(size_t)i] - c[(size_t)j]);
    v.push_back(prod);
  }
  for (int i = 0; i < order; ++i) {
    double sum = 0.0;
    for (int j = 0; j < order; ++j) {
      if (i == j) continue;
      double term = 1.0 / (c[(size_t)i] - c[(size_t)j]);
      for (int k = 0; k < order; ++k)
        if (k != i && k != j) term *= (x - c[(size_t)k]) / (c[(size_t)i] - c[(size_t)k]);
      sum += term;
    }
    d.push_back(sum);
  }
  size_t npc = 1;
  for (int i = 0; i < pd; ++i) npc *= (size_t)order;
  basis.assign(npc, 1.0);
  ders.assign(npc * (size_t)pd, 1.0);
  for (size_t f = 0; f < npc; ++f) {
    int t[3] = { 0, 0, 0 };
    size_t r = f;
    for (int i = 0; i < pd; ++i) {
      t[i] = (int)(r % (size_t)order);
      r /= (size_t)order;
    }
    for (int i = 0; i < pd; ++i) basis[f] *= v[(size_t)t[i]];
    for (int dd = 0; dd < pd; ++dd)
      for (int i = 0; i < pd; ++i) ders[(size_t)dd * npc + f] *= (i == dd) ? d[(size_t)t[i]] : v[(size_t)t[i]];
  }
}

// ---- mesh passes ------------------------------------------------------------------------------------------------
struct RpMeshArgs
{
  FuncDesc f;
  int dim, pd, order, npc;
  int mode;            // 0 volume, 1 level set, 2 + f: face f of the box
  long long ncells;    // cells produced by the cut cells
  const double *pts;
  long long *conn;
  const int *cell_job;     // [ncells] job of each cell
  const double *job_box;   // [njobs][6]
  const double *basis;     // [npc] Lagrange basis at the cell midpoint
  const double *ders;      // [pd][npc]
  double tol;              // wirebasket tolerance (1e4 eps)
  int *edge_flag;          // [ncells * nedges]
  const long long *edge_off;
  long long *wires;
};

QG_HD int rp_flat(const int *t, int pd, int order)
{
  int f = 0, s = 1;
  for (int d = 0; d < pd; ++d) {
    f += t[d] * s;
    s *= order;
  }
  return f;
}

// permute_cell_directions (reparam_mesh.cpp:759-785)
QG_HD void rp_permute(long long *cc, int pd, int order)
{
  if (pd == 1) {
    for (int i = 0; i < order / 2; ++i) {
      const long long t = cc[i];
      cc[i] = cc[order - 1 - i];
      cc[order - 1 - i] = t;
    }
    return;
  }
  int t[3] = { 0, 0, 0 };
  for (t[2] = 0; t[2] < (pd > 2 ? order : 1); ++t[2])
    for (t[1] = 0; t[1] < order; ++t[1])
      for (t[0] = 0; t[0] < order; ++t[0]) {
        if (t[0] <= t[1]) continue;
        const int u[3] = { t[1], t[0], t[2] };
        const int a = rp_flat(t, pd, order), b = rp_flat(u, pd, order);
        const long long v = cc[a];
        cc[a] = cc[b];
        cc[b] = v;
      }
}

// orient_cells_positively / orient_levelset_cells_positively / orient_facet_cells_positively (impl_reparam_mesh.cpp:183-330)
template<int N> QG_HD void rp_orient_body(long long cell, const RpMeshArgs &a)
{
  if (cell >= a.ncells) return;
  long long *cc = a.conn + cell * a.npc;
  double J[3][N], ep[N];
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int d = 0; d < N; ++d) J[r][d] = 0.0;
#pragma unroll
  for (int d = 0; d < N; ++d) ep[d] = 0.0;
  for (int i = 0; i < a.npc; ++i) {
    const double *p = a.pts + cc[i] * N;
    for (int r = 0; r < a.pd; ++r)
#pragma unroll
      for (int d = 0; d < N; ++d) J[r][d] += p[d] * a.ders[r * a.npc + i];
#pragma unroll
    for (int d = 0; d < N; ++d) ep[d] += p[d] * a.basis[i];
  }
  bool flip;
  if (a.mode == 0) {
    double det;
    if (N == 3)
      det = (J[0][0] * (J[1][1] * J[2][N - 1] - J[1][N - 1] * J[2][1])) + (J[0][1] * (J[1][N - 1] * J[2][0] - J[1][0] * J[2][N - 1]))
            + (J[0][N - 1] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]));
    else
      det = (J[0][0] * J[1][1]) - (J[0][1] * J[1][0]);
    flip = det < 0;
  } else {
    double nrm[N];
    if (a.pd == 2) {
      nrm[0] = (J[0][1] * J[1][N - 1]) - (J[0][N - 1] * J[1][1]);
      nrm[1] = (J[0][N - 1] * J[1][0]) - (J[0][0] * J[1][N - 1]);
      nrm[N - 1] = (J[0][0] * J[1][1]) - (J[0][1] * J[1][0]);
    } else {
      nrm[0] = J[0][1];
      nrm[1] = -J[0][0];
    }
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < N; ++d) s += nrm[d] * nrm[d];
    s = sqrt(s);
#pragma unroll
    for (int d = 0; d < N; ++d) nrm[d] /= s;
    double tn[N];
    if (a.mode == 1) {
      phi_grad<N>(a.f, ep, tn);
      double q = 0.0;
#pragma unroll
      for (int d = 0; d < N; ++d) q += tn[d] * tn[d];
      q = sqrt(q);
#pragma unroll
      for (int d = 0; d < N; ++d) tn[d] /= q;
    } else {
      const int facet = a.mode - 2;
#pragma unroll
      for (int d = 0; d < N; ++d) tn[d] = (d == facet / 2) ? ((facet % 2) == 0 ? -1.0 : 1.0) : 0.0;
    }
    double dot = 0.0;
#pragma unroll
    for (int d = 0; d < N; ++d) dot += nrm[d] * tn[d];
    flip = dot < 0;
  }
  if (flip) rp_permute(cc, a.pd, a.order);
}

// get_edge_points + sort_edge_points (reparam_mesh.cpp:516-571, edge numbering impl_utils.cpp:59-87)
QG_HD void rp_edge_points(const long long *cc, int pd, int order, int edge, long long *out)
{
  int cdirs[2] = { 0, 0 }, sides[2] = { 0, 0 };
  if (pd == 2) {
    cdirs[0] = edge / 2;
    sides[0] = edge % 2;
  } else {
    if (edge < 4) {
      cdirs[0] = 0;
      cdirs[1] = 1;
    } else if (edge < 8) {
      cdirs[0] = 0;
      cdirs[1] = 2;
    } else {
      cdirs[0] = 1;
      cdirs[1] = 2;
    }
    sides[0] = edge % 2;
    sides[1] = (edge / 2) % 2;
  }
  int lo[3] = { 0, 0, 0 }, hi[3] = { order, order, order };
  for (int i = 0; i < pd - 1; ++i) {
    if (sides[i] == 0)
      hi[cdirs[i]] = 1;
    else
      lo[cdirs[i]] = order - 1;
  }
  int n = 0;
  int t[3];
  for (t[2] = (pd > 2 ? lo[2] : 0); t[2] < (pd > 2 ? hi[2] : 1); ++t[2])
    for (t[1] = lo[1]; t[1] < hi[1]; ++t[1])
      for (t[0] = lo[0]; t[0] < hi[0]; ++t[0]) out[n++] = cc[rp_flat(t, pd, order)];
  bool reverse = out[order - 1] < out[0];
  if (out[0] == out[order - 1] && order > 2) reverse = out[order - 2] < out[1];
  if (reverse)
    for (int i = 0; i < order / 2; ++i) {
      const long long v = out[i];
      out[i] = out[order - 1 - i];
      out[order - 1 - i] = v;
    }
}

template<int N> QG_HD bool rp_edge_in_subdomain(const RpMeshArgs &a, const long long *ep, const double *box, int sub_dim)
{
  int b0[N];
  for (int e = 0; e < a.order; ++e) {
    const double *p = a.pts + ep[e] * N;
    int nb = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      int b = -1;
      if (rp_eq(a.tol, p[d], box[d]))
        b = 0;
      else if (rp_eq(a.tol, p[d], box[3 + d]))
        b = 1;
      if (e == 0) b0[d] = b;
      if (b0[d] >= 0 && b0[d] == b) ++nb;
    }
    if (nb < sub_dim) return false;
  }
  return true;
}

// generate_wirebasket (impl_reparam_mesh.cpp:100-178) for (cell, edge) t; EMIT: writes the edge at edge_off[t]
template<int N, bool EMIT> QG_HD void rp_wire_body(long long t, const RpMeshArgs &a)
{
  const int ne = a.pd == 2 ? 4 : 12;
  if (a.pd < 2 || t >= a.ncells * ne) return;
  const long long cell = t / ne;
  const int edge = (int)(t - cell * ne);
  if (EMIT && !a.edge_flag[t]) return;
  long long ep[kRpMaxOrder];
  rp_edge_points(a.conn + cell * a.npc, a.pd, a.order, edge, ep);
  if (EMIT) {
    const long long o = a.edge_off[t] * a.order;
    for (int e = 0; e < a.order; ++e) a.wires[o + e] = ep[e];
    return;
  }
  const double *box = a.job_box + 6 * (size_t)a.cell_job[cell];
  const double *p0 = a.pts + ep[0] * N;
  bool degenerate = true;
  for (int e = 0; e < a.order; ++e) {
    const double *p = a.pts + ep[e] * N;
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (!rp_eq(a.tol, p0[d], p[d])) degenerate = false;
  }
  bool add = false;
  if (!degenerate) {
    add = rp_edge_in_subdomain<N>(a, ep, box, N - 1);
    if (!add) {
      const bool on_face = rp_edge_in_subdomain<N>(a, ep, box, 1);
      if (on_face || N == 2) {
        // on a face of the box: every point on the level set; off the faces: range-1 functions must vanish (one is given)
        add = true;
        for (int e = 0; e < a.order; ++e) add = add && fabs(phi_val<N>(a.f, a.pts + ep[e] * N)) <= a.tol;
      }
    }
  }
  a.edge_flag[t] = add ? 1 : 0;
}

// add_full_cell with its wirebasket (reparam_mesh.cpp:352-390): full cell i of the list
template<int N>
QG_HD void rp_full_cell_body(long long i, long long nfull, const long long *cells, const GridDesc &g, int order, const double *nodes,
  long long pt_base, long long cell_base, long long wire_base, double *pts, long long *conn, long long *wires)
{
  if (i >= nfull) return;
  double lo[3], hi[3];
  cell_box<N>(g, cells[i], lo, hi);
  int npc = 1;
  for (int d = 0; d < N; ++d) npc *= order;
  const long long p0 = pt_base + i * npc;
  long long *cc = conn + (cell_base + i) * npc;
  for (int fl = 0; fl < npc; ++fl) {
    int r = fl;
    for (int d = 0; d < N; ++d) {
      const int t = r % order;
      r /= order;
      pts[(p0 + fl) * N + d] = lo[d] + (hi[d] - lo[d]) * nodes[t];
    }
    cc[fl] = p0 + fl;
  }
  const int ne = N == 2 ? 4 : 12;
  for (int e = 0; e < ne; ++e) rp_edge_points(cc, N, order, e, wires + (wire_base + i * ne + e) * order);
}

}  // namespace qg
