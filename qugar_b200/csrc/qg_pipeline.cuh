// qugar_b200 -- per-thread bodies of the quadrature pipeline passes (see qg_quad.cuh for the design).
// Each *_body(tid, args) is what one CUDA thread does; qg_kernels_pipeline.cuh launches them.
#pragma once
#include "qg_quad.cuh"

namespace qg {

// stage-1 call record: one per stage-0 quadrature point
struct Call1
{
  int item;
  int nent;
  double x0;  // coordinate along the stage-0 direction
  double w;   // accumulated weight
  double ent[6];
};
// stage-2 call record: one per stage-1 quadrature point.  Self-contained for K3 (no chain-item lookup):
// kd = kind of stages 0..2 (bits 0-5, two each) | their directions (bits 6-11, two each) | entries (bits 12-15)
struct alignas(16) Call2
{
  int job;
  int kd;
  double x0, x1;  // coordinates along the stage-0 / stage-1 directions
  double w;
  double ent[4];
};
QG_HD int kd_pack(const ChainItem &it, int nent)
{
  return it.kind[0] | (it.kind[1] << 2) | (it.kind[2] << 4) | ((it.dim[0] & 3) << 6) | ((it.dim[1] & 3) << 8)
         | ((it.dim[2] & 3) << 10) | (nent << 12);
}
QG_HD int kd_kind(int kd, int s) { return (kd >> (2 * s)) & 3; }
QG_HD int kd_dim(int kd, int s) { return (kd >> (6 + 2 * s)) & 3; }

struct PipeArgs
{
  FuncDesc f;
  GridDesc g;
  int p;              // Gauss points per direction
  int sink_mode;      // SinkMode
  const double *gxw;  // (x,w) pairs of the p-point rule on [0,1]
  // jobs
  long long njobs;
  const long long *job_cell;
  const int *job_type;
  double *job_box;         // [njobs][6]: lo[3], hi[3]
  unsigned char *job_cap;  // [njobs] or null: 1 if the job's walk ran into the 16-level bisection cap (CTL count pass)
  // chain items
  int *item_cnt;           // [njobs] items per job
  const long long *item_off;  // [njobs+1] exclusive scan
  ChainItem *items;
  long long nitems;
  // stage 0
  double *ent0;            // [nitems][kMaxSeg0][2]
  int *nent0;              // [nitems]
  int *cnt0;               // [nitems] stage-1 calls per item
  const long long *off0;   // [nitems+1]
  // stage 1
  long long ncall1;
  Call1 *call1;
  int *cnt1;
  const long long *off1;
  // stage 2
  long long ncall2;
  Call2 *call2;
  int *cnt2;
  const long long *off2;
  // output rule
  long long npts;
  double *pts, *wts, *nrm;
  // packed interior rule (transport encoding of the host API, see k3_cell3_kernel<true>): per group of p nodes
  // the three reference coordinates with the slot of the line direction unused, that direction, and per node
  // the coordinate along the line
  double *pack_pc;
  unsigned char *pack_k2;
  double *pack_pk;
  int *err;
};

// largest i in [0,n) with off[i] <= t   (off is a non-decreasing exclusive scan with off[n] = total)
QG_HD long long find_owner(const long long *off, long long n, long long t)
{
  long long lo = 0, hi = n;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (off[mid] <= t)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// CTL count / emit
template<int N, bool EMIT, class F> QG_HD void ctl_body(long long job, const PipeArgs &a, const F &f)
{
  if (job >= a.njobs) return;
  double lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
  cell_box<N>(a.g, a.job_cell[job], lo, hi);
  int err = 0;
  if (EMIT) {
    const long long o = a.item_off[job];
    const int cap = (int)(a.item_off[job + 1] - o);
    ctl_walk<N>(f, lo, hi, a.job_type[job], (int)job, a.items + o, cap, &err);
  } else {
    for (int d = 0; d < 3; ++d) {
      a.job_box[6 * job + d] = lo[d];
      a.job_box[6 * job + 3 + d] = hi[d];
    }
    bool hit = false;
    a.item_cnt[job] = ctl_walk<N>(f, lo, hi, a.job_type[job], (int)job, (ChainItem *)0, 0, &err, &hit);
    if (a.job_cap) a.job_cap[job] = hit ? 1 : 0;
  }
  if (err) flag_or(a.err, err);
}

// K0: stage 0 of every chain item
template<int N, class F> QG_HD void k0_body(long long i, const PipeArgs &a, const F &f)
{
  if (i >= a.nitems) return;
  const ChainItem it = a.items[i];
  double x[N];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = 0.0;
  double ent[2 * kMaxSeg0];
  int err = 0;
  const int n = stage_eval<N, 0, kMaxSeg0>(f, it, x, ent, &err);
  for (int k = 0; k < 2 * n; ++k) a.ent0[(size_t)i * 2 * kMaxSeg0 + k] = ent[k];
  a.nent0[i] = n;
  a.cnt0[i] = stage_count(it.kind[0], n, a.p);
  if (err) flag_or(a.err, err);
}

// K1: stage 1 at stage-0 quadrature point t = r-th child of chain item i
template<int N, class F> QG_HD void k1_work(long long i, int r, long long t, const PipeArgs &a, const F &f)
{
  const ChainItem &it = a.items[i];
  double x[N];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = 0.0;
  double w = 1.0;
  stage_child<N>(f, it, 0, a.ent0 + (size_t)i * 2 * kMaxSeg0, r, a.p, a.gxw, x, w);
  Call1 c;
  c.item = (int)i;
  double x0 = 0.0;
  if (it.kind[0] != ST_PASS) {
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == it.dim[0]) x0 = x[d];
  }
  c.x0 = x0;
  c.w = w;
  int err = 0;
  c.nent = stage_eval<N, 1, 3>(f, it, x, c.ent, &err);
  for (int k = 2 * c.nent; k < 6; ++k) c.ent[k] = 0.0;
  a.call1[t] = c;
  a.cnt1[t] = stage_count(it.kind[1], c.nent, a.p);
  if (err) flag_or(a.err, err);
}
template<int N, class F> QG_HD void k1_body(long long t, const PipeArgs &a, const F &f)
{
  if (t >= a.ncall1) return;
  const long long i = find_owner(a.off0, a.nitems, t);
  k1_work<N>(i, (int)(t - a.off0[i]), t, a, f);
}

// K2: stage 2 at stage-1 quadrature point u = r-th child of stage-1 call t
// (c1, it: the stage-1 call and its chain item, in global memory or staged in shared memory by the kernel)
template<int N, class F>
QG_HD void k2_work_rec(const Call1 &c1, const ChainItem &it, int r, long long u, const PipeArgs &a, const F &f)
{
  double x[N];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = 0.0;
  if (it.kind[0] != ST_PASS) {
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == it.dim[0]) x[d] = c1.x0;
  }
  double w = c1.w;
  stage_child<N>(f, it, 1, c1.ent, r, a.p, a.gxw, x, w);
  Call2 c;
  c.job = it.job;
  c.x0 = c1.x0;
  double x1 = 0.0;
  if (it.kind[1] != ST_PASS) {
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == it.dim[1]) x1 = x[d];
  }
  c.x1 = x1;
  c.w = w;
  int err = 0;
  const int nent = stage_eval<N, 2, 2>(f, it, x, c.ent, &err);
#pragma unroll
  for (int k = 0; k < 4; ++k) c.ent[k] = k >= 2 * nent ? 0.0 : c.ent[k];  // selects: a run-time index would put c in local memory
  c.kd = kd_pack(it, nent);
  a.call2[u] = c;
  a.cnt2[u] = stage_count(it.kind[2], nent, a.p);
  if (err) flag_or(a.err, err);
}
template<int N, class F> QG_HD void k2_work(long long t, int r, long long u, const PipeArgs &a, const F &f)
{
  const Call1 &c1 = a.call1[t];
  k2_work_rec<N>(c1, a.items[c1.item], r, u, a, f);
}
template<int N, class F> QG_HD void k2_body(long long u, const PipeArgs &a, const F &f)
{
  if (u >= a.ncall2) return;
  const long long t = find_owner(a.off1, a.ncall1, u);
  k2_work<N>(t, (int)(u - a.off1[t]), u, a, f);
}

// Sink: one finished node (x, w) of job `job` -> slot q of the rule.
// CELL:  CutCellsQuadWrapper::evalIntegrand      (impl_quadrature.cpp:65-69)
// SURF:  CutUnfBoundsQuadWrapper::evalIntegrand  (impl_quadrature.cpp:87-111)
// FACET: CutIsoBoundsQuadWrapper::evalIntegrand  (impl_quadrature.cpp:130-135)
// RAW:   algoim::QuadratureRule node (x, w), used by the classification passes
// Values of one node: pv[pdim] point coordinates, wv weight, nv[N] normal (SURF only).  Returns pdim.
// lo/hi: the job's cell box; jt: its job type; g_in: phi gradient at x if the caller has it already (have_g).
template<int N, class F>
QG_HD int sink_compute(int sink_mode, const F &f, const double *lo, const double *hi, int jt, double *x, double w,
  const double *g_in, bool have_g, double *pv, double &wv, double *nv)
{
  if (sink_mode == SINK_RAW) {
    if (jt >= 0) {
      const int cd = jt >> 1;
#pragma unroll
      for (int d = 0; d < N; ++d)
        if (d == cd) x[d] = (jt & 1) ? hi[d] : lo[d];
    }
#pragma unroll
    for (int d = 0; d < N; ++d) pv[d] = x[d];
    wv = w;
    return N;
  }
  if (sink_mode == SINK_FACET) {
    const int cd = jt >> 1;
    double fvol = 1.0;
    int k = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (d != cd) {
        fvol *= (hi[d] - lo[d]);
        const double v = (x[d] - lo[d]) / (hi[d] - lo[d]);
        if (k == 0) pv[0] = v;
        if (N == 3 && k == 1) pv[N - 2] = v;
        ++k;
      }
    }
    wv = w / fvol;
    return N - 1;
  }
  double vol = 1.0;
#pragma unroll
  for (int d = 0; d < N; ++d) {
    vol *= (hi[d] - lo[d]);
    pv[d] = (x[d] - lo[d]) / (hi[d] - lo[d]);
  }
  if (sink_mode == SINK_CELL) {
    wv = w / vol;
    return N;
  }
  // SURF
  double g[N];
  if (have_g) {
#pragma unroll
    for (int d = 0; d < N; ++d) g[d] = g_in[d];
  } else {
    phi_grad<N>(f, x, g);
  }
  double s = g[0] * g[0];
#pragma unroll
  for (int d = 1; d < N; ++d) s += g[d] * g[d];
  const double nn = sqrt(s);
#pragma unroll
  for (int d = 0; d < N; ++d) g[d] /= nn;
#pragma unroll
  for (int d = 0; d < N; ++d) g[d] *= (hi[d] - lo[d]);
  double s2 = g[0] * g[0];
#pragma unroll
  for (int d = 1; d < N; ++d) s2 += g[d] * g[d];
  const double n2 = sqrt(s2);
#pragma unroll
  for (int d = 0; d < N; ++d) g[d] /= n2;
  wv = w * n2 / vol;
#pragma unroll
  for (int d = 0; d < N; ++d) nv[d] = g[d];
  return N;
}

template<int N, class F> QG_HD void sink_point(const PipeArgs &a, const F &f, int job, double *x, double w, long long q)
{
  double pv[N], nv[N], wv;
  const double *lo = a.job_box + 6 * (size_t)job;
  const int pd = sink_compute<N>(a.sink_mode, f, lo, lo + 3, a.job_type[job], x, w, x, false, pv, wv, nv);
  for (int d = 0; d < pd; ++d) a.pts[q * pd + d] = pv[d];
  a.wts[q] = wv;
  if (a.sink_mode == SINK_SURF) {
#pragma unroll
    for (int d = 0; d < N; ++d) a.nrm[q * N + d] = nv[d];
  }
}

// r-th node of stage-2 line u: coordinates and weight before the sink
template<int N, class F> QG_HD void line_node(const PipeArgs &a, const F &f, const Call2 &c, int r, double *x, double &w)
{
  const int kd = c.kd;
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = 0.0;
  if (kd_kind(kd, 0) != ST_PASS) {
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == kd_dim(kd, 0)) x[d] = c.x0;
  }
  if (kd_kind(kd, 1) != ST_PASS) {
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == kd_dim(kd, 1)) x[d] = c.x1;
  }
  w = c.w;
  stage_child_k<N>(f, kd_kind(kd, 2), kd_dim(kd, 2), c.ent, r, a.p, a.gxw, x, w, (double *)0);
}

// K3: the nodes of every stage-2 line
template<int N, class F> QG_HD void k3_body(long long u, const PipeArgs &a, const F &f, const long long *job_shift)
{
  if (u >= a.ncall2) return;
  const int n = a.cnt2[u];
  if (n == 0) return;
  const Call2 c = a.call2[u];
  // job_shift relocates the job's nodes to its place in the final rule
  const long long q0 = a.off2[u] + (job_shift ? job_shift[c.job] : 0);
  for (int r = 0; r < n; ++r) {
    double x[N], w;
    line_node<N>(a, f, c, r, x, w);
    sink_point<N>(a, f, c.job, x, w, q0 + r);
  }
}

// points per job: difference of point offsets at the job's first stage-2 call and the next job's
QG_HD void job_points_body(long long job, const PipeArgs &a, int *n_per_job, long long *job_pt_off)
{
  if (job > a.njobs) return;
  // first item of job -> first call1 -> first call2 -> first point
  const long long i = a.item_off[job];
  const long long t = (i >= a.nitems) ? a.ncall1 : a.off0[i];
  const long long u = (t >= a.ncall1) ? a.ncall2 : a.off1[t];
  const long long q = (u >= a.ncall2) ? a.npts : a.off2[u];
  job_pt_off[job] = q;
  (void)n_per_job;
}

}  // namespace qg
