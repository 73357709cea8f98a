// qugar_b200 -- kernels of the domain construction: kd-tree sign pruning, cell / facet classification
// (part of the translation unit qg_engine.cu; included there, in this order)
#pragma once

namespace qg {

// ---- kd-tree sign pruning (create_decomposition, impl_unfitted_domain.cpp:441-469) --------------------
struct KdNode
{
  int lo[3], hi[3];
};
constexpr unsigned char ST_CUT = 0, ST_FULL = 1, ST_EMPTY = 2, ST_UNDET = 3;

// One kd-tree node (create_decomposition, impl_unfitted_domain.cpp:441-469): uniform sign -> paint list; single
// undetermined cell -> marked; else split.  Returns false if a list is full (cap / ucap).
template<int N>
__device__ __forceinline__ bool kd_process_node(const FuncDesc &f, const GridDesc &g, const KdNode &nd, KdNode *next,
  int *nnext, int cap, KdNode *uniform, unsigned char *uniform_sign, int *nuniform, int ucap, unsigned char *status, int k0,
  int k1)
{
  // only sub-grids that meet the slab [k0,k1) of the last direction are visited (the reference's
  // `intersect(subgrid, target_cells)`, impl_unfitted_domain.cpp:447-450); decisions inside are unchanged
  if (nd.hi[N - 1] <= k0 || nd.lo[N - 1] >= k1) return true;
  Iv<N> xint[N];
  Dl<N> dl;
  long long cnt = 1;
#pragma unroll
  for (int d = 0; d < N; ++d) {
    // sub-grid bounding box inflated by 1000 eps (BoundBox::extend, bbox.cpp:136-147)
    double lo = g.breaks[d][nd.lo[d]], hi = g.breaks[d][nd.hi[d]];
    lo -= kEps * 1000;
    hi += kEps * 1000;
    xint[d] = iv_var<N>(0.5 * (lo + hi), d);
    dl.d[d] = 0.5 * (hi - lo);
    cnt *= (nd.hi[d] - nd.lo[d]);
  }
  const Iv<N> res = phi_val_iv<N>(f, xint, dl);
  if (iv_uniform(res, dl)) {
    const int k = atomicAdd(nuniform, 1);
    if (k >= ucap) return false;
    uniform[k] = nd;
    uniform_sign[k] = res.a < 0.0 ? ST_FULL : ST_EMPTY;
    return true;
  }
  if (cnt == 1) {
    long long id = 0, stride = 1;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      id += nd.lo[d] * stride;
      stride *= g.n[d];
    }
    status[id] = ST_UNDET;
    return true;
  }
  // TensorIndexRangeTP::split (tensor_index_tp.cpp:258-273): first direction of maximal size, at (lo+hi)/2
  int sd = 0;
#pragma unroll
  for (int d = 1; d < N; ++d)
    if (nd.hi[d] - nd.lo[d] > nd.hi[sd] - nd.lo[sd]) sd = d;
  const int mid = (nd.lo[sd] + nd.hi[sd]) / 2;
  KdNode a = nd, b = nd;
  a.hi[sd] = mid;
  b.lo[sd] = mid;
  const int k = atomicAdd(nnext, 2);
  if (k + 2 > cap) return false;
  next[k] = a;
  next[k + 1] = b;
  return true;
}
template<int N> __device__ __forceinline__ void kd_fill_node(const GridDesc &g, KdNode nd, unsigned char s, unsigned char *status,
  int k0, int k1)
{
  nd.lo[N - 1] = max(nd.lo[N - 1], k0);
  nd.hi[N - 1] = min(nd.hi[N - 1], k1);
  const long long n0 = nd.hi[0] - nd.lo[0];
  const long long n1 = nd.hi[1] - nd.lo[1];
  const long long n2 = (N == 3) ? (nd.hi[2] - nd.lo[2]) : 1;
  const long long tot = n0 * n1 * n2;
  for (long long t = threadIdx.x; t < tot; t += blockDim.x) {
    const long long i = t % n0, j = (t / n0) % n1, k = t / (n0 * n1);
    long long id = (nd.lo[0] + i) + g.n[0] * (nd.lo[1] + j);
    if (N == 3) id += g.n[0] * g.n[1] * (nd.lo[2] + k);
    status[id] = s;
  }
}

// The whole tree in ONE cooperative launch: levels are separated by grid barriers instead of host round trips
// (two dozen levels x (launch + counter download + launch) is most of a small slab's domain time).
// ctr: [0..63] nodes per level (ctr[0] = 1 on entry), [64..127] uniform nodes per level, [128] overflow flag.
constexpr int kKdMaxLevels = 64;
template<int N>
__global__ void __launch_bounds__(256) kd_tree_kernel(FuncDesc f, GridDesc g, KdNode *buf_a, KdNode *buf_b, int cap, KdNode *uniform,
  unsigned char *uniform_sign, int ucap, int *ctr, unsigned char *status, int k0, int k1)
{
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  KdNode *cur = buf_a, *nxt = buf_b;
  for (int level = 0; level + 1 < kKdMaxLevels; ++level) {
    const int ncur = ctr[level];
    if (ncur == 0 || ctr[128]) break;
    bool ok = true;
    for (int i = gtid; i < ncur; i += gsize)
      ok = kd_process_node<N>(f, g, cur[i], nxt, ctr + level + 1, cap, uniform, uniform_sign, ctr + 64 + level, ucap, status, k0, k1)
           && ok;
    if (!ok) ctr[128] = 1;
    grid.sync();
    if (!ctr[128]) {
      const int nuni = ctr[64 + level];
      for (int b = blockIdx.x; b < nuni; b += gridDim.x) kd_fill_node<N>(g, uniform[b], uniform_sign[b], status, k0, k1);
    }
    grid.sync();
    KdNode *t = cur;
    cur = nxt;
    nxt = t;
  }
}

template<int N>
__global__ void kd_level_kernel(FuncDesc f, GridDesc g, const KdNode *nodes, int nnodes, KdNode *next, int *nnext,
  KdNode *uniform, unsigned char *uniform_sign, int *nuniform, unsigned char *status, int k0, int k1)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnodes) return;
  kd_process_node<N>(f, g, nodes[i], next, nnext, 2 * nnodes, uniform, uniform_sign, nuniform, nnodes, status, k0, k1);
}

// one block per uniform node: paint its cells
template<int N>
__global__ void kd_fill_kernel(GridDesc g, const KdNode *uniform, const unsigned char *uniform_sign,
  unsigned char *status, int k0, int k1)
{
  kd_fill_node<N>(g, uniform[blockIdx.x], uniform_sign[blockIdx.x], status, k0, k1);
}

// flags -> compacted index list (ordered)
__global__ void flag_eq_kernel(const unsigned char *status, long long n, unsigned char v, int *flag)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = status[i] == v ? 1 : 0;
}
__global__ void scatter_ids_kernel(const int *flag, const long long *off, long long n, long long *out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[off[i]] = i;
}

// classify_cell_from_quad (impl_unfitted_domain.cpp:163-180) on the RAW n=1 volume rule of each job
__host__ __device__ inline bool facet_has_unf_bdry(unsigned char s)
{
  return s == QG_FACET_CUT_UNF_BDRY || s == QG_FACET_CUT_UNF_BDRY_EXT_BDRY || s == QG_FACET_FULL_UNF_BDRY
         || s == QG_FACET_UNF_BDRY || s == QG_FACET_UNF_BDRY_EXT_BDRY;
}
enum TmpStatus : unsigned char { TMP_CUT = 0, TMP_FULL = 1, TMP_EMPTY = 2, TMP_FULL_UNF = 3 };
__global__ void classify_cells_kernel(long long njobs, const long long *job_pt_off, const double *wts,
  const double *job_box, int N, unsigned char *tmp)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  double vol = 0.0;
  for (long long q = job_pt_off[j]; q < job_pt_off[j + 1]; ++q) vol += wts[q];
  double cv = 1.0;
  for (int d = 0; d < N; ++d) cv *= (job_box[6 * j + 3 + d] - job_box[6 * j + d]);
  unsigned char t;
  if (fabs(vol - 0.0) <= kNearEps)
    t = TMP_EMPTY;
  else if (fabs(vol - cv) <= kNearEps)
    t = TMP_FULL_UNF;
  else
    t = TMP_CUT;
  tmp[j] = t;
}

// classify_facet_from_quad + classify_facet_full_unfitted_boundary (impl_unfitted_domain.cpp:237-333): status of local
// facet lf of a cell with temporary status cs from the n = 1 rule [q0,q1) of its face; box6 = lo[3], hi[3] of the cell.
template<int N>
__device__ __forceinline__ unsigned char classify_facet_from_rule(const FuncDesc &f, int lf, unsigned char cs, long long q0,
  long long q1, const double *pts, const double *wts, const double *box6)
{
  const int cd = lf >> 1, side = lf & 1;
  if (q1 == q0) return cs == TMP_FULL_UNF ? QG_FACET_FULL_UNF_BDRY : QG_FACET_EMPTY;
  bool ext = false, unf = false, cut = false;
  double fvol = 0.0;
  for (long long q = q0; q < q1; ++q) {
    double x[N];
    for (int d = 0; d < N; ++d) x[d] = pts[q * N + d];
    fvol += wts[q];
    if (fabs(phi_val<N>(f, x)) <= kNearEps) {
      double g[N];
      phi_grad<N>(f, x, g);
      double s = g[0] * g[0];
      double gc = g[0];
      for (int d = 1; d < N; ++d) {
        s += g[d] * g[d];
        if (d == cd) gc = g[d];
      }
      const double nc = gc / sqrt(s);
      if (!(fabs(fabs(nc) - 1.0) <= kNearEps)) continue;
      if (side == 0 ? (nc < -kNearEps) : (kNearEps < nc))
        unf = true;
      else
        ext = true;
    } else {
      cut = true;
    }
  }
  if (cs == TMP_FULL_UNF && !cut) return QG_FACET_FULL_UNF_BDRY;
  const unsigned char values[8] = { QG_FACET_EMPTY, QG_FACET_EXT_BDRY, QG_FACET_UNF_BDRY, QG_FACET_UNF_BDRY_EXT_BDRY,
    QG_FACET_CUT, QG_FACET_CUT_EXT_BDRY, QG_FACET_CUT_UNF_BDRY, QG_FACET_CUT_UNF_BDRY_EXT_BDRY };
  unsigned char st = values[(ext ? 1 : 0) + (unf ? 2 : 0) + (cut ? 4 : 0)];
  double dom_vol = 1.0;
  for (int d = 0; d < N; ++d)
    if (d != cd) dom_vol *= (box6[3 + d] - box6[d]);
  const double max_val = fabs(fvol) < fabs(dom_vol) ? fabs(dom_vol) : fabs(fvol);
  const bool is_full = fabs(fvol - dom_vol) <= (kNearEps + (kNearEps * 1.0e3) * max_val);
  if (is_full) {
    const long long npts = q1 - q0;
    if ((st == QG_FACET_UNF_BDRY || st == QG_FACET_EXT_BDRY) && npts == 1)
      st = st == QG_FACET_UNF_BDRY ? QG_FACET_FULL_UNF_BDRY : QG_FACET_FULL_EXT_BDRY;
    else if (st == QG_FACET_CUT)
      st = QG_FACET_FULL;
  }
  return st;
}
// job_slot (may be null = identity): position of job j among the 2N facets of all cells (cell index * 2N + facet)
template<int N>
__global__ void classify_facets_kernel(FuncDesc f, long long njobs, const int *job_type, const long long *job_pt_off,
  const double *pts, const double *wts, const double *job_box, const unsigned char *cell_tmp, const long long *job_slot,
  unsigned char *fstat_all)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const long long slot = job_slot ? job_slot[j] : j;
  fstat_all[slot] = classify_facet_from_rule<N>(f, job_type[j], cell_tmp[slot / (2 * N)], job_pt_off[j], job_pt_off[j + 1], pts, wts,
    job_box + 6 * j);
}

// ---- one n = 1 rule per FACE --------------------------------------------------------------------------------------
// The upper facet (side 1) of a cell and the lower facet (side 0) of its neighbour along the same direction are the same
// face: the n = 1 rule of a face only reads the face's own coordinate (the same grid break for both cells) and the
// extents of the two free directions, so both cells get the same nodes and weights, bit for bit.  The reference computes
// the rule twice (impl_unfitted_domain.cpp:399-401 declines the shortcut); here the lower facet becomes an ALIAS of the
// neighbour's job when both are undecided after the prefilter, and is classified from that rule with its own side and
// cell status.  vidx: position of a cell among the v-cells (cells whose facets are classified), -1 otherwise, indexed by
// flat id - c0 over the slab.
__global__ void vidx_kernel(long long nv, const long long *vcells, long long c0, int *vidx)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nv) vidx[vcells[i] - c0] = (int)i;
}
template<int N>
__global__ void facet_alias_kernel(GridDesc g, long long nslots, const long long *vcells, const int *vidx, long long c0,
  long long nslab, int *flag, long long *alias)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  const int nf = 2 * N;
  const int lf = (int)(t % nf);
  if ((lf & 1) || !flag[t]) return;  // lower facets only: they read upper-facet flags, which nobody writes here
  const int cd = lf >> 1;
  const long long c = vcells[t / nf];
  long long stride = 1, idx = c;
  for (int d = 0; d < cd; ++d) {
    stride *= g.n[d];
    idx /= g.n[d];
  }
  if (idx % g.n[cd] == 0) return;     // on the lower boundary of the grid: no neighbour
  const long long cn = c - stride - c0;
  if (cn < 0 || cn >= nslab) return;  // neighbour outside this rank's slab
  const int vn = vidx[cn];
  if (vn < 0) return;
  const long long ts = (long long)vn * nf + lf + 1;
  if (!flag[ts]) return;
  alias[t] = ts;
  flag[t] = 0;
}
template<int N>
__global__ void classify_alias_facets_kernel(FuncDesc f, long long nslots, const long long *alias, const long long *slot_job,
  const long long *job_pt_off, const double *pts, const double *wts, const double *job_box, const unsigned char *cell_tmp,
  unsigned char *fstat_all)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  const long long ts = alias[t];
  if (ts < 0) return;
  const long long j = slot_job[ts];   // exclusive scan of the job flags: the owner's job index
  fstat_all[t] = classify_facet_from_rule<N>(f, (int)(t % (2 * N)), cell_tmp[t / (2 * N)], job_pt_off[j], job_pt_off[j + 1], pts, wts,
    job_box + 6 * j);
}

// Alias facets whose owner's walk ran into the bisection cap need their own rule: there the unused PsiCode slots (sign -1,
// side 0: kPsiDefaultSign) make the cap test of an UPPER facet's job read phi on the cell's LOWER face, so the two cells of a
// face do not get the same rule.  Everywhere else the walks of the two jobs are identical decision for decision.
__global__ void alias_redo_flag_kernel(long long nslots, const long long *alias, const long long *slot_job, const unsigned char *job_cap,
  int *flag)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  const long long ts = alias[t];
  flag[t] = (ts >= 0 && job_cap[slot_job[ts]]) ? 1 : 0;
}

// Facets whose status the first step of their n = 1 rule already decides (3-D): the walk starts with the interval
// test of phi over the face (ctl_walk's prune).  Uniformly negative -> the rule is the one-point Gauss rule of the whole
// face -> FULL; uniformly positive -> the rule is empty -> EMPTY (FULL_UNF_BDRY if the cell is full with unfitted
// boundary).  Same helpers as ctl_walk, so the test sees the same bits; |alpha| must clear 1e-12 so that the facet
// node cannot count as lying on the level set.  Everything else goes through the pipeline (flag = 1).
template<int N>
__global__ void facet_prefilter_kernel(FuncDesc f, GridDesc g, long long nslots, const long long *vcells,
  const unsigned char *cell_tmp, unsigned char *fstat, int *flag)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  const int lf = (int)(t % (2 * N)), cd = lf >> 1, side = lf & 1;
  double lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
  cell_box<N>(g, vcells[t / (2 * N)], lo, hi);
  Frame fr;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    fr.lo[d] = d < N ? lo[d] : 0.0;
    fr.hi[d] = d < N ? hi[d] : 0.0;
  }
  fr.freemask = (unsigned char)(((1 << N) - 1) & ~(1 << cd));
  Iv<N> xint[N];
  Dl<N> dl;
  frame_xint<N>(fr, xint, dl);
  frame_set_sides<N>(fr, side << cd, xint);
  const Iv<N> res = phi_val_iv<N>(f, xint, dl);
  int undecided = 1;
  if (iv_uniform(res, dl) && fabs(res.a) > 1.0e-12) {
    undecided = 0;
    fstat[t] = res.a < 0.0 ? (unsigned char)QG_FACET_FULL
                           : (cell_tmp[t / (2 * N)] == TMP_FULL_UNF ? (unsigned char)QG_FACET_FULL_UNF_BDRY : (unsigned char)QG_FACET_EMPTY);
  }
  flag[t] = undecided;
}
__global__ void gather_facet_slots_kernel(long long nslots, const long long *vcells, int nf, const int *flag, const long long *off,
  long long *job_cell, int *job_type, long long *job_slot)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots || !flag[t]) return;
  const long long j = off[t];
  job_cell[j] = vcells[t / nf];
  job_type[j] = (int)(t % nf);
  job_slot[j] = t;
}

// classify_undetermined_sign_cell, final part (impl_unfitted_domain.cpp:493-512)
__global__ void finalize_cells_kernel(long long nv, const long long *vcells, const unsigned char *vtmp,
  const unsigned char *fstat, int nf, unsigned char *status, int *keep)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  bool all_full = true;
  for (int k = 0; k < nf; ++k)
    if (fstat[i * nf + k] != QG_FACET_FULL) all_full = false;
  const unsigned char t = vtmp[i];
  const bool demote = (t == TMP_FULL_UNF && all_full);
  keep[i] = demote ? 0 : 1;
  status[vcells[i]] = t == TMP_CUT ? ST_CUT : ST_FULL;
}
__global__ void set_status_from_tmp_kernel(long long nu, const long long *ucells, const unsigned char *utmp,
  unsigned char *status, int *is_v)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu) return;
  const unsigned char t = utmp[i];
  is_v[i] = (t == TMP_CUT || t == TMP_FULL_UNF) ? 1 : 0;
  if (t == TMP_EMPTY) status[ucells[i]] = ST_EMPTY;
}
__global__ void gather_v_kernel(long long nu, const long long *ucells, const unsigned char *utmp, const int *is_v,
  const long long *off, long long *vcells, unsigned char *vtmp)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu || !is_v[i]) return;
  vcells[off[i]] = ucells[i];
  vtmp[off[i]] = utmp[i];
}
__global__ void make_facet_jobs_kernel(long long nv, const long long *vcells, int nf, long long *job_cell,
  int *job_type)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv * nf) return;
  job_cell[j] = vcells[j / nf];
  job_type[j] = (int)(j % nf);
}
__global__ void fill_types_kernel(long long n, int type, int *job_type)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) job_type[j] = type;
}
__global__ void gather_records_kernel(long long nv, const long long *vcells, const unsigned char *fstat, int nf,
  const int *keep, const long long *off, long long *rcells, unsigned char *rstat, int *rec_idx, int *n_unf)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv || !keep[i]) return;
  const long long r = off[i];
  rcells[r] = vcells[i];
  bool unf = false;
  for (int k = 0; k < nf; ++k) {
    const unsigned char st = fstat[i * nf + k];
    rstat[r * nf + k] = st;
    unf = unf || facet_has_unf_bdry(st);
  }
  rec_idx[vcells[i]] = (int)r;
  if (unf) atomicAdd(n_unf, 1);  // cells owning a facet with unfitted boundary (rare)
}
__global__ void status_hist_kernel(const unsigned char *status, long long n, unsigned long long *hist)
{
  __shared__ unsigned int h[4];
  if (threadIdx.x < 4) h[threadIdx.x] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned char s = status[i];
    if (s < 3) atomicAdd(&h[s], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}
__global__ void fill_i32_kernel(long long n, int v, int *out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

}  // namespace qg
