// CPU emulation harness for the kernel bodies in qugar_b200/csrc/qg_pipeline.cuh -- a DEBUG/TEST aid.
// The build container has no GPU; this file compiles the per-thread bodies with g++ (QG_HD = inline) and
// runs them thread id by thread id, with the same scans in between, so that the pass logic can be checked
// against the oracle before going to a B200.  It is built only by the tests (tests/emu/_build) and is
// not part of, nor reachable from, libqugar_b200.so: the product has no CPU path.
#include "../../qugar_b200/csrc/qg_pipeline.cuh"
#include "../../qugar_b200/csrc/qg_reparam.cuh"

#include <cstring>
#include <memory>
#include <vector>

using namespace qg;

static const double g_gauss[] = {
#include "../../qugar_b200/csrc/gauss_table.inc"
};

struct EmuRule
{
  std::vector<int> n_per;
  std::vector<double> pts, wts, nrm;
  int pdim;
  int err;
  long long nitems, ncall1, ncall2;
};

static void exscan(const std::vector<int> &cnt, std::vector<long long> &off)
{
  off.resize(cnt.size() + 1);
  long long s = 0;
  for (size_t i = 0; i < cnt.size(); ++i) {
    off[i] = s;
    s += cnt[i];
  }
  off[cnt.size()] = s;
}

template<int N>
static EmuRule *run(const FuncDesc &f, const GridDesc &g, const long long *cells, const int *types, long long nj, int p, int mode)
{
  PipeArgs a;
  std::memset(&a, 0, sizeof(a));
  a.f = f;
  a.g = g;
  a.p = p;
  a.sink_mode = mode;
  a.gxw = g_gauss + 2 * (size_t)(p * (p - 1) / 2);
  int err = 0;
  a.err = &err;
  a.njobs = nj;
  a.job_cell = cells;
  a.job_type = types;
  std::vector<double> job_box((size_t)nj * 6 + 1);
  a.job_box = job_box.data();
  std::vector<int> item_cnt((size_t)nj);
  a.item_cnt = item_cnt.data();
  for (long long j = 0; j < nj; ++j) ctl_body<N, false>(j, a, a.f);
  std::vector<long long> item_off;
  exscan(item_cnt, item_off);
  a.item_off = item_off.data();
  a.nitems = item_off[(size_t)nj];
  std::vector<ChainItem> items((size_t)a.nitems + 1);
  a.items = items.data();
  for (long long j = 0; j < nj; ++j) ctl_body<N, true>(j, a, a.f);
  std::vector<double> ent0((size_t)a.nitems * 2 * kMaxSeg0 + 1);
  std::vector<int> nent0((size_t)a.nitems), cnt0((size_t)a.nitems);
  a.ent0 = ent0.data();
  a.nent0 = nent0.data();
  a.cnt0 = cnt0.data();
  for (long long i = 0; i < a.nitems; ++i) k0_body<N>(i, a, a.f);
  std::vector<long long> off0;
  exscan(cnt0, off0);
  a.off0 = off0.data();
  a.ncall1 = off0[(size_t)a.nitems];
  std::vector<Call1> call1((size_t)a.ncall1 + 1);
  std::vector<int> cnt1((size_t)a.ncall1);
  a.call1 = call1.data();
  a.cnt1 = cnt1.data();
  for (long long t = 0; t < a.ncall1; ++t) k1_body<N>(t, a, a.f);
  std::vector<long long> off1;
  exscan(cnt1, off1);
  a.off1 = off1.data();
  a.ncall2 = off1[(size_t)a.ncall1];
  std::vector<Call2> call2((size_t)a.ncall2 + 1);
  std::vector<int> cnt2((size_t)a.ncall2);
  a.call2 = call2.data();
  a.cnt2 = cnt2.data();
  for (long long u = 0; u < a.ncall2; ++u) k2_body<N>(u, a, a.f);
  std::vector<long long> off2;
  exscan(cnt2, off2);
  a.off2 = off2.data();
  a.npts = off2[(size_t)a.ncall2];
  std::vector<long long> job_pt_off((size_t)nj + 1);
  for (long long j = 0; j <= nj; ++j) job_points_body(j, a, nullptr, job_pt_off.data());
  EmuRule *r = new EmuRule();
  r->pdim = mode == SINK_FACET ? N - 1 : N;
  r->pts.resize((size_t)a.npts * r->pdim + 1);
  r->wts.resize((size_t)a.npts + 1);
  r->nrm.resize((size_t)a.npts * N + 1);
  a.pts = r->pts.data();
  a.wts = r->wts.data();
  a.nrm = r->nrm.data();
  for (long long u = 0; u < a.ncall2; ++u) k3_body<N>(u, a, a.f, nullptr);
  r->pts.resize((size_t)a.npts * r->pdim);
  r->wts.resize((size_t)a.npts);
  r->nrm.resize((size_t)a.npts * N);
  r->n_per.resize((size_t)nj);
  for (long long j = 0; j < nj; ++j) r->n_per[(size_t)j] = (int)(job_pt_off[(size_t)j + 1] - job_pt_off[(size_t)j]);
  r->err = err;
  r->nitems = a.nitems;
  r->ncall1 = a.ncall1;
  r->ncall2 = a.ncall2;
  return r;
}

extern "C" {

void *emu_run(int dim, int kind, int negative, const double *params, int nparams, const double *breaks,
  const long long *n_breaks, const long long *cells, const int *types, long long nj, int p, int mode)
{
  FuncDesc f;
  std::memset(&f, 0, sizeof(f));
  f.kind = kind;
  f.dim = dim;
  f.negative = negative;
  static thread_local CompositeDesc comp;
  auto leaf = [&](FuncDesc &fd, int lkind, int lneg, const double *p, int np) {
    std::memset(&fd, 0, sizeof(fd));
    fd.kind = lkind;
    fd.dim = dim;
    fd.negative = lneg;
    if (lkind == QG_FUNC_BEZIER) {
      // params: order[dim] then the coefficients (they stay in the caller's array for the duration of the run)
      for (int d = 0; d < 3; ++d) fd.order[d] = d < dim ? (int)p[d] : 1;
      fd.coefs = p + dim;
    } else {
      for (int i = 0; i < np && i < 16; ++i) fd.p[i] = p[i];
      fd.same_args = tpms_same_arg_mask(lkind, dim, fd.p);
    }
  };
  if (kind == QG_FUNC_COMPOSITE) {
    // the program form, decoded by the same parser as qg_func_create (csrc/qg_funcs.cuh)
    std::memset(&comp, 0, sizeof(comp));
    const char *err = composite_parse(params, nparams, comp, [&](FuncDesc &fd, int, int lkind, int lneg, const double *p, int np) {
      leaf(fd, lkind, lneg, p, np);
      return (const char *)nullptr;
    });
    if (err) return nullptr;
    f.comp = &comp;
  } else {
    leaf(f, kind, negative, params, nparams);
  }
  GridDesc g;
  std::memset(&g, 0, sizeof(g));
  g.dim = dim;
  const double *b = breaks;
  for (int d = 0; d < 3; ++d) g.n[d] = 1;
  for (int d = 0; d < dim; ++d) {
    g.breaks[d] = b;
    g.n[d] = n_breaks[d] - 1;
    b += n_breaks[d];
  }
  return dim == 2 ? run<2>(f, g, cells, types, nj, p, mode) : run<3>(f, g, cells, types, nj, p, mode);
}
void emu_sizes(void *h, long long *npts, int *pdim, int *err, long long *stats)
{
  EmuRule *r = (EmuRule *)h;
  *npts = (long long)r->wts.size();
  *pdim = r->pdim;
  *err = r->err;
  stats[0] = r->nitems;
  stats[1] = r->ncall1;
  stats[2] = r->ncall2;
}
void emu_copy(void *h, int *n_per, double *pts, double *wts, double *nrm)
{
  EmuRule *r = (EmuRule *)h;
  std::memcpy(n_per, r->n_per.data(), r->n_per.size() * sizeof(int));
  std::memcpy(pts, r->pts.data(), r->pts.size() * sizeof(double));
  std::memcpy(wts, r->wts.data(), r->wts.size() * sizeof(double));
  if (nrm) std::memcpy(nrm, r->nrm.data(), r->nrm.size() * sizeof(double));
}
void emu_free(void *h) { delete (EmuRule *)h; }
void emu_sincos(double x, double *s, double *c) { sincos_det(x, *s, *c); }
// number of (a[i], b[i]) for which div_rcp(a, b, 1/b) differs from a / b
long long emu_div_rcp_mismatches(const double *a, const double *b, long long n)
{
  long long bad = 0;
  for (long long i = 0; i < n; ++i) bad += div_rcp(a[i], b[i], 1.0 / b[i]) != a[i] / b[i];
  return bad;
}
}

// ---- reparameterization: the bodies of qg_reparam.cuh, thread id by thread id (no merge) ----------------------------------
struct EmuMesh
{
  std::vector<double> pts;
  std::vector<long long> conn, wires;
  int err;
};

template<int N>
static EmuMesh *run_reparam(const FuncDesc &f, const GridDesc &g, const long long *cells, long long nj, const long long *full,
  long long nfull, int order, int mode)
{
  PipeArgs a;
  std::memset(&a, 0, sizeof(a));
  a.f = f;
  a.g = g;
  a.p = 1;
  int err = 0;
  a.err = &err;
  a.njobs = nj;
  a.job_cell = cells;
  std::vector<int> types((size_t)nj + 1, mode == 0 ? JOB_VOLUME : (mode == 1 ? JOB_SURFACE : mode - 2));
  a.job_type = types.data();
  std::vector<double> job_box((size_t)nj * 6 + 1);
  a.job_box = job_box.data();
  std::vector<int> item_cnt((size_t)nj);
  a.item_cnt = item_cnt.data();
  for (long long j = 0; j < nj; ++j) ctl_body<N, false>(j, a, a.f);
  std::vector<long long> item_off;
  exscan(item_cnt, item_off);
  a.item_off = item_off.data();
  a.nitems = item_off[(size_t)nj];
  std::vector<ChainItem> items((size_t)a.nitems + 1);
  a.items = items.data();
  for (long long j = 0; j < nj; ++j) ctl_body<N, true>(j, a, a.f);

  const int pd = mode == 0 ? N : N - 1;
  int npc = 1;
  for (int i = 0; i < pd; ++i) npc *= order;
  const std::vector<double> nodes = reparam_nodes(order);
  std::vector<double> basis, ders;
  lagrange_mid(pd, order, basis, ders);
  std::vector<int> leaf_cells((size_t)a.nitems), leaf_pts((size_t)a.nitems);
  auto ctx = std::make_unique<RpCtx<N>>();
  auto setup = [&](long long i, bool emit) {
    RpCtx<N> &c = *ctx;
    c.f = f;
    c.it = items[(size_t)i];
    c.order = order;
    c.nodes = nodes.data();
    c.fixdir = mode >= 2 ? (mode - 2) / 2 : -1;
    c.emit = emit;
    c.npc = npc;
    c.err = 0;
  };
  for (long long i = 0; i < a.nitems; ++i) {
    setup(i, false);
    rp_leaf<N>(*ctx);
    leaf_cells[(size_t)i] = ctx->ncells;
    leaf_pts[(size_t)i] = (int)ctx->npts;
    err |= ctx->err;
  }
  std::vector<long long> cell_off, pt_off;
  exscan(leaf_cells, cell_off);
  exscan(leaf_pts, pt_off);
  const long long ncut_cells = cell_off[(size_t)a.nitems], ncut_pts = pt_off[(size_t)a.nitems];
  EmuMesh *m = new EmuMesh();
  m->pts.assign((size_t)(ncut_pts + nfull * npc) * N, 0.0);
  m->conn.assign((size_t)(ncut_cells + nfull) * npc, 0);
  std::vector<int> cell_job((size_t)ncut_cells + 1);
  for (long long i = 0; i < a.nitems; ++i) {
    setup(i, true);
    ctx->pts = m->pts.data();
    ctx->conn = m->conn.data();
    ctx->pt_base = pt_off[(size_t)i];
    ctx->cell_base = cell_off[(size_t)i];
    rp_leaf<N>(*ctx);
    for (int k = 0; k < ctx->ncells; ++k) cell_job[(size_t)(ctx->cell_base + k)] = ctx->it.job;
  }
  RpMeshArgs ma;
  std::memset(&ma, 0, sizeof(ma));
  ma.f = f;
  ma.dim = N;
  ma.pd = pd;
  ma.order = order;
  ma.npc = npc;
  ma.mode = mode;
  ma.ncells = ncut_cells;
  ma.pts = m->pts.data();
  ma.conn = m->conn.data();
  ma.cell_job = cell_job.data();
  ma.job_box = job_box.data();
  ma.basis = basis.data();
  ma.ders = ders.data();
  ma.tol = 1.0e4 * kEps;
  for (long long c = 0; c < ncut_cells; ++c) rp_orient_body<N>(c, ma);
  const int nedges = pd == 2 ? 4 : (pd == 3 ? 12 : 0);
  std::vector<int> edge_flag((size_t)ncut_cells * nedges);
  std::vector<long long> edge_off;
  ma.edge_flag = edge_flag.data();
  for (long long t = 0; t < ncut_cells * nedges; ++t) rp_wire_body<N, false>(t, ma);
  exscan(edge_flag, edge_off);
  const long long nwire_cut = edge_off[(size_t)ncut_cells * nedges];
  m->wires.assign((size_t)(nwire_cut + nfull * nedges) * order, 0);
  ma.edge_off = edge_off.data();
  ma.wires = m->wires.data();
  for (long long t = 0; t < ncut_cells * nedges; ++t) rp_wire_body<N, true>(t, ma);
  for (long long i = 0; i < nfull; ++i)
    rp_full_cell_body<N>(i, nfull, full, g, order, nodes.data(), ncut_pts, ncut_cells, nwire_cut, m->pts.data(), m->conn.data(),
      m->wires.data());
  m->err = err;
  return m;
}

extern "C" {
void *emu_reparam(int dim, int kind, int negative, const double *params, int nparams, const double *breaks,
  const long long *n_breaks, const long long *cells, long long nj, const long long *full, long long nfull, int order, int mode)
{
  FuncDesc f;
  std::memset(&f, 0, sizeof(f));
  f.kind = kind;
  f.dim = dim;
  f.negative = negative;
  for (int i = 0; i < nparams && i < 16; ++i) f.p[i] = params[i];
  f.same_args = tpms_same_arg_mask(kind, dim, f.p);
  GridDesc g;
  std::memset(&g, 0, sizeof(g));
  g.dim = dim;
  const double *b = breaks;
  for (int d = 0; d < 3; ++d) g.n[d] = 1;
  for (int d = 0; d < dim; ++d) {
    g.breaks[d] = b;
    g.n[d] = n_breaks[d] - 1;
    b += n_breaks[d];
  }
  return dim == 2 ? run_reparam<2>(f, g, cells, nj, full, nfull, order, mode) : run_reparam<3>(f, g, cells, nj, full, nfull, order, mode);
}
void emu_mesh_sizes(void *h, long long *npts_x_dim, long long *nconn, long long *nwires, int *err)
{
  EmuMesh *m = (EmuMesh *)h;
  *npts_x_dim = (long long)m->pts.size();
  *nconn = (long long)m->conn.size();
  *nwires = (long long)m->wires.size();
  *err = m->err;
}
void emu_mesh_copy(void *h, double *pts, long long *conn, long long *wires)
{
  EmuMesh *m = (EmuMesh *)h;
  std::memcpy(pts, m->pts.data(), m->pts.size() * sizeof(double));
  std::memcpy(conn, m->conn.data(), m->conn.size() * sizeof(long long));
  std::memcpy(wires, m->wires.data(), m->wires.size() * sizeof(long long));
}
void emu_mesh_free(void *h) { delete (EmuMesh *)h; }
}
