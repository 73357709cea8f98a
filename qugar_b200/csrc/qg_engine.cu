// qugar_b200 -- CUDA engine: kernels, pass scheduling, domain classification and the C ABI.
// Compiled for sm_100a only, with -fmad=false (see qg_math.cuh).  No CPU compute path exists here:
// every entry point that computes needs a CUDA device and fails with QG_ERR_NO_DEVICE otherwise.
#include "../../include/qugar_b200.h"
#include "qg_pipeline.cuh"
#include "qg_launch_cfg.cuh"
#include "qg_coeffs.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <string>
#include <system_error>
#include <thread>
#if defined(__AVX__)
#include <immintrin.h>
#endif
#include <vector>

namespace qg {

static const double h_gauss_table[] = {
#include "gauss_table.inc"
};
constexpr int kGaussMax = 100;

}  // namespace qg

#include "qg_device_mem.cuh"
#include "qg_kernels_pipeline.cuh"
#include "qg_kernels_domain.cuh"
#include "qg_kernels_rules.cuh"
#define QG_ENGINE_UNIT 1
#include "qg_kernels_reparam.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace qg {
// per-kind launchers, one translation unit each (qg_kind.cu): the benchmark configurations' level sets have their own
// instantiation of CTL / K0 / K1 / K2 / K3; every other kind runs the run-time-kind unit
void launch_pipe_k0(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *);    // sphere / disk
void launch_pipe_k2d2(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *);  // Bezier, 2-D
void launch_pipe_k2d3(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *);  // Bezier, 3-D
void launch_pipe_k10(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *);   // Schoen gyroid
void launch_pipe_k14(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *);   // Schwarz diamond
void launch_pipe_gend2(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *); // run-time kind, 2-D
void launch_pipe_gend3(int, int, unsigned, cudaStream_t, const PipeArgs &, const long long *); // run-time kind, 3-D
struct RpLeafArgs;
void launch_rp_k0(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
void launch_rp_k2d2(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
void launch_rp_k2d3(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
void launch_rp_k10(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
void launch_rp_k14(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
void launch_rp_gend2(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
void launch_rp_gend3(int, int, unsigned, cudaStream_t, const RpLeafArgs &);
static void launch_pipe(int which, int dim, unsigned grid, cudaStream_t s, const PipeArgs &a, const long long *shift)
{
  g_launches.fetch_add(1, std::memory_order_relaxed);
  switch (a.f.kind) {
  case QG_FUNC_SPHERE: launch_pipe_k0(which, dim, grid, s, a, shift); break;
  case QG_FUNC_BEZIER: (dim == 2 ? launch_pipe_k2d2 : launch_pipe_k2d3)(which, dim, grid, s, a, shift); break;
  case QG_FUNC_SCHOEN: launch_pipe_k10(which, dim, grid, s, a, shift); break;
  case QG_FUNC_SCHWARZ_D: launch_pipe_k14(which, dim, grid, s, a, shift); break;
  default: (dim == 2 ? launch_pipe_gend2 : launch_pipe_gend3)(which, dim, grid, s, a, shift); break;
  }
}
}  // namespace qg

namespace qg {

// ------------------------------------------------------------------------------------------------
// host-side objects
// ------------------------------------------------------------------------------------------------
struct IntToLL
{
  __host__ __device__ long long operator()(const int &v) const { return (long long)v; }
};

struct Scanner
{
  DevBuf<unsigned char> tmp;
  // exclusive scan of n ints into n long longs (the caller passes n = count + 1 with a trailing zero so
  // that out[count] is the total)
  void run(const int *in, long long *out, size_t n, cudaStream_t s)
  {
    // record indices are 32-bit in the pass records (Call1::item, ...): more than 2^31-2 records of one pass in one
    // call must be split by the caller (cell list chunks); say so instead of wrapping around
    if (n >= (size_t)2147483647) throw EngineError(QG_ERR_CAPACITY, "a pass holds more than 2^31 records: pass the cells in smaller batches");
    cub::TransformInputIterator<long long, IntToLL, const int *> it(in, IntToLL());
    size_t bytes = 0;
    QG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, s));
    if (bytes > tmp.n) tmp.alloc(bytes * 2 + 1024);
    QG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, it, out, (int)n, s));
    g_launches.fetch_add(2, std::memory_order_relaxed);  // CUB: init + scan kernel
  }
};

struct PassTimes
{
  double ms[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };  // ctl, k0, k1, k2, k3, scans, other, total
  long long counts[4] = { 0, 0, 0, 0 };       // chain items, stage-1 calls, stage-2 lines, nodes (of the last pipeline run)
};

// Pass timing without host synchronisation: begin()/end(slot) only record events on the stream; resolve() turns the
// spans into milliseconds once the caller has synchronised anyway (end of the rule / domain call).  Round 1 called
// cudaEventSynchronize after every pass: ~12 host round trips per pipeline that left the GPU idle in between.
struct EventPool
{
  std::mutex m;
  std::vector<cudaEvent_t> free_;
  cudaEvent_t get()
  {
    {
      std::lock_guard<std::mutex> l(m);
      if (!free_.empty()) {
        cudaEvent_t e = free_.back();
        free_.pop_back();
        return e;
      }
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void put(cudaEvent_t e)
  {
    std::lock_guard<std::mutex> l(m);
    free_.push_back(e);
  }
};
static EventPool g_event_pool;

struct SpanClock
{
  struct Span { cudaEvent_t a, b; int slot; };
  cudaStream_t s;
  std::vector<Span> spans;
  std::vector<cudaEvent_t> open;  // begin() events not yet closed (nesting allowed)
  explicit SpanClock(cudaStream_t s_) : s(s_) {}
  SpanClock(const SpanClock &) = delete;
  ~SpanClock()
  {
    for (auto &sp : spans) { g_event_pool.put(sp.a); g_event_pool.put(sp.b); }
    for (auto e : open) g_event_pool.put(e);
  }
  void begin()
  {
    cudaEvent_t e = g_event_pool.get();
    cudaEventRecord(e, s);
    open.push_back(e);
  }
  void end(int slot)
  {
    cudaEvent_t e = g_event_pool.get();
    cudaEventRecord(e, s);
    spans.push_back(Span{ open.back(), e, slot });
    open.pop_back();
  }
  // adds every closed span to ms[slot]; blocks until the last recorded event has completed
  void resolve(double *ms)
  {
    if (!spans.empty()) cudaEventSynchronize(spans.back().b);
    for (auto &sp : spans) {
      float t = 0;
      cudaEventSynchronize(sp.b);
      cudaEventElapsedTime(&t, sp.a, sp.b);
      ms[sp.slot] += t;
      g_event_pool.put(sp.a);
      g_event_pool.put(sp.b);
    }
    spans.clear();
  }
};

}  // namespace qg

using namespace qg;

struct qg_grid
{
  int dim;
  std::vector<double> breaks[3];
  long long n[3];
};
struct qg_func
{
  FuncDesc f;                 // f.coefs / f.comp are filled per device (domain / evaluation call)
  std::vector<double> coefs;  // Bezier coefficients, direction 0 fastest
  // composite (QG_COMPOSITE): host description; leaf l's Bezier coefficients in leaf_coefs[l]
  CompositeDesc comp;
  std::vector<double> leaf_coefs[kCompMaxLeaf];
};

// Device copies of what a function descriptor points to (Bezier coefficients, composite description).
struct FuncDeviceData
{
  DevBuf<double> coefs, leaf_coefs[kCompMaxLeaf];
  DevBuf<CompositeDesc> comp;
  // returns f->f with its pointers set to device copies uploaded on stream s
  FuncDesc upload(const qg_func *f, cudaStream_t s)
  {
    FuncDesc fd = f->f;
    if (!f->coefs.empty()) {
      coefs.alloc(f->coefs.size());
      coefs.upload(f->coefs.data(), f->coefs.size(), s);
      fd.coefs = coefs.p;
    }
    if (f->f.kind == QG_FUNC_COMPOSITE) {
      CompositeDesc h = f->comp;
      for (int l = 0; l < h.nleaf; ++l)
        if (!f->leaf_coefs[l].empty()) {
          leaf_coefs[l].alloc(f->leaf_coefs[l].size());
          leaf_coefs[l].upload(f->leaf_coefs[l].data(), f->leaf_coefs[l].size(), s);
          h.leaf[l].coefs = leaf_coefs[l].p;
        }
      comp.alloc(1);
      QG_CUDA(cudaMemcpyAsync(comp.p, &h, sizeof(h), cudaMemcpyHostToDevice, s));
      QG_CUDA(cudaStreamSynchronize(s));  // h is a local
      fd.comp = comp.p;
    }
    return fd;
  }
};

struct qg_rule
{
  int device = 0;
  long long n_entities = 0, n_points = 0;
  int point_dim = 0;
  bool has_normals = false;
  DevBuf<int> d_n_per;
  DevBuf<double> d_pts, d_wts, d_nrm;
  // transport encoding of a 3-D interior rule produced for the host API (d_pts is empty then): see download_rule
  bool packed = false;
  int pack_p = 0;
  DevBuf<double> d_pc, d_pk;
  DevBuf<unsigned char> d_k2;
  // pinned host mirrors
  int *h_n_per = nullptr;
  double *h_pts = nullptr, *h_wts = nullptr, *h_nrm = nullptr;
  bool downloaded = false;
  long long d2h_bytes = 0;  // bytes the download moved over PCIe
  size_t cap[4] = { 0, 0, 0, 0 };
  PassTimes times;
  ~qg_rule();
};

// reparameterization mesh (qg_reparam)
struct qg_mesh
{
  int device = 0, dim = 0, pd = 0, order = 0;
  long long n_points = 0, n_cells = 0, n_wire_edges = 0;
  DevBuf<double> d_pts;
  DevBuf<long long> d_conn, d_wires;
  // pinned host mirrors (pooled like the rule buffers), filled by qg_mesh_view
  double *h_pts = nullptr;
  long long *h_conn = nullptr, *h_wires = nullptr;
  size_t cap[3] = { 0, 0, 0 };
  bool downloaded = false;
  ~qg_mesh();
};

// packed custom coefficients (qg_pack_custom_coeffs)
struct qg_coeffs
{
  int device = 0;
  int elem_size = 8;
  long long n_rows = 0, n_cols = 0;  // rows include the appended ones
  DevBuf<unsigned char> d;
  void *h = nullptr;                 // pinned host copy (to_host)
  size_t cap = 0;
  long long d2h_bytes = 0, h2d_bytes = 0;
  ~qg_coeffs();
};

struct qg_dcells
{
  long long n = 0;
  DevBuf<long long> d_cells;
};

struct qg_domain
{
  int device = 0;
  int dim = 0;
  FuncDesc f;
  qg_grid grid;
  long long ncells = 0;
  cudaStream_t stream = nullptr;
  // device state
  DevBuf<double> d_breaks[3];
  GridDesc g;
  DevBuf<double> d_gauss;  // whole table (x,w) pairs
  FuncDeviceData fdata;    // device copies of what f points to (Bezier coefficients, composite description)
  DevBuf<unsigned char> d_status;
  DevBuf<int> d_rec_idx;
  DevBuf<long long> d_rec_cells;
  DevBuf<unsigned char> d_rec_stat;
  DevBuf<int> d_err;
  int k0 = 0, k1 = 0;  // slab of the last direction this domain classifies
  int num_sms = 148;
  // host mirrors (filled on first query)
  mutable std::atomic<bool> mirrored{false};  // release/acquire: readers outside the mutex see the filled vectors
  mutable std::vector<unsigned char> status;
  mutable std::vector<long long> rec_cells;
  mutable std::vector<unsigned char> rec_stat;
  long long n_full = 0, n_empty = 0, n_cut = 0, n_rec = 0;
  long long n_unf_rec = 0;  // recorded cells that own at least one facet with unfitted boundary
  void mirror() const;
  mutable std::mutex mtx;
  mutable Scanner scanner;
  const double *gauss_rule(int p) const { return d_gauss.p + 2 * (size_t)(p * (p - 1) / 2); }
};

void qg_domain::mirror() const
{
  std::lock_guard<std::mutex> lock(mtx);
  if (mirrored) return;
  QG_CUDA(cudaSetDevice(device));
  const int nf = 2 * dim;
  status.resize((size_t)ncells);
  rec_cells.resize((size_t)n_rec);
  rec_stat.resize((size_t)n_rec * nf);
  d_status.download(status.data(), (size_t)ncells, stream);
  if (n_rec) {
    d_rec_cells.download(rec_cells.data(), (size_t)n_rec, stream);
    d_rec_stat.download(rec_stat.data(), (size_t)n_rec * nf, stream);
  }
  QG_CUDA(cudaStreamSynchronize(stream));
  mirrored.store(true, std::memory_order_release);
}

namespace qg {

// ------------------------------------------------------------------------------------------------
// The pass schedule for one batch of jobs
// ------------------------------------------------------------------------------------------------
struct Pipeline
{
  const qg_domain &dom;
  cudaStream_t s;
  PipeArgs a;
  long long njobs = 0;
  DevBuf<double> job_box;
  DevBuf<unsigned char> job_cap;
  bool want_cap = false;  // record per job whether its walk hit the bisection cap (a.job_cap)
  DevBuf<int> item_cnt, nent0, cnt0, cnt1, cnt2;
  DevBuf<long long> item_off, off0, off1, off2, job_pt_off;
  DevBuf<ChainItem> items;
  DevBuf<double> ent0;
  DevBuf<Call1> call1;
  DevBuf<Call2> call2;
  PassTimes times;
  SpanClock clk;

  Pipeline(const qg_domain &d, int p, int sink_mode) : dom(d), s(d.stream), clk(d.stream)
  {
    std::memset(&a, 0, sizeof(a));
    a.f = d.f;
    a.g = d.g;
    a.p = p;
    a.sink_mode = sink_mode;
    a.gxw = d.gauss_rule(p);
    a.err = d.d_err.p;
  }

  long long read_total(const long long *dptr)
  {
    long long v = 0;
    QG_CUDA(cudaMemcpyAsync(&v, dptr, sizeof(v), cudaMemcpyDeviceToHost, s));
    QG_CUDA(cudaStreamSynchronize(s));
    return v;
  }

  template<int N> void launch_ctl(bool emit)
  {
    if (!njobs) return;
    launch_pipe(emit ? PK_CTL_EMIT : PK_CTL_COUNT, N, grid_for(njobs), s, a, nullptr);
  }

  // CTL only: the chain items (= leaves of the reference's recursion) of every job, in job order.
  template<int N> void run_items(const long long *d_job_cell, const int *d_job_type, long long nj)
  {
    njobs = nj;
    a.njobs = nj;
    a.job_cell = d_job_cell;
    a.job_type = d_job_type;
    job_box.alloc((size_t)nj * 6 + 1);
    a.job_box = job_box.p;
    item_cnt.alloc((size_t)nj + 1);
    item_cnt.zero(s);
    item_off.alloc((size_t)nj + 1);
    a.item_cnt = item_cnt.p;
    a.item_off = item_off.p;
    launch_ctl<N>(false);
    dom.scanner.run(item_cnt.p, item_off.p, (size_t)nj + 1, s);
    a.nitems = read_total(item_off.p + nj);
    items.alloc((size_t)a.nitems + 1);
    a.items = items.p;
    launch_ctl<N>(true);
    QG_CUDA(cudaGetLastError());
  }

  // Runs CTL, K0, K1, K2 and the scans; afterwards the number of nodes of every job is known.
  template<int N> void run_counts(const long long *d_job_cell, const int *d_job_type, long long nj)
  {
    njobs = nj;
    a.njobs = nj;
    a.job_cell = d_job_cell;
    a.job_type = d_job_type;
    job_box.alloc((size_t)nj * 6 + 1);
    a.job_box = job_box.p;
    if (want_cap) {
      job_cap.alloc((size_t)nj + 1);
      a.job_cap = job_cap.p;
    }
    item_cnt.alloc((size_t)nj + 1);
    item_cnt.zero(s);
    item_off.alloc((size_t)nj + 1);
    a.item_cnt = item_cnt.p;
    a.item_off = item_off.p;
    clk.begin();
    launch_ctl<N>(false);
    clk.end(0);
    clk.begin();
    dom.scanner.run(item_cnt.p, item_off.p, (size_t)nj + 1, s);
    a.nitems = read_total(item_off.p + nj);
    clk.end(5);
    items.alloc((size_t)a.nitems + 1);
    a.items = items.p;
    clk.begin();
    launch_ctl<N>(true);
    clk.end(0);
    // stage 0
    ent0.alloc((size_t)a.nitems * 2 * kMaxSeg0 + 1);
    nent0.alloc((size_t)a.nitems + 1);
    cnt0.alloc((size_t)a.nitems + 1);
    cnt0.zero(s);
    off0.alloc((size_t)a.nitems + 1);
    a.ent0 = ent0.p;
    a.nent0 = nent0.p;
    a.cnt0 = cnt0.p;
    a.off0 = off0.p;
    clk.begin();
    if (a.nitems) launch_pipe(PK_K0, N, grid_for(a.nitems), s, a, nullptr);
    clk.end(1);
    clk.begin();
    dom.scanner.run(cnt0.p, off0.p, (size_t)a.nitems + 1, s);
    a.ncall1 = read_total(off0.p + a.nitems);
    clk.end(5);
    // stage 1
    call1.alloc((size_t)a.ncall1 + 1);
    cnt1.alloc((size_t)a.ncall1 + 1);
    cnt1.zero(s);
    off1.alloc((size_t)a.ncall1 + 1);
    a.call1 = call1.p;
    a.cnt1 = cnt1.p;
    a.off1 = off1.p;
    clk.begin();
    if (a.ncall1) launch_pipe(PK_K1, N, grid_for(a.nitems, kExpParents), s, a, nullptr);
    clk.end(2);
    clk.begin();
    dom.scanner.run(cnt1.p, off1.p, (size_t)a.ncall1 + 1, s);
    a.ncall2 = read_total(off1.p + a.ncall1);
    clk.end(5);
    // stage 2
    call2.alloc((size_t)a.ncall2 + 1);
    cnt2.alloc((size_t)a.ncall2 + 1);
    cnt2.zero(s);
    off2.alloc((size_t)a.ncall2 + 1);
    a.call2 = call2.p;
    a.cnt2 = cnt2.p;
    a.off2 = off2.p;
    clk.begin();
    if (a.ncall2) launch_pipe(PK_K2, N, grid_for(a.ncall1, kExpParents), s, a, nullptr);
    clk.end(3);
    clk.begin();
    dom.scanner.run(cnt2.p, off2.p, (size_t)a.ncall2 + 1, s);
    a.npts = read_total(off2.p + a.ncall2);
    job_pt_off.alloc((size_t)nj + 1);
    job_points_kernel<<<QG_CNT(grid_for(nj + 1)), kBlock, 0, s>>>(a, job_pt_off.p);
    clk.end(5);
    times.counts[0] = a.nitems;
    times.counts[1] = a.ncall1;
    times.counts[2] = a.ncall2;
    times.counts[3] = a.npts;
    QG_CUDA(cudaGetLastError());
  }

  // K3: writes the nodes.  shift (may be null) relocates each job's nodes inside the output arrays.
  template<int N> void run_write(double *pts, double *wts, double *nrm, const long long *shift)
  {
    a.pts = pts;
    a.wts = wts;
    a.nrm = nrm;
    clk.begin();
    if (a.ncall2) {
      if (N == 3 && a.sink_mode == SINK_CELL && a.p <= kMaxGaussP) {
        // opt in to > 48 KB of dynamic shared memory once per device (the attribute is per device context)
        static std::once_flag attr_once[64];
        std::call_once(attr_once[dom.device & 63], [&]() {
          QG_CUDA(cudaFuncSetAttribute(k3_cell3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3SmemBytes));
          QG_CUDA(cudaFuncSetAttribute(k3_cell3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3SmemBytes));
        });
        const long long ntiles = (a.ncall2 + kC3Lines - 1) / kC3Lines;
        const long long nblk = ntiles < (long long)dom.num_sms * 4 ? ntiles : (long long)dom.num_sms * 4;
        if (a.pack_pk)
          k3_cell3_kernel<true><<<QG_CNT((unsigned)nblk), kC3Lines, kC3SmemBytes, s>>>(a, nullptr, ntiles);
        else
          k3_cell3_kernel<false><<<QG_CNT((unsigned)nblk), kC3Lines, kC3SmemBytes, s>>>(a, shift, ntiles);
      } else {
        launch_pipe(PK_K3, N, grid_for(a.ncall2, kK3Lines), s, a, shift);
      }
    }
    clk.end(4);
    QG_CUDA(cudaGetLastError());
  }
};

static void check_device_err(const qg_domain &d)
{
  int e = 0;
  QG_CUDA(cudaMemcpyAsync(&e, d.d_err.p, sizeof(int), cudaMemcpyDeviceToHost, d.stream));
  QG_CUDA(cudaStreamSynchronize(d.stream));
  if (e) {
    QG_CUDA(cudaMemsetAsync(d.d_err.p, 0, sizeof(int), d.stream));
    std::string m = "on-device capacity exceeded:";
    if (e & ERR_ROOT_CAP) m += " roots-per-line";
    if (e & ERR_STACK) m += " recursion-stack";
    if (e & ERR_ITEM_CAP) m += " chain-items";
    if (e & ERR_RP_CAP) m += " reparameterization-cells-per-leaf";
    throw EngineError(QG_ERR_CAPACITY, m);
  }
}

// ordered compaction helper: indices i in [0,n) with flag[i] != 0 -> positions; returns the count
static long long scan_flags(const qg_domain &d, DevBuf<int> &flag, DevBuf<long long> &off, long long n)
{
  // flag has n+1 entries with flag[n] == 0
  d.scanner.run(flag.p, off.p, (size_t)n + 1, d.stream);
  long long tot = 0;
  QG_CUDA(cudaMemcpyAsync(&tot, off.p + n, sizeof(tot), cudaMemcpyDeviceToHost, d.stream));
  QG_CUDA(cudaStreamSynchronize(d.stream));
  return tot;
}

// Ascending ids of the cells of the domain's slab whose status equals v (cells outside the slab are unclassified by
// construction).  One pass of cub::DeviceSelect::If over a counting iterator (stable, so the ids come out ascending): the
// 1-byte statuses of the slab are read once and only the selected ids are written -- no flag array, no offset array (the
// flag / scan / scatter passes of round 1 ran over the WHOLE grid, N x the cells of a rank on a sharded grid, and were the
// weak-scaling loss; a scan of the slab still wrote 8 B of offsets per cell: 1.07 ms on the 512^3 grid against 0.15 ms now).
struct StatusIs
{
  const unsigned char *status;  // of the slab: status[i] belongs to cell c0 + i
  long long c0;
  unsigned char v;
  __host__ __device__ bool operator()(const long long &id) const { return status[id - c0] == v; }
};
static long long compact_status_ids(const qg_domain &d, unsigned char v, DevBuf<long long> &ids)
{
  cudaStream_t s = d.stream;
  long long per_layer = 1;
  for (int k = 0; k + 1 < d.dim; ++k) per_layer *= d.grid.n[k];
  const long long c0 = (long long)d.k0 * per_layer, n = (long long)(d.k1 - d.k0) * per_layer;
  if (n <= 0) {
    ids.alloc(1);
    return 0;
  }
  if (n >= 2147483647LL) throw EngineError(QG_ERR_CAPACITY, "a slab holds more than 2^31 cells: split it into smaller slabs");
  // upper bound of the selection = the slab; the caching allocator hands the block back afterwards
  DevBuf<long long> tmp_ids((size_t)n + 1), count(1);
  cub::CountingInputIterator<long long> first(c0);
  StatusIs pred{ d.d_status.p + c0, c0, v };
  size_t bytes = 0;
  QG_CUDA(cub::DeviceSelect::If(nullptr, bytes, first, tmp_ids.p, count.p, (int)n, pred, s));
  DevBuf<unsigned char> work(bytes + 16);
  QG_CUDA(cub::DeviceSelect::If(work.p, bytes, first, tmp_ids.p, count.p, (int)n, pred, s));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  long long tot = 0;
  QG_CUDA(cudaMemcpyAsync(&tot, count.p, sizeof(tot), cudaMemcpyDeviceToHost, s));
  QG_CUDA(cudaStreamSynchronize(s));
  ids.alloc((size_t)tot + 1);
  if (tot > 0) QG_CUDA(cudaMemcpyAsync(ids.p, tmp_ids.p, (size_t)tot * sizeof(long long), cudaMemcpyDeviceToDevice, s));
  return tot;
}

// ------------------------------------------------------------------------------------------------
// Domain construction: kd-tree pruning + classification of undetermined cells and their facets
// ------------------------------------------------------------------------------------------------
// QG_TRACE=1: per-phase host wall times (each mark synchronises the stream; debugging aid only)
struct PhaseTrace
{
  bool on;
  cudaStream_t s;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTrace(cudaStream_t st) : on(std::getenv("QG_TRACE") != nullptr), s(st), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what)
  {
    if (!on) return;
    cudaStreamSynchronize(s);
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[qg trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

template<int N> static void classify_domain(qg_domain &dm)
{
  PhaseTrace tr(dm.stream);
  cudaStream_t s = dm.stream;
  const long long nc = dm.ncells;
  dm.d_status.alloc((size_t)nc);
  QG_CUDA(cudaMemsetAsync(dm.d_status.p, 0xFF, (size_t)nc, s));
  dm.d_rec_idx.alloc((size_t)nc);
  fill_i32_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(nc, -1, dm.d_rec_idx.p);

  tr.mark("status init");
  // --- kd-tree.  Small slabs: one cooperative launch (no host round trips; measured 0.7 ms vs 1.0 ms on a 32-layer
  // slab of C3).  Large slabs, or if a node list overflows: one launch per level, which spreads the wide bottom
  // levels over the whole GPU (1.3 ms vs 4.3 ms on the full 256^3 grid).
  std::vector<KdNode> root(1);
  for (int d = 0; d < 3; ++d) {
    root[0].lo[d] = 0;
    root[0].hi[d] = d < N ? (int)dm.grid.n[d] : 1;
  }
  bool tree_done = false;
  {
    long long slab_cells = 1;
    for (int d = 0; d < N; ++d) slab_cells *= (d == N - 1) ? (long long)(dm.k1 - dm.k0) : dm.grid.n[d];
    const int cap = (int)std::min<long long>(std::max<long long>(1 << 20, slab_cells / 4), 1LL << 26);
    int blocks_per_sm = 0, coop = 0;
    QG_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dm.device));
    QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kd_tree_kernel<N>, 256, 0));
    if (coop && blocks_per_sm > 0 && slab_cells <= (1LL << 21)) {
      DevBuf<KdNode> buf_a((size_t)cap), buf_b((size_t)cap), uni((size_t)cap);
      DevBuf<unsigned char> unis((size_t)cap);
      DevBuf<int> ctr(129);
      ctr.zero(s);
      const int one = 1;
      ctr.upload(&one, 1, s);
      buf_a.upload(root.data(), 1, s);
      FuncDesc f = dm.f;
      GridDesc g = dm.g;
      KdNode *pa = buf_a.p, *pb = buf_b.p, *pu = uni.p;
      unsigned char *pus = unis.p, *pst = dm.d_status.p;
      int *pc = ctr.p;
      int c = cap, uc = cap, k0 = dm.k0, k1 = dm.k1;
      void *args[] = { &f, &g, &pa, &pb, &c, &pu, &pus, &uc, &pc, &pst, &k0, &k1 };
      const int nblk = std::min(blocks_per_sm, 4) * dm.num_sms;
      g_launches.fetch_add(1, std::memory_order_relaxed);
      QG_CUDA(cudaLaunchCooperativeKernel((const void *)kd_tree_kernel<N>, dim3((unsigned)nblk), dim3(256), args, 0, s));
      int h[129];
      ctr.download(h, 129, s);
      QG_CUDA(cudaStreamSynchronize(s));
      tree_done = h[128] == 0;
      if (!tree_done) QG_CUDA(cudaMemsetAsync(dm.d_status.p, 0xFF, (size_t)nc, s));
    }
  }
  if (!tree_done) {
    DevBuf<KdNode> cur(1), next;
    cur.upload(root.data(), 1, s);
    int ncur = 1;
    DevBuf<int> counters(2);
    while (ncur > 0) {
      next.alloc((size_t)ncur * 2);
      DevBuf<KdNode> uni((size_t)ncur);
      DevBuf<unsigned char> unis((size_t)ncur);
      counters.zero(s);
      kd_level_kernel<N><<<QG_CNT(grid_for(ncur)), kBlock, 0, s>>>(
        dm.f, dm.g, cur.p, ncur, next.p, counters.p, uni.p, unis.p, counters.p + 1, dm.d_status.p, dm.k0, dm.k1);
      int h[2] = { 0, 0 };
      counters.download(h, 2, s);
      QG_CUDA(cudaStreamSynchronize(s));
      // no second synchronisation: uni / unis go back to the stream-ordered cache, whose next user is queued on this
      // same stream behind the fill
      if (h[1] > 0) kd_fill_kernel<N><<<QG_CNT(h[1]), 256, 0, s>>>(dm.g, uni.p, unis.p, dm.d_status.p, dm.k0, dm.k1);
      cur = std::move(next);
      ncur = h[0];
    }
  }
  tr.mark("kd-tree");
  // --- undetermined cells, ascending ids
  DevBuf<long long> ucells;
  const long long nu = compact_status_ids(dm, ST_UNDET, ucells);

  DevBuf<long long> rcells;
  DevBuf<unsigned char> rstat;
  long long nrec = 0;
  tr.mark("undetermined list");
  if (nu > 0) {
    // --- n = 1 volume rule of every undetermined cell
    DevBuf<int> utype((size_t)nu);
    fill_types_kernel<<<QG_CNT(grid_for(nu)), kBlock, 0, s>>>(nu, JOB_VOLUME, utype.p);
    DevBuf<unsigned char> utmp((size_t)nu);
    {
      Pipeline pl(dm, 1, SINK_RAW);
      pl.run_counts<N>(ucells.p, utype.p, nu);
      DevBuf<double> pts((size_t)pl.a.npts * N + 1), wts((size_t)pl.a.npts + 1);
      pl.run_write<N>(pts.p, wts.p, nullptr, nullptr);
      classify_cells_kernel<<<QG_CNT(grid_for(nu)), kBlock, 0, s>>>(nu, pl.job_pt_off.p, wts.p, pl.job_box.p, N, utmp.p);
      QG_CUDA(cudaStreamSynchronize(s));
    }
    tr.mark("cell n=1 rules");
    // --- cells that are cut or full with unfitted boundary get their 2N facets classified
    DevBuf<int> isv((size_t)nu + 1);
    isv.zero(s);
    DevBuf<long long> voff((size_t)nu + 1);
    set_status_from_tmp_kernel<<<QG_CNT(grid_for(nu)), kBlock, 0, s>>>(nu, ucells.p, utmp.p, dm.d_status.p, isv.p);
    const long long nv = scan_flags(dm, isv, voff, nu);
    if (nv > 0) {
      DevBuf<long long> vcells((size_t)nv);
      DevBuf<unsigned char> vtmp((size_t)nv);
      gather_v_kernel<<<QG_CNT(grid_for(nu)), kBlock, 0, s>>>(nu, ucells.p, utmp.p, isv.p, voff.p, vcells.p, vtmp.p);
      const int nf = 2 * N;
      const long long nfj = nv * nf;
      DevBuf<unsigned char> fstat((size_t)nfj);
      if (N == 3 && !std::getenv("QG_NO_FACET_PREFILTER")) {
        // most facets of a cut cell are decided by the interval test their rule starts with; only the rest run the pipeline
        DevBuf<int> fflag((size_t)nfj + 1);
        fflag.zero(s);
        DevBuf<long long> foff((size_t)nfj + 1);
        facet_prefilter_kernel<N><<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(dm.f, dm.g, nfj, vcells.p, vtmp.p, fstat.p, fflag.p);
        // one rule per face: the lower facet of a cell becomes an alias of its neighbour's upper facet (QG_NO_FACET_ALIAS=1: off).
        // The walks of the two jobs are identical decision for decision EXCEPT at the 16-level bisection cap, whose test reads
        // the unused PsiCode slots (sign -1 / side 0: for an UPPER facet's job that is phi on the cell's LOWER face): aliases
        // whose owner hit the cap are recomputed with their own job below (336 of 3.04 M facet records on C5).
        DevBuf<long long> alias;
        const bool use_alias = !std::getenv("QG_NO_FACET_ALIAS");
        if (use_alias) {
          long long per_layer = 1;
          for (int k = 0; k + 1 < N; ++k) per_layer *= dm.grid.n[k];
          const long long c0 = (long long)dm.k0 * per_layer, nslab = (long long)(dm.k1 - dm.k0) * per_layer;
          DevBuf<int> vidx((size_t)nslab + 1);
          QG_CUDA(cudaMemsetAsync(vidx.p, 0xFF, ((size_t)nslab + 1) * sizeof(int), s));
          alias.alloc((size_t)nfj + 1);
          QG_CUDA(cudaMemsetAsync(alias.p, 0xFF, ((size_t)nfj + 1) * sizeof(long long), s));
          vidx_kernel<<<QG_CNT(grid_for(nv)), kBlock, 0, s>>>(nv, vcells.p, c0, vidx.p);
          facet_alias_kernel<N><<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(dm.g, nfj, vcells.p, vidx.p, c0, nslab, fflag.p, alias.p);
        }
        const long long nfp = scan_flags(dm, fflag, foff, nfj);
        if (nfp > 0) {
          DevBuf<long long> fcell((size_t)nfp), fslot((size_t)nfp);
          DevBuf<int> ftype((size_t)nfp);
          gather_facet_slots_kernel<<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(nfj, vcells.p, nf, fflag.p, foff.p, fcell.p, ftype.p, fslot.p);
          Pipeline pl(dm, 1, SINK_RAW);
          pl.want_cap = use_alias && kPsiDefaultSign != 0;
          pl.run_counts<N>(fcell.p, ftype.p, nfp);
          DevBuf<double> pts((size_t)pl.a.npts * N + 1), wts((size_t)pl.a.npts + 1);
          pl.run_write<N>(pts.p, wts.p, nullptr, nullptr);
          classify_facets_kernel<N><<<QG_CNT(grid_for(nfp)), kBlock, 0, s>>>(
            dm.f, nfp, ftype.p, pl.job_pt_off.p, pts.p, wts.p, pl.job_box.p, vtmp.p, fslot.p, fstat.p);
          if (use_alias)
            classify_alias_facets_kernel<N><<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(
              dm.f, nfj, alias.p, foff.p, pl.job_pt_off.p, pts.p, wts.p, pl.job_box.p, vtmp.p, fstat.p);
          if (pl.want_cap) {
            // aliases whose owner ran into the bisection cap: their own job, classified over the alias result
            DevBuf<int> rflag((size_t)nfj + 1);
            rflag.zero(s);
            DevBuf<long long> roff((size_t)nfj + 1);
            alias_redo_flag_kernel<<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(nfj, alias.p, foff.p, pl.job_cap.p, rflag.p);
            const long long nredo = scan_flags(dm, rflag, roff, nfj);
            if (nredo > 0) {
              DevBuf<long long> rcell((size_t)nredo), rslot((size_t)nredo);
              DevBuf<int> rtype((size_t)nredo);
              gather_facet_slots_kernel<<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(nfj, vcells.p, nf, rflag.p, roff.p, rcell.p, rtype.p, rslot.p);
              Pipeline pr(dm, 1, SINK_RAW);
              pr.run_counts<N>(rcell.p, rtype.p, nredo);
              DevBuf<double> rpts((size_t)pr.a.npts * N + 1), rwts((size_t)pr.a.npts + 1);
              pr.run_write<N>(rpts.p, rwts.p, nullptr, nullptr);
              classify_facets_kernel<N><<<QG_CNT(grid_for(nredo)), kBlock, 0, s>>>(
                dm.f, nredo, rtype.p, pr.job_pt_off.p, rpts.p, rwts.p, pr.job_box.p, vtmp.p, rslot.p, fstat.p);
              QG_CUDA(cudaStreamSynchronize(s));
            }
          }
          QG_CUDA(cudaStreamSynchronize(s));  // pts / wts / the pipeline's buffers are released at the end of this scope
        }
        QG_CUDA(cudaStreamSynchronize(s));
      } else {
        DevBuf<long long> fcell((size_t)nfj);
        DevBuf<int> ftype((size_t)nfj);
        make_facet_jobs_kernel<<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(nv, vcells.p, nf, fcell.p, ftype.p);
        Pipeline pl(dm, 1, SINK_RAW);
        pl.run_counts<N>(fcell.p, ftype.p, nfj);
        DevBuf<double> pts((size_t)pl.a.npts * N + 1), wts((size_t)pl.a.npts + 1);
        pl.run_write<N>(pts.p, wts.p, nullptr, nullptr);
        classify_facets_kernel<N><<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(
          dm.f, nfj, ftype.p, pl.job_pt_off.p, pts.p, wts.p, pl.job_box.p, vtmp.p, nullptr, fstat.p);
        QG_CUDA(cudaStreamSynchronize(s));
      }
      tr.mark("facet n=1 rules");
      DevBuf<int> keep((size_t)nv + 1);
      keep.zero(s);
      DevBuf<long long> koff((size_t)nv + 1);
      finalize_cells_kernel<<<QG_CNT(grid_for(nv)), kBlock, 0, s>>>(nv, vcells.p, vtmp.p, fstat.p, nf, dm.d_status.p, keep.p);
      nrec = scan_flags(dm, keep, koff, nv);
      rcells.alloc((size_t)nrec + 1);
      rstat.alloc((size_t)nrec * nf + 1);
      DevBuf<int> nunf(1);
      nunf.zero(s);
      gather_records_kernel<<<QG_CNT(grid_for(nv)), kBlock, 0, s>>>(
        nv, vcells.p, fstat.p, nf, keep.p, koff.p, rcells.p, rstat.p, dm.d_rec_idx.p, nunf.p);
      int h_nunf = 0;
      nunf.download(&h_nunf, 1, s);
      QG_CUDA(cudaStreamSynchronize(s));
      dm.n_unf_rec = h_nunf;
    }
  }
  tr.mark("records");
  check_device_err(dm);
  dm.d_rec_cells = std::move(rcells);
  dm.d_rec_stat = std::move(rstat);
  dm.n_rec = nrec;
  DevBuf<unsigned long long> hist(4);
  hist.zero(s);
  {
    long long per_layer = 1;
    for (int k = 0; k + 1 < dm.dim; ++k) per_layer *= dm.grid.n[k];
    status_hist_kernel<<<QG_CNT(592), 256, 0, s>>>(dm.d_status.p + (long long)dm.k0 * per_layer, (long long)(dm.k1 - dm.k0) * per_layer, hist.p);
  }
  unsigned long long hh[4] = { 0, 0, 0, 0 };
  hist.download(hh, 3, s);
  QG_CUDA(cudaStreamSynchronize(s));
  dm.n_cut = (long long)hh[ST_CUT];
  dm.n_full = (long long)hh[ST_FULL];
  dm.n_empty = (long long)hh[ST_EMPTY];
  tr.mark("histogram");
}

// ------------------------------------------------------------------------------------------------
// Rules
// ------------------------------------------------------------------------------------------------
static void alloc_rule_outputs(qg_rule &r, long long ne, long long npts, int pdim, bool normals)
{
  r.n_entities = ne;
  r.n_points = npts;
  r.point_dim = pdim;
  r.has_normals = normals;
  r.d_pts.alloc((size_t)npts * pdim + 1);
  r.d_wts.alloc((size_t)npts + 1);
  if (normals) r.d_nrm.alloc((size_t)npts * pdim + 1);
}

// create_quadrature / create_unfitted_bound_quadrature over device-resident cells
template<int N>
static qg_rule *quad_cells_impl(const qg_domain &dm, const long long *d_cells, long long ne, int p, bool surf, bool for_host = false)
{
  cudaStream_t s = dm.stream;
  std::unique_ptr<qg_rule> rule(new qg_rule());
  rule->device = dm.device;
  SpanClock clk(s);  // slot 6: passes outside the pipeline, slot 7: whole call
  clk.begin();       // whole call
  clk.begin();
  // entity kinds and the job list (cut cells, in input order)
  DevBuf<int> kind((size_t)ne + 1), flag((size_t)ne + 1), bad(1);
  DevBuf<long long> off((size_t)ne + 1);
  bad.zero(s);
  flag.zero(s);
  if (ne) {
    if (surf)
      entity_kind_unf_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells, dm.ncells, dm.d_status.p, dm.d_rec_idx.p, kind.p, bad.p);
    else
      entity_kind_cells_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells, dm.ncells, dm.d_status.p, dm.d_rec_idx.p, kind.p, bad.p);
    flag_kind_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, kind.p, ENT_JOB, flag.p);
  }
  const long long nj = scan_flags(dm, flag, off, ne);
  int hbad = 0;
  bad.download(&hbad, 1, s);
  QG_CUDA(cudaStreamSynchronize(s));
  if (hbad) throw EngineError(QG_ERR_INVALID, "cell id out of range");
  DevBuf<long long> job_cell((size_t)nj + 1);
  DevBuf<int> job_ent((size_t)nj + 1), job_type((size_t)nj + 1);
  if (ne) gather_jobs_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells, flag.p, off.p, job_cell.p, job_ent.p);
  if (nj) fill_types_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, surf ? JOB_SURFACE : JOB_VOLUME, job_type.p);
  clk.end(6);

  Pipeline pl(dm, p, surf ? SINK_SURF : SINK_CELL);
  pl.run_counts<N>(job_cell.p, job_type.p, nj);

  clk.begin();
  int gauss_n = 1;
  for (int d = 0; d < N; ++d) gauss_n *= p;
  rule->d_n_per.alloc((size_t)ne + 1);
  rule->d_n_per.zero(s);
  DevBuf<long long> ent_off((size_t)ne + 1), shift((size_t)nj + 1);
  // The surface rule needs the purge / compaction passes only if some cell owns a facet with unfitted boundary.
  const bool direct = !surf || dm.n_unf_rec == 0;
  if (direct) {
    if (ne && !surf) entity_counts_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, kind.p, gauss_n, rule->d_n_per.p);
    if (nj) job_counts_to_entities_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, pl.job_pt_off.p, nullptr, rule->d_n_per.p);
    dm.scanner.run(rule->d_n_per.p, ent_off.p, (size_t)ne + 1, s);
    const long long npts = pl.read_total(ent_off.p + ne);
    // Host API, 3-D interior rule, every listed cell a cut cell (no Gauss fill, no relocation): the rule is produced in
    // its transport encoding -- 0.6 of the bytes cross PCIe and the host threads of download_rule lay out the rows.
    const bool pack = for_host && N == 3 && !surf && nj == ne && npts > 0 && p <= kMaxGaussP && npts % p == 0;
    if (pack) {
      rule->n_entities = ne;
      rule->n_points = npts;
      rule->point_dim = N;
      rule->has_normals = false;
      rule->packed = true;
      rule->pack_p = p;
      const size_t ng = (size_t)(npts / p);
      rule->d_wts.alloc((size_t)npts + 1);
      rule->d_pk.alloc((size_t)npts + 1);
      rule->d_pc.alloc(ng * N + 1);
      rule->d_k2.alloc(ng + 1);
      pl.a.pack_pc = rule->d_pc.p;
      pl.a.pack_k2 = rule->d_k2.p;
      pl.a.pack_pk = rule->d_pk.p;
      clk.end(6);
      pl.run_write<N>(nullptr, rule->d_wts.p, nullptr, nullptr);
      clk.begin();
      clk.end(6);
      QG_CUDA(cudaGetLastError());
      check_device_err(dm);
      clk.end(7);
      pl.clk.resolve(pl.times.ms);
      rule->times = pl.times;
      clk.resolve(rule->times.ms);
      return rule.release();
    }
    alloc_rule_outputs(*rule, ne, npts, N, surf);
    if (nj) job_shift_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, pl.job_pt_off.p, ent_off.p, shift.p);
    clk.end(6);
    pl.run_write<N>(rule->d_pts.p, rule->d_wts.p, surf ? rule->d_nrm.p : nullptr, shift.p);
    clk.begin();
    if (ne && !surf) gauss_fill_kernel<<<QG_CNT(grid_for(ne * 32, 256)), 256, 0, s>>>(ne, kind.p, ent_off.p, N, p, dm.gauss_rule(p), rule->d_pts.p, rule->d_wts.p);
    clk.end(6);
  } else {
    // surface rule: nodes go to a staging rule first; facet-coincident nodes are purged when a cell has
    // facets carrying unfitted boundary (rare), then the rule is compacted into its final place.
    DevBuf<double> tp((size_t)pl.a.npts * N + 1), tw((size_t)pl.a.npts + 1), tn((size_t)pl.a.npts * N + 1);
    clk.end(6);
    pl.run_write<N>(tp.p, tw.p, tn.p, nullptr);
    clk.begin();
    DevBuf<unsigned char> keep((size_t)pl.a.npts + 1);
    DevBuf<int> job_keep((size_t)nj + 1);
    if (nj)
      purge_flags_surf_kernel<N><<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(dm.f, nj, job_cell.p, pl.job_pt_off.p, pl.job_box.p,
        dm.d_rec_idx.p, dm.d_rec_stat.p, tp.p, tn.p, keep.p, job_keep.p);
    if (nj) job_counts_to_entities_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, pl.job_pt_off.p, job_keep.p, rule->d_n_per.p);
    dm.scanner.run(rule->d_n_per.p, ent_off.p, (size_t)ne + 1, s);
    const long long npts = pl.read_total(ent_off.p + ne);
    alloc_rule_outputs(*rule, ne, npts, N, true);
    if (nj) {
      job_dst_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, ent_off.p, shift.p);
      compact_points_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, pl.job_pt_off.p, keep.p, shift.p, N, tp.p, tw.p, tn.p,
        rule->d_pts.p, rule->d_wts.p, rule->d_nrm.p);
    }
    QG_CUDA(cudaStreamSynchronize(s));
    clk.end(6);
  }
  QG_CUDA(cudaGetLastError());
  check_device_err(dm);
  clk.end(7);
  pl.clk.resolve(pl.times.ms);
  rule->times = pl.times;
  clk.resolve(rule->times.ms);
  return rule.release();
}

// create_facets_quadrature_{interior,exterior}_integral (cut_quadrature.cpp:178-262) and the facet part of
// create_unfitted_bound_quadrature (mode FMODE_UNF_BDRY, d_facets == nullptr: all 2N facets of every cell).
// ent_off_out (optional) receives the exclusive scan of the per-entity counts.
template<int N>
static qg_rule *quad_facets_impl(const qg_domain &dm, const long long *d_cells, const int *d_facets, long long ne, int p,
  int mode, int exclude_ext, DevBuf<long long> *ent_off_out)
{
  cudaStream_t s = dm.stream;
  std::unique_ptr<qg_rule> rule(new qg_rule());
  rule->device = dm.device;
  SpanClock clk(s);
  clk.begin();
  DevBuf<int> kind((size_t)ne + 1), flag((size_t)ne + 1), bad(1);
  DevBuf<unsigned char> pflags((size_t)ne + 1);
  DevBuf<long long> off((size_t)ne + 1);
  bad.zero(s);
  flag.zero(s);
  if (ne) {
    facet_entity_kernel<N><<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(dm.g, ne, d_cells, d_facets, 1, mode, exclude_ext, dm.ncells,
      dm.d_status.p, dm.d_rec_idx.p, dm.d_rec_stat.p, kind.p, pflags.p, bad.p);
    flag_kind_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, kind.p, ENT_JOB, flag.p);
  }
  const long long nj = scan_flags(dm, flag, off, ne);
  int hbad = 0;
  bad.download(&hbad, 1, s);
  QG_CUDA(cudaStreamSynchronize(s));
  if (hbad) throw EngineError(QG_ERR_INVALID, "cell id or local facet id out of range");
  DevBuf<long long> job_cell((size_t)nj + 1);
  DevBuf<int> job_ent((size_t)nj + 1), job_type((size_t)nj + 1);
  if (ne)
    gather_facet_jobs_kernel<N><<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells, d_facets, 1, flag.p, off.p, job_cell.p, job_ent.p,
      job_type.p);
  Pipeline pl(dm, p, SINK_FACET);
  pl.run_counts<N>(job_cell.p, job_type.p, nj);
  // nodes go to a staging rule, are purged (level-set nodes of the wrong side / cut parts), then compacted
  DevBuf<double> tp((size_t)pl.a.npts * (N - 1) + 1), tw((size_t)pl.a.npts + 1);
  pl.run_write<N>(tp.p, tw.p, nullptr, nullptr);
  DevBuf<unsigned char> keep((size_t)pl.a.npts + 1);
  DevBuf<int> job_keep((size_t)nj + 1);
  if (nj)
    purge_flags_facet_kernel<N><<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(dm.f, nj, job_type.p, job_ent.p, pflags.p, pl.job_pt_off.p,
      pl.job_box.p, tp.p, keep.p, job_keep.p);
  int gauss_n = 1;
  for (int d = 0; d < N - 1; ++d) gauss_n *= p;
  rule->d_n_per.alloc((size_t)ne + 1);
  rule->d_n_per.zero(s);
  DevBuf<long long> ent_off((size_t)ne + 1), dst((size_t)nj + 1);
  if (ne) entity_counts_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, kind.p, gauss_n, rule->d_n_per.p);
  if (nj) job_counts_to_entities_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, pl.job_pt_off.p, job_keep.p, rule->d_n_per.p);
  dm.scanner.run(rule->d_n_per.p, ent_off.p, (size_t)ne + 1, s);
  const long long npts = pl.read_total(ent_off.p + ne);
  alloc_rule_outputs(*rule, ne, npts, N - 1, false);
  if (nj) {
    job_dst_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, ent_off.p, dst.p);
    compact_points_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, pl.job_pt_off.p, keep.p, dst.p, N - 1, tp.p, tw.p, nullptr,
      rule->d_pts.p, rule->d_wts.p, nullptr);
  }
  if (ne) gauss_fill_kernel<<<QG_CNT(grid_for(ne * 32, 256)), 256, 0, s>>>(ne, kind.p, ent_off.p, N - 1, p, dm.gauss_rule(p), rule->d_pts.p, rule->d_wts.p);
  QG_CUDA(cudaStreamSynchronize(s));
  QG_CUDA(cudaGetLastError());
  check_device_err(dm);
  if (ent_off_out) *ent_off_out = std::move(ent_off);
  clk.end(7);
  pl.clk.resolve(pl.times.ms);
  rule->times = pl.times;
  clk.resolve(rule->times.ms);
  return rule.release();
}

// create_unfitted_bound_quadrature with include_facet_unf_bdry (impl_quadrature.cpp:349-389,540-592): the surface
// rule of every cell followed by the rules of its facets that carry unfitted boundary.
template<int N>
static qg_rule *quad_unf_bdry_with_facets_impl(const qg_domain &dm, const long long *d_cells, long long ne, int p, int exclude_ext)
{
  cudaStream_t s = dm.stream;
  std::unique_ptr<qg_rule> surf(quad_cells_impl<N>(dm, d_cells, ne, p, true));
  if (dm.n_unf_rec == 0) return surf.release();  // no facet carries unfitted boundary: nothing to add
  DevBuf<long long> foff;
  std::unique_ptr<qg_rule> fac(quad_facets_impl<N>(dm, d_cells, nullptr, ne * 2 * N, p, FMODE_UNF_BDRY, exclude_ext, &foff));
  std::unique_ptr<qg_rule> rule(new qg_rule());
  rule->device = dm.device;
  rule->d_n_per.alloc((size_t)ne + 1);
  rule->d_n_per.zero(s);
  DevBuf<long long> soff((size_t)ne + 1), ooff((size_t)ne + 1);
  dm.scanner.run(surf->d_n_per.p, soff.p, (size_t)ne + 1, s);
  if (ne) merge_counts_kernel<N><<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, surf->d_n_per.p, fac->d_n_per.p, rule->d_n_per.p);
  dm.scanner.run(rule->d_n_per.p, ooff.p, (size_t)ne + 1, s);
  long long npts = 0;
  QG_CUDA(cudaMemcpyAsync(&npts, ooff.p + ne, sizeof(npts), cudaMemcpyDeviceToHost, s));
  QG_CUDA(cudaStreamSynchronize(s));
  alloc_rule_outputs(*rule, ne, npts, N, true);
  if (ne)
    merge_rules_kernel<N><<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, ooff.p, soff.p, surf->d_n_per.p, surf->d_pts.p, surf->d_wts.p,
      surf->d_nrm.p, foff.p, fac->d_n_per.p, fac->d_pts.p, fac->d_wts.p, rule->d_pts.p, rule->d_wts.p, rule->d_nrm.p);
  QG_CUDA(cudaStreamSynchronize(s));
  QG_CUDA(cudaGetLastError());
  rule->times = surf->times;
  return rule.release();
}

// Pinned host blocks are expensive to create (page-locking GBs takes seconds); blocks released by freed
// rules are kept and reused by later rules of at most the same size.
struct PinnedPool
{
  std::mutex m;
  std::vector<std::pair<size_t, void *>> free_blocks;
  void *get(size_t bytes, size_t *cap)
  {
    {
      std::lock_guard<std::mutex> lock(m);
      size_t best = (size_t)-1;
      for (size_t i = 0; i < free_blocks.size(); ++i)
        if (free_blocks[i].first >= bytes && (best == (size_t)-1 || free_blocks[i].first < free_blocks[best].first)) best = i;
      if (best != (size_t)-1) {
        void *p = free_blocks[best].second;
        *cap = free_blocks[best].first;
        free_blocks.erase(free_blocks.begin() + (long)best);
        return p;
      }
    }
    void *p = nullptr;
    QG_CUDA(cudaMallocHost(&p, bytes));
    *cap = bytes;
    return p;
  }
  void put(void *p, size_t cap)
  {
    if (!p) return;
    std::lock_guard<std::mutex> lock(m);
    free_blocks.emplace_back(cap, p);
  }
  // frees every pooled (currently unused) pinned block; returns the bytes given back
  size_t release_all()
  {
    std::lock_guard<std::mutex> lock(m);
    size_t bytes = 0;
    for (auto &b : free_blocks) {
      cudaFreeHost(b.second);
      bytes += b.first;
    }
    free_blocks.clear();
    return bytes;
  }
};
static PinnedPool g_pinned;

}  // namespace qg
qg_coeffs::~qg_coeffs() { qg::g_pinned.put(h, cap); }
qg_mesh::~qg_mesh()
{
  qg::g_pinned.put(h_pts, cap[0]);
  qg::g_pinned.put(h_conn, cap[1]);
  qg::g_pinned.put(h_wires, cap[2]);
}
qg_rule::~qg_rule()
{
  qg::g_pinned.put(h_n_per, cap[0]);
  qg::g_pinned.put(h_pts, cap[1]);
  qg::g_pinned.put(h_wts, cap[2]);
  qg::g_pinned.put(h_nrm, cap[3]);
}
namespace qg {

// Rows of groups [g0,g1): row j of group g = pc[g] with pk[g*p + j] in slot k2[g].  Four rows are three 32-byte vectors;
// they are written with streaming stores (the rows are not read again here, and the arrays are tens of GB).
static void expand_groups(const double *h_pc, const unsigned char *h_k2, const double *h_pk, double *pts, size_t p, size_t g0,
  size_t g1)
{
#if defined(__AVX__)
  const bool vec = (p % 4 == 0) && ((reinterpret_cast<uintptr_t>(pts) & 31) == 0);
#else
  const bool vec = false;
#endif
  for (size_t g = g0; g < g1; ++g) {
    const double c[3] = { h_pc[g * 3], h_pc[g * 3 + 1], h_pc[g * 3 + 2] };
    const unsigned k2 = h_k2[g];
    const double *pk = h_pk + g * p;
    double *row = pts + g * p * 3;
#if defined(__AVX__)
    if (vec) {
      // Four rows (x,y,z) are the three vectors  v0 = r0.x r0.y r0.z r1.x | v1 = r1.y r1.z r2.x r2.y | v2 = r2.z r3.x r3.y r3.z:
      // slot e = 3 r + d holds pk[r] where d == k2 and the group's constant c[d] elsewhere.  The constant patterns are
      // built once per group and the varying slots blended in -- no scalar staging array (the round-1 loop wrote 12 scalars
      // and re-read them as vectors: store-forwarding stalls held a thread at 5.8 GB/s of rows; this form reaches 14.5 GB/s,
      // so that 16 host threads keep up with the PCIe link instead of trailing it).
      const __m256d c0 = _mm256_set_pd(c[0], c[2], c[1], c[0]);  // (set_pd lists the lanes high to low)
      const __m256d c1 = _mm256_set_pd(c[1], c[0], c[2], c[1]);
      const __m256d c2 = _mm256_set_pd(c[2], c[1], c[0], c[2]);
      for (size_t j = 0; j < p; j += 4, row += 12) {
        const double q0 = pk[j], q1 = pk[j + 1], q2 = pk[j + 2], q3 = pk[j + 3];
        __m256d v0, v1, v2;
        if (k2 == 0) {         // x slots: e = 0, 3, 6, 9
          v0 = _mm256_blend_pd(c0, _mm256_set_pd(q1, 0.0, 0.0, q0), 0x9);
          v1 = _mm256_blend_pd(c1, _mm256_set_pd(0.0, q2, 0.0, 0.0), 0x4);
          v2 = _mm256_blend_pd(c2, _mm256_set_pd(0.0, 0.0, q3, 0.0), 0x2);
        } else if (k2 == 1) {  // y slots: e = 1, 4, 7, 10
          v0 = _mm256_blend_pd(c0, _mm256_set_pd(0.0, 0.0, q0, 0.0), 0x2);
          v1 = _mm256_blend_pd(c1, _mm256_set_pd(q2, 0.0, 0.0, q1), 0x9);
          v2 = _mm256_blend_pd(c2, _mm256_set_pd(0.0, q3, 0.0, 0.0), 0x4);
        } else {               // z slots: e = 2, 5, 8, 11
          v0 = _mm256_blend_pd(c0, _mm256_set_pd(0.0, q0, 0.0, 0.0), 0x4);
          v1 = _mm256_blend_pd(c1, _mm256_set_pd(0.0, 0.0, q1, 0.0), 0x2);
          v2 = _mm256_blend_pd(c2, _mm256_set_pd(q3, 0.0, 0.0, q2), 0x9);
        }
        _mm256_stream_pd(row, v0);
        _mm256_stream_pd(row + 4, v1);
        _mm256_stream_pd(row + 8, v2);
      }
      continue;
    }
#endif
    for (size_t j = 0; j < p; ++j, row += 3)
      for (unsigned d = 0; d < 3; ++d) row[d] = d == k2 ? pk[j] : c[d];
  }
#if defined(__AVX__)
  _mm_sfence();
#endif
}

// Packed interior rule -> host arrays.  The encoded pieces come over in chunks; as soon as a chunk has landed a host
// thread writes its rows: row q of group g is pc[g] with pk[q] in slot k2[g].  Pure data movement: every value was
// computed on the GPU.  Weights need no decoding and go straight to their final array.
static void download_packed_rule(qg_rule &r, cudaStream_t s)
{
  const size_t ne = (size_t)r.n_entities, np = (size_t)r.n_points;
  const size_t p = (size_t)r.pack_p, ng = np / p;
  r.h_n_per = (int *)g_pinned.get((ne + 1) * sizeof(int), &r.cap[0]);
  r.h_pts = (double *)g_pinned.get((np * 3 + 1) * sizeof(double), &r.cap[1]);
  r.h_wts = (double *)g_pinned.get((np + 1) * sizeof(double), &r.cap[2]);
  size_t cap_pc = 0, cap_pk = 0, cap_k2 = 0;
  double *h_pc = (double *)g_pinned.get((ng * 3 + 1) * sizeof(double), &cap_pc);
  double *h_pk = (double *)g_pinned.get((np + 1) * sizeof(double), &cap_pk);
  unsigned char *h_k2 = (unsigned char *)g_pinned.get(ng + 1, &cap_k2);
  r.d_n_per.download(r.h_n_per, ne, s);
  // host threads: all hardware threads, shared fairly between the ranks of a one-process-per-GPU job (torchrun sets
  // LOCAL_WORLD_SIZE); QG_HOST_THREADS overrides
  unsigned nthreads = std::thread::hardware_concurrency();
  if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) nthreads /= (unsigned)std::max(1, std::atoi(e));
  if (const char *e = std::getenv("QG_HOST_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(e));
  nthreads = std::max(1u, std::min(nthreads, 64u));
  size_t chunks_per_thread = 16;  // measured on C3: 4 -> 346 ms, 16 -> 335 ms, 64 -> 385 ms for the whole call
  if (const char *e = std::getenv("QG_PACK_CHUNKS")) chunks_per_thread = (size_t)std::max(1, std::atoi(e));
  const size_t nchunks = std::max<size_t>(1, std::min<size_t>(ng, (size_t)nthreads * chunks_per_thread));
  struct EventList
  {
    std::vector<cudaEvent_t> v;
    ~EventList()
    {
      for (auto e : v)
        if (e) cudaEventDestroy(e);
    }
  } evl;
  evl.v.assign(nchunks, nullptr);
  std::vector<cudaEvent_t> &ev = evl.v;
  auto g_begin = [&](size_t c) { return ng * c / nchunks; };
  for (size_t c = 0; c < nchunks; ++c) {
    const size_t g0 = g_begin(c), g1 = g_begin(c + 1);
    QG_CUDA(cudaMemcpyAsync(h_pc + g0 * 3, r.d_pc.p + g0 * 3, (g1 - g0) * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    QG_CUDA(cudaMemcpyAsync(h_k2 + g0, r.d_k2.p + g0, g1 - g0, cudaMemcpyDeviceToHost, s));
    QG_CUDA(cudaMemcpyAsync(h_pk + g0 * p, r.d_pk.p + g0 * p, (g1 - g0) * p * sizeof(double), cudaMemcpyDeviceToHost, s));
    QG_CUDA(cudaMemcpyAsync(r.h_wts + g0 * p, r.d_wts.p + g0 * p, (g1 - g0) * p * sizeof(double), cudaMemcpyDeviceToHost, s));
    QG_CUDA(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
    QG_CUDA(cudaEventRecord(ev[c], s));
  }
  std::atomic<size_t> next{ 0 };
  std::atomic<int> failed{ 0 };
  double *pts = r.h_pts;
  const int device = r.device;
  auto worker = [&]() {
    cudaSetDevice(device);
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= nchunks) return;
      if (cudaEventSynchronize(ev[c]) != cudaSuccess) {
        failed = 1;
        return;
      }
      const size_t g0 = g_begin(c), g1 = g_begin(c + 1);
      expand_groups(h_pc, h_k2, h_pk, pts, p, g0, g1);
    }
  };
  const auto t_host0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  try {
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
  } catch (const std::system_error &) {
    // fewer host threads than asked for: the ones that started (and this one) share the chunks
  }
  worker();
  for (auto &t : pool) t.join();
  const cudaError_t e = cudaStreamSynchronize(s);
  if (std::getenv("QG_TRACE"))
    std::fprintf(stderr, "[qg trace] packed download: %zu groups, %zu chunks, %u host threads, %.1f ms\n", ng, nchunks, nthreads,
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
  g_pinned.put(h_pc, cap_pc);
  g_pinned.put(h_pk, cap_pk);
  g_pinned.put(h_k2, cap_k2);
  if (e != cudaSuccess || failed) throw EngineError(QG_ERR_CUDA, std::string("packed rule download: ") + cudaGetErrorString(e));
  // the encoded device copy is of no further use
  r.d_pc.release();
  r.d_pk.release();
  r.d_k2.release();
  r.d2h_bytes = (long long)(ne * sizeof(int) + ng * (3 * sizeof(double) + 1) + np * 2 * sizeof(double));
  r.downloaded = true;
}

static void download_rule(qg_rule &r, cudaStream_t s)
{
  if (r.downloaded) return;
  if (r.packed) {
    download_packed_rule(r, s);
    return;
  }
  const size_t ne = (size_t)r.n_entities, np = (size_t)r.n_points, pd = (size_t)r.point_dim;
  r.h_n_per = (int *)g_pinned.get((ne + 1) * sizeof(int), &r.cap[0]);
  r.h_pts = (double *)g_pinned.get((np * pd + 1) * sizeof(double), &r.cap[1]);
  r.h_wts = (double *)g_pinned.get((np + 1) * sizeof(double), &r.cap[2]);
  r.d_n_per.download(r.h_n_per, ne, s);
  r.d_pts.download(r.h_pts, np * pd, s);
  r.d_wts.download(r.h_wts, np, s);
  if (r.has_normals) {
    r.h_nrm = (double *)g_pinned.get((np * pd + 1) * sizeof(double), &r.cap[3]);
    r.d_nrm.download(r.h_nrm, np * pd, s);
  }
  QG_CUDA(cudaStreamSynchronize(s));
  r.d2h_bytes = (long long)(ne * sizeof(int) + np * (pd + 1 + (r.has_normals ? pd : 0)) * sizeof(double));
  r.downloaded = true;
}

// ------------------------------------------------------------------------------------------------
// Reparameterization (create_reparameterization, impl_reparam.cpp:46-113; reparam_general, impl_reparam_general.cpp:823-935)
// ------------------------------------------------------------------------------------------------
static void launch_rp(int emit, int dim, unsigned grid, cudaStream_t s, const RpLeafArgs &a)
{
  switch (a.f.kind) {
  case QG_FUNC_SPHERE: launch_rp_k0(emit, dim, grid, s, a); break;
  case QG_FUNC_BEZIER: (dim == 2 ? launch_rp_k2d2 : launch_rp_k2d3)(emit, dim, grid, s, a); break;
  case QG_FUNC_SCHOEN: launch_rp_k10(emit, dim, grid, s, a); break;
  case QG_FUNC_SCHWARZ_D: launch_rp_k14(emit, dim, grid, s, a); break;
  default: (dim == 2 ? launch_rp_gend2 : launch_rp_gend3)(emit, dim, grid, s, a); break;
  }
}

struct RadixTmp
{
  DevBuf<unsigned char> tmp;
  template<typename K, typename V> void sort_pairs(const K *kin, K *kout, const V *vin, V *vout, long long n, cudaStream_t s)
  {
    size_t bytes = 0;
    QG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, (int)sizeof(K) * 8, s));
    if (bytes > tmp.n) tmp.alloc(bytes * 2 + 1024);
    QG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, (int)sizeof(K) * 8, s));
    g_launches.fetch_add(4, std::memory_order_relaxed);
  }
};

// merge_coincident_points (reparam_mesh.cpp:607-626) on the device, stable sorts (see qg_kernels_reparam.cuh)
static void merge_mesh(const qg_domain &dm, qg_mesh &m, double tol)
{
  cudaStream_t s = dm.stream;
  const long long n = m.n_points;
  const int dim = m.dim;
  if (n < 2) return;
  if (n >= 2147483647LL) throw EngineError(QG_ERR_CAPACITY, "merge_points: more than 2^31 points");
  RadixTmp rt;
  DevBuf<unsigned long long> val((size_t)n), val2((size_t)n), key((size_t)n), key2((size_t)n);
  DevBuf<unsigned> skey((size_t)n), skey2((size_t)n);
  DevBuf<int> head((size_t)n + 1);
  DevBuf<long long> off((size_t)n + 1);
  head.zero(s);
  rp_merge_init_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, val.p);
  for (int d = dim - 1; d >= 0; --d) {
    rp_merge_coord_keys_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, val.p, m.d_pts.p, dim, d, key.p);
    rt.sort_pairs(key.p, key2.p, val.p, val2.p, n, s);
    rp_merge_seg_keys_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, val2.p, skey.p);
    rt.sort_pairs(skey.p, skey2.p, val2.p, val.p, n, s);
    rp_merge_heads_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, val.p, m.d_pts.p, dim, d, tol, head.p);
    dm.scanner.run(head.p, off.p, (size_t)n + 1, s);
    rp_merge_set_seg_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, head.p, off.p, val.p);
  }
  DevBuf<long long> first((size_t)n), map((size_t)n), newid((size_t)n + 1);
  DevBuf<int> uniq((size_t)n + 1);
  uniq.zero(s);
  rp_merge_first_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, head.p, val.p, first.p);
  rp_merge_map_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, val.p, first.p, map.p, uniq.p);
  const long long nu = scan_flags(dm, uniq, newid, n);
  DevBuf<double> np((size_t)nu * dim + 1);
  rp_merge_compact_kernel<<<QG_CNT(grid_for(n)), kBlock, 0, s>>>(n, dim, uniq.p, newid.p, m.d_pts.p, np.p);
  long long npc = 1;
  for (int i = 0; i < m.pd; ++i) npc *= m.order;
  if (m.n_cells) rp_merge_remap_kernel<<<QG_CNT(grid_for(m.n_cells * npc)), kBlock, 0, s>>>(m.n_cells * npc, map.p, newid.p, m.d_conn.p);
  const long long ne = m.n_wire_edges;
  if (ne) rp_merge_remap_kernel<<<QG_CNT(grid_for(ne * m.order)), kBlock, 0, s>>>(ne * m.order, map.p, newid.p, m.d_wires.p);
  m.d_pts = std::move(np);
  m.n_points = nu;
  // purge_duplicate_wirebasket_edges
  if (ne) {
    rp_edge_dir_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, m.order, m.d_wires.p);
    DevBuf<unsigned> perm((size_t)ne), perm2((size_t)ne);
    DevBuf<unsigned long long> ek((size_t)ne), ek2((size_t)ne);
    rp_edge_init_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, perm.p);
    for (int c = m.order - 1; c >= 0; --c) {
      rp_edge_keys_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, m.order, c, perm.p, m.d_wires.p, ek.p);
      rt.sort_pairs(ek.p, ek2.p, perm.p, perm2.p, ne, s);
      std::swap(perm.p, perm2.p);
    }
    DevBuf<int> eh((size_t)ne + 1);
    DevBuf<long long> eo((size_t)ne + 1);
    eh.zero(s);
    rp_edge_heads_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, m.order, perm.p, m.d_wires.p, eh.p);
    const long long nue = scan_flags(dm, eh, eo, ne);
    DevBuf<long long> nw((size_t)nue * m.order + 1);
    rp_edge_compact_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, m.order, perm.p, eh.p, eo.p, m.d_wires.p, nw.p);
    m.d_wires = std::move(nw);
    m.n_wire_edges = nue;
  }
  QG_CUDA(cudaGetLastError());
}

// mode: 0 volume, 1 zero level set, 2 + f: face f of every listed cell.  flags bit 0: merge coincident points with merge_tol;
// bit 1: reparameterize EVERY listed cell whatever its status and add no full cells (reparam_general on a box).
template<int N>
static qg_mesh *reparam_impl(const qg_domain &dm, std::vector<long long> cells, int order, int mode, int flags, double merge_tol)
{
  cudaStream_t s = dm.stream;
  std::unique_ptr<qg_mesh> mesh(new qg_mesh());
  mesh->device = dm.device;
  mesh->dim = N;
  mesh->pd = mode == 0 ? N : N - 1;
  mesh->order = order;
  const int pd = mesh->pd;
  long long npc = 1;
  for (int i = 0; i < pd; ++i) npc *= order;
  const bool every = (flags & 2) != 0;
  std::sort(cells.begin(), cells.end());  // impl_reparam.cpp:66-67
  const long long ne = (long long)cells.size();
  for (long long c : cells)
    if (c < 0 || c >= dm.ncells) throw EngineError(QG_ERR_INVALID, "cell id out of range");
  DevBuf<long long> d_cells((size_t)ne + 1);
  d_cells.upload(cells.data(), (size_t)ne, s);
  // cut cells -> jobs, full cells -> tensor-product cells (volume only)
  DevBuf<int> kind((size_t)ne + 1), flag((size_t)ne + 1), bad(1);
  DevBuf<long long> off((size_t)ne + 1);
  bad.zero(s);
  flag.zero(s);
  long long nj = 0, nfull = 0;
  DevBuf<long long> job_cell, full_cell;
  DevBuf<int> job_ent, job_type;
  if (every) {
    nj = ne;
    job_cell.alloc((size_t)nj + 1);
    if (ne) QG_CUDA(cudaMemcpyAsync(job_cell.p, d_cells.p, (size_t)ne * sizeof(long long), cudaMemcpyDeviceToDevice, s));
  } else {
    if (ne) {
      entity_kind_cells_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells.p, dm.ncells, dm.d_status.p, dm.d_rec_idx.p, kind.p, bad.p);
      flag_kind_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, kind.p, ENT_JOB, flag.p);
    }
    nj = scan_flags(dm, flag, off, ne);
    job_cell.alloc((size_t)nj + 1);
    job_ent.alloc((size_t)ne + 1);
    if (ne) gather_jobs_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells.p, flag.p, off.p, job_cell.p, job_ent.p);
    if (mode == 0) {
      if (ne) flag_kind_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, kind.p, ENT_GAUSS, flag.p);
      nfull = scan_flags(dm, flag, off, ne);
      full_cell.alloc((size_t)nfull + 1);
      if (ne) gather_jobs_kernel<<<QG_CNT(grid_for(ne)), kBlock, 0, s>>>(ne, d_cells.p, flag.p, off.p, full_cell.p, job_ent.p);
    }
  }
  job_type.alloc((size_t)nj + 1);
  if (nj) fill_types_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, mode == 0 ? JOB_VOLUME : (mode == 1 ? JOB_SURFACE : mode - 2), job_type.p);

  Pipeline pl(dm, 1, SINK_RAW);
  pl.run_items<N>(job_cell.p, job_type.p, nj);
  const long long nitems = pl.a.nitems;

  // tables
  const std::vector<double> h_nodes = reparam_nodes(order);
  std::vector<double> h_basis, h_ders;
  lagrange_mid(pd, order, h_basis, h_ders);
  DevBuf<double> nodes((size_t)order), basis(h_basis.size()), ders(h_ders.size());
  nodes.upload(h_nodes.data(), h_nodes.size(), s);
  basis.upload(h_basis.data(), h_basis.size(), s);
  ders.upload(h_ders.data(), h_ders.size(), s);

  // leaves: count, scan, emit
  RpLeafArgs la;
  std::memset(&la, 0, sizeof(la));
  la.f = dm.f;
  la.order = order;
  la.npc = (int)npc;
  la.mode = mode;
  la.nodes = nodes.p;
  la.nitems = nitems;
  la.items = pl.items.p;
  la.err = dm.d_err.p;
  DevBuf<int> leaf_cells((size_t)nitems + 1), leaf_pts((size_t)nitems + 1);
  DevBuf<long long> cell_off((size_t)nitems + 1), pt_off((size_t)nitems + 1);
  leaf_cells.zero(s);
  leaf_pts.zero(s);
  la.leaf_cells = leaf_cells.p;
  la.leaf_pts = leaf_pts.p;
  if (nitems) launch_rp(0, N, QG_CNT(grid_for(nitems, kRpBlock)), s, la);
  const long long ncut_cells = scan_flags(dm, leaf_cells, cell_off, nitems);
  const long long ncut_pts = scan_flags(dm, leaf_pts, pt_off, nitems);
  const long long ncells = ncut_cells + nfull, npts = ncut_pts + nfull * npc;
  mesh->d_pts.alloc((size_t)npts * N + 1);
  mesh->d_conn.alloc((size_t)ncells * npc + 1);
  mesh->d_conn.zero(s);  // entries no point was inserted for stay 0, as in the reference's resized vector
  DevBuf<int> cell_job((size_t)ncut_cells + 1);
  la.cell_off = cell_off.p;
  la.pt_off = pt_off.p;
  la.pts = mesh->d_pts.p;
  la.conn = mesh->d_conn.p;
  la.cell_job = cell_job.p;
  if (nitems) launch_rp(1, N, QG_CNT(grid_for(nitems, kRpBlock)), s, la);

  // orientation, wirebasket of the cut cells
  RpMeshArgs ma;
  std::memset(&ma, 0, sizeof(ma));
  ma.f = dm.f;
  ma.dim = N;
  ma.pd = pd;
  ma.order = order;
  ma.npc = (int)npc;
  ma.mode = mode;
  ma.ncells = ncut_cells;
  ma.pts = mesh->d_pts.p;
  ma.conn = mesh->d_conn.p;
  ma.cell_job = cell_job.p;
  ma.job_box = pl.job_box.p;
  ma.basis = basis.p;
  ma.ders = ders.p;
  ma.tol = 1.0e4 * kEps;
  if (ncut_cells) rp_orient_kernel<N><<<QG_CNT(grid_for(ncut_cells)), kBlock, 0, s>>>(ma);
  const int nedges = pd == 2 ? 4 : (pd == 3 ? 12 : 0);
  long long nwire_cut = 0;
  DevBuf<int> edge_flag((size_t)ncut_cells * nedges + 1);
  DevBuf<long long> edge_off((size_t)ncut_cells * nedges + 1);
  if (nedges && ncut_cells) {
    edge_flag.zero(s);
    ma.edge_flag = edge_flag.p;
    rp_wire_kernel<N, false><<<QG_CNT(grid_for(ncut_cells * nedges)), kBlock, 0, s>>>(ma);
    nwire_cut = scan_flags(dm, edge_flag, edge_off, ncut_cells * nedges);
  }
  const long long nwires = nwire_cut + nfull * nedges;
  mesh->d_wires.alloc((size_t)nwires * order + 1);
  if (nwire_cut) {
    ma.edge_off = edge_off.p;
    ma.wires = mesh->d_wires.p;
    rp_wire_kernel<N, true><<<QG_CNT(grid_for(ncut_cells * nedges)), kBlock, 0, s>>>(ma);
  }
  if (nfull)
    rp_full_cells_kernel<N><<<QG_CNT(grid_for(nfull)), kBlock, 0, s>>>(nfull, full_cell.p, dm.g, order, nodes.p, ncut_pts, ncut_cells, nwire_cut,
      mesh->d_pts.p, mesh->d_conn.p, mesh->d_wires.p);
  QG_CUDA(cudaGetLastError());
  int hbad = 0;
  bad.download(&hbad, 1, s);
  check_device_err(dm);
  if (hbad) throw EngineError(QG_ERR_INVALID, "cell id out of range");
  mesh->n_points = npts;
  mesh->n_cells = ncells;
  mesh->n_wire_edges = nwires;
  if (flags & 1) merge_mesh(dm, *mesh, merge_tol);
  QG_CUDA(cudaStreamSynchronize(s));
  return mesh.release();
}

}  // namespace qg

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define QG_TRY try {
#define QG_CATCH                                         \
  }                                                      \
  catch (const EngineError &e) { return fail(e.code, e.msg); } \
  catch (const std::exception &e) { return fail(QG_ERR_CUDA, e.what()); }

extern "C" {

const char *qg_last_error(void) { return g_last_error.c_str(); }
const char *qg_version(void) { return "qugar_b200 0.1.0 (sm_100a)"; }

int qg_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int qg_grid_create(int dim, const double *const *breaks, const int64_t *n_breaks, qg_grid **out)
{
  if (!out || !breaks || !n_breaks || (dim != 2 && dim != 3)) return fail(QG_ERR_INVALID, "qg_grid_create: bad arguments");
  std::unique_ptr<qg_grid> g(new qg_grid());
  g->dim = dim;
  for (int d = 0; d < 3; ++d) g->n[d] = 1;
  for (int d = 0; d < dim; ++d) {
    if (n_breaks[d] < 2 || !breaks[d]) return fail(QG_ERR_INVALID, "qg_grid_create: need at least two breaks per direction");
    g->breaks[d].assign(breaks[d], breaks[d] + n_breaks[d]);
    for (int64_t i = 1; i < n_breaks[d]; ++i)
      if (!(breaks[d][i - 1] < breaks[d][i])) return fail(QG_ERR_INVALID, "qg_grid_create: breaks must be strictly increasing");
    g->n[d] = n_breaks[d] - 1;
    if (g->n[d] > 2000000000LL) return fail(QG_ERR_INVALID, "qg_grid_create: too many cells per direction");
  }
  *out = g.release();
  return QG_OK;
}
void qg_grid_free(qg_grid *g) { delete g; }
int qg_grid_num_cells_dir(const qg_grid *g, int64_t n_cells[3])
{
  if (!g || !n_cells) return fail(QG_ERR_INVALID, "null argument");
  for (int d = 0; d < 3; ++d) n_cells[d] = d < g->dim ? g->n[d] : 1;
  return QG_OK;
}

// one non-composite function: fills fd and (Bezier) coefs; returns an error string or nullptr
static const char *make_leaf_desc(int kind, int dim, const double *params, long long n_params, int negative, FuncDesc &fd,
  std::vector<double> &coefs)
{
  std::memset(&fd, 0, sizeof(FuncDesc));
  fd.kind = kind;
  fd.dim = dim;
  fd.negative = negative ? 1 : 0;
  if (kind == QG_BEZIER) {
    // params: order[dim] (as doubles) followed by prod(order) coefficients, direction 0 fastest
    if (n_params < dim) return "Bezier needs order[dim] + coefficients";
    long long nc = 1;
    for (int d = 0; d < 3; ++d) fd.order[d] = 1;
    for (int d = 0; d < dim; ++d) {
      fd.order[d] = (int)params[d];
      if (fd.order[d] < 1 || fd.order[d] > kMaxBezierOrder || (double)fd.order[d] != params[d])
        return "Bezier order must be an integer in 1..8 per direction";
      nc *= fd.order[d];
    }
    if (n_params != dim + nc) return "wrong number of Bezier coefficients";
    coefs.assign(params + dim, params + n_params);
    return nullptr;
  }
  int need = 0;
  switch (kind) {
  case QG_SPHERE: need = 1 + dim; break;
  case QG_PLANE: need = 2 * dim; break;
  case QG_CONSTANT: need = 1; break;
  case QG_CYLINDER: need = 7; break;
  case QG_ELLIPSOID: need = 2 * dim + dim * dim; break;
  case QG_ANNULUS: need = 4; break;
  case QG_TORUS: need = 8; break;
  case QG_DIM_LINEAR: need = 1 << dim; break;
  case QG_SQUARE: need = 0; break;
  case QG_SCHOEN:
  case QG_SCHOEN_IWP:
  case QG_SCHOEN_FRD:
  case QG_FISCHER_KOCH_S:
  case QG_SCHWARZ_DIAMOND:
  case QG_SCHWARZ_PRIMITIVE: need = 3; break;
  default: return "unknown function kind";
  }
  if (n_params != need) return "wrong number of parameters";
  if (kind == QG_SPHERE && !(params[0] > 0.0)) return "radius must be positive";
  if ((kind == QG_CYLINDER || kind == QG_TORUS) && dim != 3) return "3-D only";
  if (kind == QG_ANNULUS && dim != 2) return "2-D only";
  if (kind == QG_CYLINDER && !(params[0] > 0.0)) return "radius must be positive";
  if (kind == QG_ANNULUS && !(0.0 < params[0] && params[0] < params[1])) return "need 0 < inner < outer radius";
  if (kind == QG_TORUS && !(0.0 < params[1] && params[1] < params[0])) return "need 0 < minor < major radius";
  if (kind == QG_ELLIPSOID)
    for (int d = 0; d < dim; ++d)
      if (!(params[d] > 0.0)) return "semi-axes must be positive";
  for (int i = 0; i < need; ++i) fd.p[i] = params[i];
  fd.same_args = tpms_same_arg_mask(kind, dim, fd.p);
  return nullptr;
}

int qg_func_create(int kind, int dim, const double *params, int n_params, int negative, qg_func **out)
{
  if (!out || (!params && n_params != 0) || n_params < 0 || (dim != 2 && dim != 3)) return fail(QG_ERR_INVALID, "qg_func_create: bad arguments");
  std::unique_ptr<qg_func> f(new qg_func());
  std::memset(&f->comp, 0, sizeof(CompositeDesc));
  if (kind == QG_COMPOSITE) {
    // params: the program form (qugar_b200/cpp.py; decoded and checked by composite_parse, csrc/qg_funcs.cuh)
    std::memset(&f->f, 0, sizeof(FuncDesc));
    f->f.kind = QG_COMPOSITE;
    f->f.dim = dim;
    f->f.negative = negative ? 1 : 0;
    bool unsupported = false;
    qg_func *fp = f.get();
    const char *err = composite_parse(params, n_params, f->comp, [&](FuncDesc &fd, int l, int lkind, int lneg, const double *p, int np) {
      const char *e = make_leaf_desc(lkind, dim, p, np, lneg, fd, fp->leaf_coefs[l]);
      if (e && (lkind > QG_SQUARE || lkind < 0)) unsupported = true;
      return e;
    });
    if (err) return fail(unsupported ? QG_ERR_UNSUPPORTED : QG_ERR_INVALID, std::string("qg_func_create: composite: ") + err);
    *out = f.release();
    return QG_OK;
  }
  const char *err = make_leaf_desc(kind, dim, params, n_params, negative, f->f, f->coefs);
  if (err) return fail(std::string(err) == "unknown function kind" ? QG_ERR_UNSUPPORTED : QG_ERR_INVALID, std::string("qg_func_create: ") + err);
  *out = f.release();
  return QG_OK;
}
void qg_func_free(qg_func *f) { delete f; }

int qg_func_eval(const qg_func *f, const double *points, int64_t n, double *values, double *grad)
{
  if (!f || !points || !values || n < 0) return fail(QG_ERR_INVALID, "qg_func_eval: bad arguments");
  if (qg_device_count() == 0) return fail(QG_ERR_NO_DEVICE, "no CUDA device");
  QG_TRY
  const int dim = f->f.dim;
  int cur_dev = 0;
  QG_CUDA(cudaGetDevice(&cur_dev));
  cudaStream_t es = device_stream(cur_dev);
  g_alloc_stream = es;
  g_alloc_device = cur_dev;
  DevBuf<double> dp((size_t)n * dim + 1), dv((size_t)n + 1), dg;
  if (grad) dg.alloc((size_t)n * dim + 1);
  dp.upload(points, (size_t)n * dim, es);
  FuncDeviceData fdata;
  const FuncDesc fd = fdata.upload(f, es);
  if (n) {
    if (dim == 2)
      func_eval_kernel<2><<<QG_CNT(grid_for(n)), kBlock, 0, es>>>(fd, dp.p, n, dv.p, grad ? dg.p : nullptr);
    else
      func_eval_kernel<3><<<QG_CNT(grid_for(n)), kBlock, 0, es>>>(fd, dp.p, n, dv.p, grad ? dg.p : nullptr);
  }
  QG_CUDA(cudaGetLastError());
  dv.download(values, (size_t)n, es);
  if (grad) dg.download(grad, (size_t)n * dim, es);
  QG_CUDA(cudaStreamSynchronize(es));
  return QG_OK;
  QG_CATCH
}

int qg_domain_create(const qg_func *f, const qg_grid *g, int device, qg_domain **out)
{
  return qg_domain_create_slab(f, g, device, 0, -1, out);
}

int qg_domain_create_slab(const qg_func *f, const qg_grid *g, int device, int64_t k_begin, int64_t k_end, qg_domain **out)
{
  if (!f || !g || !out) return fail(QG_ERR_INVALID, "qg_domain_create: null argument");
  if (k_end < 0) k_end = g->n[g->dim - 1];
  if (k_begin < 0 || k_begin > k_end || k_end > g->n[g->dim - 1]) return fail(QG_ERR_INVALID, "qg_domain_create_slab: bad slab range");
  if (f->f.dim != g->dim) return fail(QG_ERR_INVALID, "qg_domain_create: function and grid dimensions differ");
  const int ndev = qg_device_count();
  if (ndev == 0) return fail(QG_ERR_NO_DEVICE, "qg_domain_create: no CUDA device (this engine has no CPU path)");
  if (device < 0 || device >= ndev) return fail(QG_ERR_INVALID, "qg_domain_create: bad device ordinal");
  // TPMS arguments must stay inside the range of the in-tree sin/cos reduction
  if (f->f.kind >= QG_SCHOEN && f->f.kind <= QG_SCHWARZ_PRIMITIVE) {
    for (int d = 0; d < g->dim; ++d) {
      const double m = std::fabs(f->f.p[d]) * 4.0 * kPi;
      const double xm = std::max(std::fabs(g->breaks[d].front()), std::fabs(g->breaks[d].back())) + 1.0;
      if (!(m * xm <= kSinCosMaxArg)) return fail(QG_ERR_UNSUPPORTED, "qg_domain_create: TPMS argument exceeds 1e5");
    }
  }
  QG_TRY
  QG_CUDA(cudaSetDevice(device));
  std::unique_ptr<qg_domain> dm(new qg_domain());
  dm->device = device;
  dm->dim = g->dim;
  dm->f = f->f;
  dm->grid = *g;
  dm->ncells = 1;
  for (int d = 0; d < g->dim; ++d) dm->ncells *= g->n[d];
  if (dm->ncells > 2000000000LL) return fail(QG_ERR_UNSUPPORTED, "qg_domain_create: more than 2e9 cells");
  dm->stream = device_stream(device);
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) dm->num_sms = sms;
  }
  g_alloc_stream = dm->stream;
  g_alloc_device = device;
  std::memset(&dm->g, 0, sizeof(GridDesc));
  dm->g.dim = g->dim;
  for (int d = 0; d < 3; ++d) dm->g.n[d] = d < g->dim ? g->n[d] : 1;
  for (int d = 0; d < g->dim; ++d) {
    dm->d_breaks[d].alloc(g->breaks[d].size());
    dm->d_breaks[d].upload(g->breaks[d].data(), g->breaks[d].size(), dm->stream);
    dm->g.breaks[d] = dm->d_breaks[d].p;
  }
  dm->d_gauss.alloc(sizeof(h_gauss_table) / sizeof(double));
  dm->d_gauss.upload(h_gauss_table, sizeof(h_gauss_table) / sizeof(double), dm->stream);
  dm->d_err.alloc(1);
  dm->d_err.zero(dm->stream);
  dm->f = dm->fdata.upload(f, dm->stream);
  dm->k0 = (int)k_begin;
  dm->k1 = (int)k_end;
  if (g->dim == 2)
    classify_domain<2>(*dm);
  else
    classify_domain<3>(*dm);
  *out = dm.release();
  return QG_OK;
  QG_CATCH
}
void qg_domain_free(qg_domain *d)
{
  if (d) cudaSetDevice(d->device);
  delete d;
}

int qg_domain_counts(const qg_domain *d, int64_t counts[4])
{
  if (!d || !counts) return fail(QG_ERR_INVALID, "null argument");
  counts[0] = d->n_full;
  counts[1] = d->n_empty;
  counts[2] = d->n_cut;
  counts[3] = d->n_rec;
  return QG_OK;
}
int qg_domain_cells(const qg_domain *d, int status, int64_t *ids)
{
  if (!d || !ids || status < 0 || status > 2) return fail(QG_ERR_INVALID, "qg_domain_cells: bad arguments");
  if (d->mirrored.load(std::memory_order_acquire)) {
    size_t k = 0;
    for (long long i = 0; i < d->ncells; ++i)
      if (d->status[(size_t)i] == status) ids[k++] = i;
    return QG_OK;
  }
  // no host mirror yet: compact on the device and bring back only the ids (8 B per listed cell instead of the
  // whole status array plus a host scan)
  QG_TRY
  std::lock_guard<std::mutex> lock(d->mtx);
  QG_CUDA(cudaSetDevice(d->device));
  g_alloc_stream = d->stream;
  g_alloc_device = d->device;
  cudaStream_t s = d->stream;
  DevBuf<long long> out;
  const long long n = compact_status_ids(*d, (unsigned char)status, out);
  out.download((long long *)ids, (size_t)n, s);
  QG_CUDA(cudaStreamSynchronize(s));
  return QG_OK;
  QG_CATCH
}
int qg_domain_cell_status(const qg_domain *d, uint8_t *status)
{
  if (!d || !status) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  d->mirror();
  std::memcpy(status, d->status.data(), d->status.size());
  return QG_OK;
  QG_CATCH
}
int qg_domain_facet_records(const qg_domain *d, int64_t *cells, uint8_t *statuses)
{
  if (!d || !cells || !statuses) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  d->mirror();
  for (size_t i = 0; i < d->rec_cells.size(); ++i) cells[i] = d->rec_cells[i];
  std::memcpy(statuses, d->rec_stat.data(), d->rec_stat.size());
  return QG_OK;
  QG_CATCH
}

static bool facet_query_match(int query, int s)
{
  switch (query) {
  case QG_FACETS_EMPTY: return s == QG_FACET_EMPTY || s == QG_FACET_FULL_EXT_BDRY || s == QG_FACET_EXT_BDRY;
  case QG_FACETS_FULL: return s == QG_FACET_FULL;
  case QG_FACETS_CUT:
    return s == QG_FACET_CUT || s == QG_FACET_CUT_UNF_BDRY || s == QG_FACET_CUT_EXT_BDRY || s == QG_FACET_CUT_UNF_BDRY_EXT_BDRY;
  case QG_FACETS_FULL_UNF_BDRY: return s == QG_FACET_FULL_UNF_BDRY;
  case QG_FACETS_UNF_BDRY:
    return s == QG_FACET_CUT_UNF_BDRY || s == QG_FACET_CUT_UNF_BDRY_EXT_BDRY || s == QG_FACET_FULL_UNF_BDRY
           || s == QG_FACET_UNF_BDRY || s == QG_FACET_UNF_BDRY_EXT_BDRY;
  default: return false;
  }
}

int qg_domain_facets(const qg_domain *d, int query, int64_t *cells, int32_t *facets, int64_t *n)
{
  if (!d || !n || query < 0 || query > 4) return fail(QG_ERR_INVALID, "qg_domain_facets: bad arguments");
  QG_TRY
  d->mirror();
  QG_CATCH
  const int nf = 2 * d->dim;
  int64_t k = 0;
  size_t r = 0;
  const size_t nrec = d->rec_cells.size();
  // merge of the sorted record list with the plain cell statuses, ascending cell id
  for (long long c = 0; c < d->ncells; ++c) {
    const bool has_rec = r < nrec && d->rec_cells[r] == c;
    if (has_rec) {
      for (int lf = 0; lf < nf; ++lf)
        if (facet_query_match(query, d->rec_stat[r * nf + lf])) {
          if (cells) {
            cells[k] = c;
            facets[k] = lf;
          }
          ++k;
        }
      ++r;
    } else {
      const unsigned char st = d->status[(size_t)c];
      if ((query == QG_FACETS_EMPTY && st == ST_EMPTY) || (query == QG_FACETS_FULL && st == ST_FULL)) {
        for (int lf = 0; lf < nf; ++lf) {
          if (cells) {
            cells[k] = c;
            facets[k] = lf;
          }
          ++k;
        }
      }
    }
  }
  *n = k;
  return QG_OK;
}

int qg_domain_upload_cells(const qg_domain *d, const int64_t *cells, int64_t n_cells, qg_dcells **out)
{
  if (!d || !out || n_cells < 0 || (n_cells && !cells)) return fail(QG_ERR_INVALID, "qg_domain_upload_cells: bad arguments");
  QG_TRY
  QG_CUDA(cudaSetDevice(d->device));
  g_alloc_stream = d->stream;
  g_alloc_device = d->device;
  std::unique_ptr<qg_dcells> dc(new qg_dcells());
  dc->n = n_cells;
  dc->d_cells.alloc((size_t)n_cells + 1);
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  dc->d_cells.upload((const long long *)cells, (size_t)n_cells, d->stream);
  QG_CUDA(cudaStreamSynchronize(d->stream));
  *out = dc.release();
  return QG_OK;
  QG_CATCH
}
void qg_dcells_free(qg_dcells *c) { delete c; }

int qg_domain_cut_cells_device(const qg_domain *d, qg_dcells **out)
{
  if (!d || !out) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  std::lock_guard<std::mutex> lock(d->mtx);
  QG_CUDA(cudaSetDevice(d->device));
  g_alloc_stream = d->stream;
  g_alloc_device = d->device;
  cudaStream_t s = d->stream;
  std::unique_ptr<qg_dcells> dc(new qg_dcells());
  dc->n = compact_status_ids(*d, ST_CUT, dc->d_cells);
  QG_CUDA(cudaStreamSynchronize(s));
  *out = dc.release();
  return QG_OK;
  QG_CATCH
}

int qg_dcells_count(const qg_dcells *c, int64_t *n)
{
  if (!c || !n) return fail(QG_ERR_INVALID, "null argument");
  *n = c->n;
  return QG_OK;
}

struct qg_timer
{
  int device;
  cudaEvent_t a, b;
};
int qg_timer_create(int device, qg_timer **out)
{
  if (!out) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  QG_CUDA(cudaSetDevice(device));
  std::unique_ptr<qg_timer> t(new qg_timer());
  t->device = device;
  QG_CUDA(cudaEventCreate(&t->a));
  QG_CUDA(cudaEventCreate(&t->b));
  *out = t.release();
  return QG_OK;
  QG_CATCH
}
int qg_timer_start(qg_timer *t)
{
  if (!t) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  QG_CUDA(cudaSetDevice(t->device));
  cudaStream_t s = device_stream(t->device);
  QG_CUDA(cudaStreamSynchronize(s));
  QG_CUDA(cudaEventRecord(t->a, s));
  return QG_OK;
  QG_CATCH
}
int qg_timer_stop(qg_timer *t, double *ms)
{
  if (!t || !ms) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  QG_CUDA(cudaSetDevice(t->device));
  cudaStream_t s = device_stream(t->device);
  QG_CUDA(cudaEventRecord(t->b, s));
  QG_CUDA(cudaEventSynchronize(t->b));
  float f = 0;
  QG_CUDA(cudaEventElapsedTime(&f, t->a, t->b));
  *ms = f;
  return QG_OK;
  QG_CATCH
}
void qg_timer_free(qg_timer *t)
{
  if (!t) return;
  cudaEventDestroy(t->a);
  cudaEventDestroy(t->b);
  delete t;
}

static int quad_device(const qg_domain *d, const qg_dcells *c, int n_pts_dir, bool surf, qg_rule **out, bool for_host = false)
{
  if (!d || !c || !out) return fail(QG_ERR_INVALID, "null argument");
  if (n_pts_dir < 1 || n_pts_dir > kGaussMax) return fail(QG_ERR_INVALID, "n_pts_dir must be in 1..100");
  QG_TRY
  std::lock_guard<std::mutex> lock(d->mtx);
  QG_CUDA(cudaSetDevice(d->device));
  g_alloc_stream = d->stream;
  g_alloc_device = d->device;
  *out = d->dim == 2 ? quad_cells_impl<2>(*d, c->d_cells.p, c->n, n_pts_dir, surf, for_host)
                     : quad_cells_impl<3>(*d, c->d_cells.p, c->n, n_pts_dir, surf, for_host);
  return QG_OK;
  QG_CATCH
}
int qg_quad_cells_device(const qg_domain *d, const qg_dcells *c, int n_pts_dir, qg_rule **out)
{
  return quad_device(d, c, n_pts_dir, false, out);
}
int qg_quad_unf_bdry_device(const qg_domain *d, const qg_dcells *c, int n_pts_dir, qg_rule **out)
{
  return quad_device(d, c, n_pts_dir, true, out);
}

int qg_rule_download(qg_rule *r)
{
  if (!r) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  QG_CUDA(cudaSetDevice(r->device));
  download_rule(*r, device_stream(r->device));
  return QG_OK;
  QG_CATCH
}

// The transport encoding cuts the PCIe bytes of an interior rule to 0.6 but adds host-memory traffic (staging read + row
// writes).  It wins while the GPUs' PCIe links are the bottleneck (1 rank: 435 -> 264 ms for the call; 2 and 4 ranks on one
// host with the blend-based expansion: 591 -> 570 ms and 1439 -> 1389 ms per e2e step) and lost with 8 ranks sharing a
// 32-vCPU host's memory bandwidth (2.19 -> 2.53 s per step, measured with the round-1 expansion), so it is on for up to four
// ranks per host.  QG_NO_PACK=1 / QG_FORCE_PACK=1 override.
static bool use_transport_encoding()
{
  if (std::getenv("QG_NO_PACK")) return false;
  if (std::getenv("QG_FORCE_PACK")) return true;
  const char *lws = std::getenv("LOCAL_WORLD_SIZE");
  return !lws || std::atoi(lws) <= 4;
}

int qg_quad_cells(const qg_domain *d, const int64_t *cells, int64_t n_cells, int n_pts_dir, qg_rule **out)
{
  qg_dcells *dc = nullptr;
  int rc = qg_domain_upload_cells(d, cells, n_cells, &dc);
  if (rc) return rc;
  rc = quad_device(d, dc, n_pts_dir, false, out, use_transport_encoding());
  qg_dcells_free(dc);
  if (rc) return rc;
  rc = qg_rule_download(*out);
  if (rc) {
    qg_rule_free(*out);
    *out = nullptr;
  }
  return rc;
}

int qg_quad_unf_bdry(const qg_domain *d, const int64_t *cells, int64_t n_cells, int n_pts_dir,
  int include_facet_unf_bdry, int exclude_ext_bdry, qg_rule **out)
{
  if (!d || !out || (n_cells && !cells)) return fail(QG_ERR_INVALID, "null argument");
  if (n_pts_dir < 1 || n_pts_dir > kGaussMax) return fail(QG_ERR_INVALID, "n_pts_dir must be in 1..100");
  qg_dcells *dc = nullptr;
  int rc = qg_domain_upload_cells(d, cells, n_cells, &dc);
  if (rc) return rc;
  if (!include_facet_unf_bdry) {
    rc = qg_quad_unf_bdry_device(d, dc, n_pts_dir, out);
  } else {
    rc = [&]() -> int {
      QG_TRY
      std::lock_guard<std::mutex> lock(d->mtx);
      QG_CUDA(cudaSetDevice(d->device));
      g_alloc_stream = d->stream;
      g_alloc_device = d->device;
      *out = d->dim == 2 ? quad_unf_bdry_with_facets_impl<2>(*d, dc->d_cells.p, dc->n, n_pts_dir, exclude_ext_bdry)
                         : quad_unf_bdry_with_facets_impl<3>(*d, dc->d_cells.p, dc->n, n_pts_dir, exclude_ext_bdry);
      return QG_OK;
      QG_CATCH
    }();
  }
  qg_dcells_free(dc);
  if (rc) return rc;
  rc = qg_rule_download(*out);
  if (rc) {
    qg_rule_free(*out);
    *out = nullptr;
  }
  return rc;
}

int qg_quad_facets(const qg_domain *d, const int64_t *cells, const int32_t *facets, int64_t n, int n_pts_dir, int exterior,
  qg_rule **out)
{
  if (!d || !out || (n && (!cells || !facets))) return fail(QG_ERR_INVALID, "null argument");
  if (n_pts_dir < 1 || n_pts_dir > kGaussMax) return fail(QG_ERR_INVALID, "n_pts_dir must be in 1..100");
  qg_dcells *dc = nullptr;
  int rc = qg_domain_upload_cells(d, cells, n, &dc);
  if (rc) return rc;
  rc = [&]() -> int {
    QG_TRY
    std::lock_guard<std::mutex> lock(d->mtx);
    QG_CUDA(cudaSetDevice(d->device));
    g_alloc_stream = d->stream;
    g_alloc_device = d->device;
    DevBuf<int> dfac((size_t)n + 1);
    dfac.upload(facets, (size_t)n, d->stream);
    *out = d->dim == 2
             ? quad_facets_impl<2>(*d, dc->d_cells.p, dfac.p, n, n_pts_dir, exterior ? FMODE_EXTERIOR : FMODE_INTERIOR, 0, nullptr)
             : quad_facets_impl<3>(*d, dc->d_cells.p, dfac.p, n, n_pts_dir, exterior ? FMODE_EXTERIOR : FMODE_INTERIOR, 0, nullptr);
    return QG_OK;
    QG_CATCH
  }();
  qg_dcells_free(dc);
  if (rc) return rc;
  rc = qg_rule_download(*out);
  if (rc) {
    qg_rule_free(*out);
    *out = nullptr;
  }
  return rc;
}

int qg_rule_view(const qg_rule *r, int64_t *n_entities, int64_t *n_points, int *point_dim,
  const int32_t **n_pts_per_entity, const double **points, const double **weights, const double **normals)
{
  if (!r) return fail(QG_ERR_INVALID, "null argument");
  if (!r->downloaded) return fail(QG_ERR_INVALID, "rule is device-resident; call qg_rule_download first");
  if (n_entities) *n_entities = r->n_entities;
  if (n_points) *n_points = r->n_points;
  if (point_dim) *point_dim = r->point_dim;
  if (n_pts_per_entity) *n_pts_per_entity = r->h_n_per;
  if (points) *points = r->h_pts;
  if (weights) *weights = r->h_wts;
  if (normals) *normals = r->has_normals ? r->h_nrm : nullptr;
  return QG_OK;
}
int qg_rule_device_view(const qg_rule *r, int64_t *n_entities, int64_t *n_points, const int32_t **d_n_pts_per_entity,
  const double **d_points, const double **d_weights, const double **d_normals)
{
  if (r && r->packed && d_points) return fail(QG_ERR_UNSUPPORTED, "rule was produced for the host API (transport encoding); use the *_device entry points for device-resident rules");
  if (!r) return fail(QG_ERR_INVALID, "null argument");
  if (n_entities) *n_entities = r->n_entities;
  if (n_points) *n_points = r->n_points;
  if (d_n_pts_per_entity) *d_n_pts_per_entity = r->d_n_per.p;
  if (d_points) *d_points = r->d_pts.p;
  if (d_weights) *d_weights = r->d_wts.p;
  if (d_normals) *d_normals = r->has_normals ? r->d_nrm.p : nullptr;
  return QG_OK;
}
long long qg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
long long qg_rule_d2h_bytes(const qg_rule *r) { return r ? r->d2h_bytes : 0; }
int qg_memcpy_device(void *dst, const void *src, int64_t bytes, int device)
{
  if (bytes < 0 || (bytes && (!dst || !src))) return fail(QG_ERR_INVALID, "qg_memcpy_device: bad arguments");
  QG_TRY
  // on the engine stream: a copy on the legacy default stream is not ordered against the (non-blocking) engine
  // stream, so a later engine call could overwrite a freed block while the copy still reads it
  QG_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  QG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, st));
  QG_CUDA(cudaStreamSynchronize(st));
  return QG_OK;
  QG_CATCH
}
extern "C++" {
template<typename T>
static void pack_coeffs_typed(qg_coeffs &c, cudaStream_t s, const void *old_coeffs, long long n_rows, long long n_old_cols, int n_extra,
  const long long *d_custom, long long n_custom, const long long *d_empty, long long n_empty, const CoeffRules &cr,
  const long long *val_off)
{
  T *dst = reinterpret_cast<T *>(c.d.p);
  const long long n_cols = c.n_cols;
  QG_CUDA(cudaMemsetAsync(dst, 0, (size_t)c.n_rows * n_cols * sizeof(T), s));   // np.zeros
  if (n_rows && n_old_cols)                                                       // new_coeffs[:rows, :cols] = old_coeffs
    QG_CUDA(cudaMemcpy2DAsync(dst, (size_t)n_cols * sizeof(T), old_coeffs, (size_t)n_old_cols * sizeof(T), (size_t)n_old_cols * sizeof(T),
                              (size_t)n_rows, cudaMemcpyHostToDevice, s));
  if (n_rows) coeff_offsets_default_kernel<T><<<QG_CNT(grid_for(n_rows)), kBlock, 0, s>>>(dst, n_rows, n_cols, n_extra);
  if (n_custom) {
    coeff_offsets_custom_kernel<T><<<QG_CNT(grid_for(n_custom)), kBlock, 0, s>>>(dst, n_rows, n_cols, n_extra, d_custom, n_custom, val_off);
    coeff_pack_kernel<T><<<QG_CNT((unsigned)n_custom), 256, 0, s>>>(cr, dst, n_rows * n_cols, val_off, n_custom);
  }
  if (n_empty) coeff_offsets_empty_kernel<T><<<QG_CNT(grid_for(n_empty)), kBlock, 0, s>>>(dst, n_cols, n_extra, d_empty, n_empty);
}
}  // extern "C++"

int qg_pack_custom_coeffs(int device, int elem_size, const void *old_coeffs, int64_t n_rows, int64_t n_old_cols,
  const int64_t *custom_ids, int64_t n_custom, const int64_t *empty_ids, int64_t n_empty, const qg_rule *const *rules, int n_rules,
  int to_host, qg_coeffs **out)
{
  if (!out || (elem_size != 4 && elem_size != 8) || n_rows < 0 || n_old_cols < 0 || n_custom < 0 || n_empty < 0 || n_rules < 1
      || n_rules > kMaxCoeffRules || !rules || (n_custom && !custom_ids) || (n_empty && !empty_ids)
      || (n_rows * n_old_cols && !old_coeffs))
    return fail(QG_ERR_INVALID, "qg_pack_custom_coeffs: bad arguments");
  for (int64_t i = 0; i < n_custom; ++i)
    if (custom_ids[i] < 0 || custom_ids[i] >= n_rows) return fail(QG_ERR_INVALID, "qg_pack_custom_coeffs: custom entity id out of range");
  for (int64_t i = 0; i < n_empty; ++i)
    if (empty_ids[i] < 0 || empty_ids[i] >= n_rows) return fail(QG_ERR_INVALID, "qg_pack_custom_coeffs: empty entity id out of range");
  for (int q = 0; q < n_rules; ++q) {
    if (!rules[q] || rules[q]->packed || rules[q]->device != device || rules[q]->n_entities != n_custom)
      return fail(QG_ERR_INVALID, "qg_pack_custom_coeffs: every rule must be a device-resident rule of this device with one entity per custom id");
  }
  QG_TRY
  QG_CUDA(cudaSetDevice(device));
  cudaStream_t s = device_stream(device);
  g_alloc_stream = s;
  g_alloc_device = device;
  std::unique_ptr<qg_coeffs> c(new qg_coeffs());
  c->device = device;
  c->elem_size = elem_size;
  const int n_extra = (int)((sizeof(int64_t) + (size_t)elem_size - 1) / (size_t)elem_size);   // _get_n_extra_cols
  const long long n_cols = n_old_cols + n_extra;
  Scanner sc;
  CoeffRules cr;
  std::memset(&cr, 0, sizeof(cr));
  cr.n_rules = n_rules;
  std::vector<DevBuf<long long>> ent_off((size_t)n_rules);
  for (int q = 0; q < n_rules; ++q) {
    const qg_rule &r = *rules[q];
    cr.pdim[q] = r.point_dim;
    cr.has_nrm[q] = r.has_normals ? 1 : 0;
    cr.n_per[q] = r.d_n_per.p;   // allocated with one trailing zero slot
    cr.pts[q] = r.d_pts.p;
    cr.wts[q] = r.d_wts.p;
    cr.nrm[q] = r.d_nrm.p;
    ent_off[(size_t)q].alloc((size_t)n_custom + 1);
    sc.run(r.d_n_per.p, ent_off[(size_t)q].p, (size_t)n_custom + 1, s);
    cr.ent_off[q] = ent_off[(size_t)q].p;
  }
  DevBuf<int> n_vals((size_t)n_custom + 1);
  n_vals.zero(s);
  DevBuf<long long> val_off((size_t)n_custom + 1), d_custom((size_t)n_custom + 1), d_empty((size_t)n_empty + 1);
  d_custom.upload((const long long *)custom_ids, (size_t)n_custom, s);
  d_empty.upload((const long long *)empty_ids, (size_t)n_empty, s);
  if (n_custom) coeff_counts_kernel<<<QG_CNT(grid_for(n_custom)), kBlock, 0, s>>>(cr, n_custom, n_vals.p);
  sc.run(n_vals.p, val_off.p, (size_t)n_custom + 1, s);
  long long n_tot = 0;
  QG_CUDA(cudaMemcpyAsync(&n_tot, val_off.p + n_custom, sizeof(n_tot), cudaMemcpyDeviceToHost, s));
  QG_CUDA(cudaStreamSynchronize(s));
  const long long n_extra_rows = n_cols ? (n_tot + n_cols - 1) / n_cols : 0;      // np.ceil(n_tot_vals / n_cols_real)
  c->n_rows = n_rows + n_extra_rows;
  c->n_cols = n_cols;
  c->d.alloc((size_t)c->n_rows * (size_t)n_cols * (size_t)elem_size + 16);
  if (elem_size == 8)
    pack_coeffs_typed<double>(*c, s, old_coeffs, n_rows, n_old_cols, n_extra, d_custom.p, n_custom, d_empty.p, n_empty, cr, val_off.p);
  else
    pack_coeffs_typed<float>(*c, s, old_coeffs, n_rows, n_old_cols, n_extra, d_custom.p, n_custom, d_empty.p, n_empty, cr, val_off.p);
  QG_CUDA(cudaGetLastError());
  c->h2d_bytes = n_rows * n_old_cols * elem_size + (n_custom + n_empty) * 8;
  if (to_host) {
    const size_t bytes = (size_t)c->n_rows * (size_t)n_cols * (size_t)elem_size;
    c->h = g_pinned.get(bytes + 16, &c->cap);
    if (bytes) QG_CUDA(cudaMemcpyAsync(c->h, c->d.p, bytes, cudaMemcpyDeviceToHost, s));
    c->d2h_bytes = (long long)bytes;
  }
  QG_CUDA(cudaStreamSynchronize(s));
  *out = c.release();
  return QG_OK;
  QG_CATCH
}
int qg_coeffs_view(const qg_coeffs *c, int64_t *n_rows, int64_t *n_cols, int *elem_size, const void **host, const void **dev,
  int64_t *d2h_bytes)
{
  if (!c) return fail(QG_ERR_INVALID, "null argument");
  if (n_rows) *n_rows = c->n_rows;
  if (n_cols) *n_cols = c->n_cols;
  if (elem_size) *elem_size = c->elem_size;
  if (host) *host = c->h;
  if (dev) *dev = c->d.p;
  if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
  return QG_OK;
}
void qg_coeffs_free(qg_coeffs *c)
{
  if (c) cudaSetDevice(c->device);
  delete c;
}

int qg_bezier_rescale(const qg_func *f, const double *boxes, int64_t n_boxes, int device, double *coefs_out, int8_t *sign_out)
{
  if (!f || n_boxes < 0 || (n_boxes && (!boxes || !coefs_out || !sign_out))) return fail(QG_ERR_INVALID, "qg_bezier_rescale: bad arguments");
  if (f->f.kind != QG_BEZIER) return fail(QG_ERR_INVALID, "qg_bezier_rescale: not a Bezier function");
  const int dim = f->f.dim;
  long long nc = 1;
  for (int d = 0; d < dim; ++d) nc *= f->f.order[d];
  if (nc > kMaxBezierCoefs) return fail(QG_ERR_UNSUPPORTED, "qg_bezier_rescale: more than 216 coefficients");
  QG_TRY
  QG_CUDA(cudaSetDevice(device));
  cudaStream_t s = device_stream(device);
  g_alloc_stream = s;
  g_alloc_device = device;
  DevBuf<double> dc((size_t)nc), db((size_t)n_boxes * 2 * dim + 1), dout((size_t)n_boxes * nc + 1);
  DevBuf<signed char> ds((size_t)n_boxes + 1);
  dc.upload(f->coefs.data(), (size_t)nc, s);
  db.upload(boxes, (size_t)n_boxes * 2 * dim, s);
  if (n_boxes)
    bezier_rescale_kernel<<<QG_CNT(grid_for(n_boxes, 64)), 64, 0, s>>>(dim, f->f.order[0], f->f.order[1], dim > 2 ? f->f.order[2] : 1, dc.p, db.p,
      n_boxes, dout.p, ds.p);
  QG_CUDA(cudaGetLastError());
  dout.download(coefs_out, (size_t)n_boxes * nc, s);
  ds.download((signed char *)sign_out, (size_t)n_boxes, s);
  QG_CUDA(cudaStreamSynchronize(s));
  return QG_OK;
  QG_CATCH
}

int qg_cache_release(int device, int64_t *bytes_released)
{
  QG_TRY
  if (device < 0 || device >= 64) return fail(QG_ERR_INVALID, "qg_cache_release: bad device");
  QG_CUDA(cudaSetDevice(device));
  DeviceCtx &c = device_ctx(device);
  QG_CUDA(cudaStreamSynchronize(c.stream));
  std::lock_guard<std::mutex> lock(c.m);
  if (bytes_released) *bytes_released = (int64_t)c.cached_bytes;
  cache_release_all(c);
  return QG_OK;
  QG_CATCH
}
int qg_pinned_release(int64_t *bytes_released)
{
  QG_TRY
  const size_t b = g_pinned.release_all();
  if (bytes_released) *bytes_released = (int64_t)b;
  return QG_OK;
  QG_CATCH
}
int qg_rule_pass_counts(const qg_rule *r, long long counts[4])
{
  if (!r || !counts) return fail(QG_ERR_INVALID, "null argument");
  for (int i = 0; i < 4; ++i) counts[i] = r->times.counts[i];
  return QG_OK;
}

int qg_rule_pass_ms(const qg_rule *r, double ms[8])
{
  if (!r || !ms) return fail(QG_ERR_INVALID, "null argument");
  for (int i = 0; i < 8; ++i) ms[i] = r->times.ms[i];
  return QG_OK;
}
void qg_rule_free(qg_rule *r)
{
  if (r) cudaSetDevice(r->device);
  delete r;
}

int qg_gauss_01(int dim, int n, double *points, double *weights)
{
  if (dim < 1 || dim > 3 || n < 1 || n > kGaussMax || !points || !weights) return fail(QG_ERR_INVALID, "qg_gauss_01: bad arguments");
  const double *t = h_gauss_table + 2 * (size_t)(n * (n - 1) / 2);
  int tot = 1;
  for (int d = 0; d < dim; ++d) tot *= n;
  for (int k = 0; k < tot; ++k) {
    int idx = k;
    double w = 1.0;
    for (int d = 0; d < dim; ++d) {
      const int i = idx % n;
      idx /= n;
      points[(size_t)k * dim + d] = t[2 * i];
      w = d == 0 ? t[2 * i + 1] : w * t[2 * i + 1];
    }
    weights[k] = w;
  }
  return QG_OK;
}

int qg_reparam(const qg_domain *d, const int64_t *cells, int64_t n_cells, int order, int mode, int flags, double merge_tol, qg_mesh **out)
{
  if (!d || !out || (n_cells && !cells)) return fail(QG_ERR_INVALID, "null argument");
  if (order < 2 || order > kRpMaxOrder) return fail(QG_ERR_INVALID, "order must be in 2..32");
  if (mode < 0 || mode >= 2 + 2 * d->dim) return fail(QG_ERR_INVALID, "mode: 0 volume, 1 level set, 2 + local facet id");
  QG_TRY
  std::lock_guard<std::mutex> lock(d->mtx);
  QG_CUDA(cudaSetDevice(d->device));
  g_alloc_stream = d->stream;
  g_alloc_device = d->device;
  std::vector<long long> c(cells, cells + n_cells);
  *out = d->dim == 2 ? reparam_impl<2>(*d, std::move(c), order, mode, flags, merge_tol) : reparam_impl<3>(*d, std::move(c), order, mode, flags, merge_tol);
  return QG_OK;
  QG_CATCH
}

int qg_mesh_view(qg_mesh *m, int *dim, int *param_dim, int *order, int64_t *n_points, int64_t *n_cells, int64_t *n_wire_edges,
  const double **points, const int64_t **connectivity, const int64_t **wires_connectivity)
{
  if (!m) return fail(QG_ERR_INVALID, "null argument");
  QG_TRY
  if (!m->downloaded) {
    QG_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = device_stream(m->device);
    long long npc = 1;
    for (int i = 0; i < m->pd; ++i) npc *= m->order;
    const size_t n0 = (size_t)m->n_points * m->dim, n1 = (size_t)(m->n_cells * npc), n2 = (size_t)(m->n_wire_edges * m->order);
    m->h_pts = (double *)g_pinned.get(n0 * sizeof(double) + 8, &m->cap[0]);
    m->h_conn = (long long *)g_pinned.get(n1 * sizeof(long long) + 8, &m->cap[1]);
    m->h_wires = (long long *)g_pinned.get(n2 * sizeof(long long) + 8, &m->cap[2]);
    m->d_pts.download(m->h_pts, n0, s);
    m->d_conn.download(m->h_conn, n1, s);
    m->d_wires.download(m->h_wires, n2, s);
    QG_CUDA(cudaStreamSynchronize(s));
    m->downloaded = true;
  }
  if (dim) *dim = m->dim;
  if (param_dim) *param_dim = m->pd;
  if (order) *order = m->order;
  if (n_points) *n_points = m->n_points;
  if (n_cells) *n_cells = m->n_cells;
  if (n_wire_edges) *n_wire_edges = m->n_wire_edges;
  if (points) *points = m->h_pts;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  if (connectivity) *connectivity = reinterpret_cast<const int64_t *>(m->h_conn);
  if (wires_connectivity) *wires_connectivity = reinterpret_cast<const int64_t *>(m->h_wires);
  return QG_OK;
  QG_CATCH
}

int qg_mesh_device_view(const qg_mesh *m, int64_t *n_points, int64_t *n_cells, int64_t *n_wire_edges, const double **d_points,
  const int64_t **d_connectivity, const int64_t **d_wires_connectivity)
{
  if (!m) return fail(QG_ERR_INVALID, "null argument");
  if (n_points) *n_points = m->n_points;
  if (n_cells) *n_cells = m->n_cells;
  if (n_wire_edges) *n_wire_edges = m->n_wire_edges;
  if (d_points) *d_points = m->d_pts.p;
  if (d_connectivity) *d_connectivity = reinterpret_cast<const int64_t *>(m->d_conn.p);
  if (d_wires_connectivity) *d_wires_connectivity = reinterpret_cast<const int64_t *>(m->d_wires.p);
  return QG_OK;
}

void qg_mesh_free(qg_mesh *m)
{
  if (!m) return;
  cudaSetDevice(m->device);
  delete m;
}

}  // extern "C"
