// qugar_b200 -- CUDA engine: kernels, pass scheduling, domain classification and the C ABI.
// Compiled for sm_100a only, with -fmad=false (see qg_math.cuh).  No CPU compute path exists here:
// every entry point that computes needs a CUDA device and fails with QG_ERR_NO_DEVICE otherwise.
#include "../../include/qugar_b200.h"
#include "qg_pipeline.cuh"

#include <cub/device/device_scan.cuh>
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

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg)
{
  g_last_error = msg;
  return code;
}
#define QG_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      throw EngineError(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? QG_ERR_NO_DEVICE \
                                                                                    : QG_ERR_CUDA,    \
        std::string(#call) + ": " + cudaGetErrorString(e_));                                           \
    }                                                                                                  \
  } while (0)

struct EngineError
{
  int code;
  std::string msg;
  EngineError(int c, const std::string &m) : code(c), msg(m) {}
};

// ------------------------------------------------------------------------------------------------
// per-device context: one stream and a caching allocator.  Work buffers of a rule run to tens of GB and a
// steady workload asks for the same sizes again and again; freed blocks are kept per device and handed
// back to the next request of (nearly) the same size, so that steady-state calls make no driver
// allocation at all and the heap does not fragment.  All work of a device runs on its single stream,
// which orders reuse after the previous user.
// ------------------------------------------------------------------------------------------------
struct DeviceCtx
{
  cudaStream_t stream = nullptr;
  bool ready = false;
  std::mutex m;
  std::multimap<size_t, void *> free_blocks;   // size -> block
  std::unordered_map<void *, size_t> live;     // block -> size
  size_t cached_bytes = 0;
};
static DeviceCtx g_ctx[64];
static std::mutex g_ctx_mtx;

static DeviceCtx &device_ctx(int device)
{
  std::lock_guard<std::mutex> lock(g_ctx_mtx);
  DeviceCtx &c = g_ctx[device];
  if (!c.ready) {
    QG_CUDA(cudaSetDevice(device));
    QG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.ready = true;
  }
  return c;
}
static cudaStream_t device_stream(int device) { return device_ctx(device).stream; }

static void cache_release_all(DeviceCtx &c)
{
  for (auto &kv : c.free_blocks) cudaFree(kv.second);
  c.free_blocks.clear();
  c.cached_bytes = 0;
}

static void *cache_alloc(int device, size_t bytes)
{
  DeviceCtx &c = device_ctx(device);
  std::lock_guard<std::mutex> lock(c.m);
  bytes = (bytes + 511) & ~(size_t)511;
  // best fit among cached blocks, never wasting more than 25 % (+64 KB) of the block
  auto it = c.free_blocks.lower_bound(bytes);
  if (it != c.free_blocks.end() && it->first <= bytes + bytes / 4 + 65536) {
    void *p = it->second;
    c.live[p] = it->first;
    c.cached_bytes -= it->first;
    c.free_blocks.erase(it);
    return p;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    QG_CUDA(cudaStreamSynchronize(c.stream));
    cache_release_all(c);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) throw EngineError(QG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  c.live[p] = bytes;
  return p;
}
static void cache_free(int device, void *p)
{
  if (!p) return;
  DeviceCtx &c = g_ctx[device];
  std::lock_guard<std::mutex> lock(c.m);
  auto it = c.live.find(p);
  if (it == c.live.end()) return;
  c.free_blocks.emplace(it->second, p);
  c.cached_bytes += it->second;
  c.live.erase(it);
}

// device used by DevBuf allocations made on this thread (set by every API entry point)
static thread_local int g_alloc_device = 0;
static thread_local cudaStream_t g_alloc_stream = nullptr;

// ------------------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------------------
template<typename T> struct DevBuf
{
  T *p = nullptr;
  size_t n = 0;
  int dev = 0;
  DevBuf() {}
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), dev(o.dev)
  {
    o.p = nullptr;
    o.n = 0;
  }
  DevBuf &operator=(DevBuf &&o) noexcept
  {
    if (this != &o) {
      release();
      p = o.p;
      n = o.n;
      dev = o.dev;
      o.p = nullptr;
      o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release()
  {
    if (p) cache_free(dev, p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t n_)
  {
    release();
    n = n_;
    dev = g_alloc_device;
    if (n) p = (T *)cache_alloc(dev, n * sizeof(T));
  }
  void zero(cudaStream_t st)
  {
    if (n) QG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
  }
  void upload(const T *h, size_t cnt, cudaStream_t st)
  {
    if (cnt) QG_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st));
  }
  void download(T *h, size_t cnt, cudaStream_t st) const
  {
    if (cnt) QG_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
  }
};

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// every kernel launch of the engine is counted (bench.py reports the launches of its timed region)
static std::atomic<long long> g_launches{ 0 };
#define QG_CNT(x) (g_launches.fetch_add(1, std::memory_order_relaxed), (x))
constexpr int kBlock = 128;
static inline unsigned grid_for(long long n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }

// Pipeline kernels are instantiated per (dimension, function kind): FuncK<KIND> fixes the kind at compile time.
template<int N, int KIND, bool EMIT> __global__ void __launch_bounds__(kBlock) ctl_kernel(PipeArgs a)
{
  const FuncK<KIND> f(a.f);
  ctl_body<N, EMIT>((long long)blockIdx.x * blockDim.x + threadIdx.x, a, f);
}
template<int N, int KIND> __global__ void __launch_bounds__(kBlock) k0_kernel(PipeArgs a)
{
  const FuncK<KIND> f(a.f);
  k0_body<N>((long long)blockIdx.x * blockDim.x + threadIdx.x, a, f);
}
// K1 / K2 expand parents into children placed by an exclusive scan.  A block owns kExpParents consecutive
// parents (= one contiguous range of children); their offsets sit in shared memory, and each thread finds
// the parent of its child with a short binary search there instead of a 24-step search through global memory.
constexpr int kExpParents = 64;
template<int NP>
__device__ __forceinline__ void expand_block(const long long *off, long long nparents, long long (&s_off)[NP + 1],
  long long &p0, int &np)
{
  p0 = (long long)blockIdx.x * NP;
  np = (int)((nparents - p0) < NP ? (nparents - p0) : NP);
  for (int i = threadIdx.x; i <= np; i += blockDim.x) s_off[i] = off[p0 + i];
  __syncthreads();
}
template<int NP> __device__ __forceinline__ int expand_owner(const long long (&s_off)[NP + 1], int np, long long c)
{
  int lo = 0, hi = np;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_off[mid] <= c)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}
template<int N, int KIND> __global__ void __launch_bounds__(kBlock) k1_kernel(PipeArgs a)
{
  const FuncK<KIND> f(a.f);
  __shared__ long long s_off[kExpParents + 1];
  long long p0;
  int np;
  expand_block<kExpParents>(a.off0, a.nitems, s_off, p0, np);
  for (long long t = s_off[0] + threadIdx.x; t < s_off[np]; t += blockDim.x) {
    const int o = expand_owner<kExpParents>(s_off, np, t);
    k1_work<N>(p0 + o, (int)(t - s_off[o]), t, a, f);
  }
}
template<int N, int KIND> __global__ void __launch_bounds__(kBlock, 5) k2_kernel(PipeArgs a)
{
  const FuncK<KIND> f(a.f);
  __shared__ long long s_off[kExpParents + 1];
  long long p0;
  int np;
  expand_block<kExpParents>(a.off1, a.ncall1, s_off, p0, np);
  for (long long u = s_off[0] + threadIdx.x; u < s_off[np]; u += blockDim.x) {
    const int o = expand_owner<kExpParents>(s_off, np, u);
    k2_work<N>(p0 + o, (int)(u - s_off[o]), u, a, f);
  }
}
// run CALL with the compile-time constant K set to the function kind
#define QG_KIND_CASE(KK, CALL) case KK: { constexpr int K = KK; CALL; } break;
// kinds with their own instantiation (the benchmark configurations); every other kind runs the K = -1 instantiation,
// which keeps the run-time switch (one more set of kernels instead of one per rarely used function)
#define QG_KIND_DISPATCH(kind, CALL)                                                     \
  switch (kind) {                                                                        \
    QG_KIND_CASE(QG_FUNC_SPHERE, CALL) QG_KIND_CASE(QG_FUNC_BEZIER, CALL)                \
    QG_KIND_CASE(QG_FUNC_SCHOEN, CALL) QG_KIND_CASE(QG_FUNC_SCHWARZ_D, CALL)             \
  default: { constexpr int K = -1; CALL; } break;                                        \
  }
// K3, the node writer (HBM-write bound).  A block owns kK3Lines consecutive stage-2 lines, i.e. one contiguous
// range of the rule.  Phase 1: thread i stages line i (its stage-2 record, the chain's directions, the job's
// cell box and shift) into shared memory with coalesced loads.  Phase 2: ONE THREAD PER NODE finds its line by a
// binary search of the block's offsets, evaluates the node and the sink from shared memory only, and each
// warp transposes its 32 nodes through shared memory so that points / normals leave as 32 consecutive
// doubles per store instruction (full 32-byte sectors).
// Same arithmetic as k3_body (qg_pipeline.cuh), which remains the per-line statement the CPU emulation runs.
constexpr int kK3Lines = 256;
constexpr int kK3Threads = 256;
struct K3Rec
{
  double x0, x1, w, e[4];
  double lo[3], hi[3];
  long long shift;
  int kd;  // bits 0-5: kind of stages 0..2 (2 bits each); bits 6-11: their directions (2 bits each)
  int jt;  // job type
};
template<int N, int KIND> __global__ void __launch_bounds__(kK3Threads, 3) k3_kernel(PipeArgs a, const long long *job_shift)
{
  const FuncK<KIND> f(a.f);
  __shared__ long long s_off[kK3Lines + 1];
  __shared__ K3Rec s_rec[kK3Lines];
  __shared__ double s_stage[kK3Threads / 32][32 * N];
  const long long u0 = (long long)blockIdx.x * kK3Lines;
  const int nl = (int)((a.ncall2 - u0) < kK3Lines ? (a.ncall2 - u0) : kK3Lines);
  for (int i = threadIdx.x; i <= nl; i += kK3Threads) s_off[i] = a.off2[u0 + i];
  __syncthreads();
  for (int i = threadIdx.x; i < nl; i += kK3Threads) {
    if (s_off[i + 1] == s_off[i]) continue;  // empty line: never looked up
    const Call2 &c = a.call2[u0 + i];
    K3Rec &r = s_rec[i];
    r.x0 = c.x0;
    r.x1 = c.x1;
    r.w = c.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.e[k] = c.ent[k];
    const int job = c.job;
    const double *box = a.job_box + 6 * (size_t)job;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      r.lo[d] = box[d];
      r.hi[d] = box[3 + d];
    }
    r.shift = job_shift ? job_shift[job] : 0;
    r.kd = c.kd;
    r.jt = a.job_type[job];
  }
  __syncthreads();
  const long long q_begin = s_off[0], q_end = s_off[nl];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *stage = s_stage[warp];
  const int sink_mode = a.sink_mode;
  const bool surf = sink_mode == SINK_SURF;
  const int p = a.p;
  for (long long base = q_begin + warp * 32; base < q_end; base += kK3Threads) {
    const long long q = base + lane;
    const bool valid = q < q_end;
    double pv[N], nv[N], wv = 0.0;
    long long dst = 0;
    const int pd = sink_mode == SINK_FACET ? N - 1 : N;  // warp-uniform (invalid lanes take part in the stores below)
#pragma unroll
    for (int d = 0; d < N; ++d) pv[d] = nv[d] = 0.0;
    if (valid) {
      // last line whose offset is <= q (empty lines share their successor's offset and are skipped)
      int lo = 0, hi = nl;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= q)
          lo = mid;
        else
          hi = mid;
      }
      const K3Rec &L = s_rec[lo];
      const int r = (int)(q - s_off[lo]);
      const int kd = L.kd;
      double x[N];
#pragma unroll
      for (int d = 0; d < N; ++d) x[d] = 0.0;
      if ((kd & 3) != ST_PASS) {
#pragma unroll
        for (int d = 0; d < N; ++d)
          if (d == ((kd >> 6) & 3)) x[d] = L.x0;
      }
      if (((kd >> 2) & 3) != ST_PASS) {
#pragma unroll
        for (int d = 0; d < N; ++d)
          if (d == ((kd >> 8) & 3)) x[d] = L.x1;
      }
      double w = L.w;
      double g[N];
      const bool have_g = stage_child_k<N>(f, (kd >> 4) & 3, (kd >> 10) & 3, L.e, r, p, a.gxw, x, w, g);
      sink_compute<N>(sink_mode, f, L.lo, L.hi, L.jt, x, w, g, have_g, pv, wv, nv);
      dst = q + L.shift;
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int nvalid = __popc(vmask);  // valid lanes are 0..nvalid-1
    const long long dst0 = __shfl_sync(0xffffffffu, dst, 0);
    const bool contiguous = __all_sync(0xffffffffu, !valid || dst == dst0 + lane);
    if (valid) a.wts[dst] = wv;
    if (contiguous && pd > 1) {
      for (int pass = 0; pass < (surf ? 2 : 1); ++pass) {
        double *out = (pass == 0 ? a.pts : a.nrm) + dst0 * pd;
        if (valid) {
#pragma unroll
          for (int d = 0; d < N; ++d)
            if (d < pd) stage[lane * pd + d] = pass == 0 ? pv[d] : nv[d];
        }
        __syncwarp();
        for (int e = lane; e < nvalid * pd; e += 32) out[e] = stage[e];
        __syncwarp();
      }
    } else if (valid) {
      for (int d = 0; d < pd; ++d) a.pts[dst * pd + d] = pv[d];
      if (surf)
        for (int d = 0; d < N; ++d) a.nrm[dst * N + d] = nv[d];
    }
  }
}
// K3 for the 3-D interior rule (SINK_CELL), the pass that writes most of the bytes of a step.  No level-set
// evaluation happens here, so there is one instantiation for all function kinds.  A block owns kC3Lines
// consecutive stage-2 lines, i.e. one contiguous range of the rule:
//   A  thread i <-> line i: loads the line (stage-2 record, chain directions, cell box) and stores what every
//      node of the line shares -- the two reference coordinates fixed along the line, the cell volume, and the
//      exact reciprocals of the line's extent and of the volume -- in shared memory; it also enters the line
//      into the owner table of its one or two GROUPS (a group = the p nodes of one active segment; in this
//      mode every line holds 0, p or 2p nodes).
//   B  thread t <-> group t: reads its line's record once and computes the p nodes in a loop (the coordinate
//      along the line and the weight, each with an exact reciprocal-multiply-correct division, div_rcp),
//      writing them to a padded staging tile in shared memory.
//   C  the tile leaves for HBM as contiguous doubles (points) and contiguous doubles (weights).
// B and C repeat per chunk of kC3Cap nodes.  Bit-identical to k3_body + sink_compute(SINK_CELL).
// (Tried and rejected: per-thread 16-byte stores straight to HBM without the tile -- 10.7 ms instead of 6.5 ms on C3;
// the half-written sectors cost more than the staging; lane pairs completing whole sectors per store -- 6.8 ms, no gain;
// a second round of register look-ahead -- no change.  DRAM is 51 % busy, L1TEX 73 %: the tile passes are latency bound.)
constexpr int kMaxGaussP = 100;
constexpr int kC3Lines = 128;
constexpr int kC3Cap = 1024;  // nodes staged per chunk (p <= kC3Cap is guaranteed by kMaxGaussP)
struct K3CellRec
{
  double pc[3];  // reference coordinates of the node for the directions fixed on this line
  double e[4];   // up to two active segments along the line
  double lo_k, len_k, rlen_k, w, vol, rvol;
  long long shift;
  int k2, kind2;
};
constexpr int kC3PtsDoubles = kC3Cap * 3 + kC3Cap / 8 + 4;
constexpr size_t kC3SmemBytes = sizeof(K3CellRec) * kC3Lines + sizeof(double) * (kC3PtsDoubles + kC3Cap + 2 * kMaxGaussP)
                                + sizeof(long long) * (kC3Lines + 1) + 2 * kC3Lines;
__device__ __forceinline__ int c3_phys(int slot, int d) { return slot * 3 + d + (slot >> 3); }  // padded AoS tile
// PACK = true writes the transport encoding instead of the point array: per group pack_pc / pack_k2, per node pack_pk;
// the weights are written as usual.  Only for rules without relocation (job_shift == nullptr).
template<bool PACK>
__global__ void __launch_bounds__(kC3Lines, 4) k3_cell3_kernel(PipeArgs a, const long long *job_shift, long long ntiles)
{
  constexpr int N = 3;
  extern __shared__ __align__(16) unsigned char c3_smem[];
  K3CellRec *s_rec = reinterpret_cast<K3CellRec *>(c3_smem);
  double *s_pts = reinterpret_cast<double *>(s_rec + kC3Lines);
  double *s_wts = s_pts + kC3PtsDoubles;
  double *s_gxw = s_wts + kC3Cap;
  long long *s_off = reinterpret_cast<long long *>(s_gxw + 2 * kMaxGaussP);
  unsigned char *s_owner = reinterpret_cast<unsigned char *>(s_off + kC3Lines + 1);
  __shared__ long long s_shift0;
  const int tid = threadIdx.x;
  const int p = a.p;
  for (int i = tid; i < 2 * p; i += kC3Lines) s_gxw[i] = a.gxw[i];
  // Persistent blocks; the line records and offsets of the tiles one and two rounds ahead sit in two register
  // buffers, fetched while the current tile is computed and written (those loads are DRAM-latency bound, the rest
  // is not; with one round of look-ahead 16 % of the warp time was still spent waiting for them).
  struct Pref
  {
    Call2 c;
    long long off, off_last;
  };
  const long long G = gridDim.x;
  auto prefetch = [&](long long t, Pref &b) {
    if (t >= ntiles) return;
    const long long u = t * kC3Lines + tid;
    if (u < a.ncall2) {
      const double2 *src = reinterpret_cast<const double2 *>(a.call2 + u);
      double2 *dstp = reinterpret_cast<double2 *>(&b.c);
#pragma unroll
      for (int k = 0; k < (int)(sizeof(Call2) / sizeof(double2)); ++k) dstp[k] = __ldcs(src + k);
    }
    if (u <= a.ncall2) b.off = a.off2[u];
    if (tid == 0) {
      const long long ul = (t + 1) * kC3Lines < a.ncall2 ? (t + 1) * kC3Lines : a.ncall2;
      b.off_last = a.off2[ul];
    }
  };
  // one tile; consumes buffer b and refills it for the tile two rounds ahead.  Returns false on a fatal error.
  auto do_tile = [&](long long tile, Pref &b) -> bool {
    const long long u0 = tile * kC3Lines;
    const int nl = (int)((a.ncall2 - u0) < kC3Lines ? (a.ncall2 - u0) : kC3Lines);
    __syncthreads();  // the previous tile is done with shared memory
    if (tid <= nl) s_off[tid] = b.off;
    if (tid == 0) {
      s_off[nl] = b.off_last;
      s_shift0 = 0;
    }
    const Call2 c = b.c;
    __syncthreads();
    const long long q_begin = s_off[0], q_end = s_off[nl];
    // ---- A: one line per thread
    bool ok = true;
    long long my_shift = 0;
    int my_cnt = 0;
    if (tid < nl) {
      my_cnt = (int)(s_off[tid + 1] - s_off[tid]);
      if (my_cnt > 0) {
        K3CellRec &r = s_rec[tid];
        const int job = c.job;
        const double *box = a.job_box + 6 * (size_t)job;
        const int kd = c.kd;
        const int k0 = kd_dim(kd, 0), k1 = kd_dim(kd, 1), k2 = kd_dim(kd, 2);
        double x[N], len[N], lo[N];
        double vol = 1.0;
#pragma unroll
        for (int d = 0; d < N; ++d) {
          lo[d] = box[d];
          len[d] = box[3 + d] - lo[d];
          vol *= len[d];
          x[d] = 0.0;
          if (kd_kind(kd, 0) != ST_PASS && d == k0) x[d] = c.x0;
          if (kd_kind(kd, 1) != ST_PASS && d == k1) x[d] = c.x1;
        }
#pragma unroll
        for (int d = 0; d < N; ++d) {
          r.pc[d] = (x[d] - lo[d]) / len[d];
          if (d == k2) {
            r.lo_k = lo[d];
            r.len_k = len[d];
            r.rlen_k = 1.0 / len[d];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) r.e[k] = c.ent[k];
        r.w = c.w;
        r.vol = vol;
        r.rvol = 1.0 / vol;
        my_shift = job_shift ? job_shift[job] : 0;
        r.shift = my_shift;
        r.k2 = k2;
        r.kind2 = kd_kind(kd, 2);
        const int ng = my_cnt / p;
        ok = (ng * p == my_cnt) && ng <= 2 && kd_kind(kd, 2) != ST_PASS && kd_kind(kd, 2) != ST_SURF;
        if (ok) {
          const int g0 = (int)((s_off[tid] - q_begin) / p);
          for (int k = 0; k < ng; ++k) s_owner[g0 + k] = (unsigned char)tid;
        }
        s_shift0 = my_shift;  // any active line's shift (all equal in the common case)
      }
    }
    if (!__syncthreads_and(ok)) {
      if (tid == 0) *a.err |= ERR_ITEM_CAP;  // not a 3-D interior rule: the launcher chose the wrong kernel
      return false;
    }
    const long long shift0 = s_shift0;
    const bool uniform = __syncthreads_and(my_cnt == 0 || my_shift == shift0);
    prefetch(tile + 2 * G, b);
    // ---- B / C per chunk of whole groups
    const int ngroups = (int)((q_end - q_begin) / p);
    const int gchunk = kC3Cap / p;  // groups per chunk
    for (int gbase = 0; gbase < ngroups; gbase += gchunk) {
      const int gn = (ngroups - gbase) < gchunk ? (ngroups - gbase) : gchunk;
      if (gbase > 0) __syncthreads();
      for (int g = tid; g < gn; g += kC3Lines) {
        const int line = s_owner[gbase + g];
        const K3CellRec &L = s_rec[line];
        const long long q0 = q_begin + (long long)(gbase + g) * p;
        const int sg = (q0 - s_off[line]) >= p ? 1 : 0;
        const double x0 = L.e[2 * sg], dx = L.e[2 * sg + 1] - x0;
        const double lo_k = L.lo_k, len_k = L.len_k, rlen_k = L.rlen_k, w0 = L.w, vol = L.vol, rvol = L.rvol;
        const int k2 = L.k2;
        double pc[N];
#pragma unroll
        for (int d = 0; d < N; ++d) pc[d] = L.pc[d];
        const int slot0 = g * p;
        if (PACK) {
          const long long gg = q0 / p;  // global group index (no relocation in this mode)
#pragma unroll
          for (int d = 0; d < N; ++d) a.pack_pc[gg * N + d] = pc[d];
          a.pack_k2[gg] = (unsigned char)k2;
        }
        for (int j = 0; j < p; ++j) {
          const double xk = x0 + dx * s_gxw[2 * j];
          const double w = w0 * (dx * s_gxw[2 * j + 1]);
          const double pk = div_rcp(xk - lo_k, len_k, rlen_k);
          const int slot = slot0 + j;
          if (PACK) {
            s_pts[slot] = pk;
          } else {
#pragma unroll
            for (int d = 0; d < N; ++d) s_pts[c3_phys(slot, d)] = (d == k2) ? pk : pc[d];
          }
          s_wts[slot] = div_rcp(w, vol, rvol);
        }
      }
      __syncthreads();
      const int n = gn * p;
      const long long qc = q_begin + (long long)gbase * p;
      if (PACK) {
        if (!uniform || shift0 != 0) {
          if (tid == 0) *a.err |= ERR_ITEM_CAP;  // the packed encoding has no relocation
          return false;
        }
        for (int e = tid; e < n; e += kC3Lines) {
          __stcs(a.pack_pk + qc + e, s_pts[e]);
          __stcs(a.wts + qc + e, s_wts[e]);
        }
      } else if (uniform) {
        double *op = a.pts + (qc + shift0) * N;
        double *ow = a.wts + (qc + shift0);
        for (int e = tid; e < n * N; e += kC3Lines) __stcs(op + e, s_pts[e + e / 24]);
        for (int e = tid; e < n; e += kC3Lines) __stcs(ow + e, s_wts[e]);
      } else {
        for (int i = tid; i < n; i += kC3Lines) {
          const long long dst = qc + i + s_rec[s_owner[gbase + i / p]].shift;
#pragma unroll
          for (int d = 0; d < N; ++d) a.pts[dst * N + d] = s_pts[c3_phys(i, d)];
          a.wts[dst] = s_wts[i];
        }
      }
    }
    return true;
  };
  Pref b0, b1;
  long long tile = blockIdx.x;
  prefetch(tile, b0);
  prefetch(tile + G, b1);
  for (;;) {
    if (tile >= ntiles || !do_tile(tile, b0)) break;
    tile += G;
    if (tile >= ntiles || !do_tile(tile, b1)) break;
    tile += G;
  }
}
__global__ void job_points_kernel(PipeArgs a, long long *job_pt_off)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  job_points_body(j, a, nullptr, job_pt_off);
}

// phi / grad at points
template<int N> __global__ void func_eval_kernel(FuncDesc f, const double *pts, long long n, double *val, double *grd)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[N];
  for (int d = 0; d < N; ++d) x[d] = pts[i * N + d];
  val[i] = phi_val<N>(f, x);
  if (grd) {
    double g[N];
    phi_grad<N>(f, x, g);
    for (int d = 0; d < N; ++d) grd[i * N + d] = g[d];
  }
}

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

// classify_facet_from_quad + classify_facet_full_unfitted_boundary (impl_unfitted_domain.cpp:237-333)
// job_slot (may be null = identity): position of job j among the 2N facets of all cells (cell index * 2N + facet)
template<int N>
__global__ void classify_facets_kernel(FuncDesc f, long long njobs, const int *job_type, const long long *job_pt_off,
  const double *pts, const double *wts, const double *job_box, const unsigned char *cell_tmp, const long long *job_slot,
  unsigned char *fstat_all)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const long long slot = job_slot ? job_slot[j] : j;
  unsigned char *fstat = fstat_all + (slot - j);  // so that fstat[j] below addresses the facet's slot
  const int lf = job_type[j];
  const int cd = lf >> 1, side = lf & 1;
  const unsigned char cs = cell_tmp[slot / (2 * N)];
  const long long q0 = job_pt_off[j], q1 = job_pt_off[j + 1];
  if (q1 == q0) {
    fstat[j] = cs == TMP_FULL_UNF ? QG_FACET_FULL_UNF_BDRY : QG_FACET_EMPTY;
    return;
  }
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
  if (cs == TMP_FULL_UNF && !cut) {
    fstat[j] = QG_FACET_FULL_UNF_BDRY;
    return;
  }
  const unsigned char values[8] = { QG_FACET_EMPTY, QG_FACET_EXT_BDRY, QG_FACET_UNF_BDRY, QG_FACET_UNF_BDRY_EXT_BDRY,
    QG_FACET_CUT, QG_FACET_CUT_EXT_BDRY, QG_FACET_CUT_UNF_BDRY, QG_FACET_CUT_UNF_BDRY_EXT_BDRY };
  unsigned char st = values[(ext ? 1 : 0) + (unf ? 2 : 0) + (cut ? 4 : 0)];
  double dom_vol = 1.0;
  for (int d = 0; d < N; ++d)
    if (d != cd) dom_vol *= (job_box[6 * j + 3 + d] - job_box[6 * j + d]);
  const double max_val = fabs(fvol) < fabs(dom_vol) ? fabs(dom_vol) : fabs(fvol);
  const bool is_full = fabs(fvol - dom_vol) <= (kNearEps + (kNearEps * 1.0e3) * max_val);
  if (is_full) {
    const long long npts = q1 - q0;
    if ((st == QG_FACET_UNF_BDRY || st == QG_FACET_EXT_BDRY) && npts == 1)
      st = st == QG_FACET_UNF_BDRY ? QG_FACET_FULL_UNF_BDRY : QG_FACET_FULL_EXT_BDRY;
    else if (st == QG_FACET_CUT)
      st = QG_FACET_FULL;
  }
  fstat[j] = st;
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

struct Timer
{
  cudaEvent_t a, b;
  cudaStream_t s;
  explicit Timer(cudaStream_t s_) : s(s_)
  {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~Timer()
  {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
  void start() { cudaEventRecord(a, s); }
  double stop()
  {
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
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
  std::vector<double> leaf_coefs[2];
};

// Device copies of what a function descriptor points to (Bezier coefficients, composite description).
struct FuncDeviceData
{
  DevBuf<double> coefs, leaf_coefs[2];
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
          h.leaf[l].f.coefs = leaf_coefs[l].p;
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
  mutable bool mirrored = false;
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
  cudaSetDevice(device);
  const int nf = 2 * dim;
  status.resize((size_t)ncells);
  rec_cells.resize((size_t)n_rec);
  rec_stat.resize((size_t)n_rec * nf);
  d_status.download(status.data(), (size_t)ncells, stream);
  if (n_rec) {
    d_rec_cells.download(rec_cells.data(), (size_t)n_rec, stream);
    d_rec_stat.download(rec_stat.data(), (size_t)n_rec * nf, stream);
  }
  cudaStreamSynchronize(stream);
  mirrored = true;
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
  DevBuf<int> item_cnt, nent0, cnt0, cnt1, cnt2;
  DevBuf<long long> item_off, off0, off1, off2, job_pt_off;
  DevBuf<ChainItem> items;
  DevBuf<double> ent0;
  DevBuf<Call1> call1;
  DevBuf<Call2> call2;
  PassTimes times;

  Pipeline(const qg_domain &d, int p, int sink_mode) : dom(d), s(d.stream)
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
    if (emit) {
      QG_KIND_DISPATCH(a.f.kind, (ctl_kernel<N, K, true><<<QG_CNT(grid_for(njobs)), kBlock, 0, s>>>(a)))
    } else {
      QG_KIND_DISPATCH(a.f.kind, (ctl_kernel<N, K, false><<<QG_CNT(grid_for(njobs)), kBlock, 0, s>>>(a)))
    }
  }

  // Runs CTL, K0, K1, K2 and the scans; afterwards the number of nodes of every job is known.
  template<int N> void run_counts(const long long *d_job_cell, const int *d_job_type, long long nj)
  {
    njobs = nj;
    Timer t(s);
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
    t.start();
    launch_ctl<N>(false);
    times.ms[0] += t.stop();
    t.start();
    dom.scanner.run(item_cnt.p, item_off.p, (size_t)nj + 1, s);
    a.nitems = read_total(item_off.p + nj);
    times.ms[5] += t.stop();
    items.alloc((size_t)a.nitems + 1);
    a.items = items.p;
    t.start();
    launch_ctl<N>(true);
    times.ms[0] += t.stop();
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
    t.start();
    if (a.nitems) { QG_KIND_DISPATCH(a.f.kind, (k0_kernel<N, K><<<QG_CNT(grid_for(a.nitems)), kBlock, 0, s>>>(a))) }
    times.ms[1] += t.stop();
    t.start();
    dom.scanner.run(cnt0.p, off0.p, (size_t)a.nitems + 1, s);
    a.ncall1 = read_total(off0.p + a.nitems);
    times.ms[5] += t.stop();
    // stage 1
    call1.alloc((size_t)a.ncall1 + 1);
    cnt1.alloc((size_t)a.ncall1 + 1);
    cnt1.zero(s);
    off1.alloc((size_t)a.ncall1 + 1);
    a.call1 = call1.p;
    a.cnt1 = cnt1.p;
    a.off1 = off1.p;
    t.start();
    if (a.ncall1) { QG_KIND_DISPATCH(a.f.kind, (k1_kernel<N, K><<<QG_CNT(grid_for(a.nitems, kExpParents)), kBlock, 0, s>>>(a))) }
    times.ms[2] += t.stop();
    t.start();
    dom.scanner.run(cnt1.p, off1.p, (size_t)a.ncall1 + 1, s);
    a.ncall2 = read_total(off1.p + a.ncall1);
    times.ms[5] += t.stop();
    // stage 2
    call2.alloc((size_t)a.ncall2 + 1);
    cnt2.alloc((size_t)a.ncall2 + 1);
    cnt2.zero(s);
    off2.alloc((size_t)a.ncall2 + 1);
    a.call2 = call2.p;
    a.cnt2 = cnt2.p;
    a.off2 = off2.p;
    t.start();
    if (a.ncall2) { QG_KIND_DISPATCH(a.f.kind, (k2_kernel<N, K><<<QG_CNT(grid_for(a.ncall1, kExpParents)), kBlock, 0, s>>>(a))) }
    times.ms[3] += t.stop();
    t.start();
    dom.scanner.run(cnt2.p, off2.p, (size_t)a.ncall2 + 1, s);
    a.npts = read_total(off2.p + a.ncall2);
    job_pt_off.alloc((size_t)nj + 1);
    job_points_kernel<<<QG_CNT(grid_for(nj + 1)), kBlock, 0, s>>>(a, job_pt_off.p);
    times.ms[5] += t.stop();
    times.counts[0] = a.nitems;
    times.counts[1] = a.ncall1;
    times.counts[2] = a.ncall2;
    times.counts[3] = a.npts;
    QG_CUDA(cudaGetLastError());
  }

  // K3: writes the nodes.  shift (may be null) relocates each job's nodes inside the output arrays.
  template<int N> void run_write(double *pts, double *wts, double *nrm, const long long *shift)
  {
    Timer t(s);
    a.pts = pts;
    a.wts = wts;
    a.nrm = nrm;
    t.start();
    if (a.ncall2) {
      if (N == 3 && a.sink_mode == SINK_CELL && a.p <= kMaxGaussP) {
        static bool attr_done[64] = {};
        if (!attr_done[dom.device]) {
          QG_CUDA(cudaFuncSetAttribute(k3_cell3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3SmemBytes));
          QG_CUDA(cudaFuncSetAttribute(k3_cell3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3SmemBytes));
          attr_done[dom.device] = true;
        }
        const long long ntiles = (a.ncall2 + kC3Lines - 1) / kC3Lines;
        const long long nblk = ntiles < (long long)dom.num_sms * 4 ? ntiles : (long long)dom.num_sms * 4;
        if (a.pack_pk)
          k3_cell3_kernel<true><<<QG_CNT((unsigned)nblk), kC3Lines, kC3SmemBytes, s>>>(a, nullptr, ntiles);
        else
          k3_cell3_kernel<false><<<QG_CNT((unsigned)nblk), kC3Lines, kC3SmemBytes, s>>>(a, shift, ntiles);
      } else {
        QG_KIND_DISPATCH(a.f.kind, (k3_kernel<N, K><<<QG_CNT(grid_for(a.ncall2, kK3Lines)), kK3Threads, 0, s>>>(a, shift)))
      }
    }
    times.ms[4] += t.stop();
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
      if (h[1] > 0) kd_fill_kernel<N><<<QG_CNT(h[1]), 256, 0, s>>>(dm.g, uni.p, unis.p, dm.d_status.p, dm.k0, dm.k1);
      QG_CUDA(cudaStreamSynchronize(s));
      cur = std::move(next);
      ncur = h[0];
    }
  }
  tr.mark("kd-tree");
  // --- undetermined cells, ascending ids
  DevBuf<int> flag((size_t)nc + 1);
  flag.zero(s);
  DevBuf<long long> off((size_t)nc + 1);
  flag_eq_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(dm.d_status.p, nc, ST_UNDET, flag.p);
  const long long nu = scan_flags(dm, flag, off, nc);
  DevBuf<long long> ucells((size_t)nu + 1);
  scatter_ids_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(flag.p, off.p, nc, ucells.p);
  flag.release();
  off.release();

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
        const long long nfp = scan_flags(dm, fflag, foff, nfj);
        if (nfp > 0) {
          DevBuf<long long> fcell((size_t)nfp), fslot((size_t)nfp);
          DevBuf<int> ftype((size_t)nfp);
          gather_facet_slots_kernel<<<QG_CNT(grid_for(nfj)), kBlock, 0, s>>>(nfj, vcells.p, nf, fflag.p, foff.p, fcell.p, ftype.p, fslot.p);
          Pipeline pl(dm, 1, SINK_RAW);
          pl.run_counts<N>(fcell.p, ftype.p, nfp);
          DevBuf<double> pts((size_t)pl.a.npts * N + 1), wts((size_t)pl.a.npts + 1);
          pl.run_write<N>(pts.p, wts.p, nullptr, nullptr);
          classify_facets_kernel<N><<<QG_CNT(grid_for(nfp)), kBlock, 0, s>>>(
            dm.f, nfp, ftype.p, pl.job_pt_off.p, pts.p, wts.p, pl.job_box.p, vtmp.p, fslot.p, fstat.p);
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
  status_hist_kernel<<<QG_CNT(592), 256, 0, s>>>(dm.d_status.p, nc, hist.p);
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
  Timer tt(s), to(s);
  tt.start();
  to.start();
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
  double other = to.stop();

  Pipeline pl(dm, p, surf ? SINK_SURF : SINK_CELL);
  pl.run_counts<N>(job_cell.p, job_type.p, nj);

  to.start();
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
      other += to.stop();
      pl.run_write<N>(nullptr, rule->d_wts.p, nullptr, nullptr);
      to.start();
      other += to.stop();
      QG_CUDA(cudaGetLastError());
      check_device_err(dm);
      rule->times = pl.times;
      rule->times.ms[6] = other;
      rule->times.ms[7] = tt.stop();
      return rule.release();
    }
    alloc_rule_outputs(*rule, ne, npts, N, surf);
    if (nj) job_shift_kernel<<<QG_CNT(grid_for(nj)), kBlock, 0, s>>>(nj, job_ent.p, pl.job_pt_off.p, ent_off.p, shift.p);
    other += to.stop();
    pl.run_write<N>(rule->d_pts.p, rule->d_wts.p, surf ? rule->d_nrm.p : nullptr, shift.p);
    to.start();
    if (ne && !surf) gauss_fill_kernel<<<QG_CNT(grid_for(ne * 32, 256)), 256, 0, s>>>(ne, kind.p, ent_off.p, N, p, dm.gauss_rule(p), rule->d_pts.p, rule->d_wts.p);
    other += to.stop();
  } else {
    // surface rule: nodes go to a staging rule first; facet-coincident nodes are purged when a cell has
    // facets carrying unfitted boundary (rare), then the rule is compacted into its final place.
    DevBuf<double> tp((size_t)pl.a.npts * N + 1), tw((size_t)pl.a.npts + 1), tn((size_t)pl.a.npts * N + 1);
    other += to.stop();
    pl.run_write<N>(tp.p, tw.p, tn.p, nullptr);
    to.start();
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
    other += to.stop();
  }
  QG_CUDA(cudaGetLastError());
  check_device_err(dm);
  rule->times = pl.times;
  rule->times.ms[6] = other;
  rule->times.ms[7] = tt.stop();
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
  Timer tt(s);
  tt.start();
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
  rule->times = pl.times;
  rule->times.ms[7] = tt.stop();
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
};
static PinnedPool g_pinned;

}  // namespace qg
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
      // 12 doubles of 4 rows: e = 3 r + d takes pk[r] where d == k2, else c[d]
      for (size_t j = 0; j < p; j += 4, row += 12) {
        double t[12];
        for (int r = 0; r < 4; ++r)
          for (unsigned d = 0; d < 3; ++d) t[3 * r + d] = d == k2 ? pk[j + r] : c[d];
        _mm256_stream_pd(row, _mm256_loadu_pd(t));
        _mm256_stream_pd(row + 4, _mm256_loadu_pd(t + 4));
        _mm256_stream_pd(row + 8, _mm256_loadu_pd(t + 8));
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
  std::vector<cudaEvent_t> ev(nchunks);
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
  for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
  worker();
  for (auto &t : pool) t.join();
  cudaError_t e = cudaStreamSynchronize(s);
  if (std::getenv("QG_TRACE"))
    std::fprintf(stderr, "[qg trace] packed download: %zu groups, %zu chunks, %u host threads, %.1f ms\n", ng, nchunks, nthreads,
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
  for (auto &x : ev) cudaEventDestroy(x);
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
  return nullptr;
}

int qg_func_create(int kind, int dim, const double *params, int n_params, int negative, qg_func **out)
{
  if (!out || !params || (dim != 2 && dim != 3)) return fail(QG_ERR_INVALID, "qg_func_create: bad arguments");
  std::unique_ptr<qg_func> f(new qg_func());
  std::memset(&f->comp, 0, sizeof(CompositeDesc));
  if (kind == QG_COMPOSITE) {
    // params: op, nleaf, then per leaf: kind, negative, has_aff, aff[12], np, p[np]
    if (n_params < 2) return fail(QG_ERR_INVALID, "qg_func_create: composite needs op, nleaf, leaves");
    const int op = (int)params[0], nleaf = (int)params[1];
    if (op < 0 || op > 2 || nleaf != (op == 0 ? 1 : 2)) return fail(QG_ERR_INVALID, "qg_func_create: bad composite op / leaf count");
    std::memset(&f->f, 0, sizeof(FuncDesc));
    f->f.kind = QG_COMPOSITE;
    f->f.dim = dim;
    f->f.negative = negative ? 1 : 0;
    f->comp.op = op;
    f->comp.nleaf = nleaf;
    long long q = 2;
    for (int l = 0; l < nleaf; ++l) {
      if (q + 16 > n_params) return fail(QG_ERR_INVALID, "qg_func_create: truncated composite leaf");
      const int lkind = (int)params[q], lneg = (int)params[q + 1];
      const long long np = (long long)params[q + 15];
      if (lkind == QG_COMPOSITE) return fail(QG_ERR_UNSUPPORTED, "qg_func_create: nested composites are not supported");
      if (np < 0 || q + 16 + np > n_params) return fail(QG_ERR_INVALID, "qg_func_create: truncated composite leaf");
      CompositeLeaf &lf = f->comp.leaf[l];
      lf.has_aff = params[q + 2] != 0.0 ? 1 : 0;
      for (int i = 0; i < 12; ++i) lf.aff[i] = params[q + 3 + i];
      const char *err = make_leaf_desc(lkind, dim, params + q + 16, np, lneg, lf.f, f->leaf_coefs[l]);
      if (err) return fail(lkind > QG_SCHWARZ_PRIMITIVE || lkind < 0 ? QG_ERR_UNSUPPORTED : QG_ERR_INVALID, std::string("qg_func_create: composite leaf: ") + err);
      q += 16 + np;
    }
    if (q != n_params) return fail(QG_ERR_INVALID, "qg_func_create: trailing composite parameters");
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
  if (f->f.kind >= QG_SCHOEN) {
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
  if (d->mirrored) {
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
  const long long nc = d->ncells;
  DevBuf<int> flag((size_t)nc + 1);
  flag.zero(s);
  DevBuf<long long> off((size_t)nc + 1);
  flag_eq_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(d->d_status.p, nc, (unsigned char)status, flag.p);
  const long long n = scan_flags(*d, flag, off, nc);
  DevBuf<long long> out((size_t)n + 1);
  scatter_ids_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(flag.p, off.p, nc, out.p);
  out.download((long long *)ids, (size_t)n, s);
  QG_CUDA(cudaStreamSynchronize(s));
  return QG_OK;
  QG_CATCH
}
int qg_domain_cell_status(const qg_domain *d, uint8_t *status)
{
  if (!d || !status) return fail(QG_ERR_INVALID, "null argument");
  d->mirror();
  std::memcpy(status, d->status.data(), d->status.size());
  return QG_OK;
}
int qg_domain_facet_records(const qg_domain *d, int64_t *cells, uint8_t *statuses)
{
  if (!d || !cells || !statuses) return fail(QG_ERR_INVALID, "null argument");
  d->mirror();
  for (size_t i = 0; i < d->rec_cells.size(); ++i) cells[i] = d->rec_cells[i];
  std::memcpy(statuses, d->rec_stat.data(), d->rec_stat.size());
  return QG_OK;
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
  d->mirror();
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
  const long long nc = d->ncells;
  DevBuf<int> flag((size_t)nc + 1);
  flag.zero(s);
  DevBuf<long long> off((size_t)nc + 1);
  flag_eq_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(d->d_status.p, nc, ST_CUT, flag.p);
  const long long n = scan_flags(*d, flag, off, nc);
  std::unique_ptr<qg_dcells> dc(new qg_dcells());
  dc->n = n;
  dc->d_cells.alloc((size_t)n + 1);
  scatter_ids_kernel<<<QG_CNT(grid_for(nc)), kBlock, 0, s>>>(flag.p, off.p, nc, dc->d_cells.p);
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

int qg_quad_cells(const qg_domain *d, const int64_t *cells, int64_t n_cells, int n_pts_dir, qg_rule **out)
{
  qg_dcells *dc = nullptr;
  int rc = qg_domain_upload_cells(d, cells, n_cells, &dc);
  if (rc) return rc;
  rc = quad_device(d, dc, n_pts_dir, false, out, std::getenv("QG_NO_PACK") == nullptr);
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
  QG_CUDA(cudaSetDevice(device));
  QG_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice));
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

}  // extern "C"
