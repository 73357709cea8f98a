// qugar_b200 -- kind-independent kernels of the quadrature pipeline: the 3-D interior node writer (k3_cell3), job
// offsets, function evaluation; launch counter.  The kernels that evaluate the level set live in qg_kernels_kind.cuh.
// (part of the translation unit qg_engine.cu; included there, in this order)
#pragma once

namespace qg {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// every kernel launch of the engine is counted (bench.py reports the launches of its timed region)
static std::atomic<long long> g_launches{ 0 };
#define QG_CNT(x) (g_launches.fetch_add(1, std::memory_order_relaxed), (x))
static inline unsigned grid_for(long long n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }
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
//      writing them to a staging tile in shared memory that is the byte image of the output range.  The loop over
//      the nodes starts at a lane-dependent node (rotation): the lanes' p-node groups are 24 p bytes apart, so without
//      it a store instruction would hit two banks 16 times over.
//   C  the tile leaves for HBM with two TMA bulk copies (points, weights) issued by one thread; ranges that are not
//      16-byte aligned (odd offsets: odd p, relocated rules) are copied by all threads instead.
// B and C repeat per chunk of kC3Cap nodes.  Bit-identical to k3_body + sink_compute(SINK_CELL).
// (Tried and rejected: per-thread 16-byte stores straight to HBM without the tile -- 10.7 ms instead of 6.5 ms on C3;
// the half-written sectors cost more than the staging; lane pairs completing whole sectors per store -- 6.8 ms, no gain;
// a second round of register look-ahead -- no change; 128-bit tile stores / copies -- no change.  DRAM is 51 % busy, L1TEX 73 %: the tile passes are latency bound.)
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
constexpr int kC3PtsDoubles = kC3Cap * 3;  // the tiles are unpadded images of the output ranges: they leave by bulk copies
constexpr int kC3WtsDoubles = kC3Cap;
constexpr size_t kC3SmemBytes = sizeof(K3CellRec) * kC3Lines + sizeof(double) * (kC3PtsDoubles + kC3WtsDoubles + 2 * kMaxGaussP)
                                + sizeof(long long) * (kC3Lines + 1) + 2 * kC3Lines;
static_assert(sizeof(K3CellRec) * kC3Lines % 16 == 0 && (kC3PtsDoubles * sizeof(double)) % 16 == 0, "bulk copies need 16-byte aligned tiles");
// TMA bulk copy of a contiguous shared-memory range to global memory (cp.async.bulk, shared::cta -> global): one
// instruction per tile instead of one load + one store per double.  dst, src and bytes are multiples of 16.
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread waits until its bulk copies have READ their shared-memory source (the tile may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (the bulk copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
  double *s_gxw = s_wts + kC3WtsDoubles;
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
  // No predicates here: a guarded load goes to a temporary register and is moved into the buffer under the predicate, and that
  // move waits for the load -- the look-ahead would stall on the spot (17 % of the warp samples).  Out-of-range tiles and
  // lines re-read the last valid record instead; nothing uses those copies.
  auto prefetch = [&](long long t, Pref &b) {
    const long long tc = t < ntiles ? t : ntiles - 1;
    const long long u = tc * kC3Lines + tid;
    const long long ur = u < a.ncall2 ? u : a.ncall2 - 1;
    const double2 *src = reinterpret_cast<const double2 *>(a.call2 + ur);
    double2 *dstp = reinterpret_cast<double2 *>(&b.c);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Call2) / sizeof(double2)); ++k) dstp[k] = __ldcs(src + k);
    b.off = a.off2[u < a.ncall2 ? u : a.ncall2];
    const long long ul = (tc + 1) * kC3Lines < a.ncall2 ? (tc + 1) * kC3Lines : a.ncall2;
    b.off_last = a.off2[ul];
  };
  // one tile; consumes buffer b and refills it for the tile two rounds ahead.  Returns false on a fatal error.
  auto do_tile = [&](long long tile, Pref &b) -> bool {
    const long long u0 = tile * kC3Lines;
    const int nl = (int)((a.ncall2 - u0) < kC3Lines ? (a.ncall2 - u0) : kC3Lines);
    // the cell box and the relocation of this thread's line: issued before the barriers so that their latency overlaps them
    // (every thread holds a valid record, see prefetch)
    const int job_l = b.c.job;
    double bx[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) bx[d] = a.job_box[6 * (size_t)job_l + d];
    const long long shift_l = job_shift ? job_shift[job_l] : 0;
    if (tid == 0) bulk_wait_read();
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
        const double *box = bx;
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
        my_shift = shift_l;
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
      if (tid == 0) flag_or(a.err, ERR_ITEM_CAP);  // not a 3-D interior rule: the launcher chose the wrong kernel
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
      if (gbase > 0) {
        if (tid == 0) bulk_wait_read();
        __syncthreads();
      }
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
        int j = tid % p;  // rotated start, see B above
        for (int it = 0; it < p; ++it) {
          const double xk = x0 + dx * s_gxw[2 * j];
          const double w = w0 * (dx * s_gxw[2 * j + 1]);
          const double pk = div_rcp(xk - lo_k, len_k, rlen_k);
          const int slot = slot0 + j;
          if (PACK) {
            s_pts[slot] = pk;
          } else {
#pragma unroll
            for (int d = 0; d < N; ++d) s_pts[slot * N + d] = (d == k2) ? pk : pc[d];
          }
          s_wts[slot] = div_rcp(w, vol, rvol);
          j = (j + 1 == p) ? 0 : j + 1;
        }
      }
      fence_async_smem();
      __syncthreads();
      const int n = gn * p;
      const long long qc = q_begin + (long long)gbase * p;
      if (PACK) {
        if (!uniform || shift0 != 0) {
          if (tid == 0) flag_or(a.err, ERR_ITEM_CAP);  // the packed encoding has no relocation
          return false;
        }
        double *op = a.pack_pk + qc, *ow = a.wts + qc;
        if ((((size_t)op | (size_t)ow) & 15) == 0 && (n & 1) == 0) {
          if (tid == 0) {
            bulk_store(op, s_pts, (unsigned)n * 8u);
            bulk_store(ow, s_wts, (unsigned)n * 8u);
            bulk_commit();
          }
        } else {
          for (int e = tid; e < n; e += kC3Lines) {
            __stcs(op + e, s_pts[e]);
            __stcs(ow + e, s_wts[e]);
          }
        }
      } else if (uniform) {
        double *op = a.pts + (qc + shift0) * N;
        double *ow = a.wts + (qc + shift0);
        if ((((size_t)op | (size_t)ow) & 15) == 0 && (n & 1) == 0) {
          if (tid == 0) {
            bulk_store(op, s_pts, (unsigned)n * 24u);
            bulk_store(ow, s_wts, (unsigned)n * 8u);
            bulk_commit();
          }
        } else {
          for (int e = tid; e < n * N; e += kC3Lines) __stcs(op + e, s_pts[e]);
          for (int e = tid; e < n; e += kC3Lines) __stcs(ow + e, s_wts[e]);
        }
      } else {
        for (int i = tid; i < n; i += kC3Lines) {
          const long long dst = qc + i + s_rec[s_owner[gbase + i / p]].shift;
#pragma unroll
          for (int d = 0; d < N; ++d) a.pts[dst * N + d] = s_pts[i * N + d];
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
  if (tid == 0) bulk_wait_all();  // the last tiles' bulk copies complete before the block retires
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

}  // namespace qg
