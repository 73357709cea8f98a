// qugar_b200 -- the pipeline kernels that evaluate the level set (CTL, K0, K1, K2, generic K3), templated on the
// function kind, and their launcher.  Compiled once per specialised kind (qg_kind.cu with -DQG_KIND=...), in parallel
// translation units; qg_engine.cu sizes the grids and calls launch_pipe_kind through one entry point per kind.
#pragma once
#include "qg_pipeline.cuh"
#include "qg_launch_cfg.cuh"
#include "qg_kernels_reparam.cuh"

#include <cuda_runtime.h>

namespace qg {

// Pipeline kernels are instantiated per (dimension, function kind): FuncK<KIND> fixes the kind at compile time.
// Minimum resident blocks per SM (a register cap) of the one-thread-per-item kernels.  For the trigonometric level sets these
// kernels wait on fixed-latency fp64 chains with 2 blocks per SM (CTL 218, K0 240, K1 116 registers uncapped); capping the
// registers buys warps: CTL 0.63 -> 0.58 ms, K0 0.41 -> 0.33 ms, K1 1.22 -> 1.17 ms per rule on C3.  Other kinds keep
// the compiler's choice (the unrolled de Casteljau of the Bezier units would spill).
constexpr bool kind_is_tpms(int kind) { return kind >= QG_FUNC_SCHOEN && kind <= QG_FUNC_SCHWARZ_P; }
constexpr bool kind_has_arg_styles(int kind) { return kind == QG_FUNC_SCHOEN || kind == QG_FUNC_SCHOEN_IWP; }
constexpr int min_blocks_ctl(int kind) { return kind_is_tpms(kind) ? 3 : 1; }
constexpr int min_blocks_k0(int kind) { return kind_is_tpms(kind) ? 4 : 1; }
constexpr int min_blocks_k1(int kind) { return kind_is_tpms(kind) ? 5 : 1; }
template<int N, int KIND, bool SAME> struct K2Func { typedef FuncK<KIND> type; };
template<int N, int KIND> struct K2Func<N, KIND, true> { typedef FuncKSame<KIND> type; };
template<int N, int KIND, bool EMIT> __global__ void __launch_bounds__(kBlock, min_blocks_ctl(KIND)) ctl_kernel(PipeArgs a)
{
  const FuncK<KIND> f(a.f);
  ctl_body<N, EMIT>((long long)blockIdx.x * blockDim.x + threadIdx.x, a, f);
}
template<int N, int KIND, bool SAME = false> __global__ void __launch_bounds__(kBlock, min_blocks_k0(KIND)) k0_kernel(PipeArgs a)
{
  const typename K2Func<N, KIND, SAME>::type f(a.f);
  k0_body<N>((long long)blockIdx.x * blockDim.x + threadIdx.x, a, f);
}
// K1 / K2 expand parents into children placed by an exclusive scan.  A block owns kExpParents consecutive
// parents (= one contiguous range of children); their offsets sit in shared memory, and each thread finds
// the parent of its child with a short binary search there instead of a 24-step search through global memory.
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
template<int N, int KIND, bool SAME = false> __global__ void __launch_bounds__(kBlock, min_blocks_k1(KIND)) k1_kernel(PipeArgs a)
{
  const typename K2Func<N, KIND, SAME>::type f(a.f);
  __shared__ long long s_off[kExpParents + 1];
  long long p0;
  int np;
  expand_block<kExpParents>(a.off0, a.nitems, s_off, p0, np);
  for (long long t = s_off[0] + threadIdx.x; t < s_off[np]; t += blockDim.x) {
    const int o = expand_owner<kExpParents>(s_off, np, t);
    k1_work<N>(p0 + o, (int)(t - s_off[o]), t, a, f);
  }
}
// K2 register budget: 4 blocks per SM (<= 128 registers).  With the stage-2 record assembled in registers and stored with
// 128-bit stores the Schoen instantiation needs 114 registers and no spills; at 5 blocks (96 registers) it spills 68 bytes
// (C3, same-argument instantiation: 5.44 / 3.21 ms at 4 blocks, 5.59 / 3.24 ms at 5; the kernel is not occupancy bound:
// 12, 16 and 20 warps per SM are within 3 %).
#ifndef QG_K2_MIN_BLOCKS
#define QG_K2_MIN_BLOCKS 4
#endif
// SAME: Schoen / SchoenIWP with power-of-two frequencies (FuncKSame): one set of trig values per line, 96 registers
// without the 88 bytes of spills of the general instantiation (C3: interior 6.82 -> 5.82 ms, boundary 3.86 -> 3.52 ms).
// Measured and dropped: one launch per height direction with a single solver copy each (24 KB instead of 49 KB of SASS)
// -- 5.82 -> 6.96 ms: warps hold lines of several directions, every pass idles the lanes of the other two.
template<int N, int KIND, bool SAME = false>
__global__ void __launch_bounds__(kBlock, QG_K2_MIN_BLOCKS) k2_kernel(PipeArgs a)
{
  const typename K2Func<N, KIND, SAME>::type f(a.f);
  __shared__ long long s_off[kExpParents + 1];
  long long p0;
  int np;
  // the block's parents (stage-1 calls) and their chain items are staged in shared memory next to the offsets: the
  // offset, the call and -- once the call is there -- its item are requested by thread i for parent i before the one
  // barrier, instead of every child walking offset -> call -> item through L1 / L2 on its own
  __shared__ Call1 s_c1[kExpParents];
  __shared__ ChainItem s_it[kExpParents];
  static_assert(kExpParents < kBlock, "thread i stages parent i");
  p0 = (long long)blockIdx.x * kExpParents;
  np = (int)((a.ncall1 - p0) < kExpParents ? (a.ncall1 - p0) : kExpParents);
  if (np <= 0) return;
  {
    const int i = threadIdx.x;
    if (i <= np) s_off[i] = a.off1[p0 + i];
    if (i < np) {
      const Call1 c1 = a.call1[p0 + i];
      s_it[i] = a.items[c1.item];
      s_c1[i] = c1;
    }
  }
  __syncthreads();
  for (long long u = s_off[0] + threadIdx.x; u < s_off[np]; u += blockDim.x) {
    const int o = expand_owner<kExpParents>(s_off, np, u);
    k2_work_rec<N>(s_c1[o], s_it[o], (int)(u - s_off[o]), u, a, f);
  }
}
// K3, the node writer (HBM-write bound).  A block owns kK3Lines consecutive stage-2 lines, i.e. one contiguous
// range of the rule.  Phase 1: thread i stages line i (its stage-2 record, the chain's directions, the job's
// cell box and shift) into shared memory with coalesced loads.  Phase 2: ONE THREAD PER NODE finds its line by a
// binary search of the block's offsets, evaluates the node and the sink from shared memory only, and each
// warp transposes its 32 nodes through shared memory so that points / normals leave as 32 consecutive
// doubles per store instruction (full 32-byte sectors).
// Same arithmetic as k3_body (qg_pipeline.cuh), which remains the per-line statement the CPU emulation runs.
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
  if (nl <= 0) return;
  // Phase 1.  Everything line i needs is requested before the first barrier -- its offset, its record (lines past the end
  // re-read the last one: unguarded loads, nothing uses those copies) and, as soon as the record is there, the job's box,
  // shift and type -- so that the block pays two memory round trips instead of three behind two barriers.
  static_assert(kK3Lines == kK3Threads, "thread i stages line i");
  {
    const int i = threadIdx.x;
    const long long off_i = a.off2[u0 + (i < nl ? i : nl)];
    const double2 *src = reinterpret_cast<const double2 *>(a.call2 + u0 + (i < nl ? i : nl - 1));
    Call2 c;
    double2 *cp = reinterpret_cast<double2 *>(&c);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Call2) / sizeof(double2)); ++k) cp[k] = src[k];
    const int job = c.job;
    double bx[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) bx[d] = a.job_box[6 * (size_t)job + d];
    const long long shift = job_shift ? job_shift[job] : 0;
    const int jt = a.job_type[job];
    s_off[i] = off_i;
    if (i == 0) s_off[nl] = a.off2[u0 + nl];
    __syncthreads();
    if (i < nl && s_off[i + 1] != s_off[i]) {  // empty lines are never looked up
      K3Rec &r = s_rec[i];
      r.x0 = c.x0;
      r.x1 = c.x1;
      r.w = c.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) r.e[k] = c.ent[k];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        r.lo[d] = bx[d];
        r.hi[d] = bx[3 + d];
      }
      r.shift = shift;
      r.kd = c.kd;
      r.jt = jt;
    }
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
// Launches kernel `which` of the pass for dimension `dim` (2 or 3) with this unit's kind.
// static: units split by dimension instantiate DIFFERENT bodies of launch_pipe_kind<KIND>; with external linkage the
// linker would keep one of them for both
template<int KIND>
static void launch_pipe_kind(int which, int dim, unsigned grid, cudaStream_t s, const PipeArgs &a, const long long *shift)
{
#if defined(QG_ONLY_DIM)
  // this unit holds one dimension only (the slow units are split in two to shorten the parallel build)
  if (dim != QG_ONLY_DIM) return;
#endif
#if !defined(QG_ONLY_DIM) || QG_ONLY_DIM == 2
  if (dim == 2) {
    switch (which) {
    case PK_CTL_COUNT: ctl_kernel<2, KIND, false><<<grid, kBlock, 0, s>>>(a); break;
    case PK_CTL_EMIT: ctl_kernel<2, KIND, true><<<grid, kBlock, 0, s>>>(a); break;
    case PK_K0:
      if (kind_has_arg_styles(KIND) && a.f.same_args == 7)
        k0_kernel<2, KIND, kind_has_arg_styles(KIND)><<<grid, kBlock, 0, s>>>(a);
      else
        k0_kernel<2, KIND><<<grid, kBlock, 0, s>>>(a);
      break;
    case PK_K1:
      if (kind_has_arg_styles(KIND) && a.f.same_args == 7)
        k1_kernel<2, KIND, kind_has_arg_styles(KIND)><<<grid, kBlock, 0, s>>>(a);
      else
        k1_kernel<2, KIND><<<grid, kBlock, 0, s>>>(a);
      break;
    case PK_K2:
      if (kind_has_arg_styles(KIND) && a.f.same_args == 7)
        k2_kernel<2, KIND, kind_has_arg_styles(KIND)><<<grid, kBlock, 0, s>>>(a);
      else
        k2_kernel<2, KIND><<<grid, kBlock, 0, s>>>(a);
      break;
    default: k3_kernel<2, KIND><<<grid, kK3Threads, 0, s>>>(a, shift); break;
    }
  }
#endif
#if !defined(QG_ONLY_DIM) || QG_ONLY_DIM == 3
  if (dim == 3) {
    switch (which) {
    case PK_CTL_COUNT: ctl_kernel<3, KIND, false><<<grid, kBlock, 0, s>>>(a); break;
    case PK_CTL_EMIT: ctl_kernel<3, KIND, true><<<grid, kBlock, 0, s>>>(a); break;
    case PK_K0:
      if (kind_has_arg_styles(KIND) && a.f.same_args == 7)
        k0_kernel<3, KIND, kind_has_arg_styles(KIND)><<<grid, kBlock, 0, s>>>(a);
      else
        k0_kernel<3, KIND><<<grid, kBlock, 0, s>>>(a);
      break;
    case PK_K1:
      if (kind_has_arg_styles(KIND) && a.f.same_args == 7)
        k1_kernel<3, KIND, kind_has_arg_styles(KIND)><<<grid, kBlock, 0, s>>>(a);
      else
        k1_kernel<3, KIND><<<grid, kBlock, 0, s>>>(a);
      break;
    case PK_K2:
      if (kind_has_arg_styles(KIND) && a.f.same_args == 7)
        k2_kernel<3, KIND, kind_has_arg_styles(KIND)><<<grid, kBlock, 0, s>>>(a);
      else
        k2_kernel<3, KIND><<<grid, kBlock, 0, s>>>(a);
      break;
    default: k3_kernel<3, KIND><<<grid, kK3Threads, 0, s>>>(a, shift); break;
    }
  }
#endif
}

// The reparameterization leaf kernel (qg_kernels_reparam.cuh) of this unit's kind; emit: 0 count pass, 1 write pass.
template<int KIND> static void launch_rp_kind(int emit, int dim, unsigned grid, cudaStream_t s, const RpLeafArgs &a)
{
#if defined(QG_ONLY_DIM)
  if (dim != QG_ONLY_DIM) return;
#endif
  const bool same = kind_has_arg_styles(KIND) && a.f.same_args == 7;
#if !defined(QG_ONLY_DIM) || QG_ONLY_DIM == 2
  if (dim == 2) {
    if (same) {
      if (emit)
        rp_leaf_kernel<2, true, typename K2Func<2, KIND, kind_has_arg_styles(KIND)>::type><<<grid, kRpBlock, 0, s>>>(a);
      else
        rp_leaf_kernel<2, false, typename K2Func<2, KIND, kind_has_arg_styles(KIND)>::type><<<grid, kRpBlock, 0, s>>>(a);
    } else {
      if (emit)
        rp_leaf_kernel<2, true, FuncK<KIND>><<<grid, kRpBlock, 0, s>>>(a);
      else
        rp_leaf_kernel<2, false, FuncK<KIND>><<<grid, kRpBlock, 0, s>>>(a);
    }
  }
#endif
#if !defined(QG_ONLY_DIM) || QG_ONLY_DIM == 3
  if (dim == 3) {
    if (same) {
      if (emit)
        rp_leaf_kernel<3, true, typename K2Func<3, KIND, kind_has_arg_styles(KIND)>::type><<<grid, kRpBlock, 0, s>>>(a);
      else
        rp_leaf_kernel<3, false, typename K2Func<3, KIND, kind_has_arg_styles(KIND)>::type><<<grid, kRpBlock, 0, s>>>(a);
    } else {
      if (emit)
        rp_leaf_kernel<3, true, FuncK<KIND>><<<grid, kRpBlock, 0, s>>>(a);
      else
        rp_leaf_kernel<3, false, FuncK<KIND>><<<grid, kRpBlock, 0, s>>>(a);
    }
  }
#endif
}

}  // namespace qg
