// qugar_b200 -- kernels of the reparameterization passes (bodies in qg_reparam.cuh).
#pragma once
#include "qg_reparam.cuh"
#include "qg_launch_cfg.cuh"

namespace qg {

struct RpLeafArgs
{
  FuncDesc f;
  int order, npc;
  int mode;  // 0 volume, 1 level set, 2 + f: face f
  const double *nodes;
  long long nitems;
  const ChainItem *items;
  int *leaf_cells, *leaf_pts;            // count pass
  const long long *cell_off, *pt_off;    // emit pass
  double *pts;
  long long *conn;
  int *cell_job;
  int *err;
};

// one thread per chain item (= leaf of the reference's recursion); F: the level-set kind fixed at compile time (qg_kind.cu
// instantiates it per kind like the pipeline kernels: with the run-time kind switch the kernel stalled on instruction fetch,
// ncu: no_instruction 11 of 19 stall cycles per issue, profiles/r2_ncu_rp_leaf_grid128.txt)
constexpr int kRpBlock = 64;
template<int N, bool EMIT, class F> __global__ void __launch_bounds__(kRpBlock) rp_leaf_kernel(RpLeafArgs a)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nitems) return;
  RpCtx<N, F> c;
  c.f = F(a.f);
  c.it = a.items[i];
  c.order = a.order;
  c.nodes = a.nodes;
  c.fixdir = a.mode >= 2 ? (a.mode - 2) / 2 : -1;
  c.emit = EMIT;
  c.pts = a.pts;
  c.conn = a.conn;
  c.npc = a.npc;
  c.err = 0;
  c.pt_base = EMIT ? a.pt_off[i] : 0;
  c.cell_base = EMIT ? a.cell_off[i] : 0;
  rp_leaf<N>(c);
  if (EMIT) {
    for (int k = 0; k < c.ncells; ++k) a.cell_job[c.cell_base + k] = c.it.job;
  } else {
    a.leaf_cells[i] = c.ncells;
    a.leaf_pts[i] = (int)c.npts;
  }
  if (c.err) flag_or(a.err, c.err);
}

template<int N> __global__ void rp_orient_kernel(RpMeshArgs a)
{
  rp_orient_body<N>((long long)blockIdx.x * blockDim.x + threadIdx.x, a);
}
template<int N, bool EMIT> __global__ void rp_wire_kernel(RpMeshArgs a)
{
  rp_wire_body<N, EMIT>((long long)blockIdx.x * blockDim.x + threadIdx.x, a);
}
template<int N>
__global__ void rp_full_cells_kernel(long long nfull, const long long *cells, GridDesc g, int order, const double *nodes, long long pt_base,
  long long cell_base, long long wire_base, double *pts, long long *conn, long long *wires)
{
  rp_full_cell_body<N>((long long)blockIdx.x * blockDim.x + threadIdx.x, nfull, cells, g, order, nodes, pt_base, cell_base, wire_base, pts,
    conn, wires);
}

#if defined(QG_ENGINE_UNIT)  // non-template kernels: defined once, in the engine unit (the kind units include this file for rp_leaf_kernel)
// ---- merge of coincident points (point.cpp:34-134, reparam_mesh.cpp:607-626) with STABLE sorts ------------------------------
// order-preserving map double -> uint64
__device__ __forceinline__ unsigned long long rp_key(double v)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);  // -0.0 -> +0.0: equal keys for equal values
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
// val[i] = seg(i) << 32 | point index; keys for the coordinate sort / for the segment sort
__global__ void rp_merge_init_kernel(long long n, unsigned long long *val)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) val[i] = (unsigned long long)i;
}
__global__ void rp_merge_coord_keys_kernel(long long n, const unsigned long long *val, const double *pts, int dim, int d, unsigned long long *key)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = rp_key(pts[(val[i] & 0xffffffffull) * dim + d]);
}
__global__ void rp_merge_seg_keys_kernel(long long n, const unsigned long long *val, unsigned *key)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (unsigned)(val[i] >> 32);
}
// head[i] = 1 where a new group starts: other segment, or coordinate d further than tol from the predecessor's
__global__ void rp_merge_heads_kernel(long long n, const unsigned long long *val, const double *pts, int dim, int d, double tol, int *head)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = 1;
  if (i > 0) {
    const unsigned long long a = val[i - 1], b = val[i];
    h = (a >> 32) != (b >> 32) || !(fabs(pts[(a & 0xffffffffull) * dim + d] - pts[(b & 0xffffffffull) * dim + d]) <= tol);
  }
  head[i] = h;
}
// seg(i) = (number of heads up to i) - 1, from the exclusive scan of head
__global__ void rp_merge_set_seg_kernel(long long n, const int *head, const long long *off, unsigned long long *val)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) val[i] = ((unsigned long long)(off[i] + head[i] - 1) << 32) | (val[i] & 0xffffffffull);
}
// first[seg] = point index of the first element of the group
__global__ void rp_merge_first_kernel(long long n, const int *head, const unsigned long long *val, long long *first)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && head[i]) first[val[i] >> 32] = (long long)(val[i] & 0xffffffffull);
}
__global__ void rp_merge_map_kernel(long long n, const unsigned long long *val, const long long *first, long long *map, int *uniq)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long p = (long long)(val[i] & 0xffffffffull), r = first[val[i] >> 32];
  map[p] = r;
  uniq[p] = p == r ? 1 : 0;
}
__global__ void rp_merge_compact_kernel(long long n, int dim, const int *uniq, const long long *newid, const double *pts, double *out)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !uniq[i]) return;
  for (int d = 0; d < dim; ++d) out[newid[i] * dim + d] = pts[i * dim + d];
}
__global__ void rp_merge_remap_kernel(long long n, const long long *map, const long long *newid, long long *ids)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = newid[map[ids[i]]];
}
// sort_wirebasket_edges / purge_duplicate_wirebasket_edges (reparam_mesh.cpp:659-735): direction of every edge, then the
// edges in lexicographic order without repetitions
__global__ void rp_edge_dir_kernel(long long ne, int order, long long *wires)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  long long *w = wires + e * order;
  bool reverse = w[order - 1] < w[0];
  if (w[0] == w[order - 1] && order > 2) reverse = w[order - 2] < w[1];
  if (reverse)
    for (int i = 0; i < order / 2; ++i) {
      const long long v = w[i];
      w[i] = w[order - 1 - i];
      w[order - 1 - i] = v;
    }
}
__global__ void rp_edge_init_kernel(long long ne, unsigned *perm)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne) perm[e] = (unsigned)e;
}
__global__ void rp_edge_keys_kernel(long long ne, int order, int c, const unsigned *perm, const long long *wires, unsigned long long *key)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne) key[e] = (unsigned long long)wires[(long long)perm[e] * order + c];
}
__global__ void rp_edge_heads_kernel(long long ne, int order, const unsigned *perm, const long long *wires, int *head)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int h = 1;
  if (e > 0) {
    h = 0;
    const long long *a = wires + (long long)perm[e - 1] * order, *b = wires + (long long)perm[e] * order;
    for (int i = 0; i < order; ++i)
      if (a[i] != b[i]) h = 1;
  }
  head[e] = h;
}
__global__ void rp_edge_compact_kernel(long long ne, int order, const unsigned *perm, const int *head, const long long *off,
  const long long *wires, long long *out)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || !head[e]) return;
  const long long *a = wires + (long long)perm[e] * order;
  for (int i = 0; i < order; ++i) out[off[e] * order + i] = a[i];
}

#endif  // QG_ENGINE_UNIT

}  // namespace qg
