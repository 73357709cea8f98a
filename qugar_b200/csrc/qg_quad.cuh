// qugar_b200 -- cut-cell quadrature core (host/device inline; the kernels in qg_kernels_pipeline.cuh wrap these).
//
// The reference computes each cell's rule with a recursive, callback-driven dimension reduction
// (algoim::ImplicitIntegral driven from /root/reference/cpp/qugar/impl_quadrature.cpp:138-158,392-428 and
// impl_unfitted_domain.cpp:183-190,336-349; control flow as in impl_reparam_general.cpp:653-822).
// Here the same rule is produced by a level-synchronous pipeline of flat, data-parallel passes:
//
//   CTL   one thread per job (cell x rule type): an explicit-stack walk of the bisection / height-direction
//         decisions (interval arithmetic only).  Emits "chain items": one per leaf of that walk, carrying
//         the leaf box and, for each of up to three stages, the direction and the restricted-sign codes.
//   K0    one thread per chain item: roots of phi on the item's edges (stage 0, the 1-D base integral).
//   K1    one thread per stage-0 quadrature point: roots along the stage-1 direction (<= 2, monotone).
//   K2    one thread per stage-1 quadrature point: root along the stage-2 (height) direction (<= 1).
//   K3    one thread per stage-2 line: writes the Gauss points of its active segments to the rule.
//
// Stage s corresponds to the reference's level M = s+1.  Shorter chains (2-D cells, facet rules) pad the
// upper stages with PASS; leaves that the interval test proves entirely inside use PRESET stages
// (tensor-product Gauss rule).  Output order equals the reference's callback order because every
// expansion is placed with an exclusive scan of the per-parent counts.
#pragma once
#include "qg_funcs.cuh"
#ifndef QG_NEWTON_HOOK
#define QG_NEWTON_HOOK(s)
#endif

namespace qg {

enum StageKind : unsigned char { ST_PASS = 0, ST_PRESET = 1, ST_ROOTS = 2, ST_SURF = 3 };
enum JobType { JOB_VOLUME = -1, JOB_SURFACE = -2 /* >= 0: facet id = 2*const_dir + side */ };
enum SinkMode { SINK_CELL = 0, SINK_SURF = 1, SINK_FACET = 2, SINK_RAW = 3 };
enum ErrBits { ERR_ROOT_CAP = 1, ERR_STACK = 2, ERR_ITEM_CAP = 4 };

constexpr int kMaxSubdiv = 16;       // impl_reparam_general.cpp:72
constexpr int kRootFindMaxLevel = 4;
constexpr int kMaxSeg0 = 12;         // stage 0: active segments kept per item
constexpr int kMaxRoots = 24;        // roots (incl. end points) per line
constexpr int kStackCap = 48;
// Sign decoded from a default-constructed algoim::PsiCode, i.e. from the unused slots of a psi array, which the depth-cap branch
// reads (impl_reparam_general.cpp:776-800 loops over all 2^(N-1) slots).  PsiCode packs "unrestricted" as a SET bit and the sign
// of a restricted psi as one more bit (set: +1); a zero byte therefore decodes to side 0, sign -1.  The reference's golden sizes
// agree: Fischer-Koch S reparameterization 780 170 / 2 054 052 / 17 784 against 780 182 / 2 058 048 / 17 784 with -1 and
// 782 702 / 2 067 120 / 18 048 with 0.
constexpr int kPsiDefaultSign = -1;

struct GridDesc
{
  int dim;
  int pad;
  long long n[3];
  const double *breaks[3];
};

// psi code: bits 0..2 side per direction, bits 4..5 sign+1
QG_HD unsigned char psi_pack(int sides, int sign) { return (unsigned char)((sides & 7) | ((sign + 1) << 4)); }
QG_HD int psi_sides(unsigned char c) { return c & 7; }
QG_HD int psi_sign(unsigned char c) { return (int)(c >> 4) - 1; }

struct ChainItem
{
  double lo[3], hi[3];
  int job;
  unsigned char dim[3];
  unsigned char kind[3];
  unsigned char npsi[3];
  unsigned char psi[3][4];
  unsigned char pad[3];
};

struct Frame
{
  double lo[3], hi[3];
  unsigned char M, freemask, npsi, level;
  unsigned char psi[4];
  unsigned char S;     // this frame is the surface level (job is a surface rule and M is the top level)
  unsigned char surf;  // job is a surface rule
  // stages already fixed by ancestor levels (index = stage = level-1)
  unsigned char sdim[3], snpsi[3], spsi[3][4];
};

template<int N> QG_HD void cell_box(const GridDesc &g, long long id, double *lo, double *hi)
{
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const long long t = id % g.n[d];
    id /= g.n[d];
    lo[d] = g.breaks[d][t];
    hi[d] = g.breaks[d][t + 1];
  }
}

// Newton stop on a line [xmin,xmax]: 10 eps times the largest of the segment length and the end-point
// magnitudes, i.e. a few ulps of the coordinates.  (A length-only tolerance drops below one ulp of x for cells
// far from the origin; Newton then never meets it and the root would be lost -- the reference's golden
// boundary centroid, cpp/test/test_impl_general_quad_0.cpp:292-296, rules that behaviour out: DESIGN.md section 2.)
QG_HD double line_tol(double xmin, double xmax)
{
  double scale = xmax - xmin;
  const double a = fabs(xmin), b = fabs(xmax);
  if (a > scale) scale = a;
  if (b > scale) scale = b;
  return 10.0 * kEps * scale;
}

// Degenerate-segment filter of a line: 10 eps x the extent along the line
// (/root/reference/cpp/qugar/impl_reparam_general.cpp:155-160, 212-216, 250-251).
QG_HD double seg_tol(double xmin, double xmax) { return 10.0 * kEps * (xmax - xmin); }

// ---- 1-D root finding along direction k ------------------------------------------------------------

// Newton's method safeguarded by bisection on [x0,x1] (requires a sign change); tol absolute.
// Written with ONE evaluation site for the end points and ONE for the iteration (code size: every site
// carries inlined sincos bodies); the sequence of evaluations and updates is the textbook one:
// f(x0), f(x1); (f,f')(xc); then per step: new xc, stop tests, (f,f')(xc), bracket update.
template<int N, class F, int K> QG_HD bool newton_bisection(LineEval<N, F, K> &le, double x0, double x1, double tol, double &root)
{
  double f0 = 0.0, f1 = 0.0;
#pragma unroll 1
  for (int e = 0; e < 2; ++e) {
    const double v = le.val(e ? x1 : x0);
    if (e)
      f1 = v;
    else
      f0 = v;
  }
  if ((f0 > 0.0 && f1 > 0.0) || (f0 < 0.0 && f1 < 0.0)) return false;
  if (f0 == 0.0) {
    root = x0;
    return true;
  }
  if (f1 == 0.0) {
    root = x1;
    return true;
  }
  double xc = (x0 + x1) * 0.5;
  double xl, xh;
  if (f0 < 0.0) {
    xl = x0;
    xh = x1;
  } else {
    xl = x1;
    xh = x0;
  }
  double fx, fpx;
  double dxold = fabs(x1 - x0);
  double dx = dxold;
#pragma unroll 1
  for (int step = 0;; ++step) {
    le.val_der(xc, fx, fpx);
    if (step > 0) {
      if (fx < 0.0)
        xl = xc;
      else
        xh = xc;
      if (step >= 1024) break;
    }
    if (((xc - xh) * fpx - fx) * ((xc - xl) * fpx - fx) > 0.0 || fabs(2.0 * fx) > fabs(dxold * fpx)) {
      dxold = dx;
      dx = (xh - xl) * 0.5;
      xc = xl + dx;
      if (xc == xl) {
        root = xc;
        QG_NEWTON_HOOK(step);
        return true;
      }
    } else {
      dxold = dx;
      dx = fx / fpx;
      const double xold = xc;
      xc -= dx;
      if (xold == xc) {
        root = xc;
        QG_NEWTON_HOOK(step);
        return true;
      }
    }
    if (fabs(dx) < tol) {
      root = xc;
      QG_NEWTON_HOOK(step);
      return true;
    }
  }
  QG_NEWTON_HOOK(1024);
  return false;
}

// Roots on a line that is not known to be monotone (stage 0): interval test on [a,b]; bisect (left first)
// while undecided, at most 4 levels deep; Newton on the pieces.  Appends to roots[nr..].
template<int N, class F, int K>
QG_HD int root_find_general(LineEval<N, F, K> &le, double xmin, double xmax, double *roots, int nr, int *err)
{
  const F &f = le.f;
  const int k = le.dir();
  const double *x = le.x;
  double sa[kRootFindMaxLevel + 2], sb[kRootFindMaxLevel + 2];
  int sl[kRootFindMaxLevel + 2];
  int sp = 0;
  sa[0] = xmin;
  sb[0] = xmax;
  sl[0] = 0;
  sp = 1;
  while (sp > 0) {
    --sp;
    const double a = sa[sp], b = sb[sp];
    const int level = sl[sp];
    Iv<N> xi[N];
    Dl<N> dl;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      xi[d] = iv_const<N>(x[d]);
      dl.d[d] = 0.0;
    }
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if (d == k) {
        xi[d] = iv_var<N>((a + b) * 0.5, d);
        dl.d[d] = (b - a) * 0.5;
      }
    }
    const Iv<N> v = phi_val_iv<N>(f, xi, dl);
    if (iv_uniform(v, dl)) continue;
    Iv<N> g[N];
    phi_grad_iv<N>(f, xi, dl, g);
    bool mono = false;
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == k) mono = iv_uniform(g[d], dl);
    if (!mono && level < kRootFindMaxLevel) {
      const double mid = (a + b) * 0.5;
      // right half pushed first so that the left half is processed first
      sa[sp] = mid;
      sb[sp] = b;
      sl[sp] = level + 1;
      ++sp;
      sa[sp] = a;
      sb[sp] = mid;
      sl[sp] = level + 1;
      ++sp;
      continue;
    }
    double r;
    if (newton_bisection<N>(le, a, b, line_tol(a, b), r)) {
      if (nr < kMaxRoots)
        roots[nr++] = r;
      else
        *err |= ERR_ROOT_CAP;
    }
  }
  return nr;
}

// ---- stage evaluation ----------------------------------------------------------------------------------

// directions that are not free at stage s: everything except the stage directions 0..s
QG_HD int nonfree_mask(const ChainItem &it, int s, int N)
{
  int m = (1 << N) - 1;
  for (int t = 0; t <= s; ++t)
    if (it.kind[t] != ST_PASS) m &= ~(1 << it.dim[t]);
  return m;
}

template<int N> QG_HD void set_nonfree(const ChainItem &it, int mask, int sides, double *x)
{
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = (mask & (1 << d)) ? (((sides >> d) & 1) ? it.hi[d] : it.lo[d]) : x[d];
}

// register-array access with a run-time index (select chains instead of local memory)
template<int R> QG_HD void arr_put(double *a, int i, double v)
{
#pragma unroll
  for (int d = 0; d < R; ++d) a[d] = (d == i) ? v : a[d];  // selects, not an indexed store
}
template<int R> QG_HD double arr_get(const double *a, int i)
{
  double v = a[0];
#pragma unroll
  for (int d = 1; d < R; ++d)
    if (d == i) v = a[d];
  return v;
}
QG_HD void cmp_swap(double &a, double &b)
{
  if (a > b) {
    const double t = a;
    a = b;
    b = t;
  }
}

// Evaluates stage S of the chain at the partial point x (coordinates of the lower stages set).
// Writes up to CAP entries (a,b) to ent and returns their number:
//   ROOTS / PRESET: active segments [a,b];  SURF: roots a;  PASS: one dummy entry.
// A segment is active iff at its midpoint every psi has its required sign (impl_reparam_general.cpp:210-226);
// the tests run psi-major so that each psi sets its line up once (with one psi the root-finding set-up is reused).
//
// S >= 1: the lines are monotone (<= 1 root per psi, <= 2^(2-S) psi), everything lives in registers.
template<int N, int S, int CAP, int K, class F>
QG_HD int stage_eval_dir(const F &f, const ChainItem &it, int kind, int k_dyn, double xmin, double xmax,
  const double *x_in, double *ent, int *err)
{
  const int k = K >= 0 ? K : k_dyn;
  double x[N];
#pragma unroll
  for (int d = 0; d < N; ++d) x[d] = x_in[d];
  LineEval<N, F, K> le;
  const bool surf = kind == ST_SURF;
  const int mask = nonfree_mask(it, S, N);
  const int npsi = surf ? 1 : it.npsi[S];
  // top stage of a 3-D chain only: in K1 (S == 1) the extra copy of the line solver costs more than it saves; not for
  // Bezier level sets, whose unrolled de Casteljau makes every extra evaluation site expensive (C4: K2 +20 %)
  if (S >= 2 && !surf && npsi == 1 && f.k() != QG_FUNC_BEZIER) {
    // One psi on a monotone line (always the case at the top stage): at most one root r in [xmin,xmax], hence the
    // candidate segments [xmin, r] and [r, xmax] (or [xmin, xmax]) in this order -- written out without the root array,
    // its sorting network and the select chains that index it (a third of this kernel's instructions were FSEL / ISETP /
    // IMAD.MOV).  Same tests, same expressions and the same evaluation order as the general code below.
    set_nonfree<N>(it, mask, psi_sides(it.psi[S][0]), x);
    le.init(f, x, k);
    double r = xmax;
    const bool has = newton_bisection<N>(le, xmin, xmax, line_tol(xmin, xmax), r);
    const double tol = seg_tol(xmin, xmax);
    const int sg = psi_sign(it.psi[S][0]);
    const double a1 = has ? r : xmax;
    bool ok_a = !(a1 - xmin < tol);
    if (ok_a) ok_a = le.val((xmin + a1) * 0.5) > 0.0 ? (sg >= 0) : (sg <= 0);
    bool ok_b = has && !(xmax - r < tol);
    if (ok_b) ok_b = le.val((r + xmax) * 0.5) > 0.0 ? (sg >= 0) : (sg <= 0);
    // entries: A first, then B
    ent[0] = ok_a ? xmin : r;
    ent[1] = ok_a ? a1 : xmax;
    if (CAP >= 2) {
      ent[2] = r;
      ent[3] = xmax;
    }
    return (ok_a ? 1 : 0) + (ok_b ? 1 : 0);
  }
  if (S >= 2 && surf && f.k() != QG_FUNC_BEZIER) {
    // surface stage of a 3-D chain: the root of the (monotone) line itself is the node
    le.init(f, x, k);
    double r = xmax;
    if (!newton_bisection<N>(le, xmin, xmax, line_tol(xmin, xmax), r)) return 0;
    ent[0] = r;
    ent[1] = r;
    return 1;
  }
  if (S >= 1) {
    constexpr int R = (1 << (2 - (S >= 1 ? S : 1))) + 2;
    double roots[R];
#pragma unroll
    for (int i = 0; i < R; ++i) roots[i] = (double)INFINITY;
    int nr = 0;
    if (!surf) roots[nr++] = xmin;
#pragma unroll 1
    for (int i = 0; i < npsi; ++i) {
      if (!surf) set_nonfree<N>(it, mask, psi_sides(it.psi[S][i]), x);
      le.init(f, x, k);
      double r;
      if (newton_bisection<N>(le, xmin, xmax, line_tol(xmin, xmax), r)) {
        if (nr < R)
          arr_put<R>(roots, nr++, r);
        else
          *err |= ERR_ROOT_CAP;
      }
    }
    if (surf) {
      int n = 0;
#pragma unroll
      for (int i = 0; i < CAP && i < R; ++i)
        if (i < nr) {
          ent[2 * i] = roots[i];
          ent[2 * i + 1] = roots[i];
          ++n;
        }
      return n;
    }
    if (nr < R)
      arr_put<R>(roots, nr++, xmax);
    else
      *err |= ERR_ROOT_CAP;
    // ascending order (unused slots hold +inf and stay at the end)
    if (R == 3) {
      cmp_swap(roots[0], roots[1]);
      cmp_swap(roots[1], roots[R - 1]);
      cmp_swap(roots[0], roots[1]);
    } else {
      cmp_swap(roots[0], roots[1]);
      cmp_swap(roots[R - 2], roots[R - 1]);
      cmp_swap(roots[0], roots[R - 2]);
      cmp_swap(roots[1], roots[R - 1]);
      cmp_swap(roots[1], roots[R - 2]);
    }
    const double tol = seg_tol(xmin, xmax);
    unsigned bad = 0;  // bit i: segment [roots[i], roots[i+1]] is absent, degenerate or fails a sign test
#pragma unroll
    for (int i = 0; i + 1 < R; ++i)
      if (i + 1 >= nr || roots[i + 1] - roots[i] < tol) bad |= 1u << i;
#pragma unroll 1
    for (int j = 0; j < npsi; ++j) {
      if (npsi > 1) {
        set_nonfree<N>(it, mask, psi_sides(it.psi[S][j]), x);
        le.init(f, x, k);
      }
      const int sg = psi_sign(it.psi[S][j]);
#pragma unroll 1
      for (int i = 0; i + 1 < nr; ++i) {
        if (bad & (1u << i)) continue;
        const double xm = (arr_get<R>(roots, i) + arr_get<R>(roots, i + 1)) * 0.5;
        if (!(le.val(xm) > 0.0 ? (sg >= 0) : (sg <= 0))) bad |= 1u << i;
      }
    }
    int n = 0;
#pragma unroll
    for (int i = 0; i + 1 < R; ++i) {
      if (bad & (1u << i)) continue;
      if (n < CAP) {
        arr_put<2 * CAP>(ent, 2 * n, roots[i]);
        arr_put<2 * CAP>(ent, 2 * n + 1, roots[i + 1]);
        ++n;
      } else {
        *err |= ERR_ROOT_CAP;
      }
    }
    return n;
  }
  // S == 0: general lines, roots kept in a local array
  double roots[kMaxRoots];
  int nr = 0;
  roots[nr++] = xmin;
  for (int i = 0; i < npsi; ++i) {
    set_nonfree<N>(it, mask, psi_sides(it.psi[S][i]), x);
    le.init(f, x, k);
    nr = root_find_general<N>(le, xmin, xmax, roots, nr, err);
  }
  if (nr < kMaxRoots)
    roots[nr++] = xmax;
  else
    *err |= ERR_ROOT_CAP;
  // insertion sort (tiny)
  for (int i = 1; i < nr; ++i) {
    const double v = roots[i];
    int j = i - 1;
    while (j >= 0 && roots[j] > v) {
      roots[j + 1] = roots[j];
      --j;
    }
    roots[j + 1] = v;
  }
  const double tol = seg_tol(xmin, xmax);
  unsigned bad = 0;
  for (int i = 0; i + 1 < nr; ++i)
    if (roots[i + 1] - roots[i] < tol) bad |= 1u << i;
  for (int j = 0; j < npsi; ++j) {
    if (npsi > 1) {
      set_nonfree<N>(it, mask, psi_sides(it.psi[S][j]), x);
      le.init(f, x, k);
    }
    const int sg = psi_sign(it.psi[S][j]);
    for (int i = 0; i + 1 < nr; ++i) {
      if (bad & (1u << i)) continue;
      const double xm = (roots[i] + roots[i + 1]) * 0.5;
      if (!(le.val(xm) > 0.0 ? (sg >= 0) : (sg <= 0))) bad |= 1u << i;
    }
  }
  int n = 0;
  for (int i = 0; i + 1 < nr; ++i) {
    if (bad & (1u << i)) continue;
    if (n < CAP) {
      ent[2 * n] = roots[i];
      ent[2 * n + 1] = roots[i + 1];
      ++n;
    } else {
      *err |= ERR_ROOT_CAP;
    }
  }
  return n;
}

template<int N, int S, int CAP, class F>
QG_HD int stage_eval(const F &f, const ChainItem &it, const double *x_in, double *ent, int *err)
{
  const int kind = it.kind[S];
  if (kind == ST_PASS) {
    ent[0] = 0.0;
    ent[1] = 0.0;
    return 1;
  }
  const int k = it.dim[S];
  double xmin = it.lo[0], xmax = it.hi[0];
#pragma unroll
  for (int d = 1; d < N; ++d)
    if (d == k) {
      xmin = it.lo[d];
      xmax = it.hi[d];
    }
  if (kind == ST_PRESET) {
    ent[0] = xmin;
    ent[1] = xmax;
    return 1;
  }
#if defined(__CUDA_ARCH__)
  // device, top stage: one copy of the line solver per direction (no selects on the direction inside the
  // Newton loop); neighbouring stage-2 lines share their chain item, so warps rarely split here
  if (S >= 2) {
    if (k == 0) return stage_eval_dir<N, S, CAP, 0>(f, it, kind, k, xmin, xmax, x_in, ent, err);
    if (k == 1) return stage_eval_dir<N, S, CAP, 1>(f, it, kind, k, xmin, xmax, x_in, ent, err);
    return stage_eval_dir<N, S, CAP, N - 1>(f, it, kind, k, xmin, xmax, x_in, ent, err);
  }
#endif
  return stage_eval_dir<N, S, CAP, -1>(f, it, kind, k, xmin, xmax, x_in, ent, err);
}

QG_HD int stage_count(int kind, int nent, int p)
{
  if (kind == ST_PASS) return 1;
  if (kind == ST_SURF) return nent;
  return nent * p;
}

// r-th child of a stage evaluation of the given kind along direction k: sets x[k] and multiplies the weight.
// g_out (may be null): receives the phi gradient at x when the stage is a surface stage (returns true then).
template<int N, class F>
QG_HD bool stage_child_k(const F &f, int kind, int k, const double *ent, int r, int p, const double *gxw, double *x,
  double &w, double *g_out)
{
  if (kind == ST_PASS) return false;
  if (kind == ST_SURF) {
    const double root = ent[2 * r];
#pragma unroll
    for (int d = 0; d < N; ++d)
      if (d == k) x[d] = root;
    double g[N];
    phi_grad<N>(f, x, g);
    double ss = g[0] * g[0];
    double gk = g[0];
#pragma unroll
    for (int d = 1; d < N; ++d) {
      ss += g[d] * g[d];
      if (d == k) gk = g[d];
    }
    w = sqrt(ss) / fabs(gk) * w;
    if (g_out) {
#pragma unroll
      for (int d = 0; d < N; ++d) g_out[d] = g[d];
    }
    return true;
  }
  const int sg = r / p, j = r - sg * p;
  const double x0 = ent[2 * sg], x1 = ent[2 * sg + 1];
  const double xv = x0 + (x1 - x0) * gxw[2 * j];
  const double gw = (x1 - x0) * gxw[2 * j + 1];
#pragma unroll
  for (int d = 0; d < N; ++d)
    if (d == k) x[d] = xv;
  w = w * gw;
  return false;
}
template<int N, class F>
QG_HD void stage_child(const F &f, const ChainItem &it, int s, const double *ent, int r, int p,
  const double *gxw, double *x, double &w)
{
  stage_child_k<N>(f, it.kind[s], it.dim[s], ent, r, p, gxw, x, w, (double *)0);
}

// ---- CTL: the bisection / height-direction walk ---------------------------------------------------

template<int N> QG_HD void frame_xint(const Frame &fr, Iv<N> *xint, Dl<N> &dl)
{
#pragma unroll
  for (int d = 0; d < N; ++d) {
    if (fr.freemask & (1 << d)) {
      xint[d] = iv_var<N>((fr.lo[d] + fr.hi[d]) * 0.5, d);
      dl.d[d] = (fr.hi[d] - fr.lo[d]) * 0.5;
    } else {
      xint[d] = iv_const<N>(0.0);
      dl.d[d] = 0.0;
    }
  }
}
template<int N> QG_HD void frame_set_sides(const Frame &fr, int sides, Iv<N> *xint)
{
#pragma unroll
  for (int d = 0; d < N; ++d)
    if (!(fr.freemask & (1 << d))) xint[d].a = ((sides >> d) & 1) ? fr.hi[d] : fr.lo[d];
}

QG_HD void determine_signs(bool S, bool positive_above, int sign, int &bottom, int &top)
{
  if (S) {
    bottom = positive_above ? -1 : 1;
    top = -bottom;
  } else if (sign == 1) {
    bottom = positive_above ? 0 : 1;
    top = positive_above ? 1 : 0;
  } else if (sign == -1) {
    bottom = positive_above ? -1 : 0;
    top = positive_above ? 0 : -1;
  } else {
    bottom = 0;
    top = 0;
  }
}

template<int N> QG_HD void emit_item(const Frame &fr, int job, int mtop, bool full, ChainItem *out)
{
  ChainItem it;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    it.lo[d] = d < N ? fr.lo[d] : 0.0;
    it.hi[d] = d < N ? fr.hi[d] : 0.0;
  }
  it.job = job;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    it.dim[s] = 255;
    it.kind[s] = ST_PASS;
    it.npsi[s] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) it.psi[s][i] = psi_pack(0, 0);
  }
  it.pad[0] = it.pad[1] = it.pad[2] = 0;
  // stages fixed by the ancestors: levels fr.M+1 .. mtop
  for (int s = fr.M; s < mtop; ++s) {
    it.dim[s] = fr.sdim[s];
    it.kind[s] = (fr.surf && s == mtop - 1) ? ST_SURF : ST_ROOTS;
    it.npsi[s] = fr.snpsi[s];
    for (int i = 0; i < 4; ++i) it.psi[s][i] = fr.spsi[s][i];
  }
  if (full) {
    // tensor-product rule over the free directions, ascending (outermost first)
    int s = 0;
    for (int d = 0; d < N; ++d)
      if (fr.freemask & (1 << d)) {
        it.dim[s] = (unsigned char)d;
        it.kind[s] = ST_PRESET;
        ++s;
      }
  } else {
    // level 1: the remaining free direction, with this frame's codes
    int e0 = 0;
    for (int d = 0; d < N; ++d)
      if (fr.freemask & (1 << d)) e0 = d;
    it.dim[0] = (unsigned char)e0;
    it.kind[0] = ST_ROOTS;
    it.npsi[0] = fr.npsi;
    for (int i = 0; i < 4; ++i) it.psi[0][i] = fr.psi[i];
  }
  *out = it;
}

// Walks one job.  If items != nullptr writes the chain items (at most cap), returns their number.
template<int N, class F>
QG_HDN int ctl_walk(const F &f, const double *clo, const double *chi, int jtype, int job, ChainItem *items,
  int cap, int *err, bool *hit_depth_cap = nullptr)
{
  constexpr int MAXPSI = 1 << (N - 1);
  Frame st[kStackCap];
  int sp = 0;
  int mtop;
  {
    Frame &fr = st[0];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      fr.lo[d] = d < N ? clo[d] : 0.0;
      fr.hi[d] = d < N ? chi[d] : 0.0;
    }
    fr.level = 0;
    fr.npsi = 1;
    for (int i = 0; i < 4; ++i) fr.psi[i] = psi_pack(0, kPsiDefaultSign);
    for (int s = 0; s < 3; ++s) {
      fr.sdim[s] = 255;
      fr.snpsi[s] = 0;
      for (int i = 0; i < 4; ++i) fr.spsi[s][i] = psi_pack(0, 0);
    }
    if (jtype >= 0) {
      const int cd = jtype >> 1, side = jtype & 1;
      fr.M = N - 1;
      fr.freemask = (unsigned char)(((1 << N) - 1) & ~(1 << cd));
      fr.psi[0] = psi_pack(side << cd, -1);
      fr.S = 0;
      fr.surf = 0;
    } else {
      fr.M = N;
      fr.freemask = (unsigned char)((1 << N) - 1);
      fr.psi[0] = psi_pack(0, -1);
      fr.S = (jtype == JOB_SURFACE) ? 1 : 0;
      fr.surf = fr.S;
    }
    mtop = fr.M;
    sp = 1;
  }
  int nitems = 0;
  while (sp > 0) {
    Frame fr = st[--sp];
    if (fr.M == 1) {
      if (items) {
        if (nitems < cap)
          emit_item<N>(fr, job, mtop, false, items + nitems);
        else
          *err |= ERR_ITEM_CAP;
      }
      ++nitems;
      continue;
    }
    Iv<N> xint[N];
    Dl<N> dl;
    frame_xint<N>(fr, xint, dl);
    // prune (impl_reparam_general.cpp:107-130)
    bool empty = false;
    for (int i = 0; i < fr.npsi;) {
      frame_set_sides<N>(fr, psi_sides(fr.psi[i]), xint);
      const Iv<N> res = phi_val_iv<N>(f, xint, dl);
      if (iv_uniform(res, dl)) {
        const int sg = psi_sign(fr.psi[i]);
        if ((res.a >= 0.0 && sg >= 0) || (res.a <= 0.0 && sg <= 0)) {
          --fr.npsi;
          const unsigned char t = fr.psi[i];
          fr.psi[i] = fr.psi[fr.npsi];
          fr.psi[fr.npsi] = t;
        } else {
          empty = true;
          break;
        }
      } else {
        ++i;
      }
    }
    if (empty) continue;
    if (fr.npsi == 0) {
      if (!fr.S) {
        if (items) {
          if (nitems < cap)
            emit_item<N>(fr, job, mtop, true, items + nitems);
          else
            *err |= ERR_ITEM_CAP;
        }
        ++nitems;
      }
      continue;
    }
    // height direction (impl_reparam_general.cpp:704-725)
    int e0 = -1;
    double max_quan = 0.0;
    frame_set_sides<N>(fr, psi_sides(fr.psi[0]), xint);
    Iv<N> g[N];
    phi_grad_iv<N>(f, xint, dl, g);
    double g_e0_alpha = 0.0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      if ((fr.freemask & (1 << d)) && fabs(g[d].a) > 1.001 * iv_dev(g[d], dl)) {
        const double quan = fabs(g[d].a) * (fr.hi[d] - fr.lo[d]);
        if (quan > max_quan) {
          max_quan = quan;
          e0 = d;
          g_e0_alpha = g[d].a;
        }
      }
    }
    unsigned char newpsi[4];
    for (int i = 0; i < 4; ++i) newpsi[i] = psi_pack(0, kPsiDefaultSign);
    int nnew = 0;
    bool done = false;
    for (int i = 0; i < fr.npsi && !done; ++i) {
      bool direction_okay = e0 != -1;
      if (direction_okay) {
        frame_set_sides<N>(fr, psi_sides(fr.psi[i]), xint);
        Iv<N> gi[N];
        phi_grad_iv<N>(f, xint, dl, gi);
        bool u = false;
#pragma unroll
        for (int d = 0; d < N; ++d)
          if (d == e0) u = iv_uniform(gi[d], dl);
        direction_okay = u;
      }
      if (!direction_okay) {
        if (fr.level < kMaxSubdiv) {
          double max_ext = 0.0;
          int ind = -1;
#pragma unroll
          for (int d = 0; d < N; ++d) {
            if (fr.freemask & (1 << d)) {
              const double ext = fr.hi[d] - fr.lo[d];
              if (ext > max_ext) {
                max_ext = ext;
                ind = d;
              }
            }
          }
          double xmid = 0.0;
#pragma unroll
          for (int d = 0; d < N; ++d)
            if (d == ind) xmid = (fr.lo[d] + fr.hi[d]) * 0.5;
          if (sp + 2 > kStackCap) {
            *err |= ERR_STACK;
          } else {
            Frame b1 = fr;
            b1.level = fr.level + 1;
            Frame b0 = b1;
#pragma unroll
            for (int d = 0; d < N; ++d)
              if (d == ind) {
                b0.hi[d] = xmid;
                b1.lo[d] = xmid;
              }
            st[sp++] = b1;  // second half: processed after the first
            st[sp++] = b0;
          }
        } else {
          // too deep: midpoint sign test against every code (impl_reparam_general.cpp:776-800)
          if (hit_depth_cap) *hit_depth_cap = true;
          double xp[N];
#pragma unroll
          for (int d = 0; d < N; ++d) xp[d] = (fr.lo[d] + fr.hi[d]) * 0.5;
          bool okay = true;
          for (int j = 0; j < MAXPSI && okay; ++j) {
            const int sides = psi_sides(fr.psi[j]);
#pragma unroll
            for (int d = 0; d < N; ++d)
              if (!(fr.freemask & (1 << d))) xp[d] = ((sides >> d) & 1) ? fr.hi[d] : fr.lo[d];
            const int sg = psi_sign(fr.psi[j]);
            okay = okay && (phi_val<N>(f, xp) >= 0.0 ? (sg >= 0) : (sg <= 0));
          }
          if (okay && !fr.S) {
            if (items) {
              if (nitems < cap)
                emit_item<N>(fr, job, mtop, true, items + nitems);
              else
                *err |= ERR_ITEM_CAP;
            }
            ++nitems;
          }
        }
        done = true;
        break;
      }
      int bottom, top;
      determine_signs(fr.S != 0, g_e0_alpha > 0.0, psi_sign(fr.psi[i]), bottom, top);
      const int sides = psi_sides(fr.psi[i]);
      newpsi[nnew++] = psi_pack(sides & ~(1 << e0), bottom);
      newpsi[nnew++] = psi_pack(sides | (1 << e0), top);
    }
    if (done) continue;
    // dimension reduction: level M-1 on the same box, remembering this level as stage M-1
    Frame ch = fr;
    ch.M = fr.M - 1;
    ch.freemask = (unsigned char)(fr.freemask & ~(1 << e0));
    ch.level = 0;
    ch.S = 0;
    ch.npsi = (unsigned char)nnew;
    for (int i = 0; i < 4; ++i) ch.psi[i] = newpsi[i];
    const int s = fr.M - 1;
    ch.sdim[s] = (unsigned char)e0;
    ch.snpsi[s] = fr.npsi;
    for (int i = 0; i < 4; ++i) ch.spsi[s][i] = fr.psi[i];
    if (sp + 1 > kStackCap)
      *err |= ERR_STACK;
    else
      st[sp++] = ch;
  }
  return nitems;
}

}  // namespace qg
