// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of Algoim's general (Saye 2015) dimension-reduction quadrature,
// `algoim::ImplicitIntegral<M,N,F,Q,S>` + `quadGen`, as the reference calls it at
//   /root/reference/cpp/qugar/impl_quadrature.cpp:138-158 (interior / surface rule),
//   /root/reference/cpp/qugar/impl_quadrature.cpp:392-428 (facet rule),
//   /root/reference/cpp/qugar/impl_unfitted_domain.cpp:183-190, 336-349 (classification, n=1).
// Algoim itself (gh:dyc0/algoim@1.0.0) is not in /root/reference; the control flow below
// follows the in-tree modification of the same class,
//   /root/reference/cpp/qugar/impl_reparam_general.cpp:107-130 (prune), :653-822 (constructor),
//   :183-199 (rootFind call), :210-226 (active interval test), :155-160,251 (degenerate tolerance),
// and the published algorithm for the parts that file does not show (1-D evalIntegrand,
// tensor-product rule, rootFind, Newton-bisection, determineSigns).
#pragma once
#include "oracle_funcs.hpp"
#include <algorithm>
#include <vector>
#include <cstdio>

namespace orc {

extern const real g_gauss_table[];  // (x, w) pairs; rule n starts at pair n(n-1)/2
inline real gauss_x(int p, int i) { return g_gauss_table[2 * (p * (p - 1) / 2 + i)]; }
inline real gauss_w(int p, int i) { return g_gauss_table[2 * (p * (p - 1) / 2 + i) + 1]; }

template<int N> struct Box
{
  real lo[N], hi[N];
  real side(int s, int d) const { return s == 0 ? lo[d] : hi[d]; }
  real extent(int d) const { return hi[d] - lo[d]; }
  real midpoint(int d) const { return (lo[d] + hi[d]) * 0.5; }
};

// algoim::PsiCode<N>: which side of each non-free direction phi is restricted to, and the
// sign required of it there (-1: phi<0, +1: phi>0, 0: unconstrained).
template<int N> struct PsiCode
{
  unsigned char side[N];
  int sign;
};

// Q in ImplicitIntegral<..., Q, ...>: receives (x, w).
template<int N> struct Target
{
  virtual void evalIntegrand(const real *x, real w) = 0;
  virtual ~Target() {}
};

// ---- algoim::detail::newtonBisectionSearch -------------------------------------------------
// Newton's method safeguarded by bisection on a bracket with a sign change (f(xl) < 0 side kept).
template<typename F>
inline bool newton_bisection(const F &f, real x0, real x1, real tol, int maxsteps, real &root)
{
  real f0, f1, d;
  f(x0, f0, d);
  f(x1, f1, d);
  if ((f0 > 0.0 && f1 > 0.0) || (f0 < 0.0 && f1 < 0.0)) return false;
  if (f0 == 0.0) {
    root = x0;
    return true;
  }
  if (f1 == 0.0) {
    root = x1;
    return true;
  }
  real x = (x0 + x1) * 0.5;
  real xl, xh;
  if (f0 < 0.0) {
    xl = x0;
    xh = x1;
  } else {
    xl = x1;
    xh = x0;
  }
  real fx, fpx;
  f(x, fx, fpx);
  real dxold = std::fabs(x1 - x0);
  real dx = dxold;
  for (int step = 0; step < maxsteps; ++step) {
    if (((x - xh) * fpx - fx) * ((x - xl) * fpx - fx) > 0.0 || std::fabs(2.0 * fx) > std::fabs(dxold * fpx)) {
      dxold = dx;
      dx = (xh - xl) * 0.5;
      x = xl + dx;
      if (x == xl) {
        root = x;
        return true;
      }
    } else {
      dxold = dx;
      dx = fx / fpx;
      const real xold = x;
      x -= dx;
      if (xold == x) {
        root = x;
        return true;
      }
    }
    if (std::fabs(dx) < tol) {
      root = x;
      return true;
    }
    f(x, fx, fpx);
    if (fx < 0.0)
      xl = x;
    else
      xh = x;
  }
  return false;
}

// Absolute tolerance used on a line [xmin,xmax]: factor * eps * scale.  knobs().tol_scale selects the scale:
//   0: the segment length (unreachable when the box is far from the origin: below one ulp of x);
//   1: max(|xmin|, |xmax|, length), i.e. a few ulps of the coordinates themselves (default).
inline real line_tol(real factor, real xmin, real xmax, int mode)
{
  real scale = xmax - xmin;
  if (mode == 1) {
    const real a = std::fabs(xmin), b = std::fabs(xmax);
    if (a > scale) scale = a;
    if (b > scale) scale = b;
  }
  if (mode == 2 && scale < 1.0) scale = 1.0;  // Tolerance().update(10 eps x extent) keeps the larger value (tolerance.cpp:52-55)
  return factor * EPS * scale;
}

// ---- algoim::detail::rootFind ----------------------------------------------------------------
// Roots of phi along direction k through x on [xmin,xmax], appended to roots.  When the caller
// knows the restriction is monotone (height direction of a level M > 1,
// impl_reparam_general.cpp:189-195 `dim > 1`), a bracket test + Newton-bisection suffices; otherwise
// interval arithmetic excludes sign-definite segments, certifies monotone ones, and bisects the rest.
template<int N>
void root_find(const Func &phi, const real *x, int k, real xmin, real xmax, std::vector<real> &roots, bool monotone,
  int level = 0)
{
  auto f = [&](real t, real &val, real &der) {
    real xx[N];
    for (int i = 0; i < N; ++i) xx[i] = x[i];
    xx[k] = t;
    val = eval<N, real>(phi, xx);
    real g[N];
    grad<N, real>(phi, xx, g);
    der = g[k];
  };
  if (!monotone) {
    Interval<N> xint[N];
    for (int i = 0; i < N; ++i) {
      xint[i] = Interval<N>(x[i]);
      Interval<N>::delta(i) = 0.0;
    }
    xint[k] = Interval<N>::variable((xmin + xmax) * 0.5, k);
    Interval<N>::delta(k) = (xmax - xmin) * 0.5;
    const Interval<N> v = eval<N, Interval<N>>(phi, xint);
    if (knobs().rf_use_sign && v.uniformSign()) return;
    Interval<N> g[N];
    grad<N, Interval<N>>(phi, xint, g);
    if (!g[k].uniformSign()) {
      if (level < knobs().rootfind_maxlevel) {
        const real xmid = (xmin + xmax) * 0.5;
        root_find<N>(phi, x, k, xmin, xmid, roots, false, level + 1);
        root_find<N>(phi, x, k, xmid, xmax, roots, false, level + 1);
        return;
      }
      // too deep: fall through to the bracketed search on this small segment
      if (knobs().rf_cap_mode == 1) { roots.push_back((xmin + xmax) * 0.5); return; }
      if (knobs().rf_cap_mode == 2) return;
    }
  }
  real root;
  const real tol = line_tol(knobs().newton_tol_factor, xmin, xmax, knobs().tol_scale);
  if (newton_bisection(f, xmin, xmax, tol, knobs().newton_maxsteps, root)) roots.push_back(root);
}

// ---- algoim::detail::determineSigns<S> ---------------------------------------------------------
inline void determine_signs(bool S, bool positive_above, int sign, int &bottom, int &top)
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
    bottom = top = 0;
  }
}

// Per-cell work counters (used by DESIGN.md's cost model and by the tests' structure checks).
struct Stats
{
  long long n_point_evals = 0, n_interval_evals = 0, n_bisections = 0, n_leaves = 0, max_level = 0;
};

// ---- algoim::ImplicitIntegral<M,N,F,Q,S> --------------------------------------------------------
inline int &trace_on() { static int t = 0; return t; }
#define ORC_TRACE(...) do { if (trace_on()) fprintf(stderr, __VA_ARGS__); } while (0)

template<int N> struct Integral : Target<N>
{
  static const int MAXPSI = 1 << (N - 1);
  static const int MAX_SUBDIV = 16;  // impl_reparam_general.cpp:72

  const Func &phi;
  Target<N> &f;
  int M;
  bool S;
  bool free_[N];
  PsiCode<N> psi[MAXPSI];
  int psiCount;
  Box<N> xrange;
  int p;
  int e0;
  std::vector<real> roots;

  void set_nonfree(const PsiCode<N> &c, real *x) const
  {
    for (int d = 0; d < N; ++d)
      if (!free_[d]) x[d] = xrange.side(c.side[d], d);
  }

  bool prune(Interval<N> *xint)
  {
    for (int i = 0; i < psiCount;) {
      for (int d = 0; d < N; ++d)
        if (!free_[d]) xint[d].alpha = xrange.side(psi[i].side[d], d);
      const Interval<N> res = eval<N, Interval<N>>(phi, xint);
      if (res.uniformSign()) {
        if ((res.alpha >= 0.0 && psi[i].sign >= 0) || (res.alpha <= 0.0 && psi[i].sign <= 0)) {
          --psiCount;
          std::swap(psi[i], psi[psiCount]);
        } else {
          return false;
        }
      } else {
        ++i;
      }
    }
    return true;
  }

  // Gaussian rule over the free directions of the whole box (MultiLoop: last free index fastest).
  void tensor_product_integral()
  {
    int idx[N];
    int fd[N], nf = 0;
    for (int d = 0; d < N; ++d)
      if (free_[d]) fd[nf++] = d;
    for (int i = 0; i < nf; ++i) idx[i] = 0;
    while (true) {
      real x[N];
      for (int d = 0; d < N; ++d) x[d] = 0.0;
      // non-free coordinates: the facet's own side (quadGen fixes them for the caller)
      set_nonfree(psi[0], x);
      real w = 1.0;
      for (int i = 0; i < nf; ++i) {
        const int d = fd[i];
        x[d] = xrange.lo[d] + xrange.extent(d) * gauss_x(p, idx[i]);
        w *= xrange.extent(d) * gauss_w(p, idx[i]);
      }
      f.evalIntegrand(x, w);
      int i = nf - 1;
      for (; i >= 0; --i) {
        if (++idx[i] < p) break;
        idx[i] = 0;
      }
      if (i < 0) break;
    }
  }

  // x is valid in all free variables barring e0.
  void evalIntegrand(const real *x_, real w) override
  {
    real x[N];
    for (int d = 0; d < N; ++d) x[d] = x_[d];
    const real xmin = xrange.lo[e0], xmax = xrange.hi[e0];
    if (S) {
      // surface integral, M == N: every root of phi on the line is a node, weight |grad|/|d_e0 phi|
      roots.clear();
      root_find<N>(phi, x, e0, xmin, xmax, roots, true);
      if (trace_on()) { fprintf(stderr, "      surf line e0=%d x=", e0); for (int d = 0; d < N; ++d) fprintf(stderr, "%.17g ", x[d]); fprintf(stderr, "roots=%zu\n", roots.size()); }
      for (size_t i = 0; i < roots.size(); ++i) {
        x[e0] = roots[i];
        real g[N];
        grad<N, real>(phi, x, g);
        real s = g[0] * g[0];
        for (int d = 1; d < N; ++d) s += g[d] * g[d];
        f.evalIntegrand(x, std::sqrt(s) / std::fabs(g[e0]) * w);
      }
      return;
    }
    roots.clear();
    roots.push_back(xmin);
    for (int i = 0; i < psiCount; ++i) {
      set_nonfree(psi[i], x);
      root_find<N>(phi, x, e0, xmin, xmax, roots, M > 1);
    }
    roots.push_back(xmax);
    std::sort(roots.begin(), roots.end());
    if (trace_on()) { fprintf(stderr, "      line M=%d e0=%d x=", M, e0); for (int d = 0; d < N; ++d) fprintf(stderr, "%.17g ", x[d]); fprintf(stderr, "R:"); for (real r : roots) fprintf(stderr, " %.17g", r); fprintf(stderr, "\n"); }
    // degenerate segments are filtered with tolerance 10 eps * extent (impl_reparam_general.cpp:155-160)
    const real tol = line_tol(knobs().seg_tol_factor, xmin, xmax, knobs().seg_tol_scale);
    for (size_t i = 0; i + 1 < roots.size(); ++i) {
      const real x0 = roots[i], x1 = roots[i + 1];
      if (knobs().seg_le ? (x1 - x0 <= tol) : (x1 - x0 < tol)) continue;
      bool okay = true;
      x[e0] = (x0 + x1) * 0.5;
      for (int j = 0; j < psiCount && okay; ++j) {
        set_nonfree(psi[j], x);
        okay &= eval<N, real>(phi, x) > 0.0 ? (psi[j].sign >= 0) : (psi[j].sign <= 0);
      }
      ORC_TRACE("        seg [%.17g,%.17g] %s\n", x0, x1, okay ? "active" : "inactive");
      if (!okay) continue;
      for (int j = 0; j < p; ++j) {
        x[e0] = x0 + (x1 - x0) * gauss_x(p, j);
        const real gw = (x1 - x0) * gauss_w(p, j);
        f.evalIntegrand(x, w * gw);
      }
    }
  }

  Integral(const Func &phi_, Target<N> &f_, int M_, bool S_, const bool *free, const PsiCode<N> *psi_, int psiCount_,
    const Box<N> &xrange_, int p_, Stats *st, int level = 0)
    : phi(phi_), f(f_), M(M_), S(S_), psiCount(psiCount_), xrange(xrange_), p(p_), e0(-1)
  {
    for (int d = 0; d < N; ++d) free_[d] = free[d];
    for (int i = 0; i < MAXPSI; ++i) psi[i] = psi_[i];
    if (st && level > st->max_level) st->max_level = level;

    if (M == 1) {
      for (int d = 0; d < N; ++d)
        if (free_[d]) e0 = d;
      real x0[N];
      for (int d = 0; d < N; ++d) x0[d] = 0.0;
      if (st) st->n_leaves++;
      evalIntegrand(x0, 1.0);
      return;
    }

    Interval<N> xint[N];
    for (int d = 0; d < N; ++d) {
      if (free_[d]) {
        xint[d] = Interval<N>::variable(xrange.midpoint(d), d);
        Interval<N>::delta(d) = xrange.extent(d) * 0.5;
      } else {
        xint[d] = Interval<N>(0.0);
        Interval<N>::delta(d) = 0.0;
      }
    }

    if (trace_on()) {
      fprintf(stderr, "%*sctor M=%d S=%d lvl=%d box", level * 2, "", M, (int)S, level);
      for (int d = 0; d < N; ++d) fprintf(stderr, " [%.6g,%.6g]%s", xrange.lo[d], xrange.hi[d], free_[d] ? "" : "x");
      fprintf(stderr, " psi=%d\n", psiCount);
    }
    if (!prune(xint)) { ORC_TRACE("%*s -> pruned empty\n", level * 2, ""); return; }

    if (psiCount == 0) {
      ORC_TRACE("%*s -> all psi pruned (TP if !S)\n", level * 2, "");
      if (!S) {
        if (st) st->n_leaves++;
        tensor_product_integral();
      }
      return;
    }

    // height direction (impl_reparam_general.cpp:704-725)
    real max_quan = 0.0;
    for (int d = 0; d < N; ++d)
      if (!free_[d]) xint[d].alpha = xrange.side(psi[0].side[d], d);
    Interval<N> g[N];
    grad<N, Interval<N>>(phi, xint, g);
    for (int d = 0; d < N; ++d) {
      if (free_[d] && std::fabs(g[d].alpha) > 1.001 * g[d].maxDeviation()) {
        const real quan = std::fabs(g[d].alpha) * xrange.extent(d);
        if (quan > max_quan) {
          max_quan = quan;
          e0 = d;
        }
      }
    }

    PsiCode<N> newPsi[MAXPSI];
    int newPsiCount = 0;
    for (int i = 0; i < psiCount; ++i) {
      for (int d = 0; d < N; ++d)
        if (!free_[d]) xint[d].alpha = xrange.side(psi[i].side[d], d);
      bool direction_okay = e0 != -1;
      if (direction_okay) {
        Interval<N> gi[N];
        grad<N, Interval<N>>(phi, xint, gi);
        direction_okay = gi[e0].uniformSign();
      }
      if (!direction_okay) {
        if (level < MAX_SUBDIV) {
          real max_ext = 0.0;
          int ind = -1;
          for (int d = 0; d < N; ++d) {
            if (free_[d]) {
              const real ext = xrange.extent(d);
              if (ext > max_ext) {
                max_ext = ext;
                ind = d;
              }
            }
          }
          const real xmid = xrange.midpoint(ind);
          ORC_TRACE("%*s -> bisect dir %d (e0=%d psi %d)\n", level * 2, "", ind, e0, i);
          if (st) st->n_bisections++;
          Box<N> b0 = xrange, b1 = xrange;
          b0.hi[ind] = xmid;
          b1.lo[ind] = xmid;
          { Integral<N> sub(phi, f, M, S, free_, psi, psiCount, b0, p, st, level + 1); }
          // the two halves are independent; Interval::delta is re-established by each constructor
          { Integral<N> sub(phi, f, M, S, free_, psi, psiCount, b1, p, st, level + 1); }
          return;
        } else {
          // subdivided too deep: midpoint sign test against all 2^(N-1) codes (impl_reparam_general.cpp:776-800)
          real xp[N];
          for (int d = 0; d < N; ++d) xp[d] = xrange.midpoint(d);
          bool okay = true;
          for (int j = 0; j < MAXPSI && okay; ++j) {
            for (int d = 0; d < N; ++d)
              if (!free_[d]) xp[d] = xrange.side(psi[j].side[d], d);
            okay &= eval<N, real>(phi, xp) >= 0.0 ? (psi[j].sign >= 0) : (psi[j].sign <= 0);
          }
          if (okay && !S) {
            if (st) st->n_leaves++;
            tensor_product_integral();
          }
          return;
        }
      }
      int bottom, top;
      determine_signs(S, g[e0].alpha > 0.0, psi[i].sign, bottom, top);
      newPsi[newPsiCount] = psi[i];
      newPsi[newPsiCount].side[e0] = 0;
      newPsi[newPsiCount].sign = bottom;
      ++newPsiCount;
      newPsi[newPsiCount] = psi[i];
      newPsi[newPsiCount].side[e0] = 1;
      newPsi[newPsiCount].sign = top;
      ++newPsiCount;
    }

    // dimension reduction: the (M-1)-dimensional integral feeds this level's evalIntegrand
    bool nfree[N];
    for (int d = 0; d < N; ++d) nfree[d] = free_[d];
    nfree[e0] = false;
    // slots the loop above did not fill stay as a default-constructed PsiCode would be
    for (int i = 0; i < MAXPSI; ++i) {
      if (i < newPsiCount) continue;
      for (int d = 0; d < N; ++d) newPsi[i].side[d] = 0;
      newPsi[i].sign = knobs().psi_default_sign;
    }
    ORC_TRACE("%*s -> height dir %d, %d new psi\n", level * 2, "", e0, newPsiCount);
    Integral<N> sub(phi, *this, M - 1, false, nfree, newPsi, newPsiCount, xrange, p, st);
  }
};

// ---- algoim::QuadratureRule<N> + quadGen ---------------------------------------------------------
template<int N> struct QuadRule : Target<N>
{
  std::vector<real> x;  // N per node
  std::vector<real> w;
  void evalIntegrand(const real *xx, real ww) override
  {
    for (int d = 0; d < N; ++d) x.push_back(xx[d]);
    w.push_back(ww);
  }
  real sumWeights() const
  {
    real s = 0.0;
    for (size_t i = 0; i < w.size(); ++i) s += w[i];
    return s;
  }
  size_t size() const { return w.size(); }
};

// dim < 0: volume {phi<0}; dim == N: surface {phi=0}; 0<=dim<N: face x_dim = side of the box.
template<int N> void quad_gen(const Func &phi, const Box<N> &box, int dim, int side, int qo, Target<N> &q, Stats *st = 0)
{
  PsiCode<N> psi[1 << (N - 1)];
  for (int i = 0; i < (1 << (N - 1)); ++i) {
    for (int d = 0; d < N; ++d) psi[i].side[d] = 0;
    psi[i].sign = knobs().psi_default_sign;
  }
  psi[0].sign = -1;
  bool free[N];
  for (int d = 0; d < N; ++d) free[d] = true;
  if (dim >= 0 && dim < N) {
    psi[0].side[dim] = (unsigned char)side;
    free[dim] = false;
    Integral<N> I(phi, q, N - 1, false, free, psi, 1, box, qo, st);
  } else if (dim == N) {
    Integral<N> I(phi, q, N, true, free, psi, 1, box, qo, st);
  } else {
    Integral<N> I(phi, q, N, false, free, psi, 1, box, qo, st);
  }
}

}  // namespace orc
