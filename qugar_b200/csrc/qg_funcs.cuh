// qugar_b200 -- level-set functions evaluated on double and on Iv<N> (value and gradient).
//
// One formula per function, written against a small algebra (AlgD: doubles, AlgI: intervals) so
// that both evaluations keep the reference's expression order:
//   Sphere/Disk  phi = |x-c|^2 - r^2          /root/reference/cpp/qugar/primitive_funcs_lib.cpp:66-74
//   Plane/Line   phi = (x-o).n                /root/reference/cpp/qugar/primitive_funcs_lib.cpp:849-857
//   TPMS family  Schoen, SchoenIWP, SchoenFRD, FischerKochS, SchwarzDiamond, SchwarzPrimitive
//                                             /root/reference/cpp/qugar/tpms_lib.cpp:83-108,137-166,205-240,283-322,360-391,430-452
//   Negative     -phi                         /root/reference/cpp/qugar/impl_funcs_lib.cpp:207-227
// 2-D TPMS are the 3-D formula at fixed z with period 1 in z (tpms_lib.hpp:41-62).
#pragma once
#include "qg_math.cuh"

namespace qg {

enum FuncKind {
  QG_FUNC_SPHERE = 0,
  QG_FUNC_PLANE = 1,
  QG_FUNC_SCHOEN = 10,
  QG_FUNC_SCHOEN_IWP = 11,
  QG_FUNC_SCHOEN_FRD = 12,
  QG_FUNC_FISCHER_KOCH_S = 13,
  QG_FUNC_SCHWARZ_D = 14,
  QG_FUNC_SCHWARZ_P = 15
};

// params: SPHERE {r, c[dim]}; PLANE {o[dim], n[dim]}; TPMS 3-D {m,n,q}; TPMS 2-D {m,n,z}
struct FuncDesc
{
  int kind;
  int dim;
  int negative;
  int pad;
  double p[8];
};

template<int N> struct AlgD
{
  typedef double T;
  QG_HD T cst(double c) const { return c; }
  QG_HD T add(T a, T b) const { return a + b; }
  QG_HD T sub(T a, T b) const { return a - b; }
  QG_HD T neg(T a) const { return -a; }
  QG_HD T subc(T a, double c) const { return a - c; }
  QG_HD T scale(double c, T a) const { return a * c; }
  QG_HD T mul(T a, T b) const { return a * b; }
  QG_HD void sincos(T a, T &s, T &c) const { sincos_det(a, s, c); }
};

template<int N> struct AlgI
{
  typedef Iv<N> T;
  Dl<N> dl;
  QG_HD T cst(double c) const { return iv_const<N>(c); }
  QG_HD T add(const T &a, const T &b) const { return iv_add(a, b); }
  QG_HD T sub(const T &a, const T &b) const { return iv_sub(a, b); }
  QG_HD T neg(const T &a) const { return iv_neg(a); }
  QG_HD T subc(const T &a, double c) const { return iv_subc(a, c); }
  QG_HD T scale(double c, const T &a) const { return iv_scale(c, a); }
  QG_HD T mul(const T &a, const T &b) const { return iv_mul(a, b, dl); }
  QG_HD void sincos(const T &a, T &s, T &c) const { iv_sincos(a, dl, s, c); }
};

// TPMS arguments.  eval_style: (2 pi) * (m * x)  [Schoen, SchoenIWP values];  otherwise ((2 pi) m) * x.
template<int N, class A>
QG_HD void tpms_args(const A &alg, const FuncDesc &f, const typename A::T *x, bool eval_style, typename A::T *a,
  double *pm)
{
  const double two_pi = 2.0 * kPi;
  double m[3];
  typename A::T xx[3];
  xx[0] = x[0];
  xx[1] = x[1];
  if (N == 3) {
    m[0] = f.p[0];
    m[1] = f.p[1];
    m[2] = f.p[2];
    xx[2] = x[N - 1];
  } else {
    m[0] = f.p[0];
    m[1] = f.p[1];
    m[2] = 1.0;
    xx[2] = alg.cst(f.p[2]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pm[i] = two_pi * m[i];
    a[i] = eval_style ? alg.scale(two_pi, alg.scale(m[i], xx[i])) : alg.scale(pm[i], xx[i]);
  }
}

template<int N, class A> QG_HDN typename A::T phi_eval_raw(const A &alg, const FuncDesc &f, const typename A::T *x)
{
  typedef typename A::T T;
  switch (f.kind) {
  case QG_FUNC_SPHERE: {
    T d0 = alg.subc(x[0], f.p[1]);
    T s = alg.mul(d0, d0);
#pragma unroll
    for (int i = 1; i < N; ++i) {
      T di = alg.subc(x[i], f.p[1 + i]);
      s = alg.add(s, alg.mul(di, di));
    }
    return alg.subc(s, f.p[0] * f.p[0]);
  }
  case QG_FUNC_PLANE: {
    T s = alg.scale(f.p[N], alg.subc(x[0], f.p[0]));
#pragma unroll
    for (int i = 1; i < N; ++i) s = alg.add(s, alg.scale(f.p[N + i], alg.subc(x[i], f.p[i])));
    return s;
  }
  default:
    break;
  }
  // TPMS
  T a[3], s[3], c[3];
  double pm[3];
  const bool eval_style = (f.kind == QG_FUNC_SCHOEN || f.kind == QG_FUNC_SCHOEN_IWP);
  tpms_args<N, A>(alg, f, x, eval_style, a, pm);
#pragma unroll
  for (int i = 0; i < 3; ++i) alg.sincos(a[i], s[i], c[i]);
  switch (f.kind) {
  case QG_FUNC_SCHOEN:
    return alg.add(alg.add(alg.mul(s[0], c[1]), alg.mul(s[1], c[2])), alg.mul(s[2], c[0]));
  case QG_FUNC_SCHWARZ_D:
    return alg.sub(alg.mul(alg.mul(c[0], c[1]), c[2]), alg.mul(alg.mul(s[0], s[1]), s[2]));
  case QG_FUNC_SCHWARZ_P:
    return alg.add(alg.add(c[0], c[1]), c[2]);
  default:
    break;
  }
  // functions that also need the doubled arguments
  T s2[3], c2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) alg.sincos(alg.scale(2.0, a[i]), s2[i], c2[i]);
  switch (f.kind) {
  case QG_FUNC_SCHOEN_IWP: {
    T t = alg.scale(2.0, alg.add(alg.add(alg.mul(c[0], c[1]), alg.mul(c[1], c[2])), alg.mul(c[2], c[0])));
    return alg.sub(alg.sub(alg.sub(t, c2[0]), c2[1]), c2[2]);
  }
  case QG_FUNC_SCHOEN_FRD: {
    T t = alg.mul(alg.mul(alg.scale(4.0, c[0]), c[1]), c[2]);
    return alg.sub(alg.sub(alg.sub(t, alg.mul(c2[0], c2[1])), alg.mul(c2[1], c2[2])), alg.mul(c2[2], c2[0]));
  }
  case QG_FUNC_FISCHER_KOCH_S:
    return alg.add(alg.add(alg.mul(alg.mul(c2[0], s[1]), c[2]), alg.mul(alg.mul(c[0], c2[1]), s[2])),
      alg.mul(alg.mul(s[0], c[1]), c2[2]));
  default:
    return alg.cst(0.0);
  }
}

template<int N, class A>
QG_HDN void phi_grad_raw(const A &alg, const FuncDesc &f, const typename A::T *x, typename A::T *g)
{
  typedef typename A::T T;
  switch (f.kind) {
  case QG_FUNC_SPHERE:
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.scale(2.0, alg.subc(x[i], f.p[1 + i]));
    return;
  case QG_FUNC_PLANE:
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.cst(f.p[N + i]);
    return;
  default:
    break;
  }
  T a[3], s[3], c[3];
  double pm[3];
  tpms_args<N, A>(alg, f, x, false, a, pm);
#pragma unroll
  for (int i = 0; i < 3; ++i) alg.sincos(a[i], s[i], c[i]);
  switch (f.kind) {
  case QG_FUNC_SCHOEN:
    g[0] = alg.scale(pm[0], alg.sub(alg.mul(c[0], c[1]), alg.mul(s[0], s[2])));
    g[1] = alg.scale(pm[1], alg.sub(alg.mul(c[1], c[2]), alg.mul(s[1], s[0])));
    if (N == 3) g[N - 1] = alg.scale(pm[2], alg.sub(alg.mul(c[2], c[0]), alg.mul(s[2], s[1])));
    return;
  case QG_FUNC_SCHWARZ_D:
    g[0] = alg.scale(pm[0], alg.neg(alg.add(alg.mul(alg.mul(s[0], c[1]), c[2]), alg.mul(alg.mul(c[0], s[1]), s[2]))));
    g[1] = alg.scale(pm[1], alg.neg(alg.add(alg.mul(alg.mul(c[0], s[1]), c[2]), alg.mul(alg.mul(s[0], c[1]), s[2]))));
    if (N == 3)
      g[N - 1] =
        alg.scale(pm[2], alg.neg(alg.add(alg.mul(alg.mul(c[0], c[1]), s[2]), alg.mul(alg.mul(s[0], s[1]), c[2]))));
    return;
  case QG_FUNC_SCHWARZ_P:
    g[0] = alg.scale(pm[0], alg.neg(s[0]));
    g[1] = alg.scale(pm[1], alg.neg(s[1]));
    if (N == 3) g[N - 1] = alg.scale(pm[2], alg.neg(s[2]));
    return;
  default:
    break;
  }
  T s2[3], c2[3];
  double p4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    alg.sincos(alg.scale(2.0, a[i]), s2[i], c2[i]);
    p4[i] = 2.0 * pm[i];
  }
  switch (f.kind) {
  case QG_FUNC_SCHOEN_IWP:
    g[0] = alg.scale(p4[0], alg.sub(s2[0], alg.mul(s[0], alg.add(c[1], c[2]))));
    g[1] = alg.scale(p4[0], alg.sub(s2[1], alg.mul(s[1], alg.add(c[0], c[2]))));  // sic: tpms_lib.cpp:162
    if (N == 3) g[N - 1] = alg.scale(p4[2], alg.sub(s2[2], alg.mul(s[2], alg.add(c[0], c[1]))));
    return;
  case QG_FUNC_SCHOEN_FRD:
    g[0] = alg.scale(p4[0],
      alg.add(alg.add(alg.mul(alg.mul(alg.scale(-2.0, s[0]), c[1]), c[2]), alg.mul(s2[0], c2[1])), alg.mul(c2[2], s2[0])));
    g[1] = alg.scale(p4[1],
      alg.add(alg.add(alg.mul(alg.mul(alg.scale(-2.0, c[0]), s[1]), c[2]), alg.mul(c2[0], s2[1])), alg.mul(s2[1], c2[2])));
    if (N == 3)
      g[N - 1] = alg.scale(p4[2],
        alg.add(alg.add(alg.mul(alg.mul(alg.scale(-2.0, c[0]), c[1]), s[2]), alg.mul(c2[1], s2[2])), alg.mul(s2[2], c2[0])));
    return;
  case QG_FUNC_FISCHER_KOCH_S:
    g[0] = alg.scale(pm[0],
      alg.add(alg.sub(alg.mul(alg.mul(alg.scale(-2.0, s2[0]), s[1]), c[2]), alg.mul(alg.mul(s[0], c2[1]), s[2])),
        alg.mul(alg.mul(c[0], c[1]), c2[2])));
    g[1] = alg.scale(pm[1],
      alg.sub(alg.sub(alg.mul(alg.mul(c2[0], c[1]), c[2]), alg.mul(alg.mul(alg.scale(2.0, c[0]), s2[1]), s[2])),
        alg.mul(alg.mul(s[0], s[1]), c2[2])));
    if (N == 3)
      g[N - 1] = alg.scale(pm[2],
        alg.sub(alg.add(alg.mul(alg.mul(alg.neg(c2[0]), s[1]), s[2]), alg.mul(alg.mul(c[0], c2[1]), c[2])),
          alg.mul(alg.mul(alg.scale(2.0, s[0]), c[1]), s2[2])));
    return;
  default:
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.cst(0.0);
  }
}

template<int N> QG_HD double phi_val(const FuncDesc &f, const double *x)
{
  AlgD<N> alg;
  const double v = phi_eval_raw<N, AlgD<N>>(alg, f, x);
  return f.negative ? -v : v;
}
template<int N> QG_HD void phi_grad(const FuncDesc &f, const double *x, double *g)
{
  AlgD<N> alg;
  phi_grad_raw<N, AlgD<N>>(alg, f, x, g);
  if (f.negative) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = -g[i];
  }
}
template<int N> QG_HD Iv<N> phi_val_iv(const FuncDesc &f, const Iv<N> *x, const Dl<N> &dl)
{
  AlgI<N> alg;
  alg.dl = dl;
  const Iv<N> v = phi_eval_raw<N, AlgI<N>>(alg, f, x);
  return f.negative ? iv_neg(v) : v;
}
template<int N> QG_HD void phi_grad_iv(const FuncDesc &f, const Iv<N> *x, const Dl<N> &dl, Iv<N> *g)
{
  AlgI<N> alg;
  alg.dl = dl;
  phi_grad_raw<N, AlgI<N>>(alg, f, x, g);
  if (f.negative) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = iv_neg(g[i]);
  }
}

}  // namespace qg
