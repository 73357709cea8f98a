// qugar_b200 -- level-set functions evaluated on double and on Iv<N> (value and gradient).
//
// One formula per function, written against a small algebra (AlgD: doubles, AlgI: intervals) so
// that both evaluations keep the reference's expression order:
//   Sphere/Disk  phi = |x-c|^2 - r^2          /root/reference/cpp/qugar/primitive_funcs_lib.cpp:66-74
//   Plane/Line   phi = (x-o).n                /root/reference/cpp/qugar/primitive_funcs_lib.cpp:849-857
//   TPMS family  Schoen, SchoenIWP, SchoenFRD, FischerKochS, SchwarzDiamond, SchwarzPrimitive
//                                             /root/reference/cpp/qugar/tpms_lib.cpp:83-108,137-166,205-240,283-322,360-391,430-452
//   Negative     -phi                         /root/reference/cpp/qugar/impl_funcs_lib.cpp:207-227
//   Square       phi = max_d |x_d| - 1        /root/reference/cpp/qugar/impl_funcs_lib.cpp:48-98
// 2-D TPMS are the 3-D formula at fixed z with period 1 in z (tpms_lib.hpp:41-62).
#pragma once
#include "qg_math.cuh"

namespace qg {

enum FuncKind {
  QG_FUNC_SPHERE = 0,
  QG_FUNC_PLANE = 1,
  QG_FUNC_BEZIER = 2,
  QG_FUNC_CONSTANT = 3,
  QG_FUNC_CYLINDER = 4,
  QG_FUNC_ELLIPSOID = 5,
  QG_FUNC_ANNULUS = 6,
  QG_FUNC_TORUS = 7,
  QG_FUNC_COMPOSITE = 8,
  QG_FUNC_DIM_LINEAR = 9,
  QG_FUNC_SCHOEN = 10,
  QG_FUNC_SCHOEN_IWP = 11,
  QG_FUNC_SCHOEN_FRD = 12,
  QG_FUNC_FISCHER_KOCH_S = 13,
  QG_FUNC_SCHWARZ_D = 14,
  QG_FUNC_SCHWARZ_P = 15,
  QG_FUNC_SQUARE = 16
};
QG_HD bool is_tpms(int kind) { return kind >= QG_FUNC_SCHOEN && kind <= QG_FUNC_SCHWARZ_P; }

constexpr int kMaxBezierOrder = 8;  // coefficients per direction (degree 7)

// params: SPHERE {r, c[dim]}; PLANE {o[dim], n[dim]}; TPMS 3-D {m,n,q}; TPMS 2-D {m,n,z};
// CONSTANT {value}; CYLINDER {r, origin[3], unit axis[3]}; ELLIPSOID {semi_axes[dim], origin[dim], basis[dim][dim]};
// SQUARE {} (the box [-1,1]^dim; placed by an affine map through a composite leaf);
// ANNULUS {r_in, r_out, center[2]}; TORUS {R, r, center[3], unit axis[3]}; DIM_LINEAR {2^dim vertex values, direction 0 fastest};
// BEZIER: order[dim] and the coefficient array (direction 0 fastest, bezier_tp.cpp:53-61) in coefs
// COMPOSITE: comp points to a CompositeDesc (a postfix program over leaf functions; see below)
struct CompositeDesc;
struct FuncDesc
{
  int kind;
  int dim;
  int negative;
  int same_args;  // TPMS: bit i set if the value-style and gradient-style arguments of coordinate i coincide (tpms_same_arg_mask)
  double p[16];
  const double *coefs;
  int order[3];
  int pad2;
  const CompositeDesc *comp;
  QG_HD int k() const { return kind; }
};
// TransformedFunction / Negative / AddFunctions / SubtractFunctions (impl_funcs_lib.cpp:168-290) nest arbitrarily upstream.
// A composite is a postfix program over leaf functions, evaluated with a value stack and a point stack:
//   LEAF l   push leaf l at the current point (its own negative flag applied)
//   XPUSH a  push the point T_a x   (T_a = the INVERSE of the user's affine map: row-major matrix, then translation;
//            affine_transf.cpp:345-377)
//   XPOP a   pop the point; the gradient on top of the value stack becomes T_a^t gradient
//   NEG, ADD, SUB on the value stack.
// Rarely used: evaluated by the run-time-kind instantiation.
constexpr int kCompMaxLeaf = 6, kCompMaxAff = 6, kCompMaxInstr = 24, kCompStack = 4;
enum CompositeInstr { CI_LEAF = 0, CI_XPUSH = 1, CI_XPOP = 2, CI_NEG = 3, CI_ADD = 4, CI_SUB = 5 };
struct CompositeDesc
{
  int nleaf, naff, ninstr, pad;
  FuncDesc leaf[kCompMaxLeaf];        // non-composite functions (leaf.negative applies to the leaf)
  double aff[kCompMaxAff][12];        // dim*dim matrix entries (row major), then dim translation entries
  unsigned char op[kCompMaxInstr], arg[kCompMaxInstr];
};
// Host side: decodes the parameter array of a QG_COMPOSITE function (qugar_b200/cpp.py, "program form"):
//   3, n_leaf, n_aff, n_instr, leaves (kind, negative, np, p[np]) ..., affine maps (12 doubles each) ..., instructions (op, arg) ...
// make_leaf(FuncDesc &, leaf index, kind, negative, p, np) builds one leaf and returns an error text or nullptr.
// Checks the stack discipline so that the device evaluator never indexes outside its stacks.
template<class MakeLeaf>
inline const char *composite_parse(const double *params, long long n, CompositeDesc &cd, MakeLeaf make_leaf)
{
  if (n < 4 || (int)params[0] != 3) return "composite parameters are not in program form";
  const int nleaf = (int)params[1], naff = (int)params[2], ninstr = (int)params[3];
  if (nleaf < 1 || nleaf > kCompMaxLeaf || naff < 0 || naff > kCompMaxAff || ninstr < 1 || ninstr > kCompMaxInstr)
    return "composite too large (leaves / affine maps / instructions)";
  cd.nleaf = nleaf;
  cd.naff = naff;
  cd.ninstr = ninstr;
  cd.pad = 0;
  long long q = 4;
  for (int l = 0; l < nleaf; ++l) {
    if (q + 3 > n) return "truncated composite leaf";
    const int lkind = (int)params[q], lneg = (int)params[q + 1];
    const long long np = (long long)params[q + 2];
    if (lkind == QG_FUNC_COMPOSITE) return "a composite leaf cannot be a composite (flatten it into the program)";
    if (np < 0 || q + 3 + np > n) return "truncated composite leaf";
    const char *err = make_leaf(cd.leaf[l], l, lkind, lneg, params + q + 3, (int)np);
    if (err) return err;
    q += 3 + np;
  }
  for (int a = 0; a < naff; ++a) {
    if (q + 12 > n) return "truncated composite affine map";
    for (int i = 0; i < 12; ++i) cd.aff[a][i] = params[q + i];
    q += 12;
  }
  int sp = 0, pp = 0;
  for (int i = 0; i < ninstr; ++i) {
    if (q + 2 > n) return "truncated composite program";
    const int op = (int)params[q], arg = (int)params[q + 1];
    q += 2;
    switch (op) {
    case CI_LEAF:
      if (arg < 0 || arg >= nleaf) return "composite program: leaf index out of range";
      if (++sp > kCompStack) return "composite program: value stack overflow";
      break;
    case CI_XPUSH:
      if (arg < 0 || arg >= naff) return "composite program: affine index out of range";
      if (++pp + 1 > kCompStack) return "composite program: point stack overflow";
      break;
    case CI_XPOP:
      if (arg < 0 || arg >= naff || pp < 1 || sp < 1) return "composite program: unbalanced XPOP";
      --pp;
      break;
    case CI_NEG:
      if (sp < 1) return "composite program: NEG on an empty stack";
      break;
    case CI_ADD:
    case CI_SUB:
      if (sp < 2) return "composite program: binary operator needs two operands";
      --sp;
      break;
    default: return "composite program: unknown instruction";
    }
    cd.op[i] = (unsigned char)op;
    cd.arg[i] = (unsigned char)arg;
  }
  if (sp != 1 || pp != 0) return "composite program does not reduce to one value";
  if (q != n) return "trailing composite parameters";
  return nullptr;
}

// The same descriptor with the kind fixed at compile time: kernels are instantiated per kind so that the
// switch statements below fold away (smaller code, no dead trig arrays in registers).  Every evaluation
// routine is a template on the descriptor class F and asks f.k() for the kind.
template<int KIND> struct FuncK : FuncDesc
{
  static constexpr bool kSameArgs = false;
  QG_HD FuncK() {}
  QG_HD explicit FuncK(const FuncDesc &d) : FuncDesc(d) {}
  QG_HD int k() const { return KIND >= 0 ? KIND : kind; }  // KIND = -1: run-time kind (rarely used functions)
};
// FuncK for a Schoen / SchoenIWP whose three frequencies are powers of two (same_args == 7): the value-style and the
// gradient-style arguments are the same doubles, known at compile time -- the line evaluator then carries ONE set of trig
// values (12 fewer live doubles in the Newton loop).  Chosen by the launcher from FuncDesc::same_args.
template<int KIND> struct FuncKSame : FuncK<KIND>
{
  static constexpr bool kSameArgs = true;
  QG_HD FuncKSame() {}
  QG_HD explicit FuncKSame(const FuncDesc &d) : FuncK<KIND>(d) {}
};
template<class F> struct func_same_args { static constexpr bool value = false; };
template<int KIND> struct func_same_args<FuncKSame<KIND>> { static constexpr bool value = true; };

template<int N> struct AlgD
{
  typedef double T;
  QG_HD T cst(double c) const { return c; }
  QG_HD T add(T a, T b) const { return a + b; }
  QG_HD T sub(T a, T b) const { return a - b; }
  QG_HD T neg(T a) const { return -a; }
  QG_HD T subc(T a, double c) const { return a - c; }
  QG_HD T addc(T a, double c) const { return a + c; }
  QG_HD T scale(double c, T a) const { return a * c; }
  QG_HD T divc(T a, double c) const { return a / c; }
  QG_HD T mul(T a, T b) const { return a * b; }
  QG_HD void sincos(T a, T &s, T &c) const { sincos_det(a, s, c); }
  QG_HD double mid(T a) const { return a; }
  QG_HD int sgn(T a) const { return (0.0 < a) - (a < 0.0); }
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
  QG_HD T addc(const T &a, double c) const { return iv_subc(a, -c); }
  QG_HD T scale(double c, const T &a) const { return iv_scale(c, a); }
  QG_HD T divc(const T &a, double c) const { return iv_divc(a, c); }
  QG_HD T mul(const T &a, const T &b) const { return iv_mul(a, b, dl); }
  QG_HD void sincos(const T &a, T &s, T &c) const { iv_sincos(a, dl, s, c); }
  QG_HD double mid(const T &a) const { return a.a; }
  // Interval::sign() (absent Algoim dependency): restated as the sign of the centre value, see oracle/oracle_funcs.hpp
  QG_HD int sgn(const T &a) const { return (0.0 < a.a) - (a.a < 0.0); }
};

// ---- TPMS family: trigonometric stage + combination stage ---------------------------------------------
// Every TPMS value / gradient is a polynomial in sin, cos of the three arguments a_i (and of 2 a_i for
// IWP / FRD / Fischer-Koch S).  The two stages are kept apart so that a line evaluator (LineEval below)
// can keep the trig of the coordinates that do not move along the line; the combination code is shared,
// so every evaluation performs the same fp64 operation sequence.
// eval_style: a_i = (2 pi) * (m_i * x_i)  [Schoen, SchoenIWP values];  otherwise a_i = ((2 pi) m_i) * x_i.
template<class T> struct TpmsTrig
{
  T s[3], c[3], s2[3], c2[3];
};
QG_HD bool tpms_eval_style(int kind) { return kind == QG_FUNC_SCHOEN || kind == QG_FUNC_SCHOEN_IWP; }
QG_HD bool tpms_doubled(int kind)
{
  return kind == QG_FUNC_SCHOEN_IWP || kind == QG_FUNC_SCHOEN_FRD || kind == QG_FUNC_FISCHER_KOCH_S;
}
// The reference evaluates a Schoen / SchoenIWP VALUE with the arguments (2 pi) * (m_i x_i) and its GRADIENT with
// ((2 pi) m_i) * x_i (tpms_lib.cpp:83-108, 137-170): two roundings of the same number, hence two sincos per point when
// both are needed.  When m_i is a power of two (or zero) both products are exact scalings of fl(2 pi) x_i and the two
// arguments are the SAME double: bit i of the mask says so and the line evaluator reuses the first sincos (bit-identical).
inline int tpms_same_arg_mask(int kind, int dim, const double *p)
{
  if (!(kind == QG_FUNC_SCHOEN || kind == QG_FUNC_SCHOEN_IWP)) return 0;
  int mask = 0;
  for (int i = 0; i < 3; ++i) {
    const double m = (dim == 3 || i < 2) ? p[i] : 1.0;
    int e = 0;
    const double fr = frexp(fabs(m), &e);
    if (m == 0.0 || (fr == 0.5 && e > -900 && e < 900)) mask |= 1 << i;
  }
  return mask;
}
// frequency of TPMS argument i (2-D: the third argument is the fixed z with period 1, tpms_lib.hpp:41-62)
template<int N, class F> QG_HD double tpms_freq(const F &f, int i) { return (N == 3 || i < 2) ? f.p[i] : 1.0; }

// sin/cos (and doubled) of argument I from the coordinate value xi
template<int N, int I, class A, class F>
QG_HD void tpms_trig_coord(const A &alg, const F &f, const typename A::T &xi, bool eval_style, bool doubled,
  TpmsTrig<typename A::T> &tr)
{
  const double two_pi = 2.0 * kPi;
  const double m = tpms_freq<N>(f, I);
  const typename A::T a = eval_style ? alg.scale(two_pi, alg.scale(m, xi)) : alg.scale(two_pi * m, xi);
  alg.sincos(a, tr.s[I], tr.c[I]);
  if (doubled) alg.sincos(alg.scale(2.0, a), tr.s2[I], tr.c2[I]);
}
template<int N, class A, class F>
QG_HD void tpms_trig(const A &alg, const F &f, const typename A::T *x, bool eval_style,
  TpmsTrig<typename A::T> &tr)
{
  const bool dbl = tpms_doubled(f.k());
  tpms_trig_coord<N, 0>(alg, f, x[0], eval_style, dbl, tr);
  tpms_trig_coord<N, 1>(alg, f, x[1], eval_style, dbl, tr);
  if (N == 3)
    tpms_trig_coord<N, 2>(alg, f, x[N - 1], eval_style, dbl, tr);
  else
    tpms_trig_coord<N, 2>(alg, f, alg.cst(f.p[2]), eval_style, dbl, tr);
}

template<class A> QG_HD typename A::T tpms_combine_val(const A &alg, int kind, const TpmsTrig<typename A::T> &tr)
{
  typedef typename A::T T;
  const T *s = tr.s, *c = tr.c, *s2 = tr.s2, *c2 = tr.c2;
  (void)s2;
  switch (kind) {
  case QG_FUNC_SCHOEN:
    return alg.add(alg.add(alg.mul(s[0], c[1]), alg.mul(s[1], c[2])), alg.mul(s[2], c[0]));
  case QG_FUNC_SCHWARZ_D:
    return alg.sub(alg.mul(alg.mul(c[0], c[1]), c[2]), alg.mul(alg.mul(s[0], s[1]), s[2]));
  case QG_FUNC_SCHWARZ_P:
    return alg.add(alg.add(c[0], c[1]), c[2]);
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

template<int N, class A, class F>
QG_HD void tpms_combine_grad(const A &alg, const F &f, const TpmsTrig<typename A::T> &tr, typename A::T *g)
{
  typedef typename A::T T;
  const T *s = tr.s, *c = tr.c, *s2 = tr.s2, *c2 = tr.c2;
  const double two_pi = 2.0 * kPi;
  double pm[3], p4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pm[i] = two_pi * tpms_freq<N>(f, i);
    p4[i] = 2.0 * pm[i];
  }
  switch (f.k()) {
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

// diff = x - origin, dd = dot(diff, diff), proj = dot(diff, axis)   (3-D primitives; x has >= 3 entries only if N == 3)
template<int N, class A>
QG_HD void prim_diff_dots(const A &alg, const typename A::T *x, const double *origin, const double *axis, typename A::T *diff,
  typename A::T &dd, typename A::T &proj)
{
#pragma unroll
  for (int i = 0; i < 3; ++i) diff[i] = alg.subc(x[i < N ? i : 0], origin[i]);
  dd = alg.mul(diff[0], diff[0]);
  proj = alg.scale(axis[0], diff[0]);
#pragma unroll
  for (int i = 1; i < 3; ++i) {
    dd = alg.add(dd, alg.mul(diff[i], diff[i]));
    proj = alg.add(proj, alg.scale(axis[i], diff[i]));
  }
}

// ---- BezierTP<dim,1> (bezier_tp.cpp:165-255): de Casteljau on the algebra's type -------------------------
// Value: casteljau<dim> evaluates the last direction over values of casteljau<dim-1>, consuming the
// coefficients in storage order (direction 0 fastest).  Same recursion for double and Interval.
template<int D, class A> struct BezierEval
{
  typedef typename A::T T;
  // value at x[0..D-1]; c advances over the coefficients it consumes
  static QG_HD T value(const A &alg, const T *x, const double *&c, const int *order)
  {
    const int lo = order[D - 1];
    T beta[kMaxBezierOrder];
    for (int i = 0; i < lo; ++i) beta[i] = BezierEval<D - 1, A>::value(alg, x, c, order);
    const T &xd = x[D - 1];
    for (int j = 1; j < lo; ++j)
      for (int k = 0; k < lo - j; ++k) beta[k] = alg.add(beta[k], alg.mul(alg.sub(beta[k + 1], beta[k]), xd));
    return beta[0];
  }
  // out[0] = value, out[1..D] = derivatives w.r.t. directions 0..D-1  (casteljau_der)
  static QG_HD void value_der(const A &alg, const T *x, const double *&c, const int *order, T *out)
  {
    const int lo = order[D - 1];
    const int degree = lo - 1;
    T beta[kMaxBezierOrder][D];  // [i][0] value, [i][1..D-1] derivatives of the lower directions
    T gamma[kMaxBezierOrder];
    for (int i = 0; i < lo; ++i) {
      BezierEval<D - 1, A>::value_der(alg, x, c, order, beta[i]);
      gamma[i] = alg.cst(0.0);
    }
    if (lo > 1) {
      // derivative along the last direction, degree-elevated to the same Bernstein basis, accumulated
      // in the reference's order
      gamma[0] = alg.scale(-(double)degree, beta[0][0]);
      gamma[1] = alg.neg(beta[0][0]);
      for (int i = 1; i < degree; ++i) {
        const T b = beta[i][0];
        gamma[i - 1] = alg.add(gamma[i - 1], alg.scale((double)(lo - i), b));
        gamma[i] = alg.add(gamma[i], alg.scale((double)(2 * i - degree), b));
        gamma[i + 1] = alg.sub(gamma[i + 1], alg.scale((double)(i + 1), b));
      }
      gamma[degree - 1] = alg.add(gamma[degree - 1], beta[degree][0]);
      gamma[degree] = alg.add(gamma[degree], alg.scale((double)degree, beta[degree][0]));
      const T &xd = x[D - 1];
      for (int j = 1; j < lo; ++j)
        for (int k = 0; k < lo - j; ++k) {
#pragma unroll
          for (int e = 0; e < D; ++e) beta[k][e] = alg.add(beta[k][e], alg.mul(alg.sub(beta[k + 1][e], beta[k][e]), xd));
          gamma[k] = alg.add(gamma[k], alg.mul(alg.sub(gamma[k + 1], gamma[k]), xd));
        }
    }
#pragma unroll
    for (int e = 0; e < D; ++e) out[e] = beta[0][e];
    out[D] = gamma[0];
  }
};
template<class A> struct BezierEval<0, A>
{
  typedef typename A::T T;
  static QG_HD T value(const A &alg, const T *, const double *&c, const int *) { return alg.cst(*c++); }
  static QG_HD void value_der(const A &alg, const T *, const double *&c, const int *, T *out) { out[0] = alg.cst(*c++); }
};

// The same recursion with every loop unrolled up to MO coefficients per direction and guarded by the actual order:
// the work arrays stay in registers (the generic version above indexes local memory).  Same operations in the same
// order, so the values are bit-identical.  Used for doubles when no direction has more than MO coefficients.
template<int D, class A, int MO> struct BezierEvalU
{
  typedef typename A::T T;
  static QG_HD T value(const A &alg, const T *x, const double *&c, const int *order)
  {
    const int lo = order[D - 1];
    T beta[MO];
#pragma unroll
    for (int i = 0; i < MO; ++i)
      if (i < lo) beta[i] = BezierEvalU<D - 1, A, MO>::value(alg, x, c, order);
    const T &xd = x[D - 1];
#pragma unroll
    for (int j = 1; j < MO; ++j) {
      if (j < lo) {
#pragma unroll
        for (int k = 0; k < MO - 1; ++k)
          if (k < lo - j) beta[k] = alg.add(beta[k], alg.mul(alg.sub(beta[k + 1], beta[k]), xd));
      }
    }
    return beta[0];
  }
  static QG_HD void value_der(const A &alg, const T *x, const double *&c, const int *order, T *out)
  {
    const int lo = order[D - 1];
    const int degree = lo - 1;
    T beta[MO][D];
    T gamma[MO];
#pragma unroll
    for (int i = 0; i < MO; ++i) {
      gamma[i] = alg.cst(0.0);
      if (i < lo) BezierEvalU<D - 1, A, MO>::value_der(alg, x, c, order, beta[i]);
    }
    if (lo > 1) {
      gamma[0] = alg.scale(-(double)degree, beta[0][0]);
      gamma[1] = alg.neg(beta[0][0]);
#pragma unroll
      for (int i = 1; i < MO - 1; ++i) {
        if (i < degree) {
          const T b = beta[i][0];
          gamma[i - 1] = alg.add(gamma[i - 1], alg.scale((double)(lo - i), b));
          gamma[i] = alg.add(gamma[i], alg.scale((double)(2 * i - degree), b));
          gamma[i + 1] = alg.sub(gamma[i + 1], alg.scale((double)(i + 1), b));
        }
      }
      // gamma[degree-1] += beta[degree][0]; gamma[degree] += degree * beta[degree][0]   (degree is a run-time value)
#pragma unroll
      for (int i = 1; i < MO; ++i) {
        if (i == degree) {
          gamma[i - 1] = alg.add(gamma[i - 1], beta[i][0]);
          gamma[i] = alg.add(gamma[i], alg.scale((double)degree, beta[i][0]));
        }
      }
      const T &xd = x[D - 1];
#pragma unroll
      for (int j = 1; j < MO; ++j) {
        if (j < lo) {
#pragma unroll
          for (int k = 0; k < MO - 1; ++k) {
            if (k < lo - j) {
#pragma unroll
              for (int e = 0; e < D; ++e) beta[k][e] = alg.add(beta[k][e], alg.mul(alg.sub(beta[k + 1][e], beta[k][e]), xd));
              gamma[k] = alg.add(gamma[k], alg.mul(alg.sub(gamma[k + 1], gamma[k]), xd));
            }
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < D; ++e) out[e] = beta[0][e];
    out[D] = gamma[0];
  }
};
template<class A, int MO> struct BezierEvalU<0, A, MO>
{
  typedef typename A::T T;
  static QG_HD T value(const A &alg, const T *, const double *&c, const int *) { return alg.cst(*c++); }
  static QG_HD void value_der(const A &alg, const T *, const double *&c, const int *, T *out) { out[0] = alg.cst(*c++); }
};
template<class T> struct IsDouble { static constexpr bool value = false; };
template<> struct IsDouble<double> { static constexpr bool value = true; };
// QG_BEZIER_GENERIC_ONLY (set for the run-time-kind unit and the engine unit, where Bezier is one of many switch cases):
// keep only the loop version there -- same bits, a fifth of the compile time
template<int N, class F> QG_HD bool bezier_small(const F &f)
{
#if defined(QG_BEZIER_GENERIC_ONLY)
  (void)f;
  return false;
#else
  return f.order[0] <= 4 && f.order[1] <= 4 && (N < 3 || f.order[2] <= 4);
#endif
}

// Square<dim>: direction of the largest |centre value| - 1, first maximum wins (impl_funcs_lib.cpp:50-63, algoim::argmax)
template<int N, class A> QG_HD int square_dir(const A &alg, const typename A::T *x)
{
  int ind = 0;
  double best = fabs(alg.mid(x[0])) - 1.0;
#pragma unroll
  for (int d = 1; d < N; ++d) {
    const double v = fabs(alg.mid(x[d])) - 1.0;
    if (v > best) {
      best = v;
      ind = d;
    }
  }
  return ind;
}
template<int N, class T> QG_HD T pick_dir(const T *x, int ind)
{
  T r = x[0];
#pragma unroll
  for (int d = 1; d < N; ++d)
    if (d == ind) r = x[d];
  return r;
}

template<int N, class A, class F> QG_HD typename A::T phi_eval_raw(const A &alg, const F &f, const typename A::T *x)
{
  typedef typename A::T T;
  switch (f.k()) {
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
  case QG_FUNC_BEZIER: {
    const double *c = f.coefs;
    if (IsDouble<T>::value && bezier_small<N>(f)) return BezierEvalU<N, A, 4>::value(alg, x, c, f.order);
    return BezierEval<N, A>::value(alg, x, c, f.order);
  }
  case QG_FUNC_CONSTANT:
    return alg.cst(f.p[0]);
  case QG_FUNC_CYLINDER: {  // primitive_funcs_lib.cpp:211-217: dot(diff,diff) - proj^2 - r^2
    T diff[3], dd, proj;
    prim_diff_dots<N, A>(alg, x, f.p + 1, f.p + 4, diff, dd, proj);
    return alg.subc(alg.sub(dd, alg.mul(proj, proj)), f.p[0] * f.p[0]);
  }
  case QG_FUNC_ELLIPSOID: {  // :327-337
    T diff[N];
#pragma unroll
    for (int i = 0; i < N; ++i) diff[i] = alg.subc(x[i], f.p[N + i]);
    T val = alg.cst(-1.0);
#pragma unroll
    for (int dir = 0; dir < N; ++dir) {
      const double *ax = f.p + 2 * N + N * dir;
      T d = alg.scale(ax[0], diff[0]);
#pragma unroll
      for (int i = 1; i < N; ++i) d = alg.add(d, alg.scale(ax[i], diff[i]));
      const T proj = alg.divc(d, f.p[dir]);
      val = alg.add(val, alg.mul(proj, proj));
    }
    return val;
  }
  case QG_FUNC_ANNULUS: {  // :484-490: ((dist2 - Ri2 - Ro2) * dist2) + Ri2 * Ro2
    T d0 = alg.subc(x[0], f.p[2]), d1 = alg.subc(x[1], f.p[3]);
    const T dist2 = alg.add(alg.mul(d0, d0), alg.mul(d1, d1));
    const double ri2 = f.p[0] * f.p[0], ro2 = f.p[1] * f.p[1];
    return alg.add(alg.mul(alg.subc(alg.subc(dist2, ri2), ro2), dist2), alg.cst(ri2 * ro2));
  }
  case QG_FUNC_SQUARE: {  // impl_funcs_lib.cpp:48-66: x_ind * sgn(x_ind) - 1 (on a double this is max_d |x_d| - 1, bit for bit)
    const T xi = pick_dir<N, T>(x, square_dir<N, A>(alg, x));
    return alg.subc(alg.scale((double)alg.sgn(xi), xi), 1.0);
  }
  case QG_FUNC_DIM_LINEAR: {  // impl_funcs_lib.cpp:108-124: bi- / tri-linear interpolation of the vertex values
    const T u0 = alg.addc(alg.neg(x[0]), 1.0), u1 = alg.addc(alg.neg(x[1]), 1.0);
    const T e0 = alg.add(alg.scale(f.p[0], u0), alg.scale(f.p[1], x[0]));
    const T e1 = alg.add(alg.scale(f.p[2], u0), alg.scale(f.p[3], x[0]));
    const T lower = alg.add(alg.mul(e0, u1), alg.mul(e1, x[1]));
    if (N == 2) return lower;
    const T u2 = alg.addc(alg.neg(x[N - 1]), 1.0);
    const T e2 = alg.add(alg.scale(f.p[4], u0), alg.scale(f.p[5], x[0]));
    const T e3 = alg.add(alg.scale(f.p[6], u0), alg.scale(f.p[7], x[0]));
    const T upper = alg.add(alg.mul(e2, u1), alg.mul(e3, x[1]));
    return alg.add(alg.mul(lower, u2), alg.mul(upper, x[N - 1]));
  }
  case QG_FUNC_TORUS: {  // :705-717
    const double R2 = f.p[0] * f.p[0], r2 = f.p[1] * f.p[1];
    T diff[3], dist2, proj;
    prim_diff_dots<N, A>(alg, x, f.p + 2, f.p + 5, diff, dist2, proj);
    T v = alg.sub(alg.mul(dist2, dist2), alg.scale(2.0 * (R2 + r2), dist2));
    v = alg.add(v, alg.mul(alg.scale(4.0 * R2, proj), proj));
    return alg.add(v, alg.cst((R2 - r2) * (R2 - r2)));
  }
  default:
    break;
  }
  TpmsTrig<T> tr;
  tpms_trig<N>(alg, f, x, tpms_eval_style(f.k()), tr);
  return tpms_combine_val<A>(alg, f.k(), tr);
}

template<int N, class A, class F>
QG_HD void phi_grad_raw(const A &alg, const F &f, const typename A::T *x, typename A::T *g)
{
  typedef typename A::T T;
  switch (f.k()) {
  case QG_FUNC_SPHERE:
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.scale(2.0, alg.subc(x[i], f.p[1 + i]));
    return;
  case QG_FUNC_PLANE:
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.cst(f.p[N + i]);
    return;
  case QG_FUNC_BEZIER: {
    const double *c = f.coefs;
    T out[N + 1];
    if (IsDouble<T>::value && bezier_small<N>(f))
      BezierEvalU<N, A, 4>::value_der(alg, x, c, f.order, out);
    else
      BezierEval<N, A>::value_der(alg, x, c, f.order, out);
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = out[i + 1];
    return;
  }
  case QG_FUNC_CONSTANT:
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.cst(0.0);
    return;
  case QG_FUNC_CYLINDER: {  // :219-225: 2 diff - (2 proj) axis
    T diff[3], dd, proj;
    prim_diff_dots<N, A>(alg, x, f.p + 1, f.p + 4, diff, dd, proj);
    const T p2 = alg.scale(2.0, proj);
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.sub(alg.scale(2.0, diff[i < 3 ? i : 0]), alg.scale(f.p[4 + i], p2));
    return;
  }
  case QG_FUNC_ELLIPSOID: {  // :339-350
    T diff[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      diff[i] = alg.subc(x[i], f.p[N + i]);
      g[i] = alg.cst(0.0);
    }
#pragma unroll
    for (int dir = 0; dir < N; ++dir) {
      const double coef = 2.0 / f.p[dir] / f.p[dir];
      const double *ax = f.p + 2 * N + N * dir;
      T d = alg.scale(ax[0], diff[0]);
#pragma unroll
      for (int i = 1; i < N; ++i) d = alg.add(d, alg.scale(ax[i], diff[i]));
      const T cd = alg.scale(coef, d);
#pragma unroll
      for (int i = 0; i < N; ++i) g[i] = alg.add(g[i], alg.scale(ax[i], cd));
    }
    return;
  }
  case QG_FUNC_ANNULUS: {  // :492-498: (4 |diff|^2 - Ri2 - Ro2) diff
    T d0 = alg.subc(x[0], f.p[2]), d1 = alg.subc(x[1], f.p[3]);
    const double ri2 = f.p[0] * f.p[0], ro2 = f.p[1] * f.p[1];
    const T c = alg.subc(alg.subc(alg.scale(4.0, alg.add(alg.mul(d0, d0), alg.mul(d1, d1))), ri2), ro2);
    g[0] = alg.mul(c, d0);
    g[1] = alg.mul(c, d1);
    return;
  }
  case QG_FUNC_SQUARE: {  // impl_funcs_lib.cpp:68-82: sgn(x_ind) e_ind
    const int ind = square_dir<N, A>(alg, x);
    const double sg = (double)alg.sgn(pick_dir<N, T>(x, ind));
#pragma unroll
    for (int d = 0; d < N; ++d) g[d] = alg.cst(d == ind ? sg : 0.0);
    return;
  }
  case QG_FUNC_DIM_LINEAR: {  // impl_funcs_lib.cpp:127-152
    const double *c = f.p;
    const T u0 = alg.addc(alg.neg(x[0]), 1.0), u1 = alg.addc(alg.neg(x[1]), 1.0);
    if (N == 2) {
      g[0] = alg.add(alg.scale(c[1] - c[0], u1), alg.scale(c[3] - c[2], x[1]));
      g[1] = alg.add(alg.scale(c[2] - c[0], u0), alg.scale(c[3] - c[1], x[0]));
      return;
    }
    const T u2 = alg.addc(alg.neg(x[N - 1]), 1.0);
    const T &x2 = x[N - 1];
    g[0] = alg.add(alg.mul(alg.add(alg.scale(c[1] - c[0], u1), alg.scale(c[3] - c[2], x[1])), u2),
      alg.mul(alg.add(alg.scale(c[5] - c[4], u1), alg.scale(c[7] - c[6], x[1])), x2));
    g[1] = alg.add(alg.mul(alg.add(alg.scale(c[2] - c[0], u0), alg.scale(c[3] - c[1], x[0])), u2),
      alg.mul(alg.add(alg.scale(c[6] - c[4], u0), alg.scale(c[7] - c[5], x[0])), x2));
    g[N - 1] = alg.add(alg.mul(alg.add(alg.scale(c[4] - c[0], u0), alg.scale(c[5] - c[1], x[0])), u1),
      alg.mul(alg.add(alg.scale(c[6] - c[2], u0), alg.scale(c[7] - c[3], x[0])), x[1]));
    return;
  }
  case QG_FUNC_TORUS: {  // :719-729: 4 (dist2 - R2 - r2) diff + (8 R2 proj) axis
    const double R2 = f.p[0] * f.p[0], r2 = f.p[1] * f.p[1];
    T diff[3], dist2, proj;
    prim_diff_dots<N, A>(alg, x, f.p + 2, f.p + 5, diff, dist2, proj);
    const T c = alg.scale(4.0, alg.subc(alg.subc(dist2, R2), r2));
    const T pa = alg.scale(8.0 * R2, proj);
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = alg.add(alg.mul(c, diff[i < 3 ? i : 0]), alg.scale(f.p[5 + i], pa));
    return;
  }
  default:
    break;
  }
  TpmsTrig<T> tr;
  tpms_trig<N>(alg, f, x, false, tr);
  tpms_combine_grad<N>(alg, f, tr, g);
}

// ---- composite functions -----------------------------------------------------------------------------------
template<int N, class A> QG_HD void affine_point(const A &alg, const double *c, const typename A::T *x, typename A::T *y)
{
  // coefs(0) * x0 + coefs(1) * x1 (+ coefs(2) * x2) + translation   (affine_transf.cpp:345-358)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    typename A::T v = alg.scale(c[N * i], x[0]);
#pragma unroll
    for (int j = 1; j < N; ++j) v = alg.add(v, alg.scale(c[N * i + j], x[j]));
    y[i] = alg.addc(v, c[N * N + i]);
  }
}
template<int N, class A> QG_HD void affine_vector(const A &alg, const double *c, const typename A::T *v, typename A::T *out)
{
  // transpose of the matrix (affine_transf.cpp:363-377)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    typename A::T r = alg.scale(c[i], v[0]);
#pragma unroll
    for (int j = 1; j < N; ++j) r = alg.add(r, alg.scale(c[N * j + i], v[j]));
    out[i] = r;
  }
}
// value of the composite at x
template<int N, class A> QG_HDN typename A::T composite_eval(const A &alg, const CompositeDesc &cd, const typename A::T *x)
{
  typedef typename A::T T;
  T pt[kCompStack][N], v[kCompStack];
  int sp = 0, pp = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) pt[0][i] = x[i];
  for (int ip = 0; ip < cd.ninstr; ++ip) {
    const int a = cd.arg[ip];
    switch (cd.op[ip]) {
    case CI_LEAF: {
      const T r = phi_eval_raw<N>(alg, cd.leaf[a], pt[pp]);
      v[sp++] = cd.leaf[a].negative ? alg.neg(r) : r;
      break;
    }
    case CI_XPUSH:
      affine_point<N, A>(alg, cd.aff[a], pt[pp], pt[pp + 1]);
      ++pp;
      break;
    case CI_XPOP: --pp; break;
    case CI_NEG: v[sp - 1] = alg.neg(v[sp - 1]); break;
    case CI_ADD:
      v[sp - 2] = alg.add(v[sp - 2], v[sp - 1]);
      --sp;
      break;
    default:  // CI_SUB
      v[sp - 2] = alg.sub(v[sp - 2], v[sp - 1]);
      --sp;
      break;
    }
  }
  return v[0];
}
// gradient of the composite at x
template<int N, class A>
QG_HDN void composite_grad(const A &alg, const CompositeDesc &cd, const typename A::T *x, typename A::T *g)
{
  typedef typename A::T T;
  T pt[kCompStack][N], gs[kCompStack][N];
  int sp = 0, pp = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) pt[0][i] = x[i];
  for (int ip = 0; ip < cd.ninstr; ++ip) {
    const int a = cd.arg[ip];
    switch (cd.op[ip]) {
    case CI_LEAF: {
      phi_grad_raw<N>(alg, cd.leaf[a], pt[pp], gs[sp]);
      if (cd.leaf[a].negative) {
#pragma unroll
        for (int i = 0; i < N; ++i) gs[sp][i] = alg.neg(gs[sp][i]);
      }
      ++sp;
      break;
    }
    case CI_XPUSH:
      affine_point<N, A>(alg, cd.aff[a], pt[pp], pt[pp + 1]);
      ++pp;
      break;
    case CI_XPOP: {
      T t[N];
      affine_vector<N, A>(alg, cd.aff[a], gs[sp - 1], t);
#pragma unroll
      for (int i = 0; i < N; ++i) gs[sp - 1][i] = t[i];
      --pp;
      break;
    }
    case CI_NEG:
#pragma unroll
      for (int i = 0; i < N; ++i) gs[sp - 1][i] = alg.neg(gs[sp - 1][i]);
      break;
    case CI_ADD:
#pragma unroll
      for (int i = 0; i < N; ++i) gs[sp - 2][i] = alg.add(gs[sp - 2][i], gs[sp - 1][i]);
      --sp;
      break;
    default:  // CI_SUB
#pragma unroll
      for (int i = 0; i < N; ++i) gs[sp - 2][i] = alg.sub(gs[sp - 2][i], gs[sp - 1][i]);
      --sp;
      break;
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) g[i] = gs[0][i];
}

template<int N, class F> QG_HD double phi_val(const F &f, const double *x)
{
  AlgD<N> alg;
  const double v = f.k() == QG_FUNC_COMPOSITE ? composite_eval<N>(alg, *f.comp, x) : phi_eval_raw<N>(alg, f, x);
  return f.negative ? -v : v;
}
template<int N, class F> QG_HD void phi_grad(const F &f, const double *x, double *g)
{
  AlgD<N> alg;
  if (f.k() == QG_FUNC_COMPOSITE)
    composite_grad<N>(alg, *f.comp, x, g);
  else
    phi_grad_raw<N>(alg, f, x, g);
  if (f.negative) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = -g[i];
  }
}
template<int N, class F> QG_HD Iv<N> phi_val_iv(const F &f, const Iv<N> *x, const Dl<N> &dl)
{
  AlgI<N> alg;
  alg.dl = dl;
  const Iv<N> v = f.k() == QG_FUNC_COMPOSITE ? composite_eval<N>(alg, *f.comp, x) : phi_eval_raw<N>(alg, f, x);
  return f.negative ? iv_neg(v) : v;
}
template<int N, class F> QG_HD void phi_grad_iv(const F &f, const Iv<N> *x, const Dl<N> &dl, Iv<N> *g)
{
  AlgI<N> alg;
  alg.dl = dl;
  if (f.k() == QG_FUNC_COMPOSITE)
    composite_grad<N>(alg, *f.comp, x, g);
  else
    phi_grad_raw<N>(alg, f, x, g);
  if (f.negative) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = iv_neg(g[i]);
  }
}

// phi and d(phi)/dx_k along the line x + t e_k.  For the TPMS family the sin/cos of the coordinates that stay
// fixed are computed once (init) and only argument k is re-evaluated per point: 1-2 sincos per evaluation
// instead of 3 (value) + 3 (gradient).  Values are bit-identical to phi_val / phi_grad: sincos_det is a pure
// function of its argument and the combination code is the same.
// The argument index is a run-time value here (one sincos body per call site, results placed with selects).
// K >= 0 fixes the direction at compile time (every select on the direction folds away); K = -1: run-time k.
template<int N, class F, int K = -1> struct LineEval
{
  F f;
  int k;
  double x[N];
  QG_HD int dir() const { return K >= 0 ? K : k; }
  TpmsTrig<double> tv, tg;  // trig of the value-style and gradient-style arguments (tg unused if they coincide)

  QG_HD static void put(double *arr, int i, double v)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d) arr[d] = (d == i) ? v : arr[d];  // selects, not an indexed store
  }
  // trig of argument i at coordinate value xi; grad_too: also the gradient-style argument (when it differs)
  QG_HD void coord(int i, double xi, bool grad_too)
  {
    const int kind = f.k();
    const bool evs = tpms_eval_style(kind), dbl = tpms_doubled(kind);
    const double two_pi = 2.0 * kPi;
    double m = 1.0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d == i && (N == 3 || d < 2)) m = f.p[d];
    const double a = evs ? (m * xi) * two_pi : xi * (two_pi * m);
    double sn, cs;
    sincos_det(a, sn, cs);
    put(tv.s, i, sn);
    put(tv.c, i, cs);
    const double sn0 = sn, cs0 = cs;
    if (dbl) {
      sincos_det(a * 2.0, sn, cs);
      put(tv.s2, i, sn);
      put(tv.c2, i, cs);
    }
    if (evs && grad_too && !func_same_args<F>::value) {
      if ((f.same_args >> i) & 1) {
        // power-of-two frequency: the gradient-style argument xi * (two_pi * m) is the same double as `a`
        // (warp-uniform branch: the mask belongs to the function, not to the point)
        put(tg.s, i, sn0);
        put(tg.c, i, cs0);
        if (dbl) {
          put(tg.s2, i, sn);
          put(tg.c2, i, cs);
        }
      } else {
        const double ag = xi * (two_pi * m);
        sincos_det(ag, sn, cs);
        put(tg.s, i, sn);
        put(tg.c, i, cs);
        if (dbl) {
          sincos_det(ag * 2.0, sn, cs);
          put(tg.s2, i, sn);
          put(tg.c2, i, cs);
        }
      }
    }
  }
  QG_HD void init(const F &f_, const double *x_in, int k_)
  {
    f = f_;
    k = k_;
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = x_in[d];
    if (!is_tpms(f.k())) return;
    if (K >= 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i == K) continue;
        coord(i, (N == 3 || i < 2) ? x[i < N ? i : 0] : f.p[2], true);
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        if (i == k) continue;
        double xi = f.p[2];  // 2-D: the third argument is the fixed z
#pragma unroll
        for (int d = 0; d < N; ++d)
          if (d == i) xi = x[d];
        coord(i, xi, true);
      }
    }
  }
  QG_HD void move(double t)
  {
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = (d == dir()) ? t : x[d];
  }
  QG_HD double val(double t)
  {
    move(t);
    if (!is_tpms(f.k())) return phi_val<N>(f, x);
    coord(dir(), t, false);
    AlgD<N> alg;
    const double v = tpms_combine_val<AlgD<N>>(alg, f.k(), tv);
    return f.negative ? -v : v;
  }
  QG_HD void val_der(double t, double &v, double &der)
  {
    move(t);
    double g[N];
    if (f.k() == QG_FUNC_BEZIER) {
      // one de Casteljau pass gives the value (component 0, the same recurrence as the value-only pass) and the gradient
      AlgD<N> alg;
      const double *c = f.coefs;
      double out[N + 1];
      if (bezier_small<N>(f))
        BezierEvalU<N, AlgD<N>, 4>::value_der(alg, x, c, f.order, out);
      else
        BezierEval<N, AlgD<N>>::value_der(alg, x, c, f.order, out);
      v = f.negative ? -out[0] : out[0];
#pragma unroll
      for (int d = 0; d < N; ++d) g[d] = f.negative ? -out[d + 1] : out[d + 1];
    } else if (!is_tpms(f.k())) {
      v = phi_val<N>(f, x);
      phi_grad<N>(f, x, g);
    } else {
      coord(dir(), t, true);
      AlgD<N> alg;
      v = tpms_combine_val<AlgD<N>>(alg, f.k(), tv);
      tpms_combine_grad<N>(alg, f, (tpms_eval_style(f.k()) && !func_same_args<F>::value) ? tg : tv, g);
      if (f.negative) {
        v = -v;
#pragma unroll
        for (int d = 0; d < N; ++d) g[d] = -g[d];
      }
    }
    double dk = g[0];
#pragma unroll
    for (int d = 1; d < N; ++d)
      if (d == dir()) dk = g[d];
    der = dk;
  }
};

}  // namespace qg
