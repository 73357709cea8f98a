// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path (qugar_b200/, include/)
// may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it.
//
// Scalar math shared by the oracle: deterministic sin/cos and Algoim's first-order
// Taylor interval type.
//
// PARITY STATUS.  The arithmetic of the reference's hot path lives in the third-party
// header library Algoim (gh:dyc0/algoim@1.0.0, /root/reference/cmake/Dependencies.cmake:14-16),
// which is NOT in /root/reference and not reachable offline.  Interval<N> below restates
// Algoim's published algorithm (R. Saye, SIAM J. Sci. Comput. 37(2), 2015, and the call
// sites in /root/reference/cpp/qugar/impl_reparam_general.cpp:107-130,676-685,715-740 and
// impl_unfitted_domain.cpp:136-159).  Its free constants were calibrated against the
// reference's own golden counts (cpp/test/test_impl_general_quad_0.cpp:61-306); see
// DESIGN.md "Oracle calibration".  Point/weight parity to the last bit against a real
// QUGaR+Algoim build is UNPINNED (no stored vectors exist upstream).
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

typedef double real;
constexpr real EPS = 2.220446049250313080847263336181640625e-16;  // numbers::eps
constexpr real NEAR_EPS = 10.0 * EPS;                              // numbers::near_eps
constexpr real PI = 3.141592653589793238462643383279502884;

// ---------------------------------------------------------------------------------
// sin/cos.  Two variants, selected at compile time:
//   ORACLE_LIBM  : glibc sin/cos, what the reference calls (std::sin on double).
//   default      : a deterministic fdlibm-style implementation written only with
//                  IEEE +,-,* and explicit fma(), so that the CUDA path can reproduce it
//                  bit for bit (CUDA's libdevice sin/cos differ from glibc in the last ulp,
//                  and classification thresholds are sensitive to exactly that).
// Domain of the deterministic variant: |x| <= 1e5 (3-term Cody-Waite); beyond that it
// falls back to libm on the host (the device path rejects such inputs).
// ---------------------------------------------------------------------------------
inline void sincos_det(real x, real *sn, real *cs)
{
  const real TWO_OVER_PI = 0x1.45f306dc9c883p-1;
  const real PIO2_HI = 0x1.921fb54442d18p+0;
  const real PIO2_MID = 0x1.1a62633145c07p-54;
  const real PIO2_LO = -0x1.f1976b7ed8fbcp-110;
  if (!(std::fabs(x) <= 1.0e5)) {
    *sn = std::sin(x);
    *cs = std::cos(x);
    return;
  }
  const real fn = std::rint(x * TWO_OVER_PI);
  real r = std::fma(-fn, PIO2_HI, x);
  r = std::fma(-fn, PIO2_MID, r);
  r = std::fma(-fn, PIO2_LO, r);
  const int q = static_cast<int>(fn) & 3;
  const real z = r * r;
  // fdlibm __kernel_sin / __kernel_cos minimax coefficients
  const real S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
             S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const real C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
             C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  real ps = std::fma(z, S6, S5);
  ps = std::fma(z, ps, S4);
  ps = std::fma(z, ps, S3);
  ps = std::fma(z, ps, S2);
  ps = std::fma(z, ps, S1);
  const real s = std::fma(r * z, ps, r);
  real pc = std::fma(z, C6, C5);
  pc = std::fma(z, pc, C4);
  pc = std::fma(z, pc, C3);
  pc = std::fma(z, pc, C2);
  pc = std::fma(z, pc, C1);
  const real hz = 0.5 * z;
  const real w = 1.0 - hz;
  const real c = w + (((1.0 - w) - hz) + (z * z) * pc);
  real so, co;
  switch (q) {
  case 0: so = s; co = c; break;
  case 1: so = c; co = -s; break;
  case 2: so = -s; co = -c; break;
  default: so = -c; co = s; break;
  }
  *sn = so;
  *cs = co;
}

#ifdef ORACLE_LIBM
inline real sin_(real x) { return std::sin(x); }
inline real cos_(real x) { return std::cos(x); }
#else
inline real sin_(real x) { real s, c; sincos_det(x, &s, &c); return s; }
inline real cos_(real x) { real s, c; sincos_det(x, &s, &c); return c; }
#endif

// ---------------------------------------------------------------------------------
// Calibration knobs (frozen after calibration; kept switchable so tests can show the
// golden counts discriminate between candidate restatements).
// ---------------------------------------------------------------------------------
struct Knobs
{
  int trig_bound = 0;      // 0: |f''|<=1 for sin/cos remainder; 1: tight max|f''| on the range
  int rootfind_maxlevel = 4;
  int dev_order = 0;       // 0: maxDeviation starts from eps; 1: adds eps last
  real seg_tol_factor = 10.0;  // degenerate-segment filter: factor * eps * extent
  real newton_tol_factor = 10.0;  // Newton-bisection tolerance: factor * eps * extent
  int seg_le = 0;          // 1: filter with <= instead of <
  int rf_cap_mode = 0;     // at rootfind cap: 0 newton anyway, 1 push midpoint, 2 nothing
  int rf_use_sign = 1;     // 1: interval sign test excludes segments
  real trig_C = 1.0;       // multiplier on the sin/cos remainder bound (when trig_bound==0)
  int newton_maxsteps = 1024;
  int tol_scale = 1;       // see line_tol() (Newton)
  int psi_default_sign = -1;  // sign decoded from a default-constructed algoim::PsiCode (unused array slots, read at the depth cap): a zero byte is side 0, sign -1
  int seg_tol_scale = 0;   // see line_tol() (degenerate-segment filter): the in-tree rule, 10 eps x extent
};
inline Knobs g_knobs_storage;
inline Knobs &knobs() { return g_knobs_storage; }

// ---------------------------------------------------------------------------------
// Interval<N>: f(x_c + y) = alpha + beta.y + [-eps, eps], |y_i| <= delta_i  (Algoim interval.hpp;
// fields alpha/maxDeviation()/uniformSign() as used at impl_reparam_general.cpp:116-117,718,740).
// ---------------------------------------------------------------------------------
template<int N> struct Interval
{
  real alpha;
  real beta[N];
  real eps;

  static real *delta_ptr()
  {
    static thread_local real d[N] = {};
    return d;
  }
  static real &delta(int i) { return delta_ptr()[i]; }

  Interval() {}
  Interval(real a) : alpha(a), eps(0.0)
  {
    for (int i = 0; i < N; ++i) beta[i] = 0.0;
  }
  // unit interval variable along direction dir centred at a
  static Interval variable(real a, int dir)
  {
    Interval r(a);
    r.beta[dir] = 1.0;
    return r;
  }

  real maxDeviation() const
  {
    const real *d = delta_ptr();
    if (knobs().dev_order == 0) {
      real b = eps;
      for (int i = 0; i < N; ++i) b += std::fabs(beta[i]) * d[i];
      return b;
    } else {
      real b = 0.0;
      for (int i = 0; i < N; ++i) b += std::fabs(beta[i]) * d[i];
      return b + eps;
    }
  }
  bool uniformSign() const { return std::fabs(alpha) > maxDeviation(); }
};

template<int N> inline Interval<N> operator-(const Interval<N> &x)
{
  Interval<N> r;
  r.alpha = -x.alpha;
  for (int i = 0; i < N; ++i) r.beta[i] = -x.beta[i];
  r.eps = x.eps;
  return r;
}
template<int N> inline Interval<N> operator+(const Interval<N> &x, const Interval<N> &y)
{
  Interval<N> r;
  r.alpha = x.alpha + y.alpha;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] + y.beta[i];
  r.eps = x.eps + y.eps;
  return r;
}
template<int N> inline Interval<N> operator-(const Interval<N> &x, const Interval<N> &y)
{
  Interval<N> r;
  r.alpha = x.alpha - y.alpha;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] - y.beta[i];
  r.eps = x.eps + y.eps;
  return r;
}
template<int N> inline Interval<N> operator+(const Interval<N> &x, real c)
{
  Interval<N> r = x;
  r.alpha = x.alpha + c;
  return r;
}
template<int N> inline Interval<N> operator-(const Interval<N> &x, real c)
{
  Interval<N> r = x;
  r.alpha = x.alpha - c;
  return r;
}
template<int N> inline Interval<N> operator*(real c, const Interval<N> &x)
{
  Interval<N> r;
  r.alpha = x.alpha * c;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] * c;
  r.eps = x.eps * std::fabs(c);
  return r;
}
template<int N> inline Interval<N> operator*(const Interval<N> &x, real c) { return c * x; }
// division by a scalar: every Taylor term divided, remainder by |c|
template<int N> inline Interval<N> operator/(const Interval<N> &x, real c)
{
  Interval<N> r;
  r.alpha = x.alpha / c;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] / c;
  r.eps = x.eps / std::fabs(c);
  return r;
}
template<int N> inline Interval<N> operator*(const Interval<N> &x, const Interval<N> &y)
{
  // (a + b.y + e1)(c + d.y + e2) = ac + (ad + cb).y + [remainder],
  // |remainder| <= dev(x) dev(y) + |a| e2 + |c| e1
  Interval<N> r;
  const real dx = x.maxDeviation(), dy = y.maxDeviation();
  r.alpha = x.alpha * y.alpha;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] * y.alpha + y.beta[i] * x.alpha;
  r.eps = dx * dy + std::fabs(x.alpha) * y.eps + std::fabs(y.alpha) * x.eps;
  return r;
}

// max over [a-d, a+d] of |sin| (used by the tight trig bound)
inline real max_abs_sin_on(real a, real d)
{
  if (d >= 0.5 * PI) return 1.0;
  // |sin| attains 1 at pi/2 + k pi
  const real lo = a - d, hi = a + d;
  const real k = std::ceil((lo - 0.5 * PI) / PI);
  if (0.5 * PI + k * PI <= hi) return 1.0;
  const real s0 = std::fabs(sin_(lo)), s1 = std::fabs(sin_(hi));
  return s0 > s1 ? s0 : s1;
}
inline real max_abs_cos_on(real a, real d) { return max_abs_sin_on(a + 0.5 * PI, d); }

template<int N> inline Interval<N> sin_(const Interval<N> &x)
{
  const real s = sin_(x.alpha), c = cos_(x.alpha);
  const real d = x.maxDeviation();
  const real C = knobs().trig_bound == 0 ? knobs().trig_C : knobs().trig_C * max_abs_sin_on(x.alpha, d);
  Interval<N> r;
  r.alpha = s;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] * c;
  r.eps = std::fabs(c) * x.eps + 0.5 * C * d * d;
  return r;
}
template<int N> inline Interval<N> cos_(const Interval<N> &x)
{
  const real s = sin_(x.alpha), c = cos_(x.alpha);
  const real d = x.maxDeviation();
  const real C = knobs().trig_bound == 0 ? knobs().trig_C : knobs().trig_C * max_abs_cos_on(x.alpha, d);
  Interval<N> r;
  r.alpha = c;
  for (int i = 0; i < N; ++i) r.beta[i] = x.beta[i] * (-s);
  r.eps = std::fabs(s) * x.eps + 0.5 * C * d * d;
  return r;
}

}  // namespace orc
