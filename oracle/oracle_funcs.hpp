// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// Level-set library restated from the reference, evaluated on double and on Interval<N> with
// the SAME expression order as the reference templates (operation order fixes rounding):
//   Sphere<dim>        /root/reference/cpp/qugar/primitive_funcs_lib.cpp:66-74
//   Plane<dim>         /root/reference/cpp/qugar/primitive_funcs_lib.cpp:849-857
//   tpms::Schoen       /root/reference/cpp/qugar/tpms_lib.cpp:83-108
//   tpms::SchoenIWP    /root/reference/cpp/qugar/tpms_lib.cpp:137-166   (incl. the pi_4_m(0) quirk at :162)
//   tpms::SchoenFRD    /root/reference/cpp/qugar/tpms_lib.cpp:205-240
//   tpms::FischerKochS /root/reference/cpp/qugar/tpms_lib.cpp:283-322
//   tpms::SchwarzDiamond   /root/reference/cpp/qugar/tpms_lib.cpp:360-391
//   tpms::SchwarzPrimitive /root/reference/cpp/qugar/tpms_lib.cpp:430-452
//   Constant, Cylinder, Ellipsoid, Annulus, Torus   /root/reference/cpp/qugar/primitive_funcs_lib.cpp:772-780,211-225,327-350,484-498,705-729
//   Negative<dim>      /root/reference/cpp/qugar/impl_funcs_lib.cpp:207-227
//   BezierTP<dim,1>    /root/reference/cpp/qugar/bezier_tp.cpp:165-255 (casteljau, casteljau_der on T = real / Interval)
// 2-D TPMS are the 3-D formula at a fixed z with mnq_(2) = 1 (tpms_lib.hpp:41-62, tpms_lib.cpp:35-44).
#pragma once
#include "oracle_math.hpp"

#include <vector>

namespace orc {

enum FuncKind {
  K_SPHERE = 0,
  K_PLANE = 1,
  K_BEZIER = 2,
  K_CONSTANT = 3,
  K_CYLINDER = 4,
  K_ELLIPSOID = 5,
  K_ANNULUS = 6,
  K_TORUS = 7,
  K_COMPOSITE = 8,
  K_DIM_LINEAR = 9,
  K_SCHOEN = 10,
  K_SCHOEN_IWP = 11,
  K_SCHOEN_FRD = 12,
  K_FISCHER_KOCH_S = 13,
  K_SCHWARZ_D = 14,
  K_SCHWARZ_P = 15,
  K_SQUARE = 16
};

// Square<dim> helpers (impl_funcs_lib.cpp:48-98): the centre value of the argument (real: the value, Interval: alpha) and
// its sign (real: (0 < v) - (v < 0), :91-98; Interval: Algoim's Interval::sign()).  Interval::sign() lives in the absent
// Algoim dependency and nothing in the reference tree exercises Square; it is restated here as the sign of the centre
// value alpha [UNVERIFIED].  (The other reading -- 0 unless the interval is uniformly signed -- makes every box whose
// centre lies near the square's centre evaluate to the constant -1, i.e. "inside", so that a grid centred on the square
// would be classified full at the root of the decomposition.)
inline real square_mid(real v) { return v; }
template<int N> inline real square_mid(const Interval<N> &v) { return v.alpha; }
inline int square_sgn(real v) { return (0.0 < v) - (v < 0.0); }
template<int N> inline int square_sgn(const Interval<N> &v) { return (0.0 < v.alpha) - (v.alpha < 0.0); }
// argmax over the directions of |centre| - 1, first maximum wins (algoim::argmax)
template<int N, typename T> inline int square_dir(const T *x)
{
  real diff[N];
  for (int d = 0; d < N; ++d) diff[d] = std::fabs(square_mid(x[d])) - 1.0;
  int ind = 0;
  for (int d = 1; d < N; ++d)
    if (diff[d] > diff[ind]) ind = d;
  return ind;
}

// params: SPHERE: {radius, center[dim]}; PLANE: {origin[dim], normal[dim]};
//         TPMS 3-D: {m, n, q}; TPMS 2-D: {m, n, z}
//         BEZIER: order[dim] + coefficients (direction 0 fastest), held by the caller (coefs points to them)
struct Composite;
struct Func
{
  int kind;
  int dim;
  int negative;
  real p[16];
  const real *coefs;
  int order[3];
  const Composite *comp;  // K_COMPOSITE
};
// TransformedFunction / Negative / AddFunctions / SubtractFunctions (impl_funcs_lib.cpp:168-290): an expression tree
// over leaf functions, as upstream (each node evaluates its children and combines them; a transformed node evaluates
// its child at transf^-1 x and maps the gradient back with the transposed matrix).
struct CompNode
{
  int op;          // 0 leaf, 1 transformed, 2 negative, 3 add, 4 subtract
  int leaf;        // op 0: leaf index
  int aff;         // op 1: affine map index (the INVERSE of the user's transformation, as the reference stores it)
  int child[2];
};
struct Composite
{
  std::vector<Func> leaves;
  std::vector<std::vector<real>> affs;  // dim*dim matrix (row major), then the translation
  std::vector<CompNode> nodes;
  int root;
};

// BezierTP::casteljau (bezier_tp.cpp:165-195): value by de Casteljau, last direction outermost; `c` walks the
// coefficient array in storage order
template<int D, typename T> struct Casteljau
{
  static T value(const T *x, const real *&c, const int *order)
  {
    const int last_order = order[D - 1];
    std::vector<T> beta((size_t)last_order);
    for (auto &b : beta) b = Casteljau<D - 1, T>::value(x, c, order);
    const T xd = x[D - 1];
    for (int j = 1; j < last_order; ++j)
      for (int k = 0; k < last_order - j; ++k) beta[(size_t)k] = beta[(size_t)k] + (beta[(size_t)k + 1] - beta[(size_t)k]) * xd;
    return beta.front();
  }
  // casteljau_der (bezier_tp.cpp:198-255): r[0] value, r[1..D] derivatives
  static std::vector<T> value_der(const T *x, const real *&c, const int *order)
  {
    const int last_order = order[D - 1];
    const int degree = last_order - 1;
    std::vector<std::vector<T>> beta((size_t)last_order);
    std::vector<T> gamma((size_t)last_order, T(0.0));
    for (auto &b : beta) b = Casteljau<D - 1, T>::value_der(x, c, order);
    if (last_order > 1) {
      gamma[0] = real(-degree) * beta.front()[0];
      gamma[1] = -beta.front()[0];
      for (int i = 1; i < degree; ++i) {
        const T b = beta[(size_t)i][0];
        gamma[(size_t)i - 1] = gamma[(size_t)i - 1] + real(last_order - i) * b;
        gamma[(size_t)i] = gamma[(size_t)i] + real(2 * i - degree) * b;
        gamma[(size_t)i + 1] = gamma[(size_t)i + 1] - real(i + 1) * b;
      }
      gamma[(size_t)degree - 1] = gamma[(size_t)degree - 1] + beta.back()[0];
      gamma[(size_t)degree] = gamma[(size_t)degree] + real(degree) * beta.back()[0];
      const T xd = x[D - 1];
      for (int j = 1; j < last_order; ++j)
        for (int k = 0; k < last_order - j; ++k) {
          for (size_t e = 0; e < beta[(size_t)k].size(); ++e)
            beta[(size_t)k][e] = beta[(size_t)k][e] + (beta[(size_t)k + 1][e] - beta[(size_t)k][e]) * xd;
          gamma[(size_t)k] = gamma[(size_t)k] + (gamma[(size_t)k + 1] - gamma[(size_t)k]) * xd;
        }
    }
    std::vector<T> r = beta.front();
    r.push_back(gamma.front());
    return r;
  }
};
template<typename T> struct Casteljau<0, T>
{
  static T value(const T *, const real *&c, const int *) { return T(*c++); }
  static std::vector<T> value_der(const T *, const real *&c, const int *) { return std::vector<T>(1, T(*c++)); }
};

template<typename T> struct TPMSArgs
{
  T a[3];      // 2 pi m x, ...
  real pm[3];  // 2 pi m, ...
};

// eval-style arguments: (2 pi) * (mnq * point)   (tpms_lib.cpp:85, 139)
template<int N, typename T> inline void tpms_args_eval(const Func &f, const T *x, T *a)
{
  const real two_pi = 2.0 * PI;
  if (N == 3) {
    for (int i = 0; i < 3; ++i) a[i] = two_pi * (f.p[i] * x[i]);
  } else {
    a[0] = two_pi * (f.p[0] * x[0]);
    a[1] = two_pi * (f.p[1] * x[1]);
    a[2] = two_pi * (1.0 * T(f.p[2]));
  }
}
// grad-style arguments: ((2 pi) * mnq) * point    (tpms_lib.cpp:92-93)
template<int N, typename T> inline void tpms_args_grad(const Func &f, const T *x, T *a, real *pm)
{
  const real two_pi = 2.0 * PI;
  if (N == 3) {
    for (int i = 0; i < 3; ++i) {
      pm[i] = two_pi * f.p[i];
      a[i] = pm[i] * x[i];
    }
  } else {
    pm[0] = two_pi * f.p[0];
    pm[1] = two_pi * f.p[1];
    pm[2] = two_pi * 1.0;
    a[0] = pm[0] * x[0];
    a[1] = pm[1] * x[1];
    a[2] = pm[2] * T(f.p[2]);
  }
}

template<int N, typename T> inline T eval_raw(const Func &f, const T *x)
{
  switch (f.kind) {
  case K_SPHERE: {
    // sqrnorm(point - center) - radius * radius
    T d0 = x[0] - f.p[1];
    T s = d0 * d0;
    for (int i = 1; i < N; ++i) {
      T di = x[i] - f.p[1 + i];
      s = s + di * di;
    }
    return s - (f.p[0] * f.p[0]);
  }
  case K_PLANE: {
    // dot(point - origin, normal)
    T s = (x[0] - f.p[0]) * f.p[N];
    for (int i = 1; i < N; ++i) s = s + (x[i] - f.p[i]) * f.p[N + i];
    return s;
  }
  case K_BEZIER: {
    const real *c = f.coefs;
    return Casteljau<N, T>::value(x, c, f.order);
  }
  case K_CONSTANT:
    return T(f.p[0]);
  case K_CYLINDER: {
    // params: radius, origin[3], unit axis[3];  dot(diff, diff) - proj * proj - radius^2
    T diff[3];
    for (int i = 0; i < 3; ++i) diff[i] = x[i < N ? i : 0] - f.p[1 + i];
    T dd = diff[0] * diff[0];
    T proj = diff[0] * f.p[4];
    for (int i = 1; i < 3; ++i) {
      dd = dd + diff[i] * diff[i];
      proj = proj + diff[i] * f.p[4 + i];
    }
    return (dd - proj * proj) - (f.p[0] * f.p[0]);
  }
  case K_ELLIPSOID: {
    // params: semi_axes[N], origin[N], basis[N][N];  -1 + sum (dot(axis_d, diff) / a_d)^2
    T diff[N];
    for (int i = 0; i < N; ++i) diff[i] = x[i] - f.p[N + i];
    T val(-1.0);
    for (int dir = 0; dir < N; ++dir) {
      const real *ax = f.p + 2 * N + N * dir;
      T d = diff[0] * ax[0];
      for (int i = 1; i < N; ++i) d = d + diff[i] * ax[i];
      const T proj = d / f.p[dir];
      val = val + proj * proj;
    }
    return val;
  }
  case K_ANNULUS: {
    // params: inner radius, outer radius, center[2]
    const T d0 = x[0] - f.p[2], d1 = x[1] - f.p[3];
    const T dist2 = d0 * d0 + d1 * d1;
    const real ri2 = f.p[0] * f.p[0], ro2 = f.p[1] * f.p[1];
    return (((dist2 - ri2) - ro2) * dist2) + T(ri2 * ro2);
  }
  case K_SQUARE: {
    // Square<dim>::eval_ (impl_funcs_lib.cpp:48-66): max_d |x_d| - 1 on [-1,1]^dim; real: max(diff) = |x_ind| - 1, which is the
    // Interval expression x_ind * sgn(x_ind) - 1 evaluated on a double (the product by +-1 / 0 is exact)
    const int ind = square_dir<N, T>(x);
    return x[ind] * real(square_sgn(x[ind])) - 1.0;
  }
  case K_DIM_LINEAR: {
    // DimLinear<dim>::eval_ (impl_funcs_lib.cpp:108-124): params = 2^dim vertex values
    const real *c = f.p;
    const T one(1.0);
    const T lower = (c[0] * (one - x[0]) + c[1] * x[0]) * (one - x[1]) + (c[2] * (one - x[0]) + c[3] * x[0]) * x[1];
    if (N == 2) return lower;
    const T upper = (c[4] * (one - x[0]) + c[5] * x[0]) * (one - x[1]) + (c[6] * (one - x[0]) + c[7] * x[0]) * x[1];
    return lower * (one - x[N - 1]) + upper * x[N - 1];
  }
  case K_TORUS: {
    // params: major radius, minor radius, center[3], unit axis[3]
    const real R2 = f.p[0] * f.p[0], r2 = f.p[1] * f.p[1];
    T diff[3];
    for (int i = 0; i < 3; ++i) diff[i] = x[i < N ? i : 0] - f.p[2 + i];
    T dist2 = diff[0] * diff[0];
    T proj = diff[0] * f.p[5];
    for (int i = 1; i < 3; ++i) {
      dist2 = dist2 + diff[i] * diff[i];
      proj = proj + diff[i] * f.p[5 + i];
    }
    T v = dist2 * dist2 - dist2 * (2.0 * (R2 + r2));
    v = v + (proj * (4.0 * R2)) * proj;
    return v + T((R2 - r2) * (R2 - r2));
  }
  case K_SCHOEN: {
    T a[3];
    tpms_args_eval<N, T>(f, x, a);
    return sin_(a[0]) * cos_(a[1]) + sin_(a[1]) * cos_(a[2]) + sin_(a[2]) * cos_(a[0]);
  }
  case K_SCHOEN_IWP: {
    T a[3], b[3];
    tpms_args_eval<N, T>(f, x, a);
    for (int i = 0; i < 3; ++i) b[i] = 2.0 * a[i];
    return 2.0 * (cos_(a[0]) * cos_(a[1]) + cos_(a[1]) * cos_(a[2]) + cos_(a[2]) * cos_(a[0])) - cos_(b[0])
           - cos_(b[1]) - cos_(b[2]);
  }
  case K_SCHOEN_FRD: {
    T a[3], b[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    for (int i = 0; i < 3; ++i) b[i] = 2.0 * a[i];
    return 4.0 * cos_(a[0]) * cos_(a[1]) * cos_(a[2]) - cos_(b[0]) * cos_(b[1]) - cos_(b[1]) * cos_(b[2])
           - cos_(b[2]) * cos_(b[0]);
  }
  case K_FISCHER_KOCH_S: {
    T a[3], b[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    for (int i = 0; i < 3; ++i) b[i] = 2.0 * a[i];
    return cos_(b[0]) * sin_(a[1]) * cos_(a[2]) + cos_(a[0]) * cos_(b[1]) * sin_(a[2])
           + sin_(a[0]) * cos_(a[1]) * cos_(b[2]);
  }
  case K_SCHWARZ_D: {
    T a[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    return cos_(a[0]) * cos_(a[1]) * cos_(a[2]) - sin_(a[0]) * sin_(a[1]) * sin_(a[2]);
  }
  case K_SCHWARZ_P: {
    T a[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    return cos_(a[0]) + cos_(a[1]) + cos_(a[2]);
  }
  default:
    return T(0.0);
  }
}

template<int N, typename T> inline void grad_raw(const Func &f, const T *x, T *g)
{
  switch (f.kind) {
  case K_SPHERE:
    for (int i = 0; i < N; ++i) g[i] = 2.0 * (x[i] - f.p[1 + i]);
    return;
  case K_PLANE:
    for (int i = 0; i < N; ++i) g[i] = T(f.p[N + i]);
    return;
  case K_BEZIER: {
    const real *c = f.coefs;
    const std::vector<T> r = Casteljau<N, T>::value_der(x, c, f.order);
    for (int i = 0; i < N; ++i) g[i] = r[(size_t)i + 1];
    return;
  }
  case K_CONSTANT:
    for (int i = 0; i < N; ++i) g[i] = T(0.0);
    return;
  case K_CYLINDER: {
    T diff[3];
    for (int i = 0; i < 3; ++i) diff[i] = x[i < N ? i : 0] - f.p[1 + i];
    T proj = diff[0] * f.p[4];
    for (int i = 1; i < 3; ++i) proj = proj + diff[i] * f.p[4 + i];
    const T p2 = proj * 2.0;
    for (int i = 0; i < N; ++i) g[i] = diff[i < 3 ? i : 0] * 2.0 - p2 * f.p[4 + i];
    return;
  }
  case K_ELLIPSOID: {
    T diff[N];
    for (int i = 0; i < N; ++i) {
      diff[i] = x[i] - f.p[N + i];
      g[i] = T(0.0);
    }
    for (int dir = 0; dir < N; ++dir) {
      const real coef = 2.0 / f.p[dir] / f.p[dir];
      const real *ax = f.p + 2 * N + N * dir;
      T d = diff[0] * ax[0];
      for (int i = 1; i < N; ++i) d = d + diff[i] * ax[i];
      const T cd = d * coef;
      for (int i = 0; i < N; ++i) g[i] = g[i] + cd * ax[i];
    }
    return;
  }
  case K_ANNULUS: {
    const T d0 = x[0] - f.p[2], d1 = x[1] - f.p[3];
    const real ri2 = f.p[0] * f.p[0], ro2 = f.p[1] * f.p[1];
    const T c = ((d0 * d0 + d1 * d1) * 4.0 - ri2) - ro2;
    g[0] = c * d0;
    g[1] = c * d1;
    return;
  }
  case K_SQUARE: {
    // Square<dim>::grad_ (impl_funcs_lib.cpp:68-82): sgn(x_ind) e_ind
    const int ind = square_dir<N, T>(x);
    for (int d = 0; d < N; ++d) g[d] = T(0.0);
    g[ind] = T(real(square_sgn(x[ind])));
    return;
  }
  case K_DIM_LINEAR: {
    // DimLinear<dim>::grad_ (impl_funcs_lib.cpp:127-152)
    const real *c = f.p;
    const T one(1.0);
    if (N == 2) {
      g[0] = (c[1] - c[0]) * (one - x[1]) + (c[3] - c[2]) * x[1];
      g[1] = (c[2] - c[0]) * (one - x[0]) + (c[3] - c[1]) * x[0];
      return;
    }
    const T x2 = x[N - 1];
    g[0] = ((c[1] - c[0]) * (one - x[1]) + (c[3] - c[2]) * x[1]) * (one - x2) + ((c[5] - c[4]) * (one - x[1]) + (c[7] - c[6]) * x[1]) * x2;
    g[1] = ((c[2] - c[0]) * (one - x[0]) + (c[3] - c[1]) * x[0]) * (one - x2) + ((c[6] - c[4]) * (one - x[0]) + (c[7] - c[5]) * x[0]) * x2;
    g[N - 1] = ((c[4] - c[0]) * (one - x[0]) + (c[5] - c[1]) * x[0]) * (one - x[1]) + ((c[6] - c[2]) * (one - x[0]) + (c[7] - c[3]) * x[0]) * x[1];
    return;
  }
  case K_TORUS: {
    const real R2 = f.p[0] * f.p[0], r2 = f.p[1] * f.p[1];
    T diff[3];
    for (int i = 0; i < 3; ++i) diff[i] = x[i < N ? i : 0] - f.p[2 + i];
    T dist2 = diff[0] * diff[0];
    T proj = diff[0] * f.p[5];
    for (int i = 1; i < 3; ++i) {
      dist2 = dist2 + diff[i] * diff[i];
      proj = proj + diff[i] * f.p[5 + i];
    }
    const T c = ((dist2 - R2) - r2) * 4.0;
    const T pa = proj * (8.0 * R2);
    for (int i = 0; i < N; ++i) g[i] = c * diff[i < 3 ? i : 0] + pa * f.p[5 + i];
    return;
  }
  case K_SCHOEN: {
    T a[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    g[0] = (cos_(a[0]) * cos_(a[1]) - sin_(a[0]) * sin_(a[2])) * pm[0];
    g[1] = (cos_(a[1]) * cos_(a[2]) - sin_(a[1]) * sin_(a[0])) * pm[1];
    if (N == 3) g[N - 1] = (cos_(a[2]) * cos_(a[0]) - sin_(a[2]) * sin_(a[1])) * pm[2];
    return;
  }
  case K_SCHOEN_IWP: {
    T a[3], b[3];
    real pm[3], p4[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    for (int i = 0; i < 3; ++i) {
      p4[i] = 2.0 * pm[i];
      b[i] = 2.0 * a[i];
    }
    g[0] = (sin_(b[0]) - sin_(a[0]) * (cos_(a[1]) + cos_(a[2]))) * p4[0];
    g[1] = (sin_(b[1]) - sin_(a[1]) * (cos_(a[0]) + cos_(a[2]))) * p4[0];  // sic: pi_4_m(0), tpms_lib.cpp:162
    if (N == 3) g[N - 1] = (sin_(b[2]) - sin_(a[2]) * (cos_(a[0]) + cos_(a[1]))) * p4[2];
    return;
  }
  case K_SCHOEN_FRD: {
    T a[3], b[3];
    real pm[3], p4[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    for (int i = 0; i < 3; ++i) {
      p4[i] = 2.0 * pm[i];
      b[i] = 2.0 * a[i];
    }
    g[0] = (-2.0 * sin_(a[0]) * cos_(a[1]) * cos_(a[2]) + sin_(b[0]) * cos_(b[1]) + cos_(b[2]) * sin_(b[0])) * p4[0];
    g[1] = (-2.0 * cos_(a[0]) * sin_(a[1]) * cos_(a[2]) + cos_(b[0]) * sin_(b[1]) + sin_(b[1]) * cos_(b[2])) * p4[1];
    if (N == 3)
      g[N - 1] =
        (-2.0 * cos_(a[0]) * cos_(a[1]) * sin_(a[2]) + cos_(b[1]) * sin_(b[2]) + sin_(b[2]) * cos_(b[0])) * p4[2];
    return;
  }
  case K_FISCHER_KOCH_S: {
    T a[3], b[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    for (int i = 0; i < 3; ++i) b[i] = 2.0 * a[i];
    g[0] = (-2.0 * sin_(b[0]) * sin_(a[1]) * cos_(a[2]) - sin_(a[0]) * cos_(b[1]) * sin_(a[2])
            + cos_(a[0]) * cos_(a[1]) * cos_(b[2]))
           * pm[0];
    g[1] = (cos_(b[0]) * cos_(a[1]) * cos_(a[2]) - 2.0 * cos_(a[0]) * sin_(b[1]) * sin_(a[2])
            - sin_(a[0]) * sin_(a[1]) * cos_(b[2]))
           * pm[1];
    if (N == 3)
      g[N - 1] = (-cos_(b[0]) * sin_(a[1]) * sin_(a[2]) + cos_(a[0]) * cos_(b[1]) * cos_(a[2])
                  - 2.0 * sin_(a[0]) * cos_(a[1]) * sin_(b[2]))
                 * pm[2];
    return;
  }
  case K_SCHWARZ_D: {
    T a[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    g[0] = -(sin_(a[0]) * cos_(a[1]) * cos_(a[2]) + cos_(a[0]) * sin_(a[1]) * sin_(a[2])) * pm[0];
    g[1] = -(cos_(a[0]) * sin_(a[1]) * cos_(a[2]) + sin_(a[0]) * cos_(a[1]) * sin_(a[2])) * pm[1];
    if (N == 3) g[N - 1] = -(cos_(a[0]) * cos_(a[1]) * sin_(a[2]) + sin_(a[0]) * sin_(a[1]) * cos_(a[2])) * pm[2];
    return;
  }
  case K_SCHWARZ_P: {
    T a[3];
    real pm[3];
    tpms_args_grad<N, T>(f, x, a, pm);
    g[0] = -sin_(a[0]) * pm[0];
    g[1] = -sin_(a[1]) * pm[1];
    if (N == 3) g[N - 1] = -sin_(a[2]) * pm[2];
    return;
  }
  default:
    for (int i = 0; i < N; ++i) g[i] = T(0.0);
  }
}

// AffineTransf::transform_point / transform_vector (affine_transf.cpp:345-377)
template<int N, typename T> inline void affine_point(const real *c, const T *x, T *y)
{
  for (int i = 0; i < N; ++i) {
    T v = c[N * i] * x[0];
    for (int j = 1; j < N; ++j) v = v + c[N * i + j] * x[j];
    y[i] = v + c[N * N + i];
  }
}
template<int N, typename T> inline void affine_vector(const real *c, const T *v, T *out)
{
  for (int i = 0; i < N; ++i) {
    T r = c[i] * v[0];
    for (int j = 1; j < N; ++j) r = r + c[N * j + i] * v[j];
    out[i] = r;
  }
}
// recursive evaluation of the expression tree, node by node like the reference's virtual calls
template<int N, typename T> inline T comp_eval(const Composite &c, int node, const T *x)
{
  const CompNode &nd = c.nodes[(size_t)node];
  switch (nd.op) {
  case 0: {
    const Func &lf = c.leaves[(size_t)nd.leaf];
    const T v = eval_raw<N, T>(lf, x);
    return lf.negative ? -v : v;
  }
  case 1: {
    T y[N];
    affine_point<N, T>(c.affs[(size_t)nd.aff].data(), x, y);
    return comp_eval<N, T>(c, nd.child[0], y);
  }
  case 2: return -comp_eval<N, T>(c, nd.child[0], x);
  case 3: return comp_eval<N, T>(c, nd.child[0], x) + comp_eval<N, T>(c, nd.child[1], x);
  default: return comp_eval<N, T>(c, nd.child[0], x) - comp_eval<N, T>(c, nd.child[1], x);
  }
}
template<int N, typename T> inline void comp_grad(const Composite &c, int node, const T *x, T *g)
{
  const CompNode &nd = c.nodes[(size_t)node];
  switch (nd.op) {
  case 0: {
    const Func &lf = c.leaves[(size_t)nd.leaf];
    grad_raw<N, T>(lf, x, g);
    if (lf.negative)
      for (int i = 0; i < N; ++i) g[i] = -g[i];
    return;
  }
  case 1: {
    T y[N], gl[N];
    affine_point<N, T>(c.affs[(size_t)nd.aff].data(), x, y);
    comp_grad<N, T>(c, nd.child[0], y, gl);
    affine_vector<N, T>(c.affs[(size_t)nd.aff].data(), gl, g);
    return;
  }
  case 2:
    comp_grad<N, T>(c, nd.child[0], x, g);
    for (int i = 0; i < N; ++i) g[i] = -g[i];
    return;
  default: {
    T g1[N];
    comp_grad<N, T>(c, nd.child[0], x, g);
    comp_grad<N, T>(c, nd.child[1], x, g1);
    for (int i = 0; i < N; ++i) g[i] = nd.op == 3 ? g[i] + g1[i] : g[i] - g1[i];
    return;
  }
  }
}

template<int N, typename T> inline T eval(const Func &f, const T *x)
{
  if (f.kind == K_COMPOSITE) {
    const T v = comp_eval<N, T>(*f.comp, f.comp->root, x);
    return f.negative ? -v : v;
  }
  return f.negative ? -eval_raw<N, T>(f, x) : eval_raw<N, T>(f, x);
}
template<int N, typename T> inline void grad(const Func &f, const T *x, T *g)
{
  if (f.kind == K_COMPOSITE)
    comp_grad<N, T>(*f.comp, f.comp->root, x, g);
  else
    grad_raw<N, T>(f, x, g);
  if (f.negative)
    for (int i = 0; i < N; ++i) g[i] = -g[i];
}

}  // namespace orc
