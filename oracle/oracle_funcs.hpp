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
//   Negative<dim>      /root/reference/cpp/qugar/impl_funcs_lib.cpp:207-227
// 2-D TPMS are the 3-D formula at a fixed z with mnq_(2) = 1 (tpms_lib.hpp:41-62, tpms_lib.cpp:35-44).
#pragma once
#include "oracle_math.hpp"

namespace orc {

enum FuncKind {
  K_SPHERE = 0,
  K_PLANE = 1,
  K_SCHOEN = 10,
  K_SCHOEN_IWP = 11,
  K_SCHOEN_FRD = 12,
  K_FISCHER_KOCH_S = 13,
  K_SCHWARZ_D = 14,
  K_SCHWARZ_P = 15
};

// params: SPHERE: {radius, center[dim]}; PLANE: {origin[dim], normal[dim]};
//         TPMS 3-D: {m, n, q}; TPMS 2-D: {m, n, z}
struct Func
{
  int kind;
  int dim;
  int negative;
  real p[8];
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

template<int N, typename T> inline T eval(const Func &f, const T *x)
{
  return f.negative ? -eval_raw<N, T>(f, x) : eval_raw<N, T>(f, x);
}
template<int N, typename T> inline void grad(const Func &f, const T *x, T *g)
{
  grad_raw<N, T>(f, x, g);
  if (f.negative)
    for (int i = 0; i < N; ++i) g[i] = -g[i];
}

}  // namespace orc
