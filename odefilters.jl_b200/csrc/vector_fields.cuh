// Built-in vector-field catalogue and the truncated Taylor arithmetic used for the exact
// initial state (reference: src/state_initialization.jl:15-42 uses TaylorSeries.jl on the host;
// src/jacobian.jl:6-22 builds symbolic Jacobians with ModelingToolkit -- here f and J are
// device code selected by the PNDE_VF_* enum of include/pnde.h).
#pragma once
#ifndef PNDE_UNROLL  // (cov_engine.cuh defines it; stand-alone includes unroll)
#define PNDE_UNROLL _Pragma("unroll")
#endif
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

namespace pnde {

// ---------------------------------------------------------------------------------------------
// Jet<N>: c[0] + c[1] tau + ... + c[N-1] tau^(N-1)
// ---------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double c[N];
};

template <int N>
__device__ __forceinline__ Jet<N> operator+(const Jet<N>& a, const Jet<N>& b) {
  Jet<N> r;
PNDE_UNROLL
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] + b.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator-(const Jet<N>& a, const Jet<N>& b) {
  Jet<N> r;
PNDE_UNROLL
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] - b.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator-(const Jet<N>& a) {
  Jet<N> r;
PNDE_UNROLL
  for (int i = 0; i < N; ++i) r.c[i] = -a.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator+(const Jet<N>& a, double b) {
  Jet<N> r = a;
  r.c[0] += b;
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator+(double b, const Jet<N>& a) {
  return a + b;
}
template <int N>
__device__ __forceinline__ Jet<N> operator-(const Jet<N>& a, double b) {
  Jet<N> r = a;
  r.c[0] -= b;
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator-(double b, const Jet<N>& a) {
  Jet<N> r = -a;
  r.c[0] += b;
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator*(const Jet<N>& a, double b) {
  Jet<N> r;
PNDE_UNROLL
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] * b;
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator*(double b, const Jet<N>& a) {
  return a * b;
}
template <int N>
__device__ __forceinline__ Jet<N> operator/(const Jet<N>& a, double b) {
  Jet<N> r;
PNDE_UNROLL
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] / b;
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator*(const Jet<N>& a, const Jet<N>& b) {  // Cauchy product
  Jet<N> r;
PNDE_UNROLL
  for (int k = 0; k < N; ++k) {
    double s = a.c[0] * b.c[k];
PNDE_UNROLL
    for (int i = 1; i <= k; ++i) s = fma(a.c[i], b.c[k - i], s);
    r.c[k] = s;
  }
  return r;
}


// keep the double versions visible next to the Jet overloads declared in this namespace
using ::cos;
using ::exp;
using ::log;
using ::sin;
using ::sqrt;

// Quotients and elementary functions of jets (standard power-series recurrences), so that run-time
// compiled user vector fields (rtc_model.cu) can use / exp log sin cos sqrt and still get the exact
// Taylor-mode initialisation.  The double overloads are CUDA's own.
template <int N>
__device__ __forceinline__ Jet<N> operator/(const Jet<N>& a, const Jet<N>& b) {
  Jet<N> r;
  const double ib = 1.0 / b.c[0];
PNDE_UNROLL
  for (int k = 0; k < N; ++k) {
    double s = a.c[k];
PNDE_UNROLL
    for (int i = 1; i <= k; ++i) s = fma(-b.c[i], r.c[k - i], s);
    r.c[k] = s * ib;
  }
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> operator/(double a, const Jet<N>& b) {
  Jet<N> x;
PNDE_UNROLL
  for (int i = 0; i < N; ++i) x.c[i] = 0.0;
  x.c[0] = a;
  return x / b;
}
template <int N>
__device__ __forceinline__ Jet<N> exp(const Jet<N>& a) {
  Jet<N> r;
  r.c[0] = ::exp(a.c[0]);
PNDE_UNROLL
  for (int k = 1; k < N; ++k) {
    double s = 0.0;
PNDE_UNROLL
    for (int i = 1; i <= k; ++i) s = fma(double(i) * a.c[i], r.c[k - i], s);
    r.c[k] = s / double(k);
  }
  return r;
}
template <int N>
__device__ __forceinline__ Jet<N> log(const Jet<N>& a) {
  Jet<N> r;
  r.c[0] = ::log(a.c[0]);
  const double ia = 1.0 / a.c[0];
PNDE_UNROLL
  for (int k = 1; k < N; ++k) {
    double s = 0.0;
PNDE_UNROLL
    for (int i = 1; i < k; ++i) s = fma(double(i) * r.c[i], a.c[k - i], s);
    r.c[k] = (a.c[k] - s / double(k)) * ia;
  }
  return r;
}
template <int N>
__device__ __forceinline__ void sincos_jet(const Jet<N>& a, Jet<N>& sn, Jet<N>& cs) {
  sn.c[0] = ::sin(a.c[0]);
  cs.c[0] = ::cos(a.c[0]);
PNDE_UNROLL
  for (int k = 1; k < N; ++k) {
    double ss = 0.0, cc = 0.0;
PNDE_UNROLL
    for (int i = 1; i <= k; ++i) {
      ss = fma(double(i) * a.c[i], cs.c[k - i], ss);
      cc = fma(double(i) * a.c[i], sn.c[k - i], cc);
    }
    sn.c[k] = ss / double(k);
    cs.c[k] = -cc / double(k);
  }
}
template <int N>
__device__ __forceinline__ Jet<N> sin(const Jet<N>& a) {
  Jet<N> s, c;
  sincos_jet(a, s, c);
  return s;
}
template <int N>
__device__ __forceinline__ Jet<N> cos(const Jet<N>& a) {
  Jet<N> s, c;
  sincos_jet(a, s, c);
  return c;
}
template <int N>
__device__ __forceinline__ Jet<N> sqrt(const Jet<N>& a) {
  Jet<N> r;
  r.c[0] = ::sqrt(a.c[0]);
  const double ih = 0.5 / r.c[0];
PNDE_UNROLL
  for (int k = 1; k < N; ++k) {
    double s = a.c[k];
PNDE_UNROLL
    for (int i = 1; i < k; ++i) s = fma(-r.c[i], r.c[k - i], s);
    r.c[k] = s * ih;
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Catalogue.  f is generic in the scalar type (double or Jet); all fields are autonomous (the
// reference asserts this at src/state_initialization.jl:20-22).  jac is row-major J[i][j]=df_i/du_j.
// ---------------------------------------------------------------------------------------------
struct VfFhnReadme {  // README.md:36-40
  static constexpr int d = 2, np = 3, kind = 0;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    const double b = p[1], c = p[2], a = p[0];
    du[0] = c * (u[0] - u[0] * u[0] * u[0] / 3.0 + u[1]);
    du[1] = -(1.0 / c) * (u[0] - a - b * u[1]);
  }
  __device__ __forceinline__ static void jac(const double* u, const double* p, double (*J)[2]) {
    const double b = p[1], c = p[2];
    J[0][0] = c * (1.0 - u[0] * u[0]);
    J[0][1] = c;
    J[1][0] = -(1.0 / c);
    J[1][1] = b / c;
  }
};

struct VfFhnLib {  // DiffEqProblemLibrary prob_ode_fitzhughnagumo
  static constexpr int d = 2, np = 4, kind = 1;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    const double a = p[0], b = p[1], tauinv = p[2], l = p[3];
    du[0] = u[0] - u[0] * u[0] * u[0] / 3.0 - u[1] + l;
    du[1] = tauinv * (u[0] + a - b * u[1]);
  }
  __device__ __forceinline__ static void jac(const double* u, const double* p, double (*J)[2]) {
    J[0][0] = 1.0 - u[0] * u[0];
    J[0][1] = -1.0;
    J[1][0] = p[2];
    J[1][1] = -p[2] * p[1];
  }
};

struct VfLotkaVolterra {  // prob_ode_lotkavoltera
  static constexpr int d = 2, np = 4, kind = 2;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
    du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
  }
  __device__ __forceinline__ static void jac(const double* u, const double* p, double (*J)[2]) {
    J[0][0] = p[0] - p[1] * u[1];
    J[0][1] = -p[1] * u[0];
    J[1][0] = p[3] * u[1];
    J[1][1] = -p[2] + p[3] * u[0];
  }
};

struct VfVanDerPol {  // prob_ode_vanstiff, u = (y, x)
  static constexpr int d = 2, np = 1, kind = 3;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    du[0] = p[0] * ((1.0 - u[1] * u[1]) * u[0] - u[1]);
    du[1] = u[0];
  }
  __device__ __forceinline__ static void jac(const double* u, const double* p, double (*J)[2]) {
    J[0][0] = p[0] * (1.0 - u[1] * u[1]);
    J[0][1] = p[0] * (-2.0 * u[1] * u[0] - 1.0);
    J[1][0] = 1.0;
    J[1][1] = 0.0;
  }
};

struct VfLinear2 {  // du_i = p_i u_i (test/state_init.jl:15)
  static constexpr int d = 2, np = 2, kind = 4;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    du[0] = p[0] * u[0];
    du[1] = p[1] * u[1];
  }
  __device__ __forceinline__ static void jac(const double*, const double* p, double (*J)[2]) {
    J[0][0] = p[0];
    J[0][1] = 0.0;
    J[1][0] = 0.0;
    J[1][1] = p[1];
  }
};

struct VfLogistic {  // du = p u (1 - u)  (test/specific_problems.jl:62)
  static constexpr int d = 1, np = 1, kind = 5;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    du[0] = p[0] * u[0] * (1.0 - u[0]);
  }
  __device__ __forceinline__ static void jac(const double* u, const double* p, double (*J)[1]) {
    J[0][0] = p[0] * (1.0 - 2.0 * u[0]);
  }
};

struct VfLinear1 {  // du = p u (test/convergence.jl:10)
  static constexpr int d = 1, np = 1, kind = 7;
  template <class T>
  __device__ __forceinline__ static void f(const T* u, const double* p, T* du) {
    du[0] = p[0] * u[0];
  }
  __device__ __forceinline__ static void jac(const double*, const double* p, double (*J)[1]) { J[0][0] = p[0]; }
};

// ---------------------------------------------------------------------------------------------
// initial_update! (src/state_initialization.jl:2-14): mu0 = [u0; u'(t0); ...; u^(q)(t0)] in
// natural coordinates.  Time-Taylor coefficients c_{k+1} = [f(c(tau))]_k / (k+1); u^(k) = k! c_k.
// ---------------------------------------------------------------------------------------------
template <class VF, int q>
__device__ __forceinline__ void taylor_init(const double* u0, const double* p, double* m /* [d(q+1)] */) {
  constexpr int d = VF::d;
  Jet<q + 1> x[d];
PNDE_UNROLL
  for (int i = 0; i < d; ++i) {
PNDE_UNROLL
    for (int k = 0; k <= q; ++k) x[i].c[k] = 0.0;
    x[i].c[0] = u0[i];
  }
PNDE_UNROLL
  for (int k = 0; k < q; ++k) {
    Jet<q + 1> fx[d];
    VF::template f<Jet<q + 1>>(x, p, fx);
PNDE_UNROLL
    for (int i = 0; i < d; ++i) x[i].c[k + 1] = fx[i].c[k] / double(k + 1);
  }
  double fact = 1.0;
PNDE_UNROLL
  for (int k = 0; k <= q; ++k) {
    if (k > 0) fact *= double(k);
PNDE_UNROLL
    for (int i = 0; i < d; ++i) m[k * d + i] = fact * x[i].c[k];
  }
}

}  // namespace pnde
