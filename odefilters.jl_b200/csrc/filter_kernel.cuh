// The persistent ensemble filter kernel: one trajectory per thread, the whole solve! loop of
// OrdinaryDiffEq on the device with lockstep, masked per-trajectory step control.
//
// Reference path being replaced (relative to the reference checkout):
//   initialize!      src/perform_step.jl:2-12  (+ src/state_initialization.jl:2-53)
//   perform_step!    src/perform_step.jl:27-93
//   measure!         src/perform_step.jl:95-132
//   estimate_errors  src/perform_step.jl:148-158
//   estimate_diffusion src/diffusions.jl:11-153
//   savevalues!      src/integrator_utils.jl:33-48
//   controller       src/alg_utils.jl:13-24 + OrdinaryDiffEq (external)
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#ifndef __CUDACC_RTC__
#include <math.h>
#endif

#include "cov_engine.cuh"
#include "vector_fields.cuh"

namespace pnde {

enum { DIFF_DYNAMIC = 0, DIFF_FIXED = 1, DIFF_FIXED_MAP = 2, DIFF_DYNAMIC_MV = 3, DIFF_FIXED_MV = 4 };
enum { SAVE_FINAL = 0, SAVE_EVERY = 1, SAVE_STRIDE = 2 };
enum { RET_SUCCESS = 0, RET_MAXITERS = 1, RET_DTNAN = 2, RET_NONFINITE = 3, RET_HISTORY_FULL = 4, RET_DTMIN = 5,
       RET_ZERO_RESIDUAL = 6 };
enum { FLAG_REFERENCE_QUIRKS = 1, FLAG_ONE_THREAD = 2 };

struct CtrlParams {
  double abstol, reltol, dt, t0, t1;
  double qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit, beta1, beta2, dtmin, dtmax;
  long long maxiters;
};

// IEKS (src/ieks.jl, src/perform_step.jl:111-113): the previous iterate's filtered + smoothed history, whose dense
// output sol(t + dt) is where the Jacobian of the current iterate is evaluated.  hist == nullptr: first iterate.
struct LinParams {
  const double* hist;
  const double* smooth;
  const int* n_saved;
  const double* final_diff;
  long long max_saved;
  int calibrate, is_mv;
};

struct FilterParams {
  long long n;       // trajectories in the ensemble = stride of every structure-of-arrays buffer
  long long first;   // this launch handles trajectories [first, first + count) (pipelined host transfers)
  long long count;
  const double* u0;  // [d][n]
  const double* p;   // [np][n]
  double* mean;      // [D][n]
  double* cov;       // [D(D+1)/2][n]
  double* t_final;   // [n]
  double* loglik;    // [n]
  double* final_diff;  // [ND][n]
  int* retcode;
  int* naccept;
  int* nreject;
  int* nf;
  int* njacs;
  int* n_saved;
  double* hist;  // [max_saved][REC][n]
  long long max_saved;
  int save_mode, save_stride, diffusion;
  int flags;  // FLAG_*
  IwpConsts C;
  CtrlParams K;
  LinParams lin;
};

// Linearisation-point policy of filter_kernel: the default evaluates J at the predicted mean (EK1).
struct NoLin {
  static constexpr bool enabled = false;
};

// P(h) block scales h^(k-q-1/2) (src/preconditioning.jl:4-13) and their inverses.
template <int q>
__device__ __forceinline__ void precond_scales(double h, double (&P)[q + 1], double (&PI)[q + 1]) {
  // PI_k = h^(q+1/2-k): PI_q = sqrt(h), PI_{k-1} = PI_k * h
  double v = sqrt(h);
  PI[q] = v;
PNDE_UNROLL
  for (int k = q - 1; k >= 0; --k) {
    v *= h;
    PI[k] = v;
  }
  // P_k = 1 / PI_k with two divisions instead of q + 1: P_q = 1 / sqrt(h), P_{k-1} = P_k / h
  const double ih = 1.0 / h;
  double w = 1.0 / PI[q];
  P[q] = w;
PNDE_UNROLL
  for (int k = q - 1; k >= 0; --k) {
    w *= ih;
    P[k] = w;
  }
}

// Step-size factor of the PI controller (OrdinaryDiffEq stepsize_controller!):
//   qc = clamp(EEst^beta1 / qold^beta2 / gamma, 1/qmax, 1/qmin),  qc = 1/qmax for EEst == 0.
// Two arithmetic variants, selected at build time:
//   PNDE_CTRL_POW = 1  two pow calls, the way the reference evaluates it (Julia's ^ is the libm pow);
//   PNDE_CTRL_POW = 0  exp/log with log qold carried by the caller: one log and one exp per attempted step (2-3 ulp
//                      instead of <= 2, a third of the instructions).
// The carried state `qs` is qold (pow) or log qold (exp/log); `lE` is scratch for the second variant.  Which one ships
// was decided by the ensemble-scale count statistics of tests/test_gpu_parity.py::test_adaptive_ensemble_count_parity
// (DESIGN.md section 2).
#ifndef PNDE_CTRL_POW
#define PNDE_CTRL_POW 0
#endif
__device__ __forceinline__ double ctrl_state_init(const CtrlParams& K) {
  return PNDE_CTRL_POW ? K.qoldinit : log(K.qoldinit);
}
__device__ __forceinline__ double controller_factor(double EEst, const CtrlParams& K, double qs, double& lE) {
  lE = __longlong_as_double(0xfff0000000000000LL);
  if (EEst == 0.0) return 1.0 / K.qmax;
#if PNDE_CTRL_POW
  const double qc = pow(EEst, K.beta1) / pow(qs, K.beta2);
#else
  lE = log(EEst);
  const double qc = exp(K.beta1 * lE - K.beta2 * qs);
#endif
  return fmax(1.0 / K.qmax, fmin(1.0 / K.qmin, qc / K.gamma));
}
// qold = max(EEst, qoldinit) after an accepted step
__device__ __forceinline__ double ctrl_state_accept(double EEst, double lE, double qs0, const CtrlParams& K) {
  return PNDE_CTRL_POW ? fmax(EEst, K.qoldinit) : fmax(lE, qs0);
}
// q11 = EEst^beta1, formed only after a rejection
__device__ __forceinline__ double ctrl_q11(double EEst, double lE, const CtrlParams& K) {
#if PNDE_CTRL_POW
  return pow(EEst, K.beta1);
#else
  return exp(K.beta1 * lE);
#endif
}

__device__ __forceinline__ double ulp_of(double x) {  // Julia eps(x)
  x = fabs(x);
  if (x == 0.0) return 4.9406564584124654e-324;
  const long long bits = __double_as_longlong(x);
  return __longlong_as_double(bits + 1) - x;
}

// "Sliver" intervals.  OrdinaryDiffEq advances t += dt and snaps to t1 only within 10 ulp, so a fixed-step run whose dt
// is not a binary fraction ends with one more step over what the recursion lost to rounding (test/diffusions.jl with
// dt = 1e-4: a last step of 9.4e-14; an adaptive run can end the same way).  The filter takes that step exactly like the
// reference does (it is a genuine re-measurement at the updated mean).  With a dynamic diffusion model the backward
// recursions (smoother, sampler, dense output) take it as well: the sliver's own local diffusion is huge and makes the
// gain harmless.  With a STATIC model the backward gain across the sliver is, to rounding, the identity on the range
// of the covariance, evaluated in P(h) coordinates through a matrix with condition number ~h^-(2q+1): the reference's
// dense arithmetic returns the next state (oracle: smoothed[N-1] = filtered[N] in every component), the triangular
// solves of these kernels amplified rounding by 1e10 in the highest derivative.  An interval no longer than the
// rounding error that ns additions can accumulate is therefore treated, for static models, like the reference
// treats h == 0 (src/smoothing.jl:13-16): the next state is carried across.
__device__ __forceinline__ bool sliver_interval(double h, double ta, double tb, int ns, int static_model) {
  return h == 0.0 || (static_model && h <= 2.0 * double(ns) * ulp_of(fmax(fabs(ta), fabs(tb))));
}

// ---------------------------------------------------------------------------------------------
// Model: EK1 with the full D x D covariance (src/perform_step.jl, alg isa EK1)
// ---------------------------------------------------------------------------------------------
template <class VF_, int q_>
struct DenseEK1 {
  using VF = VF_;
  static constexpr int d = VF::d, q = q_, D = d * (q + 1), ND = 1;
  static constexpr bool IS_EK1 = true;
  using Fac = Factor<d, q>;
  static constexpr int REC = 1 + ND + D + Fac::LEN;  // t, diffusion, mean, factor
  static constexpr int STATE_LEN = D + Fac::LEN;

  struct State {
    double m[D];
    Fac F;
  };

  __device__ __forceinline__ static void scale(State& s, const double (&sc)[q + 1]) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) s.m[i] *= sc[i / d];
    s.F.scale_blocks(sc);
  }

  // One attempted step in P(h) coordinates.  diffusion in {dynamic, fixed, fixedMAP}.
  // quad = z' S^-1 z and detS = sqrt(det S) = |prod diag(R00)| feed the log-likelihood (:66), which the
  // kernel accumulates without a per-step log.
  __device__ __forceinline__ static void step(State& s, const double* p, double pi0, double pi1, double ipi1,
                                              int diffusion, const IwpConsts& C, double (&u_new)[d], double (&err)[d],
                                              double (&local)[ND], double& quad, double& detS,
                                              const double* ulin = nullptr) {
    apply_A<d, q>(s.m);  // predict_mean!  src/filtering.jl:22-25
    double uhat[d], fu[d], J[d][d], Jp[d][d], z[d];
PNDE_UNROLL
    for (int i = 0; i < d; ++i) uhat[i] = pi0 * s.m[i];  // src/perform_step.jl:44
    VF::template f<double>(uhat, p, fu);                  // :106
    VF::jac(ulin ? ulin : uhat, p, J);                    // :111-122 (IEKS: at the previous iterate's sol(t))
PNDE_UNROLL
    for (int i = 0; i < d; ++i) {
      z[i] = fma(pi1, s.m[d + i], -fu[i]);  // :108
PNDE_UNROLL
      for (int j = 0; j < d; ++j) Jp[i][j] = pi0 * J[i][j];
    }
    // B = H Q H' with H = (E1 - J E0) P^-1  (:125; src/diffusions.jl:77)
    double B[d][d];
PNDE_UNROLL
    for (int i = 0; i < d; ++i) {
PNDE_UNROLL
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
PNDE_UNROLL
        for (int k = 0; k < d; ++k) acc = fma(Jp[i][k], Jp[j][k], acc);
        acc *= C.Qt[0][0];
        acc = fma(-pi1 * C.Qt[0][1], Jp[i][j] + Jp[j][i], acc);
        if (i == j) acc = fma(pi1 * pi1, C.Qt[1][1], acc);
        B[i][j] = acc;
        B[j][i] = acc;
      }
    }
    double sig = 1.0;
    if (diffusion == DIFF_DYNAMIC) {
      // sigma^2 = z' B^-1 z / d via Cholesky of the d x d matrix (src/diffusions.jl:77-79)
      double Lb[d][d], yb[d];
      double ss = 0.0;
PNDE_UNROLL
      for (int j = 0; j < d; ++j) {
        double djj = B[j][j];
PNDE_UNROLL
        for (int k = 0; k < j; ++k) djj = fma(-Lb[j][k], Lb[j][k], djj);
        const double il = (djj > 0.0) ? fast_rsqrt(djj) : 0.0;
        Lb[j][j] = djj * il;
PNDE_UNROLL
        for (int i = j + 1; i < d; ++i) {
          double v = B[i][j];
PNDE_UNROLL
          for (int k = 0; k < j; ++k) v = fma(-Lb[i][k], Lb[j][k], v);
          Lb[i][j] = v * il;
        }
        double yy = z[j];
PNDE_UNROLL
        for (int k = 0; k < j; ++k) yy = fma(-Lb[j][k], yb[k], yy);
        yb[j] = yy * il;
        ss = fma(yb[j], yb[j], ss);
      }
      local[0] = ss * (1.0 / double(d));
      sig = (local[0] > 0.0) ? local[0] * fast_rsqrt(local[0]) : 0.0;
    }
    double Rtop[d][D], Rinv[d];
    cov_filter_step<d, q, true>(s.F, Jp, sig, pi1, ipi1, C, Rtop, Rinv);  // predict_cov! + update!
    // innovation: S_z = G G', G = Rtop[:, :d]' lower triangular; y = G^-1 z
    double y[d];
    double yy2 = 0.0, dets = 1.0;
PNDE_UNROLL
    for (int a = 0; a < d; ++a) {
      double acc = z[a];
PNDE_UNROLL
      for (int b = 0; b < a; ++b) acc = fma(-Rtop[b][a], y[b], acc);
      y[a] = acc * Rinv[a];
      yy2 = fma(y[a], y[a], yy2);
      dets *= fabs(Rtop[a][a]);
    }
    quad = yy2;
    detS = dets;
    if (diffusion != DIFF_DYNAMIC) local[0] = yy2 * (1.0 / double(d));  // src/diffusions.jl:25,52
    // mean update: mu+ = mu- - K z  (src/filtering.jl:87) in primed coordinates
    double m0old[d];
PNDE_UNROLL
    for (int b = 0; b < d; ++b) {
      m0old[b] = s.m[b];
      double acc = s.m[b];
PNDE_UNROLL
      for (int a = 0; a < d; ++a) acc = fma(-Rtop[a][d + b], y[a], acc);
      s.m[b] = acc;
    }
PNDE_UNROLL
    for (int i = 2 * d; i < D; ++i) {
      double acc = s.m[i];
PNDE_UNROLL
      for (int a = 0; a < d; ++a) acc = fma(-Rtop[a][i], y[a], acc);
      s.m[i] = acc;
    }
PNDE_UNROLL
    for (int b = 0; b < d; ++b) {
      double acc = fu[b];
PNDE_UNROLL
      for (int bb = 0; bb < d; ++bb) acc = fma(Jp[b][bb], s.m[bb] - m0old[bb], acc);
      s.m[d + b] = acc * ipi1;
    }
PNDE_UNROLL
    for (int i = 0; i < d; ++i) {
      u_new[i] = pi0 * s.m[i];              // src/perform_step.jl:70
      err[i] = sqrt(local[0] * B[i][i]);    // :155
    }
  }

  __device__ __forceinline__ static void store(const State& s, double* base, long long stride) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) base[(long long)i * stride] = s.m[i];
    s.F.store(base + (long long)D * stride, stride);
  }
  __device__ __forceinline__ static void load(State& s, const double* base, long long stride) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) s.m[i] = base[(long long)i * stride];
    s.F.load(base + (long long)D * stride, stride);
  }
  __device__ __forceinline__ static void final_cov(const State& s, const double (&sc)[q + 1], double* cov,
                                                   long long stride) {
    s.F.cov_entry_all(sc, cov, stride);
  }
};

// ---------------------------------------------------------------------------------------------
// Model: EK0 with the Kronecker-factored covariance Sigma = Ctilde (x) I_d (not in
// the reference, which is dense: src/caches.jl:73).  MV = dynamicMV keeps one factor per dimension.
// ---------------------------------------------------------------------------------------------
template <class VF_, int q_, bool MVDYN>
struct KronEK0 {
  using VF = VF_;
  static constexpr int d = VF::d, q = q_, D = d * (q + 1);
  static constexpr bool IS_EK1 = false;
  static constexpr int NF = MVDYN ? d : 1;  // covariance factors
  static constexpr int ND = d;              // diffusion slots (scalar models use slot 0, MV all)
  using Fac = Factor<1, q>;
  static constexpr int REC = 1 + ND + D + NF * Fac::LEN;
  static constexpr int STATE_LEN = D + NF * Fac::LEN;

  struct State {
    double m[D];
    Fac F[NF];
  };

  __device__ __forceinline__ static void scale(State& s, const double (&sc)[q + 1]) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) s.m[i] *= sc[i / d];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) s.F[f].scale_blocks(sc);
  }

  __device__ __forceinline__ static void step(State& s, const double* p, double pi0, double pi1, double ipi1,
                                              int diffusion, const IwpConsts& C, double (&u_new)[d], double (&err)[d],
                                              double (&local)[ND], double& quad, double& detS,
                                              const double* = nullptr) {
    apply_A<d, q>(s.m);
    double uhat[d], fu[d], z[d];
PNDE_UNROLL
    for (int i = 0; i < d; ++i) uhat[i] = pi0 * s.m[i];
    VF::template f<double>(uhat, p, fu);
    double zz = 0.0;
PNDE_UNROLL
    for (int i = 0; i < d; ++i) {
      z[i] = fma(pi1, s.m[d + i], -fu[i]);
      zz = fma(z[i], z[i], zz);
    }
    const double B = pi1 * pi1 * C.Qt[1][1];  // H Q H' = B I_d for EK0 (src/diffusions.jl:101-103)
    const double Jp0[1][1] = {{0.0}};
    double R[NF][1][q + 1], Ri[NF][1];
    if (MVDYN) {
      // src/diffusions.jl:104-108: Sigma_ii = max(z_i^2 / Q0_11, eps)
PNDE_UNROLL
      for (int a = 0; a < d; ++a) {
        local[a] = fmax(z[a] * z[a] / B, 2.220446049250313e-16);
        cov_filter_step<1, q, false>(s.F[a < NF ? a : 0], Jp0, sqrt(local[a]), pi1, ipi1, C, R[a < NF ? a : 0],
                                     Ri[a < NF ? a : 0]);
      }
    } else {
      double sig = 1.0;
      if (diffusion == DIFF_DYNAMIC) {
        local[0] = zz / (double(d) * B);
        sig = sqrt(local[0]);
      }
      cov_filter_step<1, q, false>(s.F[0], Jp0, sig, pi1, ipi1, C, R[0], Ri[0]);
    }
    double yy2 = 0.0, dets = 1.0;
PNDE_UNROLL
    for (int a = 0; a < d; ++a) {
      const int f = MVDYN ? a : 0;
      const double ya = z[a] * Ri[f][0];
      yy2 = fma(ya, ya, yy2);
      dets *= fabs(R[f][0][0]);
      s.m[a] = fma(-R[f][0][1], ya, s.m[a]);
PNDE_UNROLL
      for (int k = 2; k <= q; ++k) s.m[k * d + a] = fma(-R[f][0][k], ya, s.m[k * d + a]);
      s.m[d + a] = fu[a] * ipi1;
    }
    quad = yy2;
    detS = dets;
    if (diffusion == DIFF_FIXED || diffusion == DIFF_FIXED_MAP) local[0] = yy2 / double(d);
    if (diffusion == DIFF_FIXED_MV) {
      const double S11 = R[0][0][0] * R[0][0][0];  // src/diffusions.jl:136-138
PNDE_UNROLL
      for (int a = 0; a < d; ++a) local[a] = z[a] * z[a] / S11;
    }
PNDE_UNROLL
    for (int i = 0; i < d; ++i) {
      u_new[i] = pi0 * s.m[i];
      const bool mv = (diffusion == DIFF_DYNAMIC_MV || diffusion == DIFF_FIXED_MV);
      err[i] = sqrt((mv ? local[i] : local[0]) * B);
    }
  }

  __device__ __forceinline__ static void store(const State& s, double* base, long long stride) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) base[(long long)i * stride] = s.m[i];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) s.F[f].store(base + (long long)(D + f * Fac::LEN) * stride, stride);
  }
  __device__ __forceinline__ static void load(State& s, const double* base, long long stride) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) s.m[i] = base[(long long)i * stride];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) s.F[f].load(base + (long long)(D + f * Fac::LEN) * stride, stride);
  }
  // calibration by the final global diffusion (src/integrator_utils.jl:7-12): scalar or per dimension
  // (only ever called for static models, where NF == 1); gmv: per-dimension scales used at output.
  __device__ __forceinline__ static void final_cov(const State& s, const double (&sc)[q + 1], double* cov,
                                                   long long stride, const double (&dimscale)[d]) {
    // Sigma[(k,a),(k',a')] = delta_aa' * dimscale[a] * sc[k] sc[k'] * (F_a F_a')[k][k']
PNDE_UNROLL
    for (int i = 0; i < D; ++i) {
PNDE_UNROLL
      for (int j = 0; j <= i; ++j) {
        const int ki = i / d, ai = i % d, kj = j / d, aj = j % d;
        double v = 0.0;
        if (ai == aj) {
          const Fac& F = s.F[MVDYN ? ai : 0];
          // rows of the 1-d factor: row 0 = block 0, row 1 = block 1, ...
PNDE_UNROLL
          for (int c = 0; c < 1; ++c) v = fma(F.W[c][ki], F.W[c][kj], v);
PNDE_UNROLL
          for (int c = 0; c < Fac::NZ; ++c)
            if (ki >= 2 + c && kj >= 2 + c) v = fma(F.Lz[Fac::lz(c, ki - 2)], F.Lz[Fac::lz(c, kj - 2)], v);
          v *= sc[ki] * sc[kj] * dimscale[ai];
        }
        cov[(long long)(i * (i + 1) / 2 + j) * stride] = v;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// ode_determine_initdt (Hairer; OrdinaryDiffEq).  Two f evaluations.
// ---------------------------------------------------------------------------------------------
template <class VF, int q>
__device__ __forceinline__ double initdt(const double* u0, const double* p, const CtrlParams& K) {
  constexpr int d = VF::d;
  double f0[d], f1[d], u1[d], sk[d];
  VF::template f<double>(u0, p, f0);
  double d0 = 0.0, d1 = 0.0;
PNDE_UNROLL
  for (int i = 0; i < d; ++i) {
    sk[i] = K.abstol + fabs(u0[i]) * K.reltol;
    const double a = u0[i] / sk[i], b = f0[i] / sk[i];
    d0 = fma(a, a, d0);
    d1 = fma(b, b, d1);
  }
  d0 = sqrt(d0 / double(d));
  d1 = sqrt(d1 / double(d));
  double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
  dt0 = fmin(dt0, K.dtmax);
  if (dt0 < 10.0 * 2.220446049250313e-16) return 1e-6;
PNDE_UNROLL
  for (int i = 0; i < d; ++i) u1[i] = fma(dt0, f0[i], u0[i]);
  VF::template f<double>(u1, p, f1);
  double d2 = 0.0;
PNDE_UNROLL
  for (int i = 0; i < d; ++i) {
    const double c = (f1[i] - f0[i]) / sk[i];
    d2 = fma(c, c, d2);
  }
  d2 = sqrt(d2 / double(d)) / dt0;
  const double mx = fmax(d1, d2);
  double dt1;
  if (mx <= 1e-15)
    dt1 = fmax(1e-6, dt0 * 1e-3);
  else
    dt1 = pow(10.0, -(2.0 + log10(mx)) / double(q + 1));
  return fmin(fmin(100.0 * dt0, dt1), K.dtmax);
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
#ifndef PNDE_FILTER_MINB
#define PNDE_FILTER_MINB 1
#endif
#ifndef PNDE_FILTER_BLOCK
#define PNDE_FILTER_BLOCK 128
#endif
template <class M, bool ADAPTIVE, class LIN = NoLin>
__global__ void __launch_bounds__(PNDE_FILTER_BLOCK, PNDE_FILTER_MINB) filter_kernel(const FilterParams prm) {
  using VF = typename M::VF;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC;
  const long long lid_ = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lid_ >= prm.count) return;
  const long long tid = prm.first + lid_;
  const long long n = prm.n;
  const CtrlParams& K = prm.K;
  const int diffusion = prm.diffusion;
  const bool is_static = (diffusion == DIFF_FIXED || diffusion == DIFF_FIXED_MAP || diffusion == DIFF_FIXED_MV);
  const bool is_mv = (diffusion == DIFF_DYNAMIC_MV || diffusion == DIFF_FIXED_MV);

  double p[VF::np], u0[d];
PNDE_UNROLL
  for (int i = 0; i < VF::np; ++i) p[i] = prm.p[(long long)i * n + tid];
PNDE_UNROLL
  for (int i = 0; i < d; ++i) u0[i] = prm.u0[(long long)i * n + tid];

#ifdef PNDE_ROLLED
  typename M::State saved;  // general-(d, q) fallback: the state is in local memory anyway, so is its pre-step copy
#else
  extern __shared__ double stash[];  // ADAPTIVE only: STATE_LEN x blockDim doubles
#endif
  typename M::State st;  // natural coordinates between steps when ADAPTIVE, P(hcur) coordinates otherwise
  taylor_init<VF, q>(u0, p, st.m);
  if constexpr (M::IS_EK1) {
    st.F.zero();
  } else {
PNDE_UNROLL
    for (int f = 0; f < M::NF; ++f) st.F[f].zero();
  }

  double t = K.t0;
  int iter = 0, nacc = 0, nrej = 0, nfe = 0, ret = RET_SUCCESS, nsaved = 0;
  double gsaved[ND];  // last saved global diffusion (sol.diffusions[end])
PNDE_UNROLL
  for (int i = 0; i < ND; ++i) gsaved[i] = 1.0;  // initial_diffusion, src/diffusions.jl:8
  double uprev[d];
PNDE_UNROLL
  for (int i = 0; i < d; ++i) uprev[i] = u0[i];
  // log-likelihood (src/perform_step.jl:66,91) = -1/2 sum (quad + 2 log detS + d log 2pi) over committed
  // steps; the log of the running product of detS is taken lazily (mantissa / exponent split).
  double ll_quad = 0.0, ll_log = 0.0, ll_mant = 1.0;
  long long ll_exp = 0;
  int ll_n = 0;

  auto save = [&](const typename M::State& sv, double tt, const double (&g)[ND]) {
    if (nsaved >= prm.max_saved) {
      ret = RET_HISTORY_FULL;
      return;
    }
    double* base = prm.hist + ((long long)nsaved * REC) * n + tid;
    base[0] = tt;
PNDE_UNROLL
    for (int i = 0; i < ND; ++i) base[(long long)(1 + i) * n] = g[i];
    M::store(sv, base + (long long)(1 + ND) * n, n);
    ++nsaved;
  };
  if (prm.save_mode != SAVE_FINAL) save(st, t, gsaved);

  double dt;
  if (ADAPTIVE) {
    if (K.dt > 0.0) {
      dt = K.dt;
    } else {
      dt = initdt<VF, q>(u0, p, K);
      nfe += 2;
    }
  } else {
    dt = K.dt;
  }
  const double lqold0 = ctrl_state_init(K);
  double dtpropose = dt, lqold = lqold0, q11 = 1.0;
  bool accepted_prev = true;
  double hcur = -1.0;  // fixed-step mode: the h the state is currently preconditioned with (<0: natural)
  double Pk[q + 1], PIk[q + 1];
PNDE_UNROLL
  for (int k = 0; k <= q; ++k) Pk[k] = PIk[k] = 1.0;

  while (t < K.t1) {
    // ---- loopheader! ----
    if (iter > 0) {
      if (accepted_prev)
        dt = dtpropose;
      else
        dt = dt / fmin(1.0 / K.qmin, q11 / K.gamma);
    }
    ++iter;
    if (iter > K.maxiters) {
      ret = RET_MAXITERS;
      break;
    }
    if (ADAPTIVE) {
      dt = fmin(dt, K.dtmax);
      dt = fmax(dt, K.dtmin);
      dt = fmin(dt, K.t1 - t);
    } else {
      dt = fmin(K.dt, K.t1 - t);
    }
    if (dt != dt) {
      ret = RET_DTNAN;
      break;
    }
    if (ADAPTIVE && iter > 1 && !accepted_prev && fabs(dt) <= fabs(K.dtmin)) {
      ret = RET_DTMIN;
      break;
    }
    // ---- perform_step! ----
    if (ADAPTIVE) {
      // the pre-step state is parked in shared memory ([element][thread], conflict free) instead of a second
      // register copy; it is only read back when the step is rejected
#ifdef PNDE_ROLLED
      saved = st;
#else
      M::store(st, stash + threadIdx.x, blockDim.x);
#endif
      precond_scales<q>(dt, Pk, PIk);
      M::scale(st, Pk);  // x = P * x   (src/perform_step.jl:38)
    } else if (dt != hcur) {
      double Pn[q + 1], PIn[q + 1], sc[q + 1];
      precond_scales<q>(dt, Pn, PIn);
PNDE_UNROLL
      for (int k = 0; k <= q; ++k) {
        sc[k] = Pn[k] * PIk[k];
        Pk[k] = Pn[k];
        PIk[k] = PIn[k];
      }
      M::scale(st, sc);
      hcur = dt;
    }
    double unew[d], err[d], local[ND], quad, detS;
PNDE_UNROLL
    for (int i = 0; i < ND; ++i) local[i] = 1.0;
    if constexpr (LIN::enabled) {
      double ulin[d];
      const bool have = LIN::template point<M>(prm, tid, t + dt, ulin);
      M::step(st, p, PIk[0], PIk[1], Pk[1], diffusion, prm.C, unew, err, local, quad, detS, have ? ulin : nullptr);
    } else {
      M::step(st, p, PIk[0], PIk[1], Pk[1], diffusion, prm.C, unew, err, local, quad, detS);
    }
    ++nfe;
    // global diffusion (src/diffusions.jl): success_iter == number of accepted steps so far
    double gcur[ND];
PNDE_UNROLL
    for (int i = 0; i < ND; ++i) {
      if (!is_mv && i > 0) {
        gcur[i] = 1.0;
        continue;
      }
      if (diffusion == DIFF_DYNAMIC || diffusion == DIFF_DYNAMIC_MV) {
        gcur[i] = local[i];
      } else if (diffusion == DIFF_FIXED || diffusion == DIFF_FIXED_MV) {
        gcur[i] = (nacc == 0) ? local[i] : gsaved[i] + (local[i] - gsaved[i]) / double(nacc);  // :33,:150
      } else {  // fixedMAP :46-68
        const double Nn = double(nacc + 1), al = 0.5, be = 0.5;
        if (nacc == 0) {
          gcur[i] = (be + 0.5 * local[i]) / (al + Nn * d / 2.0 + 1.0);
        } else {
          const double res_prev = (gsaved[i] * (al + (Nn - 1.0) * d / 2.0 + 1.0) - be) * 2.0;
          gcur[i] = (be + 0.5 * (res_prev + local[i])) / (al + Nn * d / 2.0 + 1.0);
        }
      }
    }
    double EEst = 0.0;
    bool finite = true;
    if (ADAPTIVE) {
      // calculate_residuals! + ODE_DEFAULT_NORM (src/perform_step.jl:78-84)
      double acc = 0.0;
PNDE_UNROLL
      for (int i = 0; i < d; ++i) {
        const double r = dt * err[i] / (K.abstol + fmax(fabs(uprev[i]), fabs(unew[i])) * K.reltol);
        acc = fma(r, r, acc);
      }
      EEst = sqrt(acc / double(d));
    }
PNDE_UNROLL
    for (int i = 0; i < d; ++i) {
      uprev[i] = unew[i];  // integ.u .= u_filt, even when rejected (:86)
      finite = finite && (fabs(unew[i]) <= 1.79769313486231570e308);
    }
    const bool commit = !ADAPTIVE || (EEst < 1.0);  // :89
    const bool accept = !ADAPTIVE || (EEst <= 1.0); // loopfooter!
    if (ADAPTIVE) {
      if (commit) {
        M::scale(st, PIk);  // PI * x_filt (:75)
      } else {
#ifdef PNDE_ROLLED
        st = saved;
#else
        M::load(st, stash + threadIdx.x, blockDim.x);
#endif
      }
    }
    if (commit) {
      ll_quad += quad;
      ++ll_n;
      if (detS > 1e-290 && detS < 1e290) {
        const long long bits = __double_as_longlong(detS);
        ll_exp += ((bits >> 52) & 0x7ff) - 1023;
        ll_mant *= __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3ff0000000000000LL);
        if (ll_mant > 1e250) {
          ll_log += log(ll_mant);
          ll_mant = 1.0;
        }
      } else {
        ll_log += log(detS);
      }
    }
    if (!finite) {
      ret = RET_NONFINITE;  // OrdinaryDiffEq check_error!: unstable_check
      break;
    }
#ifdef PNDE_QUIRK_CHECK
    // PNDE_FLAG_REFERENCE_QUIRKS (b).  Compiled only into the kernels that pnde_create builds through NVRTC for handles
    // that set the flag: in the ahead-of-time kernels this test cost 3.3 % on the headline configuration as a loop
    // exit and 2 % as a sticky flag (A/B, round 2) -- the one-thread kernel sits on the register limit.
    if (diffusion == DIFF_FIXED && quad == 0.0) {
      ret = RET_ZERO_RESIDUAL;  // the reference throws here (src/diffusions.jl:18-20)
      break;
    }
#endif
    // ---- loopfooter! ----
    const double ttmp = t + dt;
    if (ADAPTIVE) {
      double lE;
      double qc = controller_factor(EEst, K, lqold, lE);
      if (accept) {
        ++nacc;
        if (K.qsteady_min <= qc && qc <= K.qsteady_max) qc = 1.0;
        lqold = ctrl_state_accept(EEst, lE, lqold0, K);  // qold = max(EEst, qoldinit)
        const double dtnew = dt / qc;
        t = (fabs(ttmp - K.t1) < 10.0 * ulp_of(fmax(t, K.t1))) ? K.t1 : ttmp;
        dtpropose = fmax(K.dtmin, fmin(K.dtmax, dtnew));
      } else {
        ++nrej;
        q11 = ctrl_q11(EEst, lE, K);
      }
    } else {
      ++nacc;
      t = (fabs(ttmp - K.t1) < 10.0 * ulp_of(fmax(t, K.t1))) ? K.t1 : ttmp;
      dtpropose = dt;
    }
    accepted_prev = accept;
    if (accept) {
PNDE_UNROLL
      for (int i = 0; i < ND; ++i) gsaved[i] = gcur[i];
      // savevalues! (src/integrator_utils.jl:33-48)
      const bool want = (prm.save_mode == SAVE_EVERY) ||
                        (prm.save_mode == SAVE_STRIDE && (nacc % prm.save_stride == 0 || !(t < K.t1)));
      if (want) {
        if (ADAPTIVE) {
          save(st, t, gsaved);
        } else {
          typename M::State nat = st;
          M::scale(nat, PIk);
          save(nat, t, gsaved);
        }
        // fixed steps: the capacity is derived from the grid, a full history is a caller error: stop.  Adaptive: keep
        // stepping WITHOUT saving so that naccept tells the caller the capacity the run needs (one retry, api.py)
        if (!ADAPTIVE && ret == RET_HISTORY_FULL) break;
      }
    }
  }

  // ---- outputs ----
  double sc[q + 1];
PNDE_UNROLL
  for (int k = 0; k <= q; ++k) sc[k] = (ADAPTIVE || hcur < 0.0) ? 1.0 : PIk[k];
  if (prm.mean) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) prm.mean[(long long)i * n + tid] = st.m[i] * sc[i / d];
  }
  // postamble! calibration for static models (src/integrator_utils.jl:4-18)
  double dimscale[d];
PNDE_UNROLL
  for (int a = 0; a < d; ++a) dimscale[a] = 1.0;
  double ll = -0.5 * (ll_quad + 2.0 * (ll_log + log(ll_mant) + double(ll_exp) * 0.6931471805599453) +
                      double(ll_n) * double(d) * 1.8378770664093453);
  if (is_static && nacc > 0) {
PNDE_UNROLL
    for (int a = 0; a < d; ++a) dimscale[a] = is_mv ? gsaved[a < ND ? a : 0] : gsaved[0];
    ll = nan("");
  }
  if (prm.cov) {
    if constexpr (M::IS_EK1) {
      double sc2[q + 1];
      const double g = sqrt(dimscale[0]);
PNDE_UNROLL
      for (int k = 0; k <= q; ++k) sc2[k] = sc[k] * g;
      M::final_cov(st, sc2, prm.cov + tid, n);
    } else {
      M::final_cov(st, sc, prm.cov + tid, n, dimscale);
    }
  }
  if (prm.final_diff) {
PNDE_UNROLL
    for (int i = 0; i < ND; ++i) prm.final_diff[(long long)i * n + tid] = gsaved[i];
  }
  if (prm.t_final) prm.t_final[tid] = t;
  if (prm.loglik) prm.loglik[tid] = ll;
  prm.retcode[tid] = ret;
  prm.naccept[tid] = nacc;
  prm.nreject[tid] = nrej;
  prm.nf[tid] = nfe;
  prm.njacs[tid] = M::IS_EK1 ? (nfe - ((ADAPTIVE && !(K.dt > 0.0)) ? 2 : 0)) : 0;
  prm.n_saved[tid] = nsaved;
}

}  // namespace pnde
