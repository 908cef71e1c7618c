// dense_sample_states / dense_sample (src/solution_sampling.jl:63-74): backward sampling on a caller-given time grid
// tq[0..n_t) (the reference: 1000 equidistant points) instead of the solver's own grid.  The "filter states" of that
// grid are the filtering posterior extrapolated from the left saved neighbour (interp(...; smoothed=false), :66), the
// diffusion of an interval is the one of the solver interval that contains its left end (:41-42).
//   dense_sample_prep_kernel  one thread per (trajectory, grid interval k): predicted state at tq[k], then the same
//                             stage-1 sweep as sample_prep_kernel for the backward kernel x(tq[k]) | x(tq[k+1]).
//   dense_sample_draw_kernel  one thread per (trajectory, sample): the last grid point is drawn from its (predicted or
//                             stored) state, then one O(D^2) backward draw per interval.
// Scratch record of an interval: m = P m_k, m^- = A m, and per covariance factor R-, 1/diag R-, X, Y (DCOV rows: a
// predicted covariance has full rank, unlike the rank D - d filter posteriors of sample_prep_kernel).
#pragma once
#include "post_kernels.cuh"

namespace pnde {

template <class M>
struct DenseSamplePrep {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  static constexpr int D = M::D, NF = PT::NF, DCOV = PT::DCOV, NP = SC::NP;
  static constexpr int FLEN = NP + DCOV + 2 * DCOV * DCOV;  // R-, rinv, X, Y of one factor
  static constexpr int LEN = 2 * D + NF * FLEN;
};

// Filtering posterior at tval as (mean, factor columns), natural coordinates, in the convention of sample_prep_kernel
// (dense EK1 + static diffusion: calibrated; Kronecker + static diffusion: uncalibrated, the per-dimension scale is
// applied where the noise is drawn).  cols holds DCOV columns; an exact hit of the saved grid returns the stored rank
// D - d factor padded with zero columns.
template <class M>
__device__ __forceinline__ void filter_state_at(const SampleParams& sp, long long tr, int ns, double tval,
                                                double (&mean)[M::D],
                                                double (&cols)[PostTraits<M>::NF][PostTraits<M>::DCOV][PostTraits<M>::DCOV]) {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC, NF = PT::NF, DC = PT::DC, DCOV = PT::DCOV, R = PT::R;
  const long long n = sp.n;
  auto rec = [&](int slot) { return sp.hist + ((long long)slot * REC) * n + tr; };
  int lo = 0, hi = ns;  // prev = last saved index with t <= tval
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (rec(mid)[0] <= tval) lo = mid; else hi = mid;
  }
  const int prev = lo;
  const double cal = (sp.calibrate && M::IS_EK1) ? sqrt(sp.final_diff[tr]) : 1.0;
  typename M::State st;
  M::load(st, rec(prev) + (long long)(1 + ND) * n, n);
  if (rec(prev)[0] == tval || prev + 1 >= ns) {  // stored state (also for tval beyond the last saved time)
PNDE_UNROLL
    for (int i = 0; i < D; ++i) mean[i] = st.m[i];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) {
      double cf[R][DCOV];
      SC::cols_from_factor(PT::factor(st, f), cf);
PNDE_UNROLL
      for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
        for (int k = 0; k < DCOV; ++k) cols[f][c][k] = (c < R) ? cf[c < R ? c : 0][k] * cal : 0.0;
    }
    return;
  }
  // extrapolate from the left neighbour (src/solution.jl:184-189); diffusions[min(idx, end)] is stored with state idx
  const double* rd = rec(prev + 1);
  const double h1 = tval - rec(prev)[0];
  double Pk[q + 1], PIk[q + 1];
  precond_scales<q>(h1, Pk, PIk);
  M::scale(st, Pk);
  apply_A<d, q>(st.m);
PNDE_UNROLL
  for (int k = 0; k < D; ++k) mean[k] = st.m[k] * PIk[k / d];
  int status = 0;
PNDE_UNROLL
  for (int f = 0; f < NF; ++f) {
    const double g = sp.calibrate ? (M::IS_EK1 ? sp.final_diff[tr] : 1.0) : rd[(long long)(1 + (NF > 1 ? f : 0)) * n];
    const double sig = sqrt(g);
    double cf[R][DCOV];
    SC::cols_from_factor(PT::factor(st, f), cf);
PNDE_UNROLL
    for (int c = 0; c < R; ++c) {
      double w[DCOV];
PNDE_UNROLL
      for (int k = 0; k < DCOV; ++k) w[k] = cf[c][k] * cal;
      apply_A<DC, q>(w);
PNDE_UNROLL
      for (int k = 0; k < DCOV; ++k) cf[c][k] = w[k];
    }
    double Tt[DCOV][DCOV], Lp[SC::NP];
PNDE_UNROLL
    for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
      for (int k = 0; k < DCOV; ++k) Tt[c][k] = (k % DC == c % DC && k >= c) ? sig * sp.C.Lt[k / DC][c / DC] : 0.0;
    SC::template triangularize<R>(cf, Tt, Lp, status);  // factor of A S S' A' + sig^2 Q (src/filtering.jl:33-48)
PNDE_UNROLL
    for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
      for (int k = 0; k < DCOV; ++k) cols[f][c][k] = (k >= c) ? Lp[SC::tri(k, c)] * PIk[k / DC] : 0.0;
  }
}

template <class M>
__global__ void __launch_bounds__(128) dense_sample_prep_kernel(const SampleParams sp) {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  using SPp = DenseSamplePrep<M>;
  constexpr int d = M::d, q = M::q, D = M::D, REC = M::REC, NF = PT::NF, DC = PT::DC, DCOV = PT::DCOV;
  const long long ntr = sp.traj_end - sp.traj_begin;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= ntr * (sp.n_t - 1)) return;
  const long long tr = sp.traj_begin + gid % ntr;
  const int k = (int)(gid / ntr);  // interval tq[k] -> tq[k + 1]
  const long long n = sp.n;
  const int ns = sp.n_saved[tr];
  if (ns <= 0) return;
  const double ta = sp.tq[k], h = sp.tq[k + 1] - ta;
  if (!(h > 0.0)) return;  // the draw kernel copies the sample across
  auto rec = [&](int slot) { return sp.hist + ((long long)slot * REC) * n + tr; };
  double* o = sp.scratch + ((tr - sp.traj_begin) * (sp.n_t - 1) + k) * SPp::LEN;
  double m[D], cols[NF][DCOV][DCOV];
  filter_state_at<M>(sp, tr, ns, ta, m, cols);
  double Pk[q + 1], PIk[q + 1];
  precond_scales<q>(h, Pk, PIk);
  double mpred[D];
PNDE_UNROLL
  for (int i = 0; i < D; ++i) {
    m[i] *= Pk[i / d];
    mpred[i] = m[i];
  }
  apply_A<d, q>(mpred);
PNDE_UNROLL
  for (int i = 0; i < D; ++i) {
    o[i] = m[i];
    o[D + i] = mpred[i];
  }
  // diffusions[sum(sol.t .<= tq[k])] (src/solution_sampling.jl:41-42): stored with the state that ends that interval
  int lo = 0, hi = ns;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (rec(mid)[0] <= ta) lo = mid; else hi = mid;
  }
  const double* rd = rec(lo + 1 < ns ? lo + 1 : ns - 1);
PNDE_UNROLL
  for (int f = 0; f < NF; ++f) {
    const double g = sp.calibrate ? (M::IS_EK1 ? sp.final_diff[tr] : 1.0) : rd[(long long)(1 + (NF > 1 ? f : 0)) * n];
    const double sig = sqrt(g);
    double cf[DCOV][DCOV];
PNDE_UNROLL
    for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
      for (int i = 0; i < DCOV; ++i) cf[c][i] = cols[f][c][i] * Pk[i / DC];
    double Rm[SC::NP], rinv[DCOV];
    RegMat<DCOV> X;
    SC::template stage1<DCOV>(cf, sig, sp.C, Rm, rinv, X);  // cf now holds Y
    double* of = o + 2 * D + f * SPp::FLEN;
PNDE_UNROLL
    for (int i = 0; i < SC::NP; ++i) of[i] = Rm[i];
PNDE_UNROLL
    for (int i = 0; i < DCOV; ++i) of[SC::NP + i] = rinv[i];
PNDE_UNROLL
    for (int r = 0; r < DCOV; ++r)
PNDE_UNROLL
      for (int i = 0; i < DCOV; ++i) of[SC::NP + DCOV + r * DCOV + i] = X.get(r, i);
PNDE_UNROLL
    for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
      for (int i = 0; i < DCOV; ++i) of[SC::NP + DCOV + DCOV * DCOV + c * DCOV + i] = cf[c][i];
  }
}

template <class M>
__global__ void __launch_bounds__(128) dense_sample_draw_kernel(const SampleParams sp) {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  using SPp = DenseSamplePrep<M>;
  constexpr int q = M::q, D = M::D, ND = M::ND, NF = PT::NF, DC = PT::DC, DCOV = PT::DCOV;
  const long long ntr = sp.traj_end - sp.traj_begin;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= ntr * sp.n_samples) return;
  const int smp = (int)(gid % sp.n_samples);
  const long long tr = sp.traj_begin + gid / sp.n_samples;
  const long long n = sp.n;
  const int ns = sp.n_saved[tr];
  if (ns <= 0) return;
  Philox rng;
  rng.key[0] = (uint32_t)sp.seed;
  rng.key[1] = (uint32_t)(sp.seed >> 32);
  double gfin[ND];
PNDE_UNROLL
  for (int i = 0; i < ND; ++i) gfin[i] = sp.calibrate ? sp.final_diff[(long long)i * n + tr] : 1.0;
  auto outp = [&](int k) { return sp.out + (((tr - sp.traj_begin) * sp.n_t + k) * sp.n_samples + smp) * D; };
  auto kron_scale = [&](int rep) -> double {  // Kronecker + static diffusion: the scale outside the shared factor
    if (!sp.calibrate || M::IS_EK1) return 1.0;
    return sqrt(sp.is_mv ? gfin[rep < ND ? rep : 0] : gfin[0]);
  };
  const uint32_t salt = 0x40000000u;  // keeps the generator counters apart from sample_draw_kernel's
  double s[D];
  {
    // last grid point: s = mu + S xi  (src/solution_sampling.jl:31-32)
    double cols[NF][DCOV][DCOV];
    filter_state_at<M>(sp, tr, ns, sp.tq[sp.n_t - 1], s, cols);
PNDE_UNROLL
    for (int rep = 0; rep < PT::NREP; ++rep) {
      const int f = (NF > 1) ? rep : 0;
      const double cs = kron_scale(rep);
PNDE_UNROLL
      for (int c = 0; c < DCOV; c += 2) {
        double a, b;
        rng.normal2((uint32_t)(tr + sp.key_offset), (uint32_t)smp, salt | (uint32_t)(sp.n_t - 1), (uint32_t)(rep * 64 + c), a, b);
PNDE_UNROLL
        for (int i = 0; i < DCOV; ++i) {
          s[PT::idx(rep, i)] = fma(cs * cols[f][c][i], a, s[PT::idx(rep, i)]);
          if (c + 1 < DCOV) s[PT::idx(rep, i)] = fma(cs * cols[f][c + 1 < DCOV ? c + 1 : 0][i], b, s[PT::idx(rep, i)]);
        }
      }
    }
    double* o = outp((int)sp.n_t - 1);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) o[i] = s[i];
  }
  for (int k = (int)sp.n_t - 2; k >= 0; --k) {
    const double h = sp.tq[k + 1] - sp.tq[k];
    if (h > 0.0) {
      const double* pr = sp.scratch + ((tr - sp.traj_begin) * (sp.n_t - 1) + k) * SPp::LEN;
      double Pk[q + 1], PIk[q + 1];
      precond_scales<q>(h, Pk, PIk);
      double snew[D];
PNDE_UNROLL
      for (int f = 0; f < NF; ++f) {
        const double* pf = pr + 2 * D + f * SPp::FLEN;
        const double* Rm = pf;
        const double* rinv = pf + SC::NP;
        const double* X = rinv + DCOV;
        const double* Y = X + DCOV * DCOV;
        constexpr int NR_ = (NF > 1) ? 1 : PT::NREP;
PNDE_UNROLL
        for (int rr = 0; rr < NR_; ++rr) {
          const int rep = (NF > 1) ? f : rr;
          double y[DCOV];
PNDE_UNROLL
          for (int i = 0; i < DCOV; ++i) {
            double acc = fma(Pk[i / DC], s[PT::idx(rep, i)], -pr[D + PT::idx(rep, i)]);
PNDE_UNROLL
            for (int l = 0; l < i; ++l) acc = fma(-Rm[SC::tri(i, l)], y[l], acc);
            y[i] = acc * rinv[i];
          }
          const double ys = kron_scale(rep);
          double acc[DCOV];
PNDE_UNROLL
          for (int i = 0; i < DCOV; ++i) {
            double dl = 0.0;
PNDE_UNROLL
            for (int l = 0; l < DCOV; ++l) dl = fma(X[l * DCOV + i], y[l], dl);
            acc[i] = pr[PT::idx(rep, i)] + dl;
          }
PNDE_UNROLL
          for (int c = 0; c < DCOV; c += 2) {
            double a, b;
            rng.normal2((uint32_t)(tr + sp.key_offset), (uint32_t)smp, salt | (uint32_t)k, (uint32_t)(rep * 64 + c), a, b);
PNDE_UNROLL
            for (int i = 0; i < DCOV; ++i) {
              acc[i] = fma(ys * Y[c * DCOV + i], a, acc[i]);
              if (c + 1 < DCOV) acc[i] = fma(ys * Y[(c + 1 < DCOV ? c + 1 : 0) * DCOV + i], b, acc[i]);
            }
          }
PNDE_UNROLL
          for (int i = 0; i < DCOV; ++i) snew[PT::idx(rep, i)] = acc[i] * PIk[i / DC];
        }
      }
PNDE_UNROLL
      for (int i = 0; i < D; ++i) s[i] = snew[i];
    }
    double* o = outp(k);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) o[i] = s[i];
  }
}

}  // namespace pnde
