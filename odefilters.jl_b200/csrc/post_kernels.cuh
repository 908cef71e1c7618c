// Post-processing kernels on the saved history (north_star "smooth!/sample"):
//   sample_kernel  backward sampling            src/solution_sampling.jl:6-62  (sample_states)
//   dense_kernel   dense output sol(t)          src/solution.jl:165-215        (GaussianODEFilterPosterior)
// Both reuse the smoother's stage-1 sweep (smoother_kernel.cuh): the backward kernel of an interval,
//   x_i | x_{i+1} ~ N( m + G (x_{i+1} - m^-),  Y'Y ),
// does not depend on the sample, so a draw is  m + G (s_{i+1} - m^-) + Y' xi  -- the reference instead
// runs a full `smooth` against a zero-covariance "next state" per sample and per interval
// (src/solution_sampling.jl:49-58), which is the same distribution.
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#else
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#endif

#include "smoother_kernel.cuh"

namespace pnde {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator + Box-Muller (no cuRAND; reproducible per (seed, traj,
// sample, slot, draw index) regardless of launch geometry)
// ---------------------------------------------------------------------------------------------
struct Philox {
  uint32_t key[2];
  __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
  }
  __device__ __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key[0], k1 = key[1];
PNDE_UNROLL
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
PNDE_UNROLL
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
  // two independent standard normals from one block
  __device__ __forceinline__ void normal2(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double& a, double& b) const {
    uint32_t r[4];
    block(c0, c1, c2, c3, r);
    const double u1 = (double(((uint64_t)r[0] << 21) ^ (r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
    const double u2 = (double(((uint64_t)r[2] << 21) ^ (r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    a = rad * cs;
    b = rad * sn;
  }
};

struct SampleParams {
  long long n, traj_begin, traj_end, max_saved;
  const int* n_saved;
  const long long* offsets;  // CSR offsets of [traj_begin, traj_end)
  const double* hist;
  const double* final_diff;
  int calibrate, is_mv, n_samples;
  unsigned long long seed;
  long long key_offset;  // added to the trajectory index in the generator key (shards of a multi-device ensemble)
  double* out;  // [total][n_samples][D]
  double* scratch;  // [traj_end - traj_begin][max_saved - 1][SamplePrep<M>::LEN]: the per-interval backward kernels
  // dense_sample (dense_sample.cuh): backward sampling on the caller's grid tq[0..n_t) instead of the saved one; out is
  // then [traj][n_t][n_samples][D] and scratch [traj][n_t - 1][DenseSamplePrep<M>::LEN].  nullptr: the saved grid
  const double* tq;
  long long n_t;
  IwpConsts C;
};

struct DenseParams {
  long long n, traj_begin, traj_end, max_saved, n_t;
  const int* n_saved;
  const double* hist;
  const double* smooth;  // nullptr: filtering posterior only
  const double* final_diff;
  int calibrate, is_mv, smoothed;
  const double* tq;  // [n_t] query times
  double* mean;      // [ntr][n_t][D]
  double* cov;       // [ntr][n_t][D(D+1)/2]
  IwpConsts C;
};

// Helper shared by both kernels: everything that depends on whether the model is dense-EK1 or
// Kronecker-EK0 (number of covariance factors NF, their dimension DCOV, mean replicas NREP).
template <class M>
struct PostTraits {
  using SM = SmoothModel<M>;
  using SC = typename SM::SC;
  static constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, NF = SM::NF, DC = SM::DC;
  static constexpr int DCOV = DC * (q + 1), NREP = D / DCOV, R = SC::R;
  // coordinate of replica r, factor coordinate k  ->  index into the D-vector
  __device__ __forceinline__ static constexpr int idx(int rep, int k) { return M::IS_EK1 ? k : k * d + rep; }
  __device__ __forceinline__ static const Factor<DC, q>& factor(const typename M::State& st, int f) {
    if constexpr (M::IS_EK1) return st.F; else return st.F[f];
  }
};

// Backward sampling in two kernels ("reuse the smoother gain per interval instead of N n full
// smooth calls"):
//   sample_prep_kernel  one thread per (trajectory, interval): the stage-1 sweep of that interval, ONCE -- the
//                       backward kernel x_i | x_{i+1} ~ N(m + G (x_{i+1} - m^-), Y'Y) does not depend on the sample.
//                       Writes m = P m_i, m^- = A m, R-, 1/diag(R-), X (G' = R-^-1 X) and Y to a scratch record.
//   sample_draw_kernel  one thread per (trajectory, sample): per interval a forward substitution with R-, two
//                       matrix-vector products with X and Y and the normals: O(D^2) per draw instead of the O(D^3) sweep.
template <class M>
struct SamplePrep {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  static constexpr int D = M::D, NF = PT::NF, DCOV = PT::DCOV, R = PT::R, NP = SC::NP;
  static constexpr int FLEN = NP + DCOV + DCOV * DCOV + R * DCOV;  // R-, rinv, X, Y of one factor
  static constexpr int LEN = 2 * D + NF * FLEN;                     // + m, m^-
};

template <class M>
__global__ void __launch_bounds__(128) sample_prep_kernel(const SampleParams sp) {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  using SPp = SamplePrep<M>;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC, NF = PT::NF, DCOV = PT::DCOV, R = PT::R;
  const long long ntr = sp.traj_end - sp.traj_begin;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= ntr * (sp.max_saved - 1)) return;
  const long long tr = sp.traj_begin + gid % ntr;  // trajectory fastest: record loads coalesce
  const int i = (int)(gid / ntr);                  // interval t[i] -> t[i+1]
  const long long n = sp.n;
  const int ns = sp.n_saved[tr];
  if (i + 1 >= ns) return;
  auto rec = [&](int slot) { return sp.hist + ((long long)slot * REC) * n + tr; };
  double* o = sp.scratch + ((tr - sp.traj_begin) * (sp.max_saved - 1) + i) * SPp::LEN;
  const double* ri = rec(i);
  const double* rn = rec(i + 1);
  const double h = rn[0] - ri[0];
  if (sliver_interval(h, ri[0], rn[0], ns, sp.calibrate)) return;  // h == 0 or a sliver: the draw kernel copies the sample across
  double gfin[ND];
PNDE_UNROLL
  for (int k = 0; k < ND; ++k) gfin[k] = sp.calibrate ? sp.final_diff[(long long)k * n + tr] : 1.0;
  double Pk[q + 1], PIk[q + 1];
  precond_scales<q>(h, Pk, PIk);
  typename M::State st;
  M::load(st, ri + (long long)(1 + ND) * n, n);
  M::scale(st, Pk);
  double mpred[D];
PNDE_UNROLL
  for (int k = 0; k < D; ++k) mpred[k] = st.m[k];
  apply_A<d, q>(mpred);
  for (int k = 0; k < D; ++k) {
    o[k] = st.m[k];
    o[D + k] = mpred[k];
  }
PNDE_UNROLL
  for (int f = 0; f < NF; ++f) {
    // dynamic models: the interval's own diffusion; static models: the final global value.  In the Kronecker form a
    // static per-dimension scale cancels in G and multiplies Y'Y (applied by the draw kernel).
    double g;
    if (sp.calibrate)
      g = M::IS_EK1 ? gfin[0] : 1.0;
    else
      g = rn[(long long)(1 + (NF > 1 ? f : 0)) * n];
    const double sig = sqrt(g);
    double cols[R][DCOV];
    SC::cols_from_factor(PT::factor(st, f), cols);
    if (M::IS_EK1 && sp.calibrate) {
      const double cs = sqrt(gfin[0]);
PNDE_UNROLL
      for (int c = 0; c < R; ++c)
PNDE_UNROLL
        for (int k = 0; k < DCOV; ++k) cols[c][k] *= cs;
    }
    double Rm[SC::NP], rinv[DCOV];
    RegMat<DCOV> X;
    SC::template stage1<R>(cols, sig, sp.C, Rm, rinv, X);  // cols now holds Y
    double* of = o + 2 * D + f * SPp::FLEN;
    for (int k = 0; k < SC::NP; ++k) of[k] = Rm[k];
    for (int k = 0; k < DCOV; ++k) of[SC::NP + k] = rinv[k];
    for (int r = 0; r < DCOV; ++r)
      for (int k = 0; k < DCOV; ++k) of[SC::NP + DCOV + r * DCOV + k] = X.get(r, k);
    for (int c = 0; c < R; ++c)
      for (int k = 0; k < DCOV; ++k) of[SC::NP + DCOV + DCOV * DCOV + c * DCOV + k] = cols[c][k];
  }
}

// one thread per (trajectory, sample)
template <class M>
__global__ void __launch_bounds__(128) sample_draw_kernel(const SampleParams sp) {
  using PT = PostTraits<M>;
  using SC = typename PT::SC;
  using SPp = SamplePrep<M>;
  constexpr int q = M::q, D = M::D, ND = M::ND, REC = M::REC, NF = PT::NF, DC = PT::DC, DCOV = PT::DCOV, R = PT::R;
  const long long ntr = sp.traj_end - sp.traj_begin;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= ntr * sp.n_samples) return;
  const int smp = (int)(gid % sp.n_samples);  // samples of one trajectory are adjacent: the scratch loads broadcast
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
  auto rec = [&](int slot) { return sp.hist + ((long long)slot * REC) * n + tr; };
  auto outp = [&](int slot) { return sp.out + ((sp.offsets[tr - sp.traj_begin] + slot) * sp.n_samples + smp) * D; };
  auto calib = [&](int rep) -> double {  // factor scale that calibrates a filtered covariance (static models)
    if (!sp.calibrate) return 1.0;
    if (M::IS_EK1) return sqrt(gfin[0]);
    return sqrt(sp.is_mv ? gfin[rep < ND ? rep : 0] : gfin[0]);
  };
  double s[D];  // current sample (natural coordinates)
  {
    // last state: s = mu + S xi   (src/solution_sampling.jl:31-32)
    typename M::State st;
    M::load(st, rec(ns - 1) + (long long)(1 + ND) * n, n);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) s[i] = st.m[i];
PNDE_UNROLL
    for (int rep = 0; rep < PT::NREP; ++rep) {
      const int f = (NF > 1) ? rep : 0;
      double cols[R][DCOV];
      SC::cols_from_factor(PT::factor(st, f), cols);
      const double cs = calib(rep);
PNDE_UNROLL
      for (int c = 0; c < R; c += 2) {
        double a, b;
        rng.normal2((uint32_t)(tr + sp.key_offset), (uint32_t)smp, (uint32_t)(ns - 1), (uint32_t)(rep * 64 + c), a, b);
PNDE_UNROLL
        for (int k = 0; k < DCOV; ++k) {
          s[PT::idx(rep, k)] = fma(cs * cols[c][k], a, s[PT::idx(rep, k)]);
          if (c + 1 < R) s[PT::idx(rep, k)] = fma(cs * cols[c + 1 < R ? c + 1 : 0][k], b, s[PT::idx(rep, k)]);
        }
      }
    }
    double* o = outp(ns - 1);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) o[i] = s[i];
  }
  for (int i = ns - 2; i >= 0; --i) {
    const double h = rec(i + 1)[0] - rec(i)[0];
    if (!sliver_interval(h, rec(i)[0], rec(i + 1)[0], ns, sp.calibrate)) {
      const double* pr = sp.scratch + ((tr - sp.traj_begin) * (sp.max_saved - 1) + i) * SPp::LEN;
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
        constexpr int NR_ = (NF > 1) ? 1 : PT::NREP;  // replicas served by this factor
PNDE_UNROLL
        for (int rr = 0; rr < NR_; ++rr) {
          const int rep = (NF > 1) ? f : rr;
          // y = R-^-T (P s - m^-), delta = X' y  (the gain applied without ever forming it)
          double y[DCOV];
PNDE_UNROLL
          for (int k = 0; k < DCOV; ++k) {
            double acc = fma(Pk[k / DC], s[PT::idx(rep, k)], -pr[D + PT::idx(rep, k)]);
PNDE_UNROLL
            for (int l = 0; l < k; ++l) acc = fma(-Rm[SC::tri(k, l)], y[l], acc);
            y[k] = acc * rinv[k];
          }
          const double ys = (!M::IS_EK1 && sp.calibrate) ? calib(rep) : 1.0;
          double acc[DCOV];
PNDE_UNROLL
          for (int k = 0; k < DCOV; ++k) {
            double dl = 0.0;
PNDE_UNROLL
            for (int l = 0; l < DCOV; ++l) dl = fma(X[l * DCOV + k], y[l], dl);
            acc[k] = pr[PT::idx(rep, k)] + dl;
          }
PNDE_UNROLL
          for (int c = 0; c < R; c += 2) {
            double a, b;
            rng.normal2((uint32_t)(tr + sp.key_offset), (uint32_t)smp, (uint32_t)i, (uint32_t)(rep * 64 + c), a, b);
PNDE_UNROLL
            for (int k = 0; k < DCOV; ++k) {
              acc[k] = fma(ys * Y[c * DCOV + k], a, acc[k]);
              if (c + 1 < R) acc[k] = fma(ys * Y[(c + 1 < R ? c + 1 : 0) * DCOV + k], b, acc[k]);
            }
          }
PNDE_UNROLL
          for (int k = 0; k < DCOV; ++k) snew[PT::idx(rep, k)] = acc[k] * PIk[k / DC];
        }
      }
PNDE_UNROLL
      for (int k = 0; k < D; ++k) s[k] = snew[k];
    }
    double* o = outp(i);
PNDE_UNROLL
    for (int k = 0; k < D; ++k) o[k] = s[k];
  }
}

// Posterior N(mean, cov) of trajectory `tr` at time `tval` (src/solution.jl:165-215): the stored state on an exact
// hit of the grid, otherwise a prediction from the left filtered neighbour, smoothed against the right smoothed
// neighbour when dp.smoothed.  Shared by dense_kernel and the IEKS linearisation point (ieks_kernel.cuh).
template <class M>
__device__ __forceinline__ void dense_state(const DenseParams& dp, long long tr, double tval, double (&mean)[M::D],
                                            double (&cov)[M::D * (M::D + 1) / 2]) {
  using PT = PostTraits<M>;
  using SM = SmoothModel<M>;
  using SC = typename PT::SC;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC, SREC = SM::SREC, NF = PT::NF, DC = PT::DC,
                DCOV = PT::DCOV, R = PT::R;
  const long long n = dp.n;
  const int ns = dp.n_saved[tr];
  auto rec = [&](int slot) { return dp.hist + ((long long)slot * REC) * n + tr; };
  if (ns <= 0 || tval < rec(0)[0]) {  // "Invalid t<t0" (src/solution.jl:169-171): NaN, data not an exception
    for (int i = 0; i < D; ++i) mean[i] = nan("");
    for (int i = 0; i < D * (D + 1) / 2; ++i) cov[i] = nan("");
    return;
  }
  // idx = number of saved times <= tval (binary search), prev = idx - 1
  int lo = 0, hi = ns;  // invariant: t[lo] <= tval (lo valid), t[hi] > tval or hi == ns
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (rec(mid)[0] <= tval) lo = mid; else hi = mid;
  }
  const int prev = lo;
  double gfin[ND];
PNDE_UNROLL
  for (int i = 0; i < ND; ++i) gfin[i] = dp.calibrate ? dp.final_diff[(long long)i * n + tr] : 1.0;
  double dimscale[d];
PNDE_UNROLL
  for (int a = 0; a < d; ++a)
    dimscale[a] = (dp.calibrate && !M::IS_EK1) ? (dp.is_mv ? gfin[a < ND ? a : 0] : gfin[0]) : 1.0;
  const double dense_cal = (dp.calibrate && M::IS_EK1) ? sqrt(gfin[0]) : 1.0;
  // exact hit: the stored state (src/solution.jl:172-176); a query closer than a sliver (filter_kernel.cuh) to the
  // right neighbour of a smoothed solution takes that neighbour (the backward step across it is not evaluated)
  int hit = (rec(prev)[0] == tval) ? prev : -1;
  if (hit < 0 && dp.smoothed && prev + 1 < ns && sliver_interval(rec(prev + 1)[0] - tval, tval, rec(prev + 1)[0], ns, dp.calibrate))
    hit = prev + 1;
  if (hit >= 0) {
    if (dp.smoothed) {
      SM::load_cov(dp.smooth + ((long long)hit * SREC) * n + tr, n, mean, cov);
    } else {
      typename M::State st;
      M::load(st, rec(hit) + (long long)(1 + ND) * n, n);
PNDE_UNROLL
      for (int i = 0; i < D; ++i) mean[i] = st.m[i];
      double sc[q + 1];
PNDE_UNROLL
      for (int k = 0; k <= q; ++k) sc[k] = dense_cal;
      if constexpr (M::IS_EK1) M::final_cov(st, sc, cov, 1); else M::final_cov(st, sc, cov, 1, dimscale);
    }
  } else {
    // extrapolate from the left neighbour (src/solution.jl:184-189)
    const int dslot = (prev + 1 < ns) ? prev + 1 : ns - 1;  // diffusions[min(idx, end)]
    const double* rd = rec(dslot);
    const double h1 = tval - rec(prev)[0];
    double Pk[q + 1], PIk[q + 1];
    precond_scales<q>(h1, Pk, PIk);
    typename M::State st;
    M::load(st, rec(prev) + (long long)(1 + ND) * n, n);
    M::scale(st, Pk);
    apply_A<d, q>(st.m);
    double mp[D];
PNDE_UNROLL
    for (int k = 0; k < D; ++k) mp[k] = st.m[k] * PIk[k / d];  // predicted mean, natural coordinates
    double Lp[NF][SC::NP];                                    // predicted factor(s), natural coordinates
    double sig[NF];
    int status = 0;
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) {
      const double g = dp.calibrate ? (M::IS_EK1 ? gfin[0] : 1.0) : rd[(long long)(1 + (NF > 1 ? f : 0)) * n];
      sig[f] = sqrt(g);
      double cols[R][DCOV];
      SC::cols_from_factor(PT::factor(st, f), cols);
PNDE_UNROLL
      for (int c = 0; c < R; ++c) {
        double w[DCOV];
PNDE_UNROLL
        for (int k = 0; k < DCOV; ++k) w[k] = cols[c][k] * dense_cal;
        apply_A<DC, q>(w);
PNDE_UNROLL
        for (int k = 0; k < DCOV; ++k) cols[c][k] = w[k];
      }
      // factor of A S S' A' + sig^2 Q: triangularise [(A S)' ; sig Q_L'] (predict, src/filtering.jl:33-48)
      double Tt[DCOV][DCOV];
PNDE_UNROLL
      for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
        for (int k = 0; k < DCOV; ++k)
          Tt[c][k] = (k % DC == c % DC && k >= c) ? sig[f] * dp.C.Lt[k / DC][c / DC] : 0.0;  // row c of (sig Q_L)'
      SC::template triangularize<R>(cols, Tt, Lp[f], status);
PNDE_UNROLL
      for (int r = 0; r < DCOV; ++r)
PNDE_UNROLL
        for (int c = 0; c <= r; ++c) Lp[f][SC::tri(r, c)] *= PIk[r / DC];
    }
    double ms[D];
    if (dp.smoothed && prev + 1 < ns) {
      // smooth the prediction against the right smoothed neighbour (src/solution.jl:195-209)
      const double h2 = rec(prev + 1)[0] - tval;
      precond_scales<q>(h2, Pk, PIk);
      const double* sb = dp.smooth + ((long long)(prev + 1) * SREC) * n + tr;
PNDE_UNROLL
      for (int k = 0; k < D; ++k) ms[k] = sb[(long long)k * n];
      double mcur[D], mpred[D];
PNDE_UNROLL
      for (int k = 0; k < D; ++k) mcur[k] = mp[k] * Pk[k / d];
PNDE_UNROLL
      for (int k = 0; k < D; ++k) mpred[k] = mcur[k];
      apply_A<d, q>(mpred);
PNDE_UNROLL
      for (int f = 0; f < NF; ++f) {
        double Ls[SC::NP];
PNDE_UNROLL
        for (int r = 0; r < DCOV; ++r)
PNDE_UNROLL
          for (int c = 0; c <= r; ++c) {
            Ls[SC::tri(r, c)] = sb[(long long)(D + f * SC::NP + SC::tri(r, c)) * n] * Pk[r / DC];
            Lp[f][SC::tri(r, c)] *= Pk[r / DC];
          }
        double cols[DCOV][DCOV];
        SC::cols_from_lower(Lp[f], cols);
        constexpr int NR_ = (NF > 1) ? 1 : PT::NREP;
        double delta[NR_][DCOV];
PNDE_UNROLL
        for (int rr = 0; rr < NR_; ++rr) {
          const int rep = (NF > 1) ? f : rr;
PNDE_UNROLL
          for (int k = 0; k < DCOV; ++k)
            delta[rr][k] = fma(Pk[k / DC], ms[PT::idx(rep, k)], -mpred[PT::idx(rep, k)]);
        }
        SC::template step_cols<DCOV, NR_>(cols, sig[f], dp.C, Ls, delta, status);
PNDE_UNROLL
        for (int rr = 0; rr < NR_; ++rr) {
          const int rep = (NF > 1) ? f : rr;
PNDE_UNROLL
          for (int k = 0; k < DCOV; ++k)
            mp[PT::idx(rep, k)] = (mcur[PT::idx(rep, k)] + delta[rr][k]) * PIk[k / DC];
        }
PNDE_UNROLL
        for (int r = 0; r < DCOV; ++r)
PNDE_UNROLL
          for (int c = 0; c <= r; ++c) Lp[f][SC::tri(r, c)] = Ls[SC::tri(r, c)] * PIk[r / DC];
      }
    }
PNDE_UNROLL
    for (int k = 0; k < D; ++k) mean[k] = mp[k];
    // covariance from the factor(s)
PNDE_UNROLL
    for (int i = 0; i < D; ++i)
PNDE_UNROLL
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
        if constexpr (M::IS_EK1) {
PNDE_UNROLL
          for (int k = 0; k <= j; ++k) acc = fma(Lp[0][SC::tri(i, k)], Lp[0][SC::tri(j, k)], acc);
        } else {
          const int ki = i / d, ai = i % d, kj = j / d, aj = j % d;
          if (ai == aj) {
            const int f = (NF > 1) ? ai : 0;
            const int lo2 = ki < kj ? ki : kj;
PNDE_UNROLL
            for (int k = 0; k <= q; ++k)
              if (k <= lo2) acc = fma(Lp[f][SC::tri(ki, k)], Lp[f][SC::tri(kj, k)], acc);
            acc *= dimscale[ai];
          }
        }
        cov[i * (i + 1) / 2 + j] = acc;
      }
  }
}

// one thread per (trajectory, query time)
template <class M>
__global__ void __launch_bounds__(128) dense_kernel(const DenseParams dp) {
  constexpr int D = M::D;
  const long long ntr = dp.traj_end - dp.traj_begin;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= ntr * dp.n_t) return;
  const long long it = gid / ntr;  // trajectory fastest: record loads coalesce for a common query time
  const long long tr = dp.traj_begin + gid % ntr;
  double* omean = dp.mean + ((tr - dp.traj_begin) * dp.n_t + it) * D;
  double* ocov = dp.cov + ((tr - dp.traj_begin) * dp.n_t + it) * (D * (D + 1) / 2);
  double mean[D], cov[D * (D + 1) / 2];
  dense_state<M>(dp, tr, dp.tq[it], mean, cov);
  for (int i = 0; i < D; ++i) omean[i] = mean[i];
  for (int i = 0; i < D * (D + 1) / 2; ++i) ocov[i] = cov[i];
}

}  // namespace pnde
