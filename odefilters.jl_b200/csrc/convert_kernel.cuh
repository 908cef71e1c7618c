// History conversion kernel (device code only; also compiled at run time by rtc_model.cu).
#pragma once
#include "model_ops.cuh"
#include "smoother_kernel.cuh"

namespace pnde {

// History records -> (t, mean, packed covariance, diffusion) in CSR order; one thread per
// (slot, trajectory) with the trajectory index fastest so that record reads coalesce.
template <class M>
__global__ void __launch_bounds__(128) convert_kernel(const ConvertParams c) {
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC, SREC = SmoothModel<M>::SREC;
  const long long ntr = c.traj_end - c.traj_begin;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ntr * c.max_saved) return;
  const long long slot = idx / ntr;
  const long long tr = c.traj_begin + idx % ntr;
  if (slot >= c.n_saved[tr]) return;
  const long long n = c.n;
  const double* base = c.hist + (slot * REC) * n + tr;
  const long long o = c.offsets[tr - c.traj_begin] + slot;
  if (c.t) c.t[o] = base[0];
  double g[ND];
PNDE_UNROLL
  for (int i = 0; i < ND; ++i) g[i] = c.calibrate ? c.final_diff[(long long)i * n + tr] : base[(long long)(1 + i) * n];
  if (c.diffusion) {
    for (int i = 0; i < c.nd_out; ++i) c.diffusion[o * c.nd_out + i] = (i < ND) ? g[i] : g[0];
  }
  double dimscale[d];
PNDE_UNROLL
  for (int a = 0; a < d; ++a) dimscale[a] = c.calibrate ? (c.is_mv ? g[a < ND ? a : 0] : g[0]) : 1.0;
  // full packed covariance into registers/local, then emit what was asked for
  double mean[D];
  double cov[D * (D + 1) / 2];
  if (c.which == 0) {
    typename M::State st;
    M::load(st, base + (long long)(1 + ND) * n, n);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) mean[i] = st.m[i];
    double sc[q + 1];
PNDE_UNROLL
    for (int k = 0; k <= q; ++k) sc[k] = 1.0;
    if constexpr (M::IS_EK1) {
      const double gs = sqrt(dimscale[0]);
PNDE_UNROLL
      for (int k = 0; k <= q; ++k) sc[k] = gs;
      M::final_cov(st, sc, cov, 1);
    } else {
      M::final_cov(st, sc, cov, 1, dimscale);
    }
  } else {
    const double* sb = c.smooth + (slot * SREC) * n + tr;
    SmoothModel<M>::load_cov(sb, n, mean, cov);
  }
  if (c.sqrt) {
    // D x D square root, row major: filtered states are the reduced-rank factor [W | Lz] padded with d zero columns
    // (calibrated), smoothed states the lower-triangular factor of the smoother
    double* o2 = c.sqrt + o * (D * D);
    for (int i = 0; i < D * D; ++i) o2[i] = 0.0;
    using SMd = SmoothModel<M>;
    using SC = typename SMd::SC;
    constexpr int DC = SMd::DC, NF = SMd::NF, DCOV = DC * (q + 1);
    if (c.which == 0) {
      typename M::State st;
      M::load(st, base + (long long)(1 + ND) * n, n);
      for (int rep = 0; rep < D / DCOV; ++rep) {  // dense EK1: one factor; Kronecker: one replica per dimension
        const Factor<DC, q>* F;
        if constexpr (M::IS_EK1) F = &st.F; else F = &st.F[NF > 1 ? rep : 0];
        const double gs = sqrt(M::IS_EK1 ? dimscale[0] : dimscale[rep < d ? rep : 0]);
        double cols[SC::R][DCOV];
        SC::cols_from_factor(*F, cols);
        for (int cc = 0; cc < SC::R; ++cc)
          for (int k = 0; k < DCOV; ++k) {
            const int row = M::IS_EK1 ? k : k * d + rep, col = M::IS_EK1 ? cc : cc * d + rep;
            o2[row * D + col] = gs * cols[cc][k];
          }
      }
    } else {
      const double* sb = c.smooth + (slot * SREC) * n + tr;
      for (int rep = 0; rep < D / DCOV; ++rep) {
        const int f = NF > 1 ? rep : 0;
        double ds = 1.0;
        if constexpr (!M::IS_EK1) ds = sqrt(sb[(long long)(D + NF * SC::NP + rep) * n]);
        for (int k = 0; k < DCOV; ++k)
          for (int cc = 0; cc <= k; ++cc) {
            const int row = M::IS_EK1 ? k : k * d + rep, col = M::IS_EK1 ? cc : cc * d + rep;
            o2[row * D + col] = ds * sb[(long long)(D + f * SC::NP + SC::tri(k, cc)) * n];
          }
      }
    }
  }
  if (!c.marginals) {
    if (c.mean)
      for (int i = 0; i < D; ++i) c.mean[o * D + i] = mean[i];
    if (c.cov)
      for (int i = 0; i < D * (D + 1) / 2; ++i) c.cov[o * (D * (D + 1) / 2) + i] = cov[i];
  } else {
    if (c.mean)
      for (int i = 0; i < d; ++i) c.mean[o * d + i] = mean[i];
    if (c.cov)
      for (int i = 0; i < d * (d + 1) / 2; ++i) c.cov[o * (d * (d + 1) / 2) + i] = cov[i];
  }
}

}  // namespace pnde
