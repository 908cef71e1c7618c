// One attempted filter step from caller-supplied states ("teacher forcing", SURVEY 8c protocol (i); also what a
// Julia step!/callback user drives, examples/fitzhughnagumo_animation.jl:23-26): perform_step!
// (src/perform_step.jl:27-93) applied to n independent (mu, S, t, dt, p, u_prev) tuples, through exactly the device
// functions the persistent filter kernel uses (M::scale, M::step).
//
// The caller hands a full square root S (D x D, Sigma = S S', SRMatrix.squareroot of src/squarerootmatrix.jl:10-16);
// the kernels carry the reduced-rank factor [W | Lz] (cov_engine.cuh).  import_factor rotates S from the right with
// Householder reflections into that form -- possible exactly when S S' is a state the filter can be in: zero (the
// initial state) or a filter posterior (rank <= D - d, block 1 slaved to block 0 by H S = 0).  Anything else is
// reported per trajectory (status 1), not silently projected.
#pragma once
#include "model_ops.cuh"
#include "smoother_kernel.cuh"

namespace pnde {

struct StepParams {
  long long n;
  const double* mean;   // [D][n]
  const double* sqrt;   // [D * D][n], entry (i, j) at (i * D + j) * n
  const double* t;      // [n]
  const double* dt;     // [n]
  const double* p;      // [np][n]
  const double* uprev;  // [d][n]  integ.u before the step (error norm, src/perform_step.jl:80-83); may be null
  double* mean_out;     // [D][n]      x_filt.mu
  double* cov_out;      // [D(D+1)/2][n] packed lower x_filt.Sigma
  double* sigma2;       // [ND][n]     local diffusion of this step
  double* eest;         // [n]         EEst (src/perform_step.jl:84); 0 when uprev is null
  double* u_out;        // [d][n]      integ.u after the step
  double* quad_logdet;  // [2][n]      z' S^-1 z and log det S (log-likelihood term of :66)
  int* status;          // [n]  0 ok, 1 state not representable as a filter posterior, 2 non-finite
  int diffusion;
  double abstol, reltol;
  IwpConsts C;
};

// Rotate S (Dc rows = the coordinates of one covariance factor, NCOLS >= Dc - dc columns) from the right into the
// form of Factor<dc, q>: d dense columns W, then columns that vanish in blocks 0 and 1 and are lower triangular below.
// Returns the largest entry that should have vanished, relative to the largest entry of S.
template <int dc, int q, int NCOLS>
__device__ double import_factor(double (&S)[dc * (q + 1)][NCOLS], Factor<dc, q>& F) {
  constexpr int Dc = dc * (q + 1), R = Dc - dc, NZ = Dc - 2 * dc;
  double big = 0.0;
  for (int i = 0; i < Dc; ++i)
    for (int j = 0; j < NCOLS; ++j) big = fmax(big, fabs(S[i][j]));
  // elimination order of the rows: block 0, blocks 2..q (block 1 is slaved: it must come out by itself)
  for (int pcol = 0; pcol < R; ++pcol) {
    const int row = (pcol < dc) ? pcol : pcol + dc;
    double n2 = 0.0;
    for (int j = pcol; j < NCOLS; ++j) n2 = fma(S[row][j], S[row][j], n2);
    if (!(n2 > 0.0)) continue;
    const double nrm = sqrt(n2), pv = S[row][pcol];
    const double snrm = copysign(nrm, pv), v0 = pv + snrm;
    const double beta = 1.0 / fma(fabs(pv), nrm, n2);
    for (int i = 0; i < Dc; ++i) {
      if (i == row) continue;
      double w = v0 * S[i][pcol];
      for (int j = pcol + 1; j < NCOLS; ++j) w = fma(S[row][j], S[i][j], w);
      const double s = beta * w;
      S[i][pcol] = fma(-s, v0, S[i][pcol]);
      for (int j = pcol + 1; j < NCOLS; ++j) S[i][j] = fma(-s, S[row][j], S[i][j]);
    }
    S[row][pcol] = -snrm;
    for (int j = pcol + 1; j < NCOLS; ++j) S[row][j] = 0.0;
  }
  // what must be zero now: columns R.. (all rows) and block 1 in the triangular columns
  double bad = 0.0;
  for (int i = 0; i < Dc; ++i)
    for (int j = R; j < NCOLS; ++j) bad = fmax(bad, fabs(S[i][j]));
  for (int i = dc; i < 2 * dc; ++i)
    for (int j = dc; j < R; ++j) bad = fmax(bad, fabs(S[i][j]));
  for (int a = 0; a < dc; ++a)
    for (int i = 0; i < Dc; ++i) F.W[a][i] = S[i][a];
  for (int j = 0; j < NZ; ++j)
    for (int i = j; i < NZ; ++i) F.Lz[Factor<dc, q>::lz(j, i)] = S[2 * dc + i][dc + j];
  if (NZ == 0) F.Lz[0] = 0.0;
  return big > 0.0 ? bad / big : 0.0;
}

template <class M>
__global__ void __launch_bounds__(64) step_kernel(const StepParams sp) {
  using VF = typename M::VF;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= sp.n) return;
  const long long n = sp.n;
  double p[VF::np];
  for (int i = 0; i < VF::np; ++i) p[i] = sp.p[(long long)i * n + tid];
  typename M::State st;
  for (int i = 0; i < D; ++i) st.m[i] = sp.mean[(long long)i * n + tid];
  int status = 0;
  double worst = 0.0;
  if constexpr (M::IS_EK1) {
    double S[D][D];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) S[i][j] = sp.sqrt[(long long)(i * D + j) * n + tid];
    worst = import_factor<d, q, D>(S, st.F);
  } else {
    // Kronecker models (Sigma = Ctilde (x) I_d, or one Ctilde per dimension for dynamicMV): the rows (k, f), k = 0..q,
    // of S are a square root of factor f's Ctilde
    for (int f = 0; f < M::NF; ++f) {
      double S[q + 1][D];
      for (int k = 0; k <= q; ++k)
        for (int j = 0; j < D; ++j) S[k][j] = sp.sqrt[(long long)((k * d + f) * D + j) * n + tid];
      worst = fmax(worst, import_factor<1, q, D>(S, st.F[f]));
    }
  }
  if (worst > 1e-7) status = 1;
  const double dt = sp.dt[tid];
  double Pk[q + 1], PIk[q + 1];
  precond_scales<q>(dt, Pk, PIk);
  M::scale(st, Pk);  // x = P x  (src/perform_step.jl:38)
  double unew[d], err[d], local[ND], quad, detS;
  for (int i = 0; i < ND; ++i) local[i] = 1.0;
  M::step(st, p, PIk[0], PIk[1], Pk[1], sp.diffusion, sp.C, unew, err, local, quad, detS);
  M::scale(st, PIk);  // back to natural coordinates (:73-75)
  double EEst = 0.0;
  if (sp.uprev) {
    double acc = 0.0;
    for (int i = 0; i < d; ++i) {
      const double up = sp.uprev[(long long)i * n + tid];
      const double r = dt * err[i] / (sp.abstol + fmax(fabs(up), fabs(unew[i])) * sp.reltol);
      acc = fma(r, r, acc);
    }
    EEst = sqrt(acc / double(d));
  }
  for (int i = 0; i < D; ++i) {
    if (!(fabs(st.m[i]) <= 1.79769313486231570e308)) status = 2;
    if (sp.mean_out) sp.mean_out[(long long)i * n + tid] = st.m[i];
  }
  double one[q + 1];
  for (int k = 0; k <= q; ++k) one[k] = 1.0;
  if (sp.cov_out) {
    if constexpr (M::IS_EK1) {
      M::final_cov(st, one, sp.cov_out + tid, n);
    } else {
      double ds[d];
      for (int a = 0; a < d; ++a) ds[a] = 1.0;
      M::final_cov(st, one, sp.cov_out + tid, n, ds);
    }
  }
  if (sp.sigma2)
    for (int i = 0; i < ND; ++i) sp.sigma2[(long long)i * n + tid] = local[i];
  if (sp.eest) sp.eest[tid] = EEst;
  if (sp.u_out)
    for (int i = 0; i < d; ++i) sp.u_out[(long long)i * n + tid] = unew[i];
  if (sp.quad_logdet) {
    sp.quad_logdet[tid] = quad;
    sp.quad_logdet[n + tid] = 2.0 * log(detS);
  }
  sp.status[tid] = status;
}

}  // namespace pnde
