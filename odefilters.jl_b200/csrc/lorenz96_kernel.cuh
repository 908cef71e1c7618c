// Large-d single solves (BASELINE config 4): Lorenz-96, EK0 with the Kronecker-factored covariance
// Sigma = Ctilde (x) I_d.  One CTA per trajectory, `DPT` dimensions per thread:
// the mean M in R^{(q+1) x d} lives in registers (q+1 values per owned dimension), the (q+1) x q
// covariance factor is replicated in every thread (its step costs ~60 flops), the vector field reads
// its three neighbours through shared memory, and ||z||^2 (and EEst when adaptive) are block
// reductions.  Per step: 3 block-wide syncs fixed-step, 5 adaptive -- the kernel is latency bound
// (one sequential chain), reported as microseconds per step, not as a roofline fraction.
//
// Reference path: identical to filter_kernel.cuh (perform_step! src/perform_step.jl:27-93 with
// alg isa EK0, H = E1 PI, src/perform_step.jl:127); the reference itself is dense D x D.
#pragma once
#include "filter_kernel.cuh"
#include "smoother_kernel.cuh"

namespace pnde {

struct LorenzParams {
  long long n;        // trajectories (one CTA each)
  int d;              // ODE dimension
  const double* u0;   // [d][n]
  const double* p;    // [1][n]  forcing F
  double* mean;       // [D][n]
  double* cov;        // [(q+1)(q+2)/2][n]  packed lower triangle of Ctilde (Sigma = Ctilde (x) I_d)
  double* t_final;
  double* loglik;
  double* final_diff;
  int* retcode;
  int* naccept;
  int* nreject;
  int* nf;
  int* n_saved;
  double* hist;        // [max_saved][1 + 1 + D + LEN][n]
  long long max_saved;
  int save_mode, save_stride, diffusion, adaptive;
  IwpConsts C;
  CtrlParams K;
};

constexpr int LORENZ_THREADS = 256;

// sum over the block; every thread gets the total.  red: >= 8 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();  // protect red from the previous reduction's readers
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < LORENZ_THREADS / 32; ++i) tot += red[i];
  return tot;
}

template <int q, int DPT>
__global__ void __launch_bounds__(LORENZ_THREADS) lorenz96_kernel(const LorenzParams prm) {
  extern __shared__ double sm[];  // [ (q+1) * d ] jets during init, then u-hat in sm[0..d)
  __shared__ double red[LORENZ_THREADS / 32];
  const int d = prm.d;
  const long long n = prm.n, tr = blockIdx.x;
  const CtrlParams& K = prm.K;
  const int diffusion = prm.diffusion;
  const bool adaptive = prm.adaptive != 0;
  const bool is_static = (diffusion == DIFF_FIXED || diffusion == DIFF_FIXED_MAP);
  const double F = prm.p[tr];
  using Fac = Factor<1, q>;
  // owned dimensions: i = threadIdx.x + j * LORENZ_THREADS (interleaved: neighbouring threads own
  // neighbouring dimensions, shared-memory neighbour reads are conflict free)
  int own[DPT];
  bool act[DPT];
#pragma unroll
  for (int j = 0; j < DPT; ++j) {
    own[j] = threadIdx.x + j * LORENZ_THREADS;
    act[j] = own[j] < d;
  }
  auto wrap = [&](int i) { return i < 0 ? i + d : (i >= d ? i - d : i); };

  // ---- initial_update!: Taylor-mode jets through shared memory (src/state_initialization.jl:15-42)
  double m[DPT][q + 1];
#pragma unroll
  for (int j = 0; j < DPT; ++j) {
#pragma unroll
    for (int k = 0; k <= q; ++k) m[j][k] = 0.0;
    if (act[j]) {
      m[j][0] = prm.u0[(long long)own[j] * n + tr];
      sm[own[j]] = m[j][0];  // coefficient 0 plane
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < q; ++k) {
    // coefficient k of f_i = [(x_{i+1} - x_{i-2}) * x_{i-1}]_k - x_i[k] + F delta_k0 ; c_{k+1} = f_k / (k+1)
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      if (act[j]) {
        const int i = own[j];
        const int ip = wrap(i + 1), im1 = wrap(i - 1), im2 = wrap(i - 2);
        double acc = 0.0;
        for (int a = 0; a <= k; ++a)
          acc = fma(sm[a * d + ip] - sm[a * d + im2], sm[(k - a) * d + im1], acc);
        acc -= sm[k * d + i];
        if (k == 0) acc += F;
        m[j][k + 1] = acc / double(k + 1);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DPT; ++j)
      if (act[j]) sm[(k + 1) * d + own[j]] = m[j][k + 1];
    __syncthreads();
  }
  {
    double fact = 1.0;
#pragma unroll
    for (int k = 1; k <= q; ++k) {
      fact *= double(k);
#pragma unroll
      for (int j = 0; j < DPT; ++j) m[j][k] *= fact;  // u^(k) = k! c_k
    }
  }
  Fac Fc;
  Fc.zero();

  double t = K.t0;
  int iter = 0, nacc = 0, nrej = 0, nfe = 0, ret = RET_SUCCESS, nsaved = 0;
  double gsaved = 1.0;
  double uprev[DPT];
#pragma unroll
  for (int j = 0; j < DPT; ++j) uprev[j] = m[j][0];
  double ll_quad = 0.0, ll_logdet = 0.0;
  int ll_n = 0;
  constexpr int REC = 2 + Fac::LEN;  // t, diffusion, factor; the mean follows as D entries
  const int D = d * (q + 1);
  auto save = [&](double tt, const double (&sc)[q + 1]) {
    if (nsaved >= prm.max_saved) {
      ret = RET_HISTORY_FULL;
      return;
    }
    double* base = prm.hist + ((long long)nsaved * (REC + D)) * n + tr;
    if (threadIdx.x == 0) {
      base[0] = tt;
      base[n] = gsaved;
      Fac tmp = Fc;
      tmp.scale_blocks(sc);
      tmp.store(base + 2 * n, n);
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j)
      if (act[j])
#pragma unroll
        for (int k = 0; k <= q; ++k) base[(long long)(REC + k * d + own[j]) * n] = m[j][k] * sc[k];
    ++nsaved;
  };
  double Pk[q + 1], PIk[q + 1], ones[q + 1];
#pragma unroll
  for (int k = 0; k <= q; ++k) Pk[k] = PIk[k] = ones[k] = 1.0;
  if (prm.save_mode != SAVE_FINAL) save(t, ones);

  double dt;
  if (adaptive && !(K.dt > 0.0)) {
    // Hairer initdt with block reductions
    double s0 = 0.0, s1 = 0.0;
    double f0[DPT], sk[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      f0[j] = sk[j] = 0.0;
      if (act[j]) {
        f0[j] = m[j][1];  // f(u0) is the first derivative of the initial state
        sk[j] = K.abstol + fabs(m[j][0]) * K.reltol;
        const double a = m[j][0] / sk[j], b = f0[j] / sk[j];
        s0 = fma(a, a, s0);
        s1 = fma(b, b, s1);
      }
    }
    const double d0 = sqrt(block_sum(s0, red) / d), d1 = sqrt(block_sum(s1, red) / d);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
    dt0 = fmin(dt0, K.dtmax);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DPT; ++j)
      if (act[j]) sm[own[j]] = fma(dt0, f0[j], m[j][0]);
    __syncthreads();
    double s2 = 0.0;
#pragma unroll
    for (int j = 0; j < DPT; ++j)
      if (act[j]) {
        const int i = own[j];
        const double f1 = (sm[wrap(i + 1)] - sm[wrap(i - 2)]) * sm[wrap(i - 1)] - sm[i] + F;
        const double c = (f1 - f0[j]) / sk[j];
        s2 = fma(c, c, s2);
      }
    const double d2 = sqrt(block_sum(s2, red) / d) / dt0;
    const double mx = fmax(d1, d2);
    const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / double(q + 1));
    dt = (dt0 < 10.0 * 2.220446049250313e-16) ? 1e-6 : fmin(fmin(100.0 * dt0, dt1), K.dtmax);
    nfe += 2;
  } else {
    dt = K.dt;
  }
  const double lqold0 = ctrl_state_init(K);
  double dtpropose = dt, lqold = lqold0, q11 = 1.0, hcur = -1.0;
  bool accepted_prev = true;

  while (t < K.t1) {
    if (iter > 0) dt = accepted_prev ? dtpropose : dt / fmin(1.0 / K.qmin, q11 / K.gamma);
    ++iter;
    if (iter > K.maxiters) {
      ret = RET_MAXITERS;
      break;
    }
    if (adaptive) {
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
    // state is kept in P(hcur) coordinates; re-scale when h changes (every step when adaptive)
    if (dt != hcur) {
      double Pn[q + 1], PIn[q + 1], sc[q + 1];
      precond_scales<q>(dt, Pn, PIn);
#pragma unroll
      for (int k = 0; k <= q; ++k) {
        sc[k] = Pn[k] * PIk[k];
        Pk[k] = Pn[k];
        PIk[k] = PIn[k];
      }
#pragma unroll
      for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int k = 0; k <= q; ++k) m[j][k] *= sc[k];
      Fc.scale_blocks(sc);
      hcur = dt;
    }
    const double pi0 = PIk[0], pi1 = PIk[1], ipi1 = Pk[1];
    // on rejection the step is undone from these copies
    double mold[DPT][q + 1];
    Fac Fold;
    if (adaptive) {
#pragma unroll
      for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int k = 0; k <= q; ++k) mold[j][k] = m[j][k];
      Fold = Fc;
    }
    // predict mean, u-hat to shared memory
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
#pragma unroll
      for (int k = 0; k <= q; ++k) {
        double acc = m[j][k];
#pragma unroll
        for (int kk = k + 1; kk <= q; ++kk)
          acc = (kk - k == 1) ? acc + m[j][kk] : fma(inv_factorial(kk - k), m[j][kk], acc);
        m[j][k] = acc;
      }
      if (act[j]) sm[own[j]] = pi0 * m[j][0];
    }
    __syncthreads();
    double z[DPT], fu[DPT], zz = 0.0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      z[j] = fu[j] = 0.0;
      if (act[j]) {
        const int i = own[j];
        fu[j] = (sm[wrap(i + 1)] - sm[wrap(i - 2)]) * sm[wrap(i - 1)] - sm[i] + F;
        z[j] = fma(pi1, m[j][1], -fu[j]);
        zz = fma(z[j], z[j], zz);
      }
    }
    ++nfe;
    zz = block_sum(zz, red);
    const double B = pi1 * pi1 * prm.C.Qt[1][1];
    double local = 1.0, sig = 1.0;
    if (diffusion == DIFF_DYNAMIC) {
      local = zz / (double(d) * B);
      sig = sqrt(local);
    }
    const double Jp0[1][1] = {{0.0}};
    double R[1][q + 1], Ri[1];
    cov_filter_step<1, q, false>(Fc, Jp0, sig, pi1, ipi1, prm.C, R, Ri);
    const double quad = zz * Ri[0] * Ri[0];  // z' S^-1 z with S = R00^2 I
    if (is_static) local = quad / double(d);
    double racc = 0.0;
    bool finite = true;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      const double ya = z[j] * Ri[0];
      m[j][0] = fma(-R[0][1], ya, m[j][0]);
#pragma unroll
      for (int k = 2; k <= q; ++k) m[j][k] = fma(-R[0][k], ya, m[j][k]);
      m[j][1] = fu[j] * ipi1;
      if (act[j]) {
        const double un = pi0 * m[j][0];
        const double e = sqrt(local * B);
        const double r = dt * e / (K.abstol + fmax(fabs(uprev[j]), fabs(un)) * K.reltol);
        racc = fma(r, r, racc);
        uprev[j] = un;
        finite = finite && (fabs(un) <= 1.79769313486231570e308);
      }
    }
    double gcur;
    if (diffusion == DIFF_DYNAMIC) {
      gcur = local;
    } else if (diffusion == DIFF_FIXED) {
      gcur = (nacc == 0) ? local : gsaved + (local - gsaved) / double(nacc);
    } else {
      const double Nn = double(nacc + 1), al = 0.5, be = 0.5;
      if (nacc == 0)
        gcur = (be + 0.5 * local) / (al + Nn * d / 2.0 + 1.0);
      else
        gcur = (be + 0.5 * ((gsaved * (al + (Nn - 1.0) * d / 2.0 + 1.0) - be) * 2.0 + local)) / (al + Nn * d / 2.0 + 1.0);
    }
    double EEst = 0.0;
    if (adaptive) EEst = sqrt(block_sum(racc, red) / double(d));
    const bool anynf = block_sum(finite ? 0.0 : 1.0, red) > 0.0;
    const bool commit = !adaptive || (EEst < 1.0);
    const bool accept = !adaptive || (EEst <= 1.0);
    if (!commit) {
#pragma unroll
      for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int k = 0; k <= q; ++k) m[j][k] = mold[j][k];
      Fc = Fold;
    } else {
      ll_quad += quad;
      ll_logdet += double(d) * log(fabs(R[0][0]));
      ++ll_n;
    }
    if (anynf) {
      ret = RET_NONFINITE;
      break;
    }
    const double ttmp = t + dt;
    if (adaptive) {
      double lE;
      double qc = controller_factor(EEst, K, lqold, lE);
      if (accept) {
        ++nacc;
        if (K.qsteady_min <= qc && qc <= K.qsteady_max) qc = 1.0;
        lqold = ctrl_state_accept(EEst, lE, lqold0, K);  // qold = max(EEst, qoldinit)
        t = (fabs(ttmp - K.t1) < 10.0 * ulp_of(fmax(t, K.t1))) ? K.t1 : ttmp;
        dtpropose = fmax(K.dtmin, fmin(K.dtmax, dt / qc));
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
      gsaved = gcur;
      const bool want = (prm.save_mode == SAVE_EVERY) ||
                        (prm.save_mode == SAVE_STRIDE && (nacc % prm.save_stride == 0 || !(t < K.t1)));
      if (want) {
        save(t, PIk);
        if (ret == RET_HISTORY_FULL) break;
      }
    }
  }
  // ---- outputs (natural coordinates) ----
  const double cal = (is_static && nacc > 0) ? gsaved : 1.0;
#pragma unroll
  for (int j = 0; j < DPT; ++j)
    if (act[j])
#pragma unroll
      for (int k = 0; k <= q; ++k) prm.mean[(long long)(k * d + own[j]) * n + tr] = m[j][k] * PIk[k];
  if (threadIdx.x == 0) {
    // packed lower triangle of Ctilde = cal * PI F F' PI
#pragma unroll
    for (int i = 0; i <= q; ++i)
#pragma unroll
      for (int jj = 0; jj <= i; ++jj) {
        double v = Fc.W[0][i] * Fc.W[0][jj];
#pragma unroll
        for (int c = 0; c < Fac::NZ; ++c)
          if (i >= 2 + c && jj >= 2 + c) v = fma(Fc.Lz[Fac::lz(c, i - 2)], Fc.Lz[Fac::lz(c, jj - 2)], v);
        prm.cov[(long long)(i * (i + 1) / 2 + jj) * n + tr] = v * PIk[i] * PIk[jj] * cal;
      }
    double ll = -0.5 * (ll_quad + 2.0 * ll_logdet + double(ll_n) * double(d) * 1.8378770664093453);
    if (is_static && nacc > 0) ll = nan("");
    prm.final_diff[tr] = gsaved;
    prm.t_final[tr] = t;
    prm.loglik[tr] = ll;
    prm.retcode[tr] = ret;
    prm.naccept[tr] = nacc;
    prm.nreject[tr] = nrej;
    prm.nf[tr] = nfe;
    prm.n_saved[tr] = nsaved;
  }
}

template <int q>
cudaError_t launch_lorenz_q(const LorenzParams& prm, cudaStream_t s) {
  const int dpt = (prm.d + LORENZ_THREADS - 1) / LORENZ_THREADS;
  const size_t smem = (size_t)(q + 1) * prm.d * sizeof(double);
  const unsigned grid = (unsigned)prm.n;
#define PNDE_LZ(DPT)                                                                                      \
  {                                                                                                       \
    cudaError_t e = cudaFuncSetAttribute(lorenz96_kernel<q, DPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                      \
    if (e != cudaSuccess) return e;                                                                       \
    lorenz96_kernel<q, DPT><<<grid, LORENZ_THREADS, smem, s>>>(prm);                                      \
    return cudaGetLastError();                                                                            \
  }
  if (dpt <= 1) PNDE_LZ(1)
  if (dpt <= 2) PNDE_LZ(2)
  if (dpt <= 4) PNDE_LZ(4)
  if (dpt <= 8) PNDE_LZ(8)
#undef PNDE_LZ
  return cudaErrorInvalidValue;
}

inline cudaError_t launch_lorenz(int q, const LorenzParams& prm, cudaStream_t s) {
  switch (q) {
    case 1: return launch_lorenz_q<1>(prm, s);
    case 2: return launch_lorenz_q<2>(prm, s);
    case 3: return launch_lorenz_q<3>(prm, s);
    case 4: return launch_lorenz_q<4>(prm, s);
    case 5: return launch_lorenz_q<5>(prm, s);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------------------------
// History of the large-d Kronecker path: RTS smoother and record conversion.
//   filtered record  [slot][2 + LEN + D][n]:  t, diffusion, factor of Ctilde (Factor<1, q>), mean (k-major: k d + i)
//   smoothed record  [slot][NP + 1 + D][n]:   packed lower factor of the smoothed Ctilde, calibration scale, mean
// Reference: smooth_all! / smooth!  src/smoothing.jl:4-63 on Sigma = Ctilde (x) I_d: the gain is G~ (x) I_d, so the
// covariance recursion is the (q+1) x (q+1) one (every thread repeats it, ~1 kflop) and the mean recursion applies the
// same small gain to every dimension independently: no communication between the threads of the CTA at all.
// ---------------------------------------------------------------------------------------------
struct LorenzSmoothParams {
  long long n;
  int d;
  long long max_saved;
  const int* n_saved;
  const double* hist;
  double* smooth;
  const double* final_diff;  // [n]
  int calibrate;             // static diffusion model: sigma^2 = 1 inside, the final global value scales the output
  int* status;
  IwpConsts C;
};

template <int q, int DPT>
__global__ void __launch_bounds__(LORENZ_THREADS) lorenz96_smoother_kernel(const LorenzSmoothParams sp) {
  using Fac = Factor<1, q>;
  using SC = SmoothCov<1, q>;
  constexpr int REC = 2 + Fac::LEN, NP = SC::NP;
  const int d = sp.d, D = d * (q + 1);
  const long long n = sp.n, tr = blockIdx.x;
  const int ns = sp.n_saved[tr];
  if (ns <= 0) {
    if (threadIdx.x == 0) sp.status[tr] = 0;
    return;
  }
  int own[DPT];
  bool act[DPT];
#pragma unroll
  for (int j = 0; j < DPT; ++j) {
    own[j] = threadIdx.x + j * LORENZ_THREADS;
    act[j] = own[j] < d;
  }
  const double gfin = sp.calibrate ? sp.final_diff[tr] : 1.0;
  auto rec = [&](int slot) { return sp.hist + ((long long)slot * (REC + D)) * n + tr; };
  auto srec = [&](int slot) { return sp.smooth + ((long long)slot * (NP + 1 + D)) * n + tr; };
  int status = 0;
  double ms[DPT][q + 1];  // smoothed mean of the own dimensions at i + 1 (natural coordinates)
  double Ls[NP];          // smoothed factor of Ctilde at i + 1 (natural coordinates), the same in every thread
  auto load_mean = [&](const double* r, double (&mm)[DPT][q + 1]) {
#pragma unroll
    for (int j = 0; j < DPT; ++j)
#pragma unroll
      for (int k = 0; k <= q; ++k) mm[j][k] = act[j] ? r[(long long)(REC + k * d + own[j]) * n] : 0.0;
  };
  auto write = [&](int slot) {
    double* o = srec(slot);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int e = 0; e < NP; ++e) o[(long long)e * n] = Ls[e];
      o[(long long)NP * n] = gfin;
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j)
      if (act[j])
#pragma unroll
        for (int k = 0; k <= q; ++k) o[(long long)(NP + 1 + k * d + own[j]) * n] = ms[j][k];
  };
  auto from_filtered = [&](int slot) {
    const double* r = rec(slot);
    Fac F;
    F.load(r + 2 * n, n);
    load_mean(r, ms);
    double Y[SC::R][q + 1], Tt[q + 1][q + 1];
    SC::cols_from_factor(F, Y);
#pragma unroll
    for (int c = 0; c <= q; ++c)
#pragma unroll
      for (int k = 0; k <= q; ++k) Tt[c][k] = 0.0;
    SC::template triangularize<SC::R>(Y, Tt, Ls, status);
  };
  from_filtered(ns - 1);
  write(ns - 1);
  for (int i = ns - 2; i >= 1; --i) {
    const double* ri = rec(i);
    const double* rn = rec(i + 1);
    const double h = rn[0] - ri[0];
    if (!sliver_interval(h, ri[0], rn[0], ns, sp.calibrate)) {  // (h == 0 or a sliver: the state is carried across, filter_kernel.cuh)
      double Pk[q + 1], PIk[q + 1];
      precond_scales<q>(h, Pk, PIk);
      Fac F;
      F.load(ri + 2 * n, n);
      F.scale_blocks(Pk);
      const double sg = sp.calibrate ? 1.0 : sqrt(rn[n]);  // the interval's diffusion is stored with state i + 1
      double m[DPT][q + 1], delta[DPT][q + 1];
      load_mean(ri, m);
#pragma unroll
      for (int j = 0; j < DPT; ++j) {
        double mp[q + 1];
#pragma unroll
        for (int k = 0; k <= q; ++k) {
          m[j][k] *= Pk[k];
          mp[k] = m[j][k];
        }
        apply_A<1, q>(mp);
#pragma unroll
        for (int k = 0; k <= q; ++k) delta[j][k] = fma(Pk[k], ms[j][k], -mp[k]);
      }
#pragma unroll
      for (int r = 0; r <= q; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) Ls[SC::tri(r, c)] *= Pk[r];
      SC::template step<DPT>(F, sg, sp.C, Ls, delta, status);
#pragma unroll
      for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int k = 0; k <= q; ++k) {
          ms[j][k] = (m[j][k] + delta[j][k]) * PIk[k];
          if (act[j] && !(ms[j][k] == ms[j][k])) status |= 1;
        }
#pragma unroll
      for (int r = 0; r <= q; ++r) {
#pragma unroll
        for (int c = 0; c <= r; ++c) Ls[SC::tri(r, c)] *= PIk[r];
        if (!(Ls[SC::tri(r, r)] == Ls[SC::tri(r, r)])) status |= 1;
      }
    }
    write(i);
  }
  if (ns >= 2) {
    from_filtered(0);  // the first state is never smoothed (src/smoothing.jl:11)
    write(0);
  }
  status = __syncthreads_or(status);
  if (threadIdx.x == 0) sp.status[tr] = status;
}

template <int q>
cudaError_t launch_lorenz_smooth_q(const LorenzSmoothParams& sp, cudaStream_t s) {
  const int dpt = (sp.d + LORENZ_THREADS - 1) / LORENZ_THREADS;
  const unsigned grid = (unsigned)sp.n;
  if (dpt <= 1) { lorenz96_smoother_kernel<q, 1><<<grid, LORENZ_THREADS, 0, s>>>(sp); return cudaGetLastError(); }
  if (dpt <= 2) { lorenz96_smoother_kernel<q, 2><<<grid, LORENZ_THREADS, 0, s>>>(sp); return cudaGetLastError(); }
  if (dpt <= 4) { lorenz96_smoother_kernel<q, 4><<<grid, LORENZ_THREADS, 0, s>>>(sp); return cudaGetLastError(); }
  if (dpt <= 8) { lorenz96_smoother_kernel<q, 8><<<grid, LORENZ_THREADS, 0, s>>>(sp); return cudaGetLastError(); }
  return cudaErrorInvalidValue;
}
inline cudaError_t launch_lorenz_smooth(int q, const LorenzSmoothParams& sp, cudaStream_t s) {
  switch (q) {
    case 1: return launch_lorenz_smooth_q<1>(sp, s);
    case 2: return launch_lorenz_smooth_q<2>(sp, s);
    case 3: return launch_lorenz_smooth_q<3>(sp, s);
    case 4: return launch_lorenz_smooth_q<4>(sp, s);
    case 5: return launch_lorenz_smooth_q<5>(sp, s);
    default: return cudaErrorInvalidValue;
  }
}

// History records -> (t, mean, Ctilde packed, diffusion) in CSR order; one CTA per (slot, trajectory).
// marginals: u [total][d] and ONE covariance entry per state, Ctilde[0][0] (Sigma_u = Ctilde[0][0] I_d).
struct LorenzConvertParams {
  long long n, traj_begin, traj_end, max_saved;
  int d;
  const int* n_saved;
  const long long* offsets;
  const double* hist;
  const double* smooth;
  const double* final_diff;
  int which, calibrate, marginals;
  double* t;
  double* mean;
  double* cov;
  double* diffusion;
};

template <int q>
__global__ void __launch_bounds__(256) lorenz_convert_kernel(const LorenzConvertParams c) {
  using Fac = Factor<1, q>;
  using SC = SmoothCov<1, q>;
  constexpr int REC = 2 + Fac::LEN, NP = SC::NP;
  const long long ntr = c.traj_end - c.traj_begin;
  const long long slot = blockIdx.x / ntr;
  const long long tr = c.traj_begin + blockIdx.x % ntr;
  if (slot >= c.n_saved[tr]) return;
  const int d = c.d, D = d * (q + 1);
  const long long n = c.n;
  const long long o = c.offsets[tr - c.traj_begin] + slot;
  const double* fr = c.hist + (slot * (REC + D)) * n + tr;
  const double* sr = c.which ? c.smooth + (slot * (NP + 1 + D)) * n + tr : nullptr;
  const double* mean = c.which ? sr + (long long)(NP + 1) * n : fr + (long long)REC * n;
  const int DM = c.marginals ? d : D;
  if (c.mean)
    for (int i = threadIdx.x; i < DM; i += blockDim.x) c.mean[o * DM + i] = mean[(long long)i * n];
  if (threadIdx.x == 0) {
    const double g = c.calibrate ? c.final_diff[tr] : fr[n];
    if (c.t) c.t[o] = fr[0];
    if (c.diffusion) c.diffusion[o] = g;
    double Ct[NP];
    if (c.which == 0) {
      Fac F;
      F.load(fr + 2 * n, n);
      const double cal = c.calibrate ? g : 1.0;
#pragma unroll
      for (int i = 0; i <= q; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          double v = F.W[0][i] * F.W[0][j];
#pragma unroll
          for (int cc = 0; cc < Fac::NZ; ++cc)
            if (i >= 2 + cc && j >= 2 + cc) v = fma(F.Lz[Fac::lz(cc, i - 2)], F.Lz[Fac::lz(cc, j - 2)], v);
          Ct[SC::tri(i, j)] = v * cal;
        }
    } else {
      double L[NP];
#pragma unroll
      for (int e = 0; e < NP; ++e) L[e] = sr[(long long)e * n];
      const double ds = sr[(long long)NP * n];
#pragma unroll
      for (int i = 0; i <= q; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          double v = 0.0;
#pragma unroll
          for (int k = 0; k <= j; ++k) v = fma(L[SC::tri(i, k)], L[SC::tri(j, k)], v);
          Ct[SC::tri(i, j)] = v * ds;
        }
    }
    if (c.cov) {
      if (c.marginals) {
        c.cov[o] = Ct[0];
      } else {
#pragma unroll
        for (int e = 0; e < NP; ++e) c.cov[o * NP + e] = Ct[e];
      }
    }
  }
}

inline cudaError_t launch_lorenz_convert(int q, const LorenzConvertParams& c, cudaStream_t s) {
  const long long blocks = (c.traj_end - c.traj_begin) * c.max_saved;
  if (blocks <= 0) return cudaSuccess;
  switch (q) {
    case 1: lorenz_convert_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(c); break;
    case 2: lorenz_convert_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(c); break;
    case 3: lorenz_convert_kernel<3><<<(unsigned)blocks, 256, 0, s>>>(c); break;
    case 4: lorenz_convert_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(c); break;
    case 5: lorenz_convert_kernel<5><<<(unsigned)blocks, 256, 0, s>>>(c); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace pnde
