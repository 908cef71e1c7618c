// Host orchestration + element-wise kernels of the large-D EK1 path (see big_dense.cuh).
// Lorenz-96 only (the one catalogue entry with large d); fixed or adaptive steps; one trajectory at a time.
#include "big_dense.cuh"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "big_dense_api.h"
#include "filter_kernel.cuh"

namespace pnde {
namespace big {

namespace {

struct StepCtx {
  int d, q, D;
  double F;
  double* m;      // [D] mean, natural coordinates
  double* mp;     // [D] predicted mean, P(h) coordinates
  double* Jp;     // [d][4] pi0 * J, columns (i-2, i-1, i, i+1)
  double* fu;     // [d]
  double* z;      // [d]
  double* y;      // [d]
  double* S;      // [D-d][D] posterior factor columns (natural coordinates), row c = column c
  double* E;      // [D][D]
  double* R;      // [D][D]
  Scalars* sc;
  int diffusion;
  IwpConsts C;
};

__device__ __forceinline__ int wrapi(int i, int d) { return i < 0 ? i + d : (i >= d ? i - d : i); }

// per step scalars: pi_k for this h (also resets the per-step accumulators)
__global__ void set_step_kernel(Scalars* sc, double h, int q) {
  double v = sqrt(h);
  for (int k = q; k > 1; --k) v *= h;  // h^(q-1/2) = pi_1
  sc->pi1 = v;
  sc->pi0 = v * h;
  sc->ipi1 = 1.0 / v;
  sc->h = h;
}

// mean predict (P(h) coordinates), u-hat, f, sparse J, z
__global__ void measure_kernel(StepCtx c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.d) return;
  const int d = c.d, q = c.q;
  const double h = c.sc->h, pi0 = c.sc->pi0, pi1 = c.sc->pi1;
  // P_k = h^(k-q-1/2): P_0 = 1/pi0, P_k = P_{k-1} * h
  auto pred0 = [&](int j) {  // block 0 of A P m at dimension j
    double Pk = 1.0 / pi0, acc = 0.0;
    for (int k = 0; k <= q; ++k) {
      acc = fma(inv_factorial(k) * Pk, c.m[k * d + j], acc);
      Pk *= h;
    }
    return acc;
  };
  double Pk = 1.0 / pi0;
  double Pks[QMAX + 1];
  for (int k = 0; k <= q; ++k) {
    Pks[k] = Pk;
    Pk *= h;
  }
  for (int k = 0; k <= q; ++k) {
    double acc = 0.0;
    for (int kk = k; kk <= q; ++kk) acc = fma(inv_factorial(kk - k) * Pks[kk], c.m[kk * d + i], acc);
    c.mp[k * d + i] = acc;
  }
  const double up = pi0 * pred0(wrapi(i + 1, d)), um1 = pi0 * pred0(wrapi(i - 1, d)), um2 = pi0 * pred0(wrapi(i - 2, d));
  const double ui = pi0 * c.mp[i];
  const double f = (up - um2) * um1 - ui + c.F;
  c.fu[i] = f;
  c.z[i] = fma(pi1, c.mp[d + i], -f);
  // J row i: d f_i / d u_{i-2} = -u_{i-1}; / d u_{i-1} = u_{i+1} - u_{i-2}; / d u_i = -1; / d u_{i+1} = u_{i-1}
  c.Jp[i * 4 + 0] = pi0 * (-um1);
  c.Jp[i * 4 + 1] = pi0 * (up - um2);
  c.Jp[i * 4 + 2] = pi0 * (-1.0);
  c.Jp[i * 4 + 3] = pi0 * um1;
}

// Jp[b][a] for the sparse Lorenz-96 Jacobian (0 outside the 4 stored columns)
__device__ __forceinline__ double jp_entry(const double* Jp, int b, int a, int d) {
  int off = a - b;
  if (off > d / 2) off -= d;
  if (off < -d / 2) off += d;
  if (off < -2 || off > 1) return 0.0;
  return Jp[b * 4 + off + 2];
}

// Cmat (2d x d) with B = H Q H' = Cmat' Cmat (src/diffusions.jl:77): rows (0,a): pi1 L10 delta - L00 Jp[b][a];
// rows (1,a): pi1 L11 delta.  Written into E (ld = D).
__global__ void build_c_kernel(StepCtx c) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = c.d;
  if (idx >= 2LL * d * d) return;
  const int row = (int)(idx / d), b = (int)(idx % d);
  const int blk = row / d, a = row % d;
  const double pi1 = c.sc->pi1;
  double v;
  if (blk == 0)
    v = ((a == b) ? pi1 * c.C.Lt[1][0] : 0.0) - c.C.Lt[0][0] * jp_entry(c.Jp, b, a, d);
  else
    v = (a == b) ? pi1 * c.C.Lt[1][1] : 0.0;
  c.E[(size_t)row * c.D + b] = v;
}

// Solve R(0:n,0:n)' y = z (R upper triangular): blocked forward substitution in one CTA.
__global__ void __launch_bounds__(1024) trsv_t_kernel(const double* R, int ld, int n, const double* z, double* y,
                                                      double* quad_out, double* logdet_out) {
  extern __shared__ double sy[];  // [n] running right-hand side / solution
  __shared__ double sd[32][33];   // diagonal block
  for (int i = threadIdx.x; i < n; i += blockDim.x) sy[i] = z[i];
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 32) {
    const int bn = min(32, n - b0);
    {
      const int r = threadIdx.x / 32, cc = threadIdx.x % 32;
      if (r < bn && cc < bn) sd[r][cc] = R[(size_t)(b0 + r) * ld + b0 + cc];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // warp-sequential solve of the diagonal block (from shared memory)
      for (int k = 0; k < bn; ++k) {
        const double rkk = sd[k][k];
        const double yk = (rkk != 0.0) ? sy[b0 + k] / rkk : 0.0;
        __syncwarp();
        if (threadIdx.x == 0) sy[b0 + k] = yk;
        const int i = k + 1 + threadIdx.x;
        if (i < bn) sy[b0 + i] -= sd[k][i] * yk;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = b0 + bn + threadIdx.x; i < n; i += blockDim.x) {
      double acc = sy[i];
      for (int k = 0; k < bn; ++k) acc = fma(-R[(size_t)(b0 + k) * ld + i], sy[b0 + k], acc);
      sy[i] = acc;
    }
    __syncthreads();
  }
  __shared__ double red[64];
  double q2 = 0.0, ld2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    y[i] = sy[i];
    q2 = fma(sy[i], sy[i], q2);
    ld2 += log(fabs(R[(size_t)i * ld + i]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    q2 += __shfl_xor_sync(0xffffffffu, q2, o);
    ld2 += __shfl_xor_sync(0xffffffffu, ld2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = q2;
    red[32 + (threadIdx.x >> 5)] = ld2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += red[w];
      b += red[32 + w];
    }
    *quad_out = a;
    *logdet_out = b;
  }
}

// sigma from the diffusion solve: sigma^2 = quad / d (dynamic); static models predict with sigma = 1
__global__ void set_sigma_kernel(Scalars* sc, int d, int dynamic) {
  if (dynamic) {
    sc->local = sc->quad / double(d);
    sc->sigma = sqrt(sc->local);
  } else {
    sc->sigma = 1.0;
  }
}

// E rows 0..d-1: the prior rows of block 0; rows d..D-1: (T A P s)' for every factor column s
__global__ void build_e_kernel(StepCtx c) {
  const int d = c.d, q = c.q, D = c.D;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)D * d) return;
  const int row = (int)(idx / d), b = (int)(idx % d);
  const double sigma = c.sc->sigma, pi1 = c.sc->pi1, pi0 = c.sc->pi0, h = c.sc->h;
  double* e = c.E + (size_t)row * D;
  if (row < d) {
    const int a = row;
    e[b] = sigma * (((a == b) ? pi1 * c.C.Lt[1][0] : 0.0) - c.C.Lt[0][0] * jp_entry(c.Jp, b, a, d));
    e[d + b] = (a == b) ? sigma * c.C.Lt[0][0] : 0.0;
    for (int k = 2; k <= q; ++k) e[k * d + b] = (a == b) ? sigma * c.C.Lt[k][0] : 0.0;
    return;
  }
  const double* s = c.S + (size_t)(row - d) * D;
  double Pks[QMAX + 1];
  double Pk = 1.0 / pi0;
  for (int k = 0; k <= q; ++k) {
    Pks[k] = Pk;
    Pk *= h;
  }
  auto w_blk = [&](int k, int j) {  // block k of A P s at dimension j
    double acc = 0.0;
    for (int kk = k; kk <= q; ++kk) acc = fma(inv_factorial(kk - k) * Pks[kk], s[kk * d + j], acc);
    return acc;
  };
  double yv = pi1 * w_blk(1, b);
  for (int off = -2; off <= 1; ++off) yv = fma(-c.Jp[b * 4 + off + 2], w_blk(0, wrapi(b + off, d)), yv);
  e[b] = yv;
  e[d + b] = w_blk(0, b);
  for (int k = 2; k <= q; ++k) e[k * d + b] = w_blk(k, b);
}

// delta[j] = sum_a R[a][j] y[a] for j in [d, D): rows split over blockIdx.y, FP64 atomics
__global__ void gain_gemv_kernel(StepCtx c, double* delta) {
  const int d = c.d, D = c.D;
  const int j = d + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const int a0 = blockIdx.y * 64, a1 = min(d, a0 + 64);
  double acc = 0.0;
  for (int a = a0; a < a1; ++a) acc = fma(c.R[(size_t)a * D + j], c.y[a], acc);
  atomicAdd(&delta[j], acc);
}

// mean update in primed coordinates + back to natural coordinates (delta from gain_gemv_kernel)
__global__ void mean_update2_kernel(StepCtx c, const double* delta) {
  const int d = c.d, q = c.q;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d) return;
  const double pi0 = c.sc->pi0, h = c.sc->h, ipi1 = c.sc->ipi1;
  double PIk = pi0;
  c.m[b] = (c.mp[b] - delta[d + b]) * PIk;
  double m1 = c.fu[b];
  for (int off = -2; off <= 1; ++off) m1 = fma(c.Jp[b * 4 + off + 2], -delta[d + wrapi(b + off, d)], m1);
  PIk /= h;
  c.m[d + b] = m1 * ipi1 * PIk;
  for (int k = 2; k <= q; ++k) {
    PIk /= h;
    c.m[k * d + b] = (c.mp[k * d + b] - delta[k * d + b]) * PIk;
  }
  if (!(fabs(c.m[b]) <= 1.79769313486231570e308)) c.sc->nonfinite = 1;
}

// posterior factor columns in natural coordinates from rows d.. of R (upper triangular: entries left of
// the diagonal are stale storage and read as zero)
__global__ void build_s_kernel(StepCtx c) {
  const int d = c.d, q = c.q, D = c.D;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(D - d) * d) return;
  const int cc = (int)(idx / d), b = (int)(idx % d);
  const int row = d + cc;
  const double pi0 = c.sc->pi0, h = c.sc->h, ipi1 = c.sc->ipi1;
  const double* r = c.R + (size_t)row * D;
  auto x0 = [&](int j) { return (d + j >= row) ? r[d + j] : 0.0; };
  double* s = c.S + (size_t)cc * D;
  double PIk = pi0;
  s[b] = x0(b) * PIk;
  double x1 = 0.0;
  for (int off = -2; off <= 1; ++off) x1 = fma(c.Jp[b * 4 + off + 2], x0(wrapi(b + off, d)), x1);
  PIk /= h;
  s[d + b] = x1 * ipi1 * PIk;
  for (int k = 2; k <= q; ++k) {
    PIk /= h;
    s[k * d + b] = ((k * d + b >= row) ? r[k * d + b] : 0.0) * PIk;
  }
}

// end of step: diffusion bookkeeping (src/diffusions.jl) and log-likelihood accumulation
__global__ void finish_step_kernel(Scalars* sc, int d, int diffusion) {
  if (diffusion != DIFF_DYNAMIC) sc->local = sc->quad / double(d);
  const int nacc = sc->nacc;
  double g;
  if (diffusion == DIFF_DYNAMIC) {
    g = sc->local;
  } else if (diffusion == DIFF_FIXED) {
    g = (nacc == 0) ? sc->local : sc->global_saved + (sc->local - sc->global_saved) / double(nacc);
  } else {
    const double Nn = double(nacc + 1), al = 0.5, be = 0.5;
    g = (nacc == 0) ? (be + 0.5 * sc->local) / (al + Nn * d / 2.0 + 1.0)
                    : (be + 0.5 * ((sc->global_saved * (al + (Nn - 1.0) * d / 2.0 + 1.0) - be) * 2.0 + sc->local)) /
                          (al + Nn * d / 2.0 + 1.0);
  }
  sc->global_saved = g;
  sc->nacc = nacc + 1;
  sc->ll_quad += sc->quad;
  sc->ll_logdet += sc->logdet;
  sc->ll_n += 1;
}

// Local error estimate and EEst of one attempted step (src/perform_step.jl:78-84,148-158), single CTA:
//   err_i = sqrt(sigma^2_loc (H Q H')_ii),  r_i = dt err_i / (abstol + max(|uprev_i|, |u_i|) reltol),  EEst = rms(r).
// integ.u is overwritten with the new solution whether or not the step is accepted (:86).
__global__ void __launch_bounds__(1024) eest_kernel(StepCtx c, double* uprev, double dt, double abstol, double reltol,
                                                    int dynamic, double* out) {
  __shared__ double red[32];
  const int d = c.d;
  const double pi1 = c.sc->pi1;
  const double local = dynamic ? c.sc->local : c.sc->quad / double(d);
  double acc = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    double jj = 0.0;
    for (int o = 0; o < 4; ++o) jj = fma(c.Jp[i * 4 + o], c.Jp[i * 4 + o], jj);
    const double Bii = fma(pi1 * pi1, c.C.Qt[1][1], fma(-2.0 * pi1 * c.C.Qt[0][1], c.Jp[i * 4 + 2], c.C.Qt[0][0] * jj));
    const double un = c.m[i];
    const double r = dt * sqrt(local * Bii) / (abstol + fmax(fabs(uprev[i]), fabs(un)) * reltol);
    acc = fma(r, r, acc);
    uprev[i] = un;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    out[0] = sqrt(tot / double(d));
  }
}

// ode_determine_initdt (Hairer) for Lorenz-96, single CTA; m holds the initial jets (m[d + i] = f(u0)_i)
__global__ void __launch_bounds__(1024) initdt_kernel(StepCtx c, double abstol, double reltol, double dtmax, double* scratch,
                                                      double* out) {
  __shared__ double red[3][32];
  const int d = c.d, q = c.q;
  auto bsum = [&](double v, int slot) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[slot][threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[slot][w];
    return tot;
  };
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const double sk = abstol + fabs(c.m[i]) * reltol;
    const double a = c.m[i] / sk, b = c.m[d + i] / sk;
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  const double d0 = sqrt(bsum(s0, 0) / d), d1 = sqrt(bsum(s1, 1) / d);
  double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
  dt0 = fmin(dt0, dtmax);
  for (int i = threadIdx.x; i < d; i += blockDim.x) scratch[i] = fma(dt0, c.m[d + i], c.m[i]);
  __syncthreads();
  double s2 = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const double f1 = (scratch[wrapi(i + 1, d)] - scratch[wrapi(i - 2, d)]) * scratch[wrapi(i - 1, d)] - scratch[i] + c.F;
    const double sk = abstol + fabs(c.m[i]) * reltol;
    const double cc = (f1 - c.m[d + i]) / sk;
    s2 = fma(cc, cc, s2);
  }
  const double d2 = sqrt(bsum(s2, 2) / d) / dt0;
  if (threadIdx.x == 0) {
    const double mx = fmax(d1, d2);
    const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / double(q + 1));
    out[0] = (dt0 < 10.0 * 2.220446049250313e-16) ? 1e-6 : fmin(fmin(100.0 * dt0, dt1), dtmax);
  }
}

__global__ void init_mean_kernel(StepCtx c, const double* u0, long long n, long long tr, double* jets /* [(q+1)][d] */) {
  // Taylor-mode jets of Lorenz-96 (same recursion as lorenz96_kernel.cuh); single CTA
  const int d = c.d, q = c.q;
  for (int i = threadIdx.x; i < d; i += blockDim.x) jets[i] = u0[(long long)i * n + tr];
  __syncthreads();
  for (int k = 0; k < q; ++k) {
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      double acc = 0.0;
      for (int a = 0; a <= k; ++a)
        acc = fma(jets[a * d + wrapi(i + 1, d)] - jets[a * d + wrapi(i - 2, d)], jets[(k - a) * d + wrapi(i - 1, d)], acc);
      acc -= jets[k * d + i];
      if (k == 0) acc += c.F;
      jets[(k + 1) * d + i] = acc / double(k + 1);
    }
    __syncthreads();
  }
  double fact = 1.0;
  for (int k = 0; k <= q; ++k) {
    if (k > 0) fact *= double(k);
    for (int i = threadIdx.x; i < d; i += blockDim.x) c.m[k * d + i] = fact * jets[k * d + i];
  }
}

// packed lower triangle of cal * S' S (Sigma = sum over factor columns) from the full D x D product
__global__ void pack_cov_kernel(const double* full, int D, double cal, double* out, long long n, long long tr) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)D * D) return;
  const int i = (int)(idx / D), j = (int)(idx % D);
  if (j > i) return;
  out[((long long)i * (i + 1) / 2 + j) * n + tr] = cal * full[(size_t)i * D + j];
}

__global__ void write_outputs_kernel(StepCtx c, double* mean, double* t_final, double* loglik, double* final_diff,
                                     int* retcode, int* naccept, int* nreject, int* nf, int* njacs, int* n_saved,
                                     long long n, long long tr, double t, int is_static, int ret_host, int nrej,
                                     int nfe, int nsaved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.D) mean[(long long)i * n + tr] = c.m[i];
  if (i == 0) {
    const Scalars* sc = c.sc;
    double ll = -0.5 * (sc->ll_quad + 2.0 * sc->ll_logdet + double(sc->ll_n) * double(c.d) * 1.8378770664093453);
    if (is_static && sc->nacc > 0) ll = nan("");
    t_final[tr] = t;
    loglik[tr] = ll;
    final_diff[tr] = sc->global_saved;
    // a loop cut short by maxiters is reported as such (filter_kernel / lorenz96_kernel do the same)
    retcode[tr] = sc->nonfinite ? RET_NONFINITE : ret_host;
    naccept[tr] = sc->nacc;
    nreject[tr] = nrej;
    nf[tr] = nfe;
    njacs[tr] = sc->nacc + nrej;
    n_saved[tr] = nsaved;
  }
}

// savevalues! (src/integrator_utils.jl:33-48) for this path: the solution marginals of one accepted state.
//   u = E0 m,  diag(Sigma_u)_b = sum over the factor columns of S[.][b]^2   (Sigma = S' S, natural coordinates)
__global__ void __launch_bounds__(256) save_marginals_kernel(StepCtx c, double* base, long long n, double t, int initial) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = c.d, D = c.D;
  if (b == 0) {
    base[0] = t;
    base[n] = initial ? 1.0 : c.sc->global_saved;  // initial_diffusion, src/diffusions.jl:8
  }
  if (b >= d) return;
  double var = 0.0;
  if (!initial)
    for (int r = 0; r < D - d; ++r) {
      const double v = c.S[(size_t)r * D + b];
      var = fma(v, v, var);
    }
  base[(long long)(2 + b) * n] = c.m[b];
  base[(long long)(2 + d + b) * n] = var;
}

double ulp_host(double x) {
  x = fabs(x);
  if (x == 0.0) return 4.9406564584124654e-324;
  return nextafter(x, INFINITY) - x;
}

}  // namespace

size_t big_work_bytes(int d, int q, bool adaptive) {
  const size_t D = (size_t)d * (q + 1);
  // m, mp, Jp, fu, z, y, jets | S | E | R | W | v0, T | scalars  (+ adaptive: second S, pre-step m, uprev, EEst)
  return (D * 2 + (size_t)d * 4 + (size_t)d * 3 + D) * 8 + (D - d) * D * 8 + D * D * 8 * 2 + (size_t)NB * D * 8 +
         2 * (NB + NB * NB) * 8 + sizeof(Scalars) + 8192 + (adaptive ? (D - d) * D * 8 + (D + d + 64) * 8 + 2048 : 0);
}

#define BCK(call)                      \
  do {                                 \
    cudaError_t e__ = (call);          \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

// RAII owner of the auxiliary high-priority stream and its events: released on every exit path of big_run
struct AuxGuard {
  QrWork& wk;
  explicit AuxGuard(QrWork& w) : wk(w) {
    wk.aux = nullptr;
    for (int i = 0; i < 2; ++i) wk.ev_narrow[i] = wk.ev_panel[i] = nullptr;
  }
  ~AuxGuard() {
    for (int i = 0; i < 2; ++i) {
      if (wk.ev_narrow[i]) cudaEventDestroy(wk.ev_narrow[i]);
      if (wk.ev_panel[i]) cudaEventDestroy(wk.ev_panel[i]);
    }
    if (wk.aux) cudaStreamDestroy(wk.aux);
  }
};

cudaError_t big_run(const BigRunArgs& A, cudaStream_t s, long long* launches) {
  const int d = A.d, q = A.q, D = d * (q + 1);
  char* base = reinterpret_cast<char*>(A.work);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base + off;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  StepCtx c;
  memset(&c, 0, sizeof(c));
  c.d = d;
  c.q = q;
  c.D = D;
  c.m = (double*)take((size_t)D * 8);
  c.mp = (double*)take((size_t)D * 8);
  c.Jp = (double*)take((size_t)d * 4 * 8);
  c.fu = (double*)take((size_t)d * 8);
  c.z = (double*)take((size_t)d * 8);
  c.y = (double*)take((size_t)d * 8);
  double* jets = (double*)take((size_t)D * 8);
  c.S = (double*)take((size_t)(D - d) * D * 8);
  c.E = (double*)take((size_t)D * D * 8);
  c.R = (double*)take((size_t)D * D * 8);
  QrWork wk;
  AuxGuard guard(wk);
  wk.W = (double*)take((size_t)NB * D * 8);
  wk.v0 = (double*)take(2 * NB * 8);
  wk.T = (double*)take(2 * NB * NB * 8);
  {
    // the panel chain is the critical path of the blocked QR: its clusters must not queue behind the CTAs of the
    // trailing update that runs beside it
    int prio_lo = 0, prio_hi = 0;
    BCK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    BCK(cudaStreamCreateWithPriority(&wk.aux, cudaStreamNonBlocking, prio_hi));
  }
  for (int i = 0; i < 2; ++i) {
    BCK(cudaEventCreateWithFlags(&wk.ev_narrow[i], cudaEventDisableTiming));
    BCK(cudaEventCreateWithFlags(&wk.ev_panel[i], cudaEventDisableTiming));
  }
  c.sc = (Scalars*)take(sizeof(Scalars));
  // adaptive: a rejected step must leave (m, S) untouched -- the new factor goes to a second buffer that is swapped in
  // on commit, the pre-step mean is kept, EEst comes back to the host controller (8 bytes per attempted step)
  double *S2 = nullptr, *m_old = nullptr, *uprev = nullptr, *d_eest = nullptr;
  if (A.adaptive) {
    S2 = (double*)take((size_t)(D - d) * D * 8);
    m_old = (double*)take((size_t)D * 8);
    uprev = (double*)take((size_t)d * 8);
    d_eest = (double*)take(64);
  }
  c.diffusion = A.diffusion;
  c.C = A.C;
  const bool dynamic = (A.diffusion == DIFF_DYNAMIC);
  const bool is_static = !dynamic;
  const int TB = 256;
  for (long long tr = 0; tr < A.n; ++tr) {
    double Fh = 0.0;
    BCK(cudaMemcpyAsync(&Fh, A.p + tr, 8, cudaMemcpyDeviceToHost, s));
    BCK(cudaStreamSynchronize(s));
    c.F = Fh;
    BCK(cudaMemsetAsync(c.sc, 0, sizeof(Scalars), s));
    BCK(cudaMemsetAsync(c.S, 0, (size_t)(D - d) * D * 8, s));  // Sigma_0 = 0 exactly
    init_mean_kernel<<<1, 1024, 0, s>>>(c, A.u0, A.n, tr, jets);
    ++*launches;
    double t = A.K.t0;
    long long iter = 0;
    int nsaved = 0, nacc_host = 0;
    bool hist_full = false;
    const int REC = 2 + 2 * d;
    auto save = [&](double tt, int initial) {
      if (!A.hist) return;
      if (nsaved >= A.max_saved) {
        hist_full = true;
        return;
      }
      save_marginals_kernel<<<(d + 255) / 256, 256, 0, s>>>(c, A.hist + ((long long)nsaved * REC) * A.n + tr, A.n, tt, initial);
      ++*launches;
      ++nsaved;
    };
    auto want_save = [&](double tt) {
      return A.save_mode == SAVE_EVERY || (A.save_mode == SAVE_STRIDE && (nacc_host % A.save_stride == 0 || !(tt < A.K.t1)));
    };
    save(t, 1);
    const int trsv_smem = d * 8;
    BCK(cudaFuncSetAttribute(trsv_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsv_smem));
    // one attempted step (fixed h): every launch below depends on the step only through h
    auto step_body = [&](double h) -> cudaError_t {
      set_step_kernel<<<1, 1, 0, s>>>(c.sc, h, q);
      measure_kernel<<<(d + TB - 1) / TB, TB, 0, s>>>(c);
      *launches += 2;
      if (dynamic) {
        // sigma^2 = z' (H Q H')^-1 z / d via the QR of the 2d x d factor of H Q H'
        build_c_kernel<<<(unsigned)((2LL * d * d + TB - 1) / TB), TB, 0, s>>>(c);
        ++*launches;
        BCK(blocked_qr(c.E, c.R, D, 2 * d, d, d, 0, c.sc, c.C, wk, s, launches));
        trsv_t_kernel<<<1, 1024, trsv_smem, s>>>(c.R, D, d, c.z, c.y, &c.sc->quad, &c.sc->logdet);
        ++*launches;
      }
      set_sigma_kernel<<<1, 1, 0, s>>>(c.sc, d, dynamic ? 1 : 0);
      build_e_kernel<<<(unsigned)(((long long)D * d + TB - 1) / TB), TB, 0, s>>>(c);
      *launches += 2;
      BCK(blocked_qr(c.E, c.R, D, D, D, d, 1, c.sc, c.C, wk, s, launches));
      trsv_t_kernel<<<1, 1024, trsv_smem, s>>>(c.R, D, d, c.z, c.y, &c.sc->quad, &c.sc->logdet);
      BCK(cudaMemsetAsync(jets, 0, (size_t)D * 8, s));  // jets is free after the initialisation: reuse as delta
      gain_gemv_kernel<<<dim3((D - d + TB - 1) / TB, (d + 63) / 64), TB, 0, s>>>(c, jets);
      mean_update2_kernel<<<(d + TB - 1) / TB, TB, 0, s>>>(c, jets);
      ++*launches;
      if (A.adaptive) {
        StepCtx c2 = c;  // same R, Jp, scalars; the posterior factor is written next to the current one
        c2.S = S2;
        build_s_kernel<<<(unsigned)(((long long)(D - d) * d + TB - 1) / TB), TB, 0, s>>>(c2);
        *launches += 3;
      } else {
        build_s_kernel<<<(unsigned)(((long long)(D - d) * d + TB - 1) / TB), TB, 0, s>>>(c);
        finish_step_kernel<<<1, 1, 0, s>>>(c.sc, d, A.diffusion);
        *launches += 4;
      }
      return cudaGetLastError();
    };
    // (capturing the ~650 launches of a step into a CUDA graph was measured: 23.5 ms/step either way -- the panel
    // chain is bound by the kernels themselves, not by launch gaps -- so the plain launches stay)
    int ret_host = RET_SUCCESS, nrej = 0, nfe = 0;
    if (!A.adaptive) {
      while (t < A.K.t1) {
        if (++iter > A.K.maxiters) {
          ret_host = RET_MAXITERS;
          break;
        }
        const double h = fmin(A.K.dt, A.K.t1 - t);
        BCK(step_body(h));
        ++nfe;
        const double ttmp = t + h;
        t = (fabs(ttmp - A.K.t1) < 10.0 * ulp_host(fmax(t, A.K.t1))) ? A.K.t1 : ttmp;
        ++nacc_host;
        if (want_save(t)) save(t, 0);
        if (hist_full) {  // fixed steps: the capacity follows from the grid; a full history is a caller error
          ret_host = RET_HISTORY_FULL;
          break;
        }
      }
    } else {
      // OrdinaryDiffEq's loop (loopheader!, perform_step!, loopfooter!, PI controller) on the host:
      // a step is ~18 ms of device work at D = 4096, the 8-byte read-back of EEst per attempt is free beside it
      const CtrlParams& K = A.K;
      double dt = K.dt;
      BCK(cudaMemcpyAsync(uprev, c.m, (size_t)d * 8, cudaMemcpyDeviceToDevice, s));  // integ.u = u0
      if (!(dt > 0.0)) {
        initdt_kernel<<<1, 1024, 0, s>>>(c, K.abstol, K.reltol, K.dtmax, jets, d_eest);
        ++*launches;
        BCK(cudaMemcpyAsync(&dt, d_eest, 8, cudaMemcpyDeviceToHost, s));
        BCK(cudaStreamSynchronize(s));
        nfe += 2;
      }
      double dtpropose = dt, qold = K.qoldinit, q11 = 1.0;
      bool accepted_prev = true;
      while (t < K.t1) {
        if (iter > 0) dt = accepted_prev ? dtpropose : dt / fmin(1.0 / K.qmin, q11 / K.gamma);
        if (++iter > K.maxiters) {
          ret_host = RET_MAXITERS;
          break;
        }
        dt = fmin(dt, K.dtmax);
        dt = fmax(dt, K.dtmin);
        dt = fmin(dt, K.t1 - t);
        if (dt != dt) {
          ret_host = RET_DTNAN;
          break;
        }
        if (iter > 1 && !accepted_prev && fabs(dt) <= fabs(K.dtmin)) {
          ret_host = RET_DTMIN;
          break;
        }
        BCK(cudaMemcpyAsync(m_old, c.m, (size_t)D * 8, cudaMemcpyDeviceToDevice, s));
        BCK(step_body(dt));
        ++nfe;
        eest_kernel<<<1, 1024, 0, s>>>(c, uprev, dt, K.abstol, K.reltol, dynamic ? 1 : 0, d_eest);
        ++*launches;
        double EEst = 0.0;
        BCK(cudaMemcpyAsync(&EEst, d_eest, 8, cudaMemcpyDeviceToHost, s));
        BCK(cudaStreamSynchronize(s));
        if (!(EEst == EEst)) {  // non-finite state: OrdinaryDiffEq's unstable check
          ret_host = RET_NONFINITE;
          break;
        }
        const bool commit = EEst < 1.0, accept = EEst <= 1.0;  // src/perform_step.jl:89 / loopfooter!
        if (commit) {
          std::swap(c.S, S2);
          finish_step_kernel<<<1, 1, 0, s>>>(c.sc, d, A.diffusion);
          ++*launches;
        } else {
          BCK(cudaMemcpyAsync(c.m, m_old, (size_t)D * 8, cudaMemcpyDeviceToDevice, s));
        }
        double qc;
        if (EEst == 0.0) {
          qc = 1.0 / K.qmax;
        } else {
          q11 = pow(EEst, K.beta1);
          qc = fmax(1.0 / K.qmax, fmin(1.0 / K.qmin, q11 / pow(qold, K.beta2) / K.gamma));
        }
        const double ttmp = t + dt;
        if (accept) {
          if (K.qsteady_min <= qc && qc <= K.qsteady_max) qc = 1.0;
          qold = fmax(EEst, K.qoldinit);
          t = (fabs(ttmp - K.t1) < 10.0 * ulp_host(fmax(t, K.t1))) ? K.t1 : ttmp;
          dtpropose = fmax(K.dtmin, fmin(K.dtmax, dt / qc));
          ++nacc_host;
          // (EEst == 1 exactly is accepted but not committed, src/perform_step.jl:89: the saved state is then the old one,
          // like in the reference)
          if (want_save(t)) save(t, 0);
          if (hist_full) ret_host = RET_HISTORY_FULL;  // keep stepping without saving: naccept + 1 is the capacity needed
        } else {
          ++nrej;
        }
        accepted_prev = accept;
      }
    }
    write_outputs_kernel<<<(D + TB - 1) / TB, TB, 0, s>>>(c, A.mean, A.t_final, A.loglik, A.final_diff, A.retcode,
                                                          A.naccept, A.nreject, A.nf, A.njacs, A.n_saved, A.n, tr, t,
                                                          is_static ? 1 : 0, ret_host, nrej, nfe, nsaved);
    ++*launches;
    if (A.cov) {
      // Sigma = S' S over the factor columns (rows of S): full D x D by the same DMMA kernel, then packed
      BCK(cudaMemsetAsync(c.E, 0, (size_t)D * D * 8, s));
      const int kchunk = 512;
      dim3 g((D + 63) / 64, D / 32, (D - d + kchunk - 1) / kchunk);
      atb_kernel<<<g, 128, 0, s>>>(c.S, D, c.S, D, c.E, D, D, D - d, kchunk);
      double cal = 1.0;
      if (is_static) {
        Scalars hs;
        BCK(cudaMemcpyAsync(&hs, c.sc, sizeof(hs), cudaMemcpyDeviceToHost, s));
        BCK(cudaStreamSynchronize(s));
        if (hs.nacc > 0) cal = hs.global_saved;
      }
      pack_cov_kernel<<<(unsigned)(((long long)D * D + TB - 1) / TB), TB, 0, s>>>(c.E, D, cal, A.cov, A.n, tr);
      *launches += 2;
    }
  }
  cudaError_t fin = cudaStreamSynchronize(s);  // the auxiliary stream's work is ordered before this point
  if (fin != cudaSuccess) return fin;
  return cudaGetLastError();
}

}  // namespace big
}  // namespace pnde
