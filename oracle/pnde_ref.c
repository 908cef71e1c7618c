/*
 * pnde_ref.c -- C restatement of the REFERENCE's dense ODE-filter algorithm (ProbNumDiffEq v0.1.5).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Used by tests/ (as a fast checker, itself checked
 * against oracle/pnde_oracle.py) and by bench.py's cpu_baseline / --impl reference legs.  The
 * product path never links or calls it.
 *
 * It deliberately keeps the reference's own (wasteful) arithmetic so that the CPU number is a fair
 * stand-in for the Julia package on the same host cores (Julia is not installed in this image):
 *   x = P*x, P^-1 round trips            src/perform_step.jl:36-38,73-75
 *   eager S*S' in every SRMatrix          src/squarerootmatrix.jl:16
 *   predict: chol([A S, sQ_L][.]') , QR fallback   src/filtering.jl:33-48
 *   measure: H = (E1 - J E0) P^-1, S = H Sigma H'  src/perform_step.jl:95-132
 *   dynamic diffusion                     src/diffusions.jl:72-80
 *   update: K = Sigma H' inv(S) (LU), (I-KH) S     src/filtering.jl:79-91
 *   error estimate / EEst                 src/perform_step.jl:78-84,148-158
 *   PI controller + loop                  src/alg_utils.jl:13-24 + OrdinaryDiffEq
 * One trajectory per OpenMP task mirrors EnsembleThreads.  It omits Julia's ~60 heap
 * allocations per step and dynamic dispatch, so it is FASTER than the real reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MAXD 16

typedef struct {
  int alg, order, vf, adaptive;
  double abstol, reltol, dt, t0, t1;
  double qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit, dtmin, dtmax;
  int64_t maxiters;
} ref_config;

/* ---------------------------------------------------------------- vector fields (d = 2) */
static void vf_f(int vf, const double* u, const double* p, double* du) {
  switch (vf) {
    case 0: /* README.md:36-40 */
      du[0] = p[2] * (u[0] - u[0] * u[0] * u[0] / 3.0 + u[1]);
      du[1] = -(1.0 / p[2]) * (u[0] - p[0] - p[1] * u[1]);
      break;
    case 1:
      du[0] = u[0] - u[0] * u[0] * u[0] / 3.0 - u[1] + p[3];
      du[1] = p[2] * (u[0] + p[0] - p[1] * u[1]);
      break;
    case 2:
      du[0] = p[0] * u[0] - p[1] * u[0] * u[1];
      du[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
      break;
    default: /* 3: van der Pol, u = (y, x) */
      du[0] = p[0] * ((1.0 - u[1] * u[1]) * u[0] - u[1]);
      du[1] = u[0];
      break;
  }
}
static void vf_jac(int vf, const double* u, const double* p, double* J /* 2x2 row major */) {
  switch (vf) {
    case 0:
      J[0] = p[2] * (1.0 - u[0] * u[0]); J[1] = p[2]; J[2] = -(1.0 / p[2]); J[3] = p[1] / p[2];
      break;
    case 1:
      J[0] = 1.0 - u[0] * u[0]; J[1] = -1.0; J[2] = p[2]; J[3] = -p[2] * p[1];
      break;
    case 2:
      J[0] = p[0] - p[1] * u[1]; J[1] = -p[1] * u[0]; J[2] = p[3] * u[1]; J[3] = -p[2] + p[3] * u[0];
      break;
    default:
      J[0] = p[0] * (1.0 - u[1] * u[1]); J[1] = p[0] * (-2.0 * u[1] * u[0] - 1.0); J[2] = 1.0; J[3] = 0.0;
      break;
  }
}
static int vf_np(int vf) { return vf == 0 ? 3 : (vf == 3 ? 1 : 4); }

/* Taylor-mode initial derivatives (src/state_initialization.jl:15-42): time-Taylor coefficients with
 * truncated Cauchy products, written out for the polynomial catalogue above. */
static void jet_mul(const double* a, const double* b, double* r, int n) {
  for (int k = 0; k < n; ++k) {
    double s = 0.0;
    for (int i = 0; i <= k; ++i) s += a[i] * b[k - i];
    r[k] = s;
  }
}
static void jet_f(int vf, double x[2][8], const double* p, double f[2][8], int n) {
  double t1[8], t2[8];
  switch (vf) {
    case 0:
      jet_mul(x[0], x[0], t1, n); jet_mul(t1, x[0], t2, n);
      for (int k = 0; k < n; ++k) {
        f[0][k] = p[2] * (x[0][k] - t2[k] / 3.0 + x[1][k]);
        f[1][k] = -(1.0 / p[2]) * (x[0][k] - (k == 0 ? p[0] : 0.0) - p[1] * x[1][k]);
      }
      break;
    case 1:
      jet_mul(x[0], x[0], t1, n); jet_mul(t1, x[0], t2, n);
      for (int k = 0; k < n; ++k) {
        f[0][k] = x[0][k] - t2[k] / 3.0 - x[1][k] + (k == 0 ? p[3] : 0.0);
        f[1][k] = p[2] * (x[0][k] + (k == 0 ? p[0] : 0.0) - p[1] * x[1][k]);
      }
      break;
    case 2:
      jet_mul(x[0], x[1], t1, n);
      for (int k = 0; k < n; ++k) {
        f[0][k] = p[0] * x[0][k] - p[1] * t1[k];
        f[1][k] = -p[2] * x[1][k] + p[3] * t1[k];
      }
      break;
    default:
      jet_mul(x[1], x[1], t1, n);
      for (int k = 0; k < n; ++k) t1[k] = (k == 0 ? 1.0 : 0.0) - t1[k];
      jet_mul(t1, x[0], t2, n);
      for (int k = 0; k < n; ++k) {
        f[0][k] = p[0] * (t2[k] - x[1][k]);
        f[1][k] = x[0][k];
      }
      break;
  }
}

/* ---------------------------------------------------------------- small dense linear algebra */
static void matmul(const double* A, const double* B, double* C, int m, int k, int n) { /* C = A B */
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}
static void matmul_nt(const double* A, const double* B, double* C, int m, int k, int n) { /* C = A B' (B n x k) */
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
      C[i * n + j] = s;
    }
}
static int chol_lower(const double* A, double* L, int n) { /* returns 1 on success */
  memset(L, 0, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 0.0) || !isfinite(s)) return 0;
    double ljj = sqrt(s);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / ljj;
    }
  }
  return 1;
}
static void inv_lu(const double* A, double* Ai, int n) { /* getrf + getri style */
  double M[MAXD * 2 * MAXD];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      M[i * 2 * n + j] = A[i * n + j];
      M[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(M[r * 2 * n + c]) > fabs(M[piv * 2 * n + c])) piv = r;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) {
        double t = M[c * 2 * n + j];
        M[c * 2 * n + j] = M[piv * 2 * n + j];
        M[piv * 2 * n + j] = t;
      }
    double pv = M[c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) M[c * 2 * n + j] /= pv;
    for (int r = 0; r < n; ++r)
      if (r != c) {
        double f = M[r * 2 * n + c];
        if (f != 0.0)
          for (int j = 0; j < 2 * n; ++j) M[r * 2 * n + j] -= f * M[c * 2 * n + j];
      }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Ai[i * n + j] = M[i * 2 * n + n + j];
}
/* R factor (n x n upper) of the m x n matrix A (row major, overwritten) */
static void qr_r(double* A, int m, int n, double* Rout) {
  for (int k = 0; k < n && k < m; ++k) {
    double n2 = 0.0;
    for (int i = k; i < m; ++i) n2 += A[i * n + k] * A[i * n + k];
    if (n2 == 0.0) continue;
    double nrm = sqrt(n2), alpha = A[k * n + k] >= 0 ? -nrm : nrm;
    double v[4 * MAXD];
    for (int i = k; i < m; ++i) v[i - k] = A[i * n + k];
    v[0] -= alpha;
    double vtv = 0.0;
    for (int i = 0; i < m - k; ++i) vtv += v[i] * v[i];
    for (int j = k; j < n; ++j) {
      double dot = 0.0;
      for (int i = k; i < m; ++i) dot += v[i - k] * A[i * n + j];
      double s = 2.0 * dot / vtv;
      for (int i = k; i < m; ++i) A[i * n + j] -= s * v[i - k];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Rout[i * n + j] = (j >= i && i < m) ? A[i * n + j] : 0.0;
}

/* ---------------------------------------------------------------- prior (src/priors.jl:7-59) */
static double factorial(int k) {
  double r = 1.0;
  for (int i = 2; i <= k; ++i) r *= i;
  return r;
}
static void ibm(int d, int q, double* A, double* QL) {
  int D = d * (q + 1);
  double Qb[MAXD * MAXD];
  memset(A, 0, sizeof(double) * D * D);
  memset(Qb, 0, sizeof(Qb));
  for (int i = 0; i < D; ++i) A[i * D + i] = 1.0;
  double val = 1.0;
  for (int i = 1; i <= q; ++i) {
    val /= i;
    for (int j = 0; j < d * (q + 1 - i); ++j) A[j * D + j + d * i] = val;
  }
  for (int c = 0; c <= q; ++c)
    for (int r = c; r <= q; ++r) {
      double v = 1.0 / ((2 * q + 1 - r - c) * factorial(q - r) * factorial(q - c));
      for (int i = 0; i < d; ++i) {
        Qb[(c * d + i) * D + r * d + i] = v;
        Qb[(r * d + i) * D + c * d + i] = v;
      }
    }
  chol_lower(Qb, QL, D);
}

static double eps_of(double x) {
  x = fabs(x);
  if (x == 0.0) return 4.9406564584124654e-324;
  return nextafter(x, INFINITY) - x;
}

/* ---------------------------------------------------------------- one trajectory */
typedef struct {
  double mean[MAXD];
  double cov[MAXD * MAXD];
  double t_final, loglik;
  int64_t naccept, nreject, nf, chol_fail;
  int retcode;
} ref_result;

static void solve_one(const ref_config* c, const double* u0, const double* p, ref_result* out) {
  const int d = 2, q = c->order, D = d * (q + 1);
  double A[MAXD * MAXD], QL[MAXD * MAXD];
  ibm(d, q, A, QL);
  const double beta2 = 2.0 / (5.0 * (q + 1)), beta1 = 7.0 / (10.0 * (q + 1));
  double mu[MAXD], S[MAXD * MAXD]; /* state: mean and D x D factor (Sigma0 = 0 exactly) */
  memset(S, 0, sizeof(S));
  { /* initial_update! */
    double x[2][8], f[2][8];
    memset(x, 0, sizeof(x));
    x[0][0] = u0[0];
    x[1][0] = u0[1];
    for (int k = 0; k < q; ++k) {
      jet_f(c->vf, x, p, f, q + 1);
      x[0][k + 1] = f[0][k] / (k + 1);
      x[1][k + 1] = f[1][k] / (k + 1);
    }
    for (int k = 0; k <= q; ++k) {
      mu[k * d] = factorial(k) * x[0][k];
      mu[k * d + 1] = factorial(k) * x[1][k];
    }
  }
  double t = c->t0, dt, dtmax = c->dtmax > 0 ? c->dtmax : c->t1 - c->t0;
  int64_t nf = 0, nacc = 0, nrej = 0, iter = 0, cf = 0;
  if (c->adaptive && !(c->dt > 0)) { /* initdt, SURVEY App. B.3 */
    double f0[2], f1[2], u1[2], sk[2], d0 = 0, d1 = 0, d2 = 0;
    vf_f(c->vf, u0, p, f0);
    for (int i = 0; i < d; ++i) {
      sk[i] = c->abstol + fabs(u0[i]) * c->reltol;
      d0 += (u0[i] / sk[i]) * (u0[i] / sk[i]);
      d1 += (f0[i] / sk[i]) * (f0[i] / sk[i]);
    }
    d0 = sqrt(d0 / d); d1 = sqrt(d1 / d);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
    dt0 = fmin(dt0, dtmax);
    for (int i = 0; i < d; ++i) u1[i] = u0[i] + dt0 * f0[i];
    vf_f(c->vf, u1, p, f1);
    nf += 2;
    for (int i = 0; i < d; ++i) d2 += ((f1[i] - f0[i]) / sk[i]) * ((f1[i] - f0[i]) / sk[i]);
    d2 = sqrt(d2 / d) / dt0;
    double mx = fmax(d1, d2);
    double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / (q + 1));
    dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
  } else {
    dt = c->dt;
  }
  double dtpropose = dt, qold = c->qoldinit, q11 = 1.0, ll = 0.0;
  int accepted_prev = 1, ret = 0;
  double uprev[2] = {u0[0], u0[1]};
  while (t < c->t1) {
    if (iter > 0) dt = accepted_prev ? dtpropose : dt / fmin(1.0 / c->qmin, q11 / c->gamma);
    if (++iter > c->maxiters) { ret = 1; break; }
    if (c->adaptive) { dt = fmin(dt, dtmax); dt = fmax(dt, c->dtmin); dt = fmin(dt, c->t1 - t); }
    else dt = fmin(c->dt, c->t1 - t);
    if (dt != dt) { ret = 2; break; }
    /* ---- perform_step! ---- */
    double P[MAXD], PI[MAXD];
    { double val = pow(dt, -q - 0.5);
      for (int j = 0; j <= q; ++j) { for (int i = 0; i < d; ++i) P[j * d + i] = val; val *= dt; }
      for (int i = 0; i < D; ++i) PI[i] = 1.0 / P[i]; }
    double xm[MAXD], xS[MAXD * MAXD], xmat[MAXD * MAXD];
    for (int i = 0; i < D; ++i) { xm[i] = P[i] * mu[i]; for (int j = 0; j < D; ++j) xS[i * D + j] = P[i] * S[i * D + j]; }
    matmul_nt(xS, xS, xmat, D, D, D); /* SRMatrix(P*S): eager mat */
    double mp[MAXD];
    matmul(A, xm, mp, D, D, 1);
    double upred[2] = {PI[0] * mp[0], PI[1] * mp[1]};
    double du[2], J[4], z[2], H[2 * MAXD];
    vf_f(c->vf, upred, p, du); ++nf;
    for (int i = 0; i < d; ++i) z[i] = PI[d + i] * mp[d + i] - du[i];
    memset(H, 0, sizeof(H));
    if (c->alg == 1) vf_jac(c->vf, upred, p, J); else memset(J, 0, sizeof(J));
    for (int i = 0; i < d; ++i) {
      H[i * D + d + i] = PI[d + i];
      for (int j = 0; j < d; ++j) H[i * D + j] += -J[i * d + j] * PI[j];
    }
    /* dynamic diffusion: sigma^2 = z' (H Q H')^-1 z / d */
    double HQL[2 * MAXD], HQH[4], HQHi[4];
    matmul(H, QL, HQL, d, D, D);
    matmul_nt(HQL, HQL, HQH, d, D, d);
    inv_lu(HQH, HQHi, d);
    double s2 = 0.0;
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) s2 += z[i] * HQHi[i * d + j] * z[j];
    s2 /= d;
    /* predict_cov!: _L = [A*S, sqrt(s2) Q_L] ; chol(_L _L') else qr(_L') */
    double AS[MAXD * MAXD], Lw[MAXD * 2 * MAXD], prod[MAXD * MAXD], Sp[MAXD * MAXD], Spmat[MAXD * MAXD];
    matmul(A, xS, AS, D, D, D);
    double sg = sqrt(s2);
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { Lw[i * 2 * D + j] = AS[i * D + j]; Lw[i * 2 * D + D + j] = sg * QL[i * D + j]; }
    matmul_nt(Lw, Lw, prod, D, 2 * D, D);
    if (!chol_lower(prod, Sp, D)) {
      ++cf;
      double Lt[2 * MAXD * MAXD], R[MAXD * MAXD];
      for (int i = 0; i < D; ++i) for (int j = 0; j < 2 * D; ++j) Lt[j * D + i] = Lw[i * 2 * D + j];
      qr_r(Lt, 2 * D, D, R);
      for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) Sp[i * D + j] = R[j * D + i];
    }
    matmul_nt(Sp, Sp, Spmat, D, D, D);
    /* S = H Sigma- H' via SRMatrix(H*Sp) */
    double HS[2 * MAXD], Sz[4], Szi[4];
    matmul(H, Sp, HS, d, D, D);
    matmul_nt(HS, HS, Sz, d, D, d);
    /* log-likelihood (src/perform_step.jl:66) */
    double lls = NAN;
    { double Lz[4];
      if (chol_lower(Sz, Lz, d)) {
        double y0 = z[0] / Lz[0], y1 = (z[1] - Lz[2] * y0) / Lz[3];
        lls = -0.5 * (y0 * y0 + y1 * y1 + 2.0 * (log(Lz[0]) + log(Lz[3])) + d * 1.8378770664093453);
      } }
    /* update!: K = Sigma- H' inv(S); mu = m - K z; Sigma = SRMatrix((I-KH) Sp) */
    double PHt[MAXD * 2], K[MAXD * 2], IKH[MAXD * MAXD], Sf[MAXD * MAXD], Sfmat[MAXD * MAXD], mf[MAXD];
    inv_lu(Sz, Szi, d);
    matmul_nt(Spmat, H, PHt, D, D, d);
    matmul(PHt, Szi, K, D, d, d);
    for (int i = 0; i < D; ++i) { mf[i] = mp[i]; for (int j = 0; j < d; ++j) mf[i] -= K[i * d + j] * z[j]; }
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int l = 0; l < d; ++l) s -= K[i * d + l] * H[l * D + j];
      IKH[i * D + j] = s; }
    matmul(IKH, Sp, Sf, D, D, D);
    matmul_nt(Sf, Sf, Sfmat, D, D, D);
    double ufilt[2] = {PI[0] * mf[0], PI[1] * mf[1]};
    /* undo preconditioning of x, x_pred, x_filt (three more eager products in the reference) */
    double tmpS[MAXD * MAXD], tmpm[MAXD * MAXD];
    for (int rep = 0; rep < 2; ++rep) { /* x and x_pred: results unused, cost kept */
      const double* src = rep == 0 ? xS : Sp;
      for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) tmpS[i * D + j] = PI[i] * src[i * D + j];
      matmul_nt(tmpS, tmpS, tmpm, D, D, D);
      if (rep == 0) for (int i = 0; i < D * D; ++i) S[i] = tmpS[i]; /* cache.x := PI*(P*x) */
    }
    for (int i = 0; i < D; ++i) mu[i] = PI[i] * xm[i];
    double fS[MAXD * MAXD], fm[MAXD];
    for (int i = 0; i < D; ++i) { fm[i] = PI[i] * mf[i]; for (int j = 0; j < D; ++j) fS[i * D + j] = PI[i] * Sf[i * D + j]; }
    matmul_nt(fS, fS, tmpm, D, D, D);
    (void)Sfmat; (void)xmat;
    double EEst = 0.0;
    if (c->adaptive) {
      double acc = 0.0;
      for (int i = 0; i < d; ++i) {
        double e = sqrt(s2 * HQH[i * d + i]);
        double r = dt * e / (c->abstol + fmax(fabs(uprev[i]), fabs(ufilt[i])) * c->reltol);
        acc += r * r;
      }
      EEst = sqrt(acc / d);
    }
    uprev[0] = ufilt[0]; uprev[1] = ufilt[1];
    if (!c->adaptive || EEst < 1.0) {
      memcpy(mu, fm, sizeof(double) * D);
      memcpy(S, fS, sizeof(double) * D * D);
      ll += lls;
    }
    if (!isfinite(ufilt[0]) || !isfinite(ufilt[1])) { ret = 3; break; }
    double ttmp = t + dt;
    if (c->adaptive) {
      double qc;
      if (EEst == 0.0) qc = 1.0 / c->qmax;
      else { q11 = pow(EEst, beta1); qc = q11 / pow(qold, beta2); qc = fmax(1.0 / c->qmax, fmin(1.0 / c->qmin, qc / c->gamma)); }
      if (EEst <= 1.0) {
        ++nacc;
        if (c->qsteady_min <= qc && qc <= c->qsteady_max) qc = 1.0;
        qold = fmax(EEst, c->qoldinit);
        double dtnew = dt / qc;
        t = (fabs(ttmp - c->t1) < 10.0 * eps_of(fmax(t, c->t1))) ? c->t1 : ttmp;
        dtpropose = fmax(c->dtmin, fmin(dtmax, dtnew));
        accepted_prev = 1;
      } else { ++nrej; accepted_prev = 0; }
    } else {
      ++nacc;
      t = (fabs(ttmp - c->t1) < 10.0 * eps_of(fmax(t, c->t1))) ? c->t1 : ttmp;
      dtpropose = dt; accepted_prev = 1;
    }
  }
  memcpy(out->mean, mu, sizeof(double) * D);
  matmul_nt(S, S, out->cov, D, D, D);
  out->t_final = t; out->loglik = ll;
  out->naccept = nacc; out->nreject = nrej; out->nf = nf; out->chol_fail = cf; out->retcode = ret;
}

/* Ensemble entry (EnsembleThreads stand-in): trajectories are handed out in chunks of 16 to
 * `nthreads` POSIX threads (libgomp is not installed in this image, hence pthreads).
 * u0: [2][n], p: [np][n] (trajectory index fastest).
 * Outputs: mean [D][n], cov [D*D][n] (full), counts [4][n] = naccept, nreject, nf, chol_fail. */
typedef struct {
  const ref_config* c;
  int64_t n;
  const double *u0, *p;
  double *mean, *cov, *t_final, *loglik;
  int64_t* counts;
  int32_t* retcode;
  volatile int64_t* next;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  const ref_config* c = j->c;
  const int64_t n = j->n;
  const int D = 2 * (c->order + 1), np = vf_np(c->vf);
  for (;;) {
    int64_t lo = __sync_fetch_and_add(j->next, 16);
    if (lo >= n) break;
    int64_t hi = lo + 16 < n ? lo + 16 : n;
    for (int64_t i = lo; i < hi; ++i) {
      double u[2] = {j->u0[i], j->u0[n + i]}, pp[4];
      for (int k = 0; k < np; ++k) pp[k] = j->p[(int64_t)k * n + i];
      ref_result r;
      memset(&r, 0, sizeof(r));
      solve_one(c, u, pp, &r);
      if (j->mean) for (int k = 0; k < D; ++k) j->mean[(int64_t)k * n + i] = r.mean[k];
      if (j->cov) for (int k = 0; k < D * D; ++k) j->cov[(int64_t)k * n + i] = r.cov[k];
      if (j->t_final) j->t_final[i] = r.t_final;
      if (j->loglik) j->loglik[i] = r.loglik;
      if (j->counts) { j->counts[i] = r.naccept; j->counts[n + i] = r.nreject; j->counts[2 * n + i] = r.nf; j->counts[3 * n + i] = r.chol_fail; }
      if (j->retcode) j->retcode[i] = r.retcode;
    }
  }
  return NULL;
}

int pnde_ref_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int pnde_ref_solve_ensemble(const ref_config* c, int64_t n, const double* u0, const double* p, double* mean,
                            double* cov, double* t_final, double* loglik, int64_t* counts, int32_t* retcode,
                            int32_t nthreads) {
  if (c->order < 1 || (c->order + 1) * 2 > MAXD) return -1;
  if (nthreads <= 0) nthreads = pnde_ref_max_threads();
  if (nthreads > 256) nthreads = 256;
  volatile int64_t next = 0;
  job_t job = {c, n, u0, p, mean, cov, t_final, loglik, counts, retcode, &next};
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; ++t)
    if (pthread_create(&th[started], NULL, worker, &job) == 0) ++started;
  worker(&job);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  return 0;
}
