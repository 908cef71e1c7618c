// Square-root covariance engine of the filter step: ONE structured Householder triangularisation
// per attempted step, held entirely in registers (everything below is unrolled at compile time).
//
// Reference arithmetic being replaced (file:line relative to the reference checkout):
//   predict_cov!  src/filtering.jl:33-48      chol([A S, sig Q_L][.]') with QR fallback
//   S = H Sigma- H'  src/perform_step.jl:54
//   update!       src/filtering.jl:79-91      K = Sigma- H' inv(S);  Sigma = (I-KH) S-
// and the triangularize!-style QR that BASELINE.json's north_star asks for.
//
// Design.  The measurement is noise free (R = 0, src/filtering.jl:81), so the
// posterior covariance has rank D-dc and is carried as a D x (D-dc) factor S = [W | Lz]:
//   W  : dc dense columns,
//   Lz : D-2dc columns that are zero in blocks 0,1 and lower triangular below.
// The stacked matrix [sig Q_L' ; (A S)'] is triangularised in MEASUREMENT-ALIGNED coordinates
//   x' = (y, x_0, x_2, ..., x_q),   y = H x = pi1 x_1 - Jp x_0,   Jp = pi0 J,
// in which conditioning on y is "drop the first dc columns of the triangular factor": the first dc
// rows of R hold the innovation factor (S_z = R00' R00) and the gain, the remaining rows ARE the
// posterior factor.  sig Q_L' is already triangular in these coordinates up to a dc x dc block, so
// every Householder vector has length <= D+1.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

// Every loop over matrix entries is written `PNDE_UNROLL for (...)`.  Ahead-of-time and small run-time compiled models
// unroll them completely (all entries become named registers: the design of these kernels).  PNDE_ROLLED keeps them as
// loops over arrays in local memory: the general-(d, q) fallback of rtc_model.cu for user ODEs whose state is too large
// to unroll (D > 16 with EK1, D > 64 with EK0) -- every kernel then works for any dimension, slowly.  Rolled means "no
// pragma", not "unroll 1": the compiler still unrolls what its own heuristics find worthwhile, and -- the reason this
// is not `unroll 1` -- NVVM 12.9 miscompiled that variant (a local array filled by a loop marked nounroll and read at
// constant indices after a large inlined call had two of its elements replaced by undef; seen as NaN constants in the
// PTX of filter_kernel, and as a wrong initial step size on the device).
#ifdef PNDE_ROLLED
#define PNDE_UNROLL
#else
#define PNDE_UNROLL _Pragma("unroll")
#endif

namespace pnde {

constexpr int QMAX = 7;

struct IwpConsts {        // src/priors.jl:7-59 with d = 1: Ltilde = chol(Qtilde), Qtilde
  double Lt[QMAX + 1][QMAX + 1];
  double Qt[QMAX + 1][QMAX + 1];
};

__host__ __device__ __forceinline__ constexpr double inv_factorial(int k) {
  return k <= 1 ? 1.0
       : k == 2 ? 0.5
       : k == 3 ? 1.0 / 6.0
       : k == 4 ? 1.0 / 24.0
       : k == 5 ? 1.0 / 120.0
       : k == 6 ? 1.0 / 720.0
                : 1.0 / 5040.0;
}

// rsqrt / reciprocal with one cubic correction step on top of the 20-bit hardware approximation
// (MUFU.RSQ64H / MUFU.RCP64H): ~1 ulp, no special-case slow path.  x must be a positive normal
// number (callers guard zero).
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);           // 1 - x y^2
  const double q = e * fma(0.375, e, 0.5);    // e/2 + 3e^2/8
  return fma(y, q, y);
}
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);  // 1 - x y
  const double q = fma(e, e, e);     // e + e^2
  return fma(y, q, y);
}

// x <- A x with A = Atilde (x) I_dc, Atilde[i][j] = 1/(j-i)!  (src/priors.jl:15-27).
template <int dc, int q>
__device__ __forceinline__ void apply_A(double (&x)[dc * (q + 1)]) {
PNDE_UNROLL
  for (int k = 0; k <= q; ++k) {
PNDE_UNROLL
    for (int a = 0; a < dc; ++a) {
      double acc = x[k * dc + a];
PNDE_UNROLL
      for (int j = k + 1; j <= q; ++j)
        acc = (j - k == 1) ? acc + x[j * dc + a] : fma(inv_factorial(j - k), x[j * dc + a], acc);
      x[k * dc + a] = acc;
    }
  }
}

template <int dc, int q>
struct Factor {
  static constexpr int D = dc * (q + 1);
  static constexpr int R = D - dc;       // number of columns
  static constexpr int NZ = D - 2 * dc;  // triangular columns
  static constexpr int NLZ = NZ > 0 ? NZ * (NZ + 1) / 2 : 1;
  static constexpr int LEN = dc * D + (NZ > 0 ? NZ * (NZ + 1) / 2 : 0);  // doubles in a record
  double W[dc][D];
  double Lz[NLZ];
  // column j (0-based among the triangular columns), row i >= j (0-based below block 1)
  __host__ __device__ static constexpr int lz(int j, int i) { return j * NZ - (j * (j - 1)) / 2 + (i - j); }

  __device__ __forceinline__ void zero() {
PNDE_UNROLL
    for (int a = 0; a < dc; ++a)
PNDE_UNROLL
      for (int i = 0; i < D; ++i) W[a][i] = 0.0;
PNDE_UNROLL
    for (int i = 0; i < NLZ; ++i) Lz[i] = 0.0;
  }
  // rows of block k scaled by s[k]  (Diagonal * SRGaussian, src/ProbNumDiffEq.jl:58)
  __device__ __forceinline__ void scale_blocks(const double (&s)[q + 1]) {
PNDE_UNROLL
    for (int a = 0; a < dc; ++a)
PNDE_UNROLL
      for (int i = 0; i < D; ++i) W[a][i] *= s[i / dc];
PNDE_UNROLL
    for (int j = 0; j < NZ; ++j)
PNDE_UNROLL
      for (int i = j; i < NZ; ++i) Lz[lz(j, i)] *= s[(2 * dc + i) / dc];
  }
  // rows of coordinate a scaled by s[a] (apply_diffusion with a Diagonal, src/ProbNumDiffEq.jl:38)
  __device__ __forceinline__ void scale_all(double s) {
PNDE_UNROLL
    for (int a = 0; a < dc; ++a)
PNDE_UNROLL
      for (int i = 0; i < D; ++i) W[a][i] *= s;
PNDE_UNROLL
    for (int i = 0; i < NLZ; ++i) Lz[i] *= s;
  }
  // packed lower triangle (by rows) of  diag(s) S S' diag(s),  s[k] per block
  __device__ __forceinline__ void cov_entry_all(const double (&s)[q + 1], double* out, long long stride) const {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) {
PNDE_UNROLL
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
PNDE_UNROLL
        for (int a = 0; a < dc; ++a) acc = fma(W[a][i], W[a][j], acc);
PNDE_UNROLL
        for (int c = 0; c < NZ; ++c) {
          if (i >= 2 * dc + c && j >= 2 * dc + c) acc = fma(Lz[lz(c, i - 2 * dc)], Lz[lz(c, j - 2 * dc)], acc);
        }
        out[(long long)(i * (i + 1) / 2 + j) * stride] = acc * s[i / dc] * s[j / dc];
      }
    }
  }
  __device__ __forceinline__ void store(double* base, long long stride) const {
    int o = 0;
PNDE_UNROLL
    for (int a = 0; a < dc; ++a)
PNDE_UNROLL
      for (int i = 0; i < D; ++i) base[(long long)(o++) * stride] = W[a][i];
PNDE_UNROLL
    for (int i = 0; i < (NZ > 0 ? NZ * (NZ + 1) / 2 : 0); ++i) base[(long long)(o++) * stride] = Lz[i];
  }
  __device__ __forceinline__ void load(const double* base, long long stride) {
    int o = 0;
PNDE_UNROLL
    for (int a = 0; a < dc; ++a)
PNDE_UNROLL
      for (int i = 0; i < D; ++i) W[a][i] = base[(long long)(o++) * stride];
PNDE_UNROLL
    for (int i = 0; i < (NZ > 0 ? NZ * (NZ + 1) / 2 : 0); ++i) Lz[i] = base[(long long)(o++) * stride];
    if (NZ == 0) Lz[0] = 0.0;
  }
};

// One filter step of the covariance: predict (with diffusion sig^2) + exact update, in place.
//   in : F = factor of Sigma (preconditioned coordinates), Jp = pi0 * J (dc x dc), sig, pi1
//   out: F = factor of Sigma+;  Rinv[a] = 1 / Rtop[a][a];  Rtop = first dc rows of R in primed column order
//        (cols 0..dc-1: innovation factor, cols dc..2dc-1: x_0 block, cols k*dc..: block k >= 2)
template <int dc, int q, bool HASJ>
__device__ __forceinline__ void cov_filter_step(Factor<dc, q>& F, const double (&Jp)[dc][dc], const double sig,
                                                const double pi1, const double ipi1, const IwpConsts& C,
                                                double (&Rtop)[dc][dc * (q + 1)], double (&Rinv)[dc]) {
  constexpr int D = dc * (q + 1);
  constexpr int NZ = D - 2 * dc;
  double sL[q + 1][2 > q + 1 ? 2 : q + 1];  // sig * Ltilde[k][k']
PNDE_UNROLL
  for (int k = 0; k <= q; ++k)
PNDE_UNROLL
    for (int kk = 0; kk <= k; ++kk) sL[k][kk] = sig * C.Lt[k][kk];

  double E[D][D];
  // rows 0..dc-1: the prior rows of block 0, the only non-triangular part of sig (T Q_L)'
PNDE_UNROLL
  for (int i = 0; i < dc; ++i) {
PNDE_UNROLL
    for (int j = 0; j < D; ++j) {
      if (j < dc) {
        double v = (i == j) ? pi1 * sL[1][0] : 0.0;
        if (HASJ) v = fma(-sL[0][0], Jp[j][i], v);
        E[i][j] = v;
      } else {
        const int k = (j < 2 * dc) ? 0 : j / dc;
        E[i][j] = (j % dc == i) ? sL[k][0] : 0.0;
      }
    }
  }
  // bottom rows: (T A s)' for every factor column s
PNDE_UNROLL
  for (int c = 0; c < dc; ++c) {
    double w[D];
PNDE_UNROLL
    for (int i = 0; i < D; ++i) w[i] = F.W[c][i];
    apply_A<dc, q>(w);
PNDE_UNROLL
    for (int b = 0; b < dc; ++b) {
      double y = pi1 * w[dc + b];
      if (HASJ) {
PNDE_UNROLL
        for (int bb = 0; bb < dc; ++bb) y = fma(-Jp[b][bb], w[bb], y);
      }
      E[dc + c][b] = y;
      E[dc + c][dc + b] = w[b];
    }
PNDE_UNROLL
    for (int i = 2 * dc; i < D; ++i) E[dc + c][i] = w[i];
  }
PNDE_UNROLL
  for (int c = 0; c < NZ; ++c) {
    double w[D];
PNDE_UNROLL
    for (int i = 0; i < D; ++i) w[i] = (i >= 2 * dc + c) ? F.Lz[Factor<dc, q>::lz(c, i - 2 * dc)] : 0.0;
    // A w, skipping the structurally zero entries below 2dc+c (c is a compile-time constant here)
PNDE_UNROLL
    for (int k = 0; k <= q; ++k) {
PNDE_UNROLL
      for (int a = 0; a < dc; ++a) {
        bool any = (k * dc + a >= 2 * dc + c);
        double acc = any ? w[k * dc + a] : 0.0;
PNDE_UNROLL
        for (int j = k + 1; j <= q; ++j) {
          if (j * dc + a >= 2 * dc + c) {
            const double cf = inv_factorial(j - k);
            if (!any) {
              acc = (j - k == 1) ? w[j * dc + a] : cf * w[j * dc + a];
              any = true;
            } else {
              acc = (j - k == 1) ? acc + w[j * dc + a] : fma(cf, w[j * dc + a], acc);
            }
          }
        }
        w[k * dc + a] = acc;
      }
    }
PNDE_UNROLL
    for (int b = 0; b < dc; ++b) {
      double y = pi1 * w[dc + b];
      if (HASJ) {
PNDE_UNROLL
        for (int bb = 0; bb < dc; ++bb) y = fma(-Jp[b][bb], w[bb], y);
      }
      E[2 * dc + c][b] = y;
      E[2 * dc + c][dc + b] = w[b];
    }
PNDE_UNROLL
    for (int i = 2 * dc; i < D; ++i) E[2 * dc + c][i] = w[i];
  }

  // Householder sweep over the primed columns
PNDE_UNROLL
  for (int c = 0; c < D; ++c) {
    const int first = (c < dc) ? 0 : (c < 2 * dc ? c - dc + 1 : dc);  // active dense rows [first, D)
    const int kc = c / dc, ac = c % dc;
    double pv;
    if (c < dc)
      pv = pi1 * sL[1][1];
    else if (c < 2 * dc)
      pv = E[c - dc][c];
    else
      pv = sL[kc][kc];
    // squared norm in two interleaved chains (halves the dependent-FMA latency)
    double n2a = pv * pv, n2b = E[first][c] * E[first][c];
PNDE_UNROLL
    for (int i = first + 1; i < D; ++i) {
      if ((i - first) & 1)
        n2a = fma(E[i][c], E[i][c], n2a);
      else
        n2b = fma(E[i][c], E[i][c], n2b);
    }
    const double n2 = n2a + n2b;
    const bool nzcol = n2 > 0.0;
    const double rn = nzcol ? fast_rsqrt(n2) : 0.0;  // 1 / ||x||
    const double nrm = n2 * rn;
    const double snrm = copysign(nrm, pv);
    const double v0 = pv + snrm;
    const double beta = nzcol ? fast_rcp(fma(fabs(pv), nrm, n2)) : 0.0;  // 1 / (||x|| (||x|| + |pv|))
    double Rrow[D];
    Rrow[c] = -snrm;
PNDE_UNROLL
    for (int j = c + 1; j < D; ++j) {
      // pivot-row entry (compile-time sparsity for the prior rows)
      bool pnz;
      double prj = 0.0;
      if (c < dc) {
        pnz = (j >= 2 * dc) && (j % dc == ac);
        if (pnz) prj = sL[j / dc][1];
      } else if (c < 2 * dc) {
        pnz = true;
        prj = E[c - dc][j];
      } else {
        pnz = (j % dc == ac);
        if (pnz) prj = sL[j / dc][kc];
      }
      double w = pnz ? v0 * prj : 0.0;
      bool started = pnz;
      {
PNDE_UNROLL
        for (int i = first; i < D; ++i) {
          if (!started) {
            w = E[i][c] * E[i][j];
            started = true;
          } else {
            w = fma(E[i][c], E[i][j], w);
          }
        }
      }
      const double s = beta * w;
      Rrow[j] = pnz ? fma(-s, v0, prj) : -s * v0;
PNDE_UNROLL
      for (int i = first; i < D; ++i) E[i][j] = fma(-s, E[i][c], E[i][j]);
    }
    // consume the finished row of R
    if (c < dc) {
PNDE_UNROLL
      for (int j = 0; j < D; ++j) Rtop[c][j] = (j >= c) ? Rrow[j] : 0.0;
      Rinv[c] = -copysign(rn, pv);  // 1 / R[c][c] (0 for a zero column)
    } else if (c < 2 * dc) {
      const int a = c - dc;
PNDE_UNROLL
      for (int b = 0; b < dc; ++b) F.W[a][b] = (b >= a) ? Rrow[dc + b] : 0.0;
PNDE_UNROLL
      for (int i = 2 * dc; i < D; ++i) F.W[a][i] = Rrow[i];
    } else {
      const int j0 = c - 2 * dc;
PNDE_UNROLL
      for (int i = j0; i < NZ; ++i) F.Lz[Factor<dc, q>::lz(j0, i)] = Rrow[2 * dc + i];
    }
  }
  // block 1 of the posterior factor is slaved to block 0: x_1 = (Jp x_0) / pi1  (H S+ = 0)
PNDE_UNROLL
  for (int a = 0; a < dc; ++a) {
PNDE_UNROLL
    for (int b = 0; b < dc; ++b) {
      double v = 0.0;
      if (HASJ) {
PNDE_UNROLL
        for (int bb = 0; bb < dc; ++bb) v = fma(Jp[b][bb], F.W[a][bb], v);
        v *= ipi1;
      }
      F.W[a][dc + b] = v;
    }
  }
}

}  // namespace pnde
