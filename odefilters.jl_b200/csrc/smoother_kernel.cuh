// Backward Rauch-Tung-Striebel pass over the saved filter history, one trajectory per thread.
//
// Reference being replaced: smooth_all! / smooth!  src/smoothing.jl:4-63  (algebra twin: smooth,
// src/filtering.jl:136-154).  The reference forms G = Sigma A' inv(Sigma-) with a dense LU inverse
// and triangularises the 3D x D stack [S'(I-GA)'; Q_L'G'; S_next'G'].  Here:
//   stage 1  one Householder sweep over  [sig Q_L' | 0 ; (A S)' | S']  (the filter's own predict QR
//            carried through D extra columns) gives  [R- X ; 0 Y]  with  G' = R-^-1 X  and
//            Y'Y = Sigma - G Sigma- G'  (= (I-GA) Sigma (I-GA)' + G Q G', the backward-kernel noise);
//   stage 2  Z = R-^-T S_next (lower triangular), T = X' Z = G S_next, mean += X' R-^-T (m_next - m-);
//   stage 3  triangularise [Y ; T'] -> the smoothed factor, lower triangular.
// G is never formed and no matrix is inverted.
#pragma once
#include "model_ops.cuh"

// The smoother's hot loop is ~100 KB of straight-line code and is sensitive to instruction-cache misses (ncu: 13 % of
// the stall samples are no_instruction): the two loops of step_cols_smem that only index shared memory are kept
// rolled (PNDE_SMOOTH_UNROLL = 1: -12 % instructions, -15 % time at q = 3).
#ifndef PNDE_SMOOTH_UNROLL
#define PNDE_SMOOTH_UNROLL 1
#endif

namespace pnde {
constexpr int kSmoothUnroll = PNDE_SMOOTH_UNROLL;

// D x D matrix either in registers or in shared memory (one column of shared memory per thread, stride =
// block size, so consecutive threads hit consecutive banks)
template <int D>
struct RegMat {
  double v[D][D];
  __device__ __forceinline__ double get(int i, int j) const { return v[i][j]; }
  __device__ __forceinline__ void set(int i, int j, double x) { v[i][j] = x; }
};
// [element][thread] layout in shared memory; the stride (= threads per CTA) is a compile-time constant so that every
// address is base + immediate (a run-time stride costs one IMAD per access: ~9 % of the smoother's instructions)
#ifndef PNDE_SMOOTH_BLOCK
#define PNDE_SMOOTH_BLOCK 128
#endif
template <int D>
struct SmemMat {
  static constexpr int st = PNDE_SMOOTH_BLOCK;
  double* p;
  __device__ __forceinline__ double get(int i, int j) const { return p[(i * D + j) * st]; }
  __device__ __forceinline__ void set(int i, int j, double x) { p[(i * D + j) * st] = x; }
};

template <int dc, int q>
struct SmoothCov {
  static constexpr int D = dc * (q + 1);
  static constexpr int R = D - dc;
  static constexpr int NP = D * (D + 1) / 2;
  __host__ __device__ static constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // j <= i

  // Householder triangularisation of the (NR1 + D) x D stack [Y ; T'] -> packed lower factor L (= R').
  // Y: NR1 x D dense rows.  Tt: D x D where Tt[c][i] = T[i][c]; rows of the stack are Tt[c][:].
  template <int NR1>
  __device__ __forceinline__ static void triangularize_impl(double (&Y)[NR1 > 0 ? NR1 : 1][D], double (&Tt)[D][D],
                                                       double (&L)[NP], int& status) {
PNDE_UNROLL
    for (int c = 0; c < D; ++c) {
      // pivot: Tt row c is NOT special here; use a virtual zero pivot row -> plain Householder on all
      // rows with the first active row as pivot.  Rows are consumed in order: Y rows then Tt rows.
      // Active rows at step c: all rows with index >= c in the concatenated order.
      double n2 = 0.0;
PNDE_UNROLL
      for (int i = c; i < NR1 + D; ++i) {
        const double v = (i < NR1) ? Y[i < NR1 ? i : 0][c] : Tt[i - NR1 >= 0 ? i - NR1 : 0][c];
        n2 = fma(v, v, n2);
      }
      const double pv = (c < NR1) ? Y[c < NR1 ? c : 0][c] : Tt[c - NR1 >= 0 ? c - NR1 : 0][c];
      const bool nz = n2 > 0.0;
      const double rn = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double snrm = copysign(nrm, pv);
      const double v0 = pv + snrm;
      const double beta = nz ? fast_rcp(fma(fabs(pv), nrm, n2)) : 0.0;
      L[tri(c, c)] = -snrm;
PNDE_UNROLL
      for (int j = c + 1; j < D; ++j) {
        const double prj = (c < NR1) ? Y[c < NR1 ? c : 0][j] : Tt[c - NR1 >= 0 ? c - NR1 : 0][j];
        double w = v0 * prj;
PNDE_UNROLL
        for (int i = c + 1; i < NR1 + D; ++i) {
          const double vc = (i < NR1) ? Y[i < NR1 ? i : 0][c] : Tt[i - NR1 >= 0 ? i - NR1 : 0][c];
          const double vj = (i < NR1) ? Y[i < NR1 ? i : 0][j] : Tt[i - NR1 >= 0 ? i - NR1 : 0][j];
          w = fma(vc, vj, w);
        }
        const double s = beta * w;
        L[tri(j, c)] = fma(-s, v0, prj);
PNDE_UNROLL
        for (int i = c + 1; i < NR1 + D; ++i) {
          if (i < NR1) {
            Y[i < NR1 ? i : 0][j] = fma(-s, Y[i < NR1 ? i : 0][c], Y[i < NR1 ? i : 0][j]);
          } else {
            Tt[i - NR1 >= 0 ? i - NR1 : 0][j] =
                fma(-s, Tt[i - NR1 >= 0 ? i - NR1 : 0][c], Tt[i - NR1 >= 0 ? i - NR1 : 0][j]);
          }
        }
      }
      if (!(n2 == n2)) status |= 1;  // NaN  (src/smoothing.jl:25)
    }
  }

  // columns of the reduced-rank filtered factor as dense D-vectors
  __device__ __forceinline__ static void cols_from_factor(const Factor<dc, q>& F, double (&cols)[R][D]) {
PNDE_UNROLL
    for (int c = 0; c < R; ++c) {
PNDE_UNROLL
      for (int i = 0; i < D; ++i) {
        if (c < dc)
          cols[c][i] = F.W[c < dc ? c : 0][i];
        else
          cols[c][i] = (i >= 2 * dc + (c - dc))
                           ? F.Lz[Factor<dc, q>::lz(c - dc >= 0 ? c - dc : 0, i - 2 * dc >= 0 ? i - 2 * dc : 0)]
                           : 0.0;
      }
    }
  }
  // columns of a packed lower-triangular D x D factor
  __device__ __forceinline__ static void cols_from_lower(const double (&L)[NP], double (&cols)[D][D]) {
PNDE_UNROLL
    for (int c = 0; c < D; ++c)
PNDE_UNROLL
      for (int i = 0; i < D; ++i) cols[c][i] = (i >= c) ? L[tri(i, c)] : 0.0;
  }

  // stage 1: Householder sweep over [sig Q_L' | 0 ; (A cols)' | cols'] in natural coordinate order.
  //   in : Er = the NR factor columns (P(h) coordinates)
  //   out: Rm = R-' packed lower (Rm[tri(j,c)] = R-[c][j]), rinv[c] = 1/R-[c][c], X = top-right block,
  //        Er = Y (backward-kernel noise factor: Y'Y = Sigma - G Sigma- G')
  template <int NR, class XV>
  __device__ __forceinline__ static void stage1_impl(double (&Er)[NR][D], const double sig, const IwpConsts& C,
                                                double (&Rm)[NP], double (&rinv)[D], XV& X) {
    double sL[q + 1][q + 1];
PNDE_UNROLL
    for (int k = 0; k <= q; ++k)
PNDE_UNROLL
      for (int kk = 0; kk <= k; ++kk) sL[k][kk] = sig * C.Lt[k][kk];
    double El[NR][D];
PNDE_UNROLL
    for (int c = 0; c < NR; ++c) {
      double w[D];
PNDE_UNROLL
      for (int i = 0; i < D; ++i) w[i] = Er[c][i];
      apply_A<dc, q>(w);
PNDE_UNROLL
      for (int i = 0; i < D; ++i) El[c][i] = w[i];
    }
PNDE_UNROLL
    for (int c = 0; c < D; ++c) {
      const int kc = c / dc, ac = c % dc;
      const double pv = sL[kc][kc];
      double n2 = pv * pv;
PNDE_UNROLL
      for (int i = 0; i < NR; ++i) n2 = fma(El[i][c], El[i][c], n2);
      const bool nz = n2 > 0.0;
      const double rn = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double v0 = pv + nrm;
      const double beta = nz ? fast_rcp(fma(pv, nrm, n2)) : 0.0;
      Rm[tri(c, c)] = -nrm;
      rinv[c] = -rn;
PNDE_UNROLL
      for (int j = c + 1; j < D; ++j) {
        const bool pnz = (j % dc == ac);
        const double prj = pnz ? sL[j / dc][kc] : 0.0;
        double w = pnz ? v0 * prj : 0.0;
PNDE_UNROLL
        for (int i = 0; i < NR; ++i) w = (i == 0 && !pnz) ? El[i][c] * El[i][j] : fma(El[i][c], El[i][j], w);
        const double s = beta * w;
        Rm[tri(j, c)] = pnz ? fma(-s, v0, prj) : -s * v0;
PNDE_UNROLL
        for (int i = 0; i < NR; ++i) El[i][j] = fma(-s, El[i][c], El[i][j]);
      }
PNDE_UNROLL
      for (int j = 0; j < D; ++j) {
        double w = El[0][c] * Er[0][j];
PNDE_UNROLL
        for (int i = 1; i < NR; ++i) w = fma(El[i][c], Er[i][j], w);
        const double s = beta * w;
        X.set(c, j, -s * v0);
PNDE_UNROLL
        for (int i = 0; i < NR; ++i) Er[i][j] = fma(-s, El[i][c], Er[i][j]);
      }
    }
  }

  // delta <- G delta = X' (R-^-T delta)
  template <int NREP, class XV>
  __device__ __forceinline__ static void apply_gain_impl(const double (&Rm)[NP], const double (&rinv)[D], const XV& X,
                                                    double (&delta)[NREP][D]) {
PNDE_UNROLL
    for (int r = 0; r < NREP; ++r) {
      double y[D];
PNDE_UNROLL
      for (int i = 0; i < D; ++i) {
        double acc = delta[r][i];
PNDE_UNROLL
        for (int k = 0; k < i; ++k) acc = fma(-Rm[tri(i, k)], y[k], acc);
        y[i] = acc * rinv[i];
      }
PNDE_UNROLL
      for (int i = 0; i < D; ++i) {
        double acc = 0.0;
PNDE_UNROLL
        for (int k = 0; k < D; ++k) acc = fma(X.get(k, i), y[k], acc);
        delta[r][i] = acc;
      }
    }
  }

  // Tt[c][i] = (G Ls)[i][c]:  Z = R-^-T Ls (lower triangular, forward substitution), T = X' Z
  __device__ __forceinline__ static void gain_times_lower_impl(const double (&Rm)[NP], const double (&rinv)[D],
                                                          const RegMat<D>& X, const double (&Ls)[NP],
                                                          double (&Tt)[D][D]) {
    double Z[NP];
PNDE_UNROLL
    for (int c = 0; c < D; ++c) {
PNDE_UNROLL
      for (int i = c; i < D; ++i) {
        double acc = Ls[tri(i, c)];
PNDE_UNROLL
        for (int k = c; k < i; ++k) acc = fma(-Rm[tri(i, k)], Z[tri(k, c)], acc);
        Z[tri(i, c)] = acc * rinv[i];
      }
    }
PNDE_UNROLL
    for (int c = 0; c < D; ++c) {
PNDE_UNROLL
      for (int i = 0; i < D; ++i) {
        double acc = X.get(c, i) * Z[tri(c, c)];
PNDE_UNROLL
        for (int k = c + 1; k < D; ++k) acc = fma(X.get(k, i), Z[tri(k, c)], acc);
        Tt[c][i] = acc;
      }
    }
  }

  // One RTS step of the covariance.
  //   cols : the NR columns of the current (filtered / predicted) factor in P(h) coordinates; destroyed
  //   Ls   : smoothed factor at i+1 in P(h) coordinates, packed lower; overwritten by the smoothed
  //          factor at i (still P(h) coordinates)
  //   delta: in  m_next_smoothed - A m  ->  out  G * delta          (NREP mean replicas)
  template <int NR, int NREP>
  __device__ __forceinline__ static void step_cols(double (&cols)[NR][D], const double sig, const IwpConsts& C,
                                                   double (&Ls)[NP], double (&delta)[NREP][D], int& status) {
    double Rm[NP], rinv[D];
    RegMat<D> X;
    stage1<NR>(cols, sig, C, Rm, rinv, X);
    apply_gain<NREP>(Rm, rinv, X, delta);
    double Tt[D][D];
    gain_times_lower(Rm, rinv, X, Ls, Tt);
    triangularize<NR>(cols, Tt, Ls, status);
  }
  // Row-block Householder update: R_acc (upper triangular, stored as its transpose: Racc[tri(j,c)] = R[c][j])
  // <- triangular factor of [R_acc ; rows].  Reflector c = [R_acc[c][c] ; rows[:, c]] (length NCH + 1).
  template <int NCH>
  __device__ __forceinline__ static void qr_update_rows_impl(double (&Racc)[NP], double (&rows)[NCH][D], int& status) {
PNDE_UNROLL
    for (int c = 0; c < D; ++c) {
      const double pv = Racc[tri(c, c)];
      double n2 = pv * pv;
PNDE_UNROLL
      for (int i = 0; i < NCH; ++i) n2 = fma(rows[i][c], rows[i][c], n2);
      const bool nz = n2 > 0.0;
      const double rn = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double snrm = copysign(nrm, pv);
      const double v0 = pv + snrm;
      const double beta = nz ? fast_rcp(fma(fabs(pv), nrm, n2)) : 0.0;
      Racc[tri(c, c)] = -snrm;
      if (!(n2 == n2)) status |= 1;
PNDE_UNROLL
      for (int j = c + 1; j < D; ++j) {
        const double prj = Racc[tri(j, c)];
        double w = v0 * prj;
PNDE_UNROLL
        for (int i = 0; i < NCH; ++i) w = fma(rows[i][c], rows[i][j], w);
        const double s = beta * w;
        Racc[tri(j, c)] = fma(-s, v0, prj);
PNDE_UNROLL
        for (int i = 0; i < NCH; ++i) rows[i][j] = fma(-s, rows[i][c], rows[i][j]);
      }
    }
  }

  // Same RTS step as step_cols, laid out for one thread per trajectory WITHOUT register spills: X (and then
  // T' in its place) and the smoothed factor L^s live in shared memory, the final triangularisation is a
  // sequence of row-block updates (Y, then T' four rows at a time) so that at most ~85 doubles are live.
  //   Xs : D x D scratch in shared memory;  Lsv: packed lower L^s_{i+1} in shared memory, natural coordinates;
  //   Pk / PIk: block scales of this interval (L^s is re-preconditioned on the fly)
  template <int NR>
  __device__ __forceinline__ static void step_cols_smem(double (&cols)[NR][D], const double sig, const IwpConsts& C,
                                                        SmemMat<D>& Xs, double* Lsv, const double (&Pk)[q + 1],
                                                        const double (&PIk)[q + 1], double (&delta)[1][D], int& status) {
    constexpr int lst = SmemMat<D>::st;
    double Rm[NP], rinv[D];
    stage1<NR>(cols, sig, C, Rm, rinv, Xs);
    apply_gain<1>(Rm, rinv, Xs, delta);
    // Z = R-^-T (P L^s): forward substitution row by row
    double Z[NP];
PNDE_UNROLL
    for (int i = 0; i < D; ++i) {
PNDE_UNROLL
      for (int c = 0; c <= i; ++c) {
        double acc = Lsv[tri(i, c) * lst] * Pk[i / dc];
PNDE_UNROLL
        for (int k = c; k < i; ++k) acc = fma(-Rm[tri(i, k)], Z[tri(k, c)], acc);
        Z[tri(i, c)] = acc * rinv[i];
      }
    }
    // T' in place of X: Tt[c][i] = sum_{k >= c} X[k][i] Z[k][c], one column i of X at a time
#pragma unroll kSmoothUnroll
    for (int i = 0; i < D; ++i) {
      double xc[D];
PNDE_UNROLL
      for (int k = 0; k < D; ++k) xc[k] = Xs.get(k, i);
PNDE_UNROLL
      for (int c = 0; c < D; ++c) {
        double acc = xc[c] * Z[tri(c, c)];
PNDE_UNROLL
        for (int k = c + 1; k < D; ++k) acc = fma(xc[k], Z[tri(k, c)], acc);
        Xs.set(c, i, acc);
      }
    }
    // triangularise [Y ; T'] by row blocks
    double Racc[NP];
PNDE_UNROLL
    for (int i = 0; i < NP; ++i) Racc[i] = 0.0;
    qr_update_rows<NR>(Racc, cols, status);
    constexpr int CH = 4;
#pragma unroll kSmoothUnroll
    for (int r0 = 0; r0 < D; r0 += CH) {
      double rows[CH][D];
PNDE_UNROLL
      for (int i = 0; i < CH; ++i)
PNDE_UNROLL
        for (int j = 0; j < D; ++j) rows[i][j] = (r0 + i < D) ? Xs.get(r0 + i < D ? r0 + i : 0, j) : 0.0;
      qr_update_rows<CH>(Racc, rows, status);
    }
    // L^s_i = R', back to natural coordinates
PNDE_UNROLL
    for (int j = 0; j < D; ++j)
PNDE_UNROLL
      for (int c = 0; c <= j; ++c) Lsv[tri(j, c) * lst] = Racc[tri(j, c)] * PIk[j / dc];
  }

  // For D >= 10 the fully inlined smoother exceeds what ptxas optimises as one function (it falls back to 32 registers
  // and a 50 KB stack: measured 10 M steps/s at q = 5); as separate functions the same pieces compile normally
  // (43 M steps/s).  Below D = 10 everything stays inlined (noinline costs 6x at q = 3).
  static constexpr bool OUTLINE = (D >= 10);
  template <int NR1>
  __device__ __noinline__ static void triangularize_ni(double (&Y)[NR1 > 0 ? NR1 : 1][D], double (&Tt)[D][D], double (&L)[NP],
                                                       int& status) {
    triangularize_impl<NR1>(Y, Tt, L, status);
  }
  template <int NR1>
  __device__ __forceinline__ static void triangularize(double (&Y)[NR1 > 0 ? NR1 : 1][D], double (&Tt)[D][D], double (&L)[NP],
                                                       int& status) {
    if constexpr (OUTLINE) triangularize_ni<NR1>(Y, Tt, L, status); else triangularize_impl<NR1>(Y, Tt, L, status);
  }
  template <int NR, class XV>
  __device__ __noinline__ static void stage1_ni(double (&Er)[NR][D], const double sig, const IwpConsts& C, double (&Rm)[NP],
                                                double (&rinv)[D], XV& X) {
    stage1_impl<NR, XV>(Er, sig, C, Rm, rinv, X);
  }
  template <int NR, class XV>
  __device__ __forceinline__ static void stage1(double (&Er)[NR][D], const double sig, const IwpConsts& C, double (&Rm)[NP],
                                                double (&rinv)[D], XV& X) {
    if constexpr (OUTLINE) stage1_ni<NR, XV>(Er, sig, C, Rm, rinv, X); else stage1_impl<NR, XV>(Er, sig, C, Rm, rinv, X);
  }
  template <int NREP, class XV>
  __device__ __noinline__ static void apply_gain_ni(const double (&Rm)[NP], const double (&rinv)[D], const XV& X,
                                                    double (&delta)[NREP][D]) {
    apply_gain_impl<NREP, XV>(Rm, rinv, X, delta);
  }
  template <int NREP, class XV>
  __device__ __forceinline__ static void apply_gain(const double (&Rm)[NP], const double (&rinv)[D], const XV& X,
                                                    double (&delta)[NREP][D]) {
    if constexpr (OUTLINE) apply_gain_ni<NREP, XV>(Rm, rinv, X, delta); else apply_gain_impl<NREP, XV>(Rm, rinv, X, delta);
  }
  __device__ __noinline__ static void gain_times_lower_ni(const double (&Rm)[NP], const double (&rinv)[D], const RegMat<D>& X,
                                                          const double (&Ls)[NP], double (&Tt)[D][D]) {
    gain_times_lower_impl(Rm, rinv, X, Ls, Tt);
  }
  __device__ __forceinline__ static void gain_times_lower(const double (&Rm)[NP], const double (&rinv)[D], const RegMat<D>& X,
                                                          const double (&Ls)[NP], double (&Tt)[D][D]) {
    if constexpr (OUTLINE) gain_times_lower_ni(Rm, rinv, X, Ls, Tt); else gain_times_lower_impl(Rm, rinv, X, Ls, Tt);
  }
  template <int NCH>
  __device__ __noinline__ static void qr_update_rows_ni(double (&Racc)[NP], double (&rows)[NCH][D], int& status) {
    qr_update_rows_impl<NCH>(Racc, rows, status);
  }
  template <int NCH>
  __device__ __forceinline__ static void qr_update_rows(double (&Racc)[NP], double (&rows)[NCH][D], int& status) {
    if constexpr (OUTLINE) qr_update_rows_ni<NCH>(Racc, rows, status); else qr_update_rows_impl<NCH>(Racc, rows, status);
  }

  template <int NREP>
  __device__ __forceinline__ static void step(const Factor<dc, q>& F, const double sig, const IwpConsts& C,
                                              double (&Ls)[NP], double (&delta)[NREP][D], int& status) {
    double cols[R][D];
    cols_from_factor(F, cols);
    step_cols<R, NREP>(cols, sig, C, Ls, delta, status);
  }
};

// ---------------------------------------------------------------------------------------------
// Smoothed-record layout and the kernel
// ---------------------------------------------------------------------------------------------
template <class M>
struct SmoothModel;

template <class VF, int q_>
struct SmoothModel<DenseEK1<VF, q_>> {
  using M = DenseEK1<VF, q_>;
  static constexpr int d = M::d, q = M::q, D = M::D, NF = 1, DC = d;
  using SC = SmoothCov<d, q>;
  static constexpr int SREC = D + SC::NP;
  static constexpr bool USE_SMEM = (D < 10);  // shared-memory scratch of the smoother kernel (see there)
  __device__ static void load_cov(const double* sb, long long n, double* mean, double* cov) {
    double L[SC::NP];
PNDE_UNROLL
    for (int i = 0; i < D; ++i) mean[i] = sb[(long long)i * n];
PNDE_UNROLL
    for (int i = 0; i < SC::NP; ++i) L[i] = sb[(long long)(D + i) * n];
PNDE_UNROLL
    for (int i = 0; i < D; ++i)
PNDE_UNROLL
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
PNDE_UNROLL
        for (int k = 0; k <= j; ++k) acc = fma(L[SC::tri(i, k)], L[SC::tri(j, k)], acc);
        cov[SC::tri(i, j)] = acc;
      }
  }
};

template <class VF, int q_, bool MVDYN>
struct SmoothModel<KronEK0<VF, q_, MVDYN>> {
  using M = KronEK0<VF, q_, MVDYN>;
  static constexpr int d = M::d, q = M::q, D = M::D, NF = M::NF, DC = 1;
  using SC = SmoothCov<1, q>;
  static constexpr int SREC = D + NF * SC::NP + d;  // mean, factors, per-dimension calibration scale
  static constexpr bool USE_SMEM = false;
  __device__ static void load_cov(const double* sb, long long n, double* mean, double* cov) {
PNDE_UNROLL
    for (int i = 0; i < D; ++i) mean[i] = sb[(long long)i * n];
    double L[NF][SC::NP];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f)
PNDE_UNROLL
      for (int i = 0; i < SC::NP; ++i) L[f][i] = sb[(long long)(D + f * SC::NP + i) * n];
    double ds[d];
PNDE_UNROLL
    for (int a = 0; a < d; ++a) ds[a] = sb[(long long)(D + NF * SC::NP + a) * n];
PNDE_UNROLL
    for (int i = 0; i < D; ++i)
PNDE_UNROLL
      for (int j = 0; j <= i; ++j) {
        const int ki = i / d, ai = i % d, kj = j / d, aj = j % d;
        double acc = 0.0;
        if (ai == aj) {
          const int f = MVDYN ? ai : 0;
          const int lo = ki < kj ? ki : kj;
PNDE_UNROLL
          for (int k = 0; k <= q; ++k)
            if (k <= lo) acc = fma(L[f][SC::tri(ki, k)], L[f][SC::tri(kj, k)], acc);
          acc *= ds[ai];
        }
        cov[i * (i + 1) / 2 + j] = acc;
      }
  }
};

#ifndef PNDE_PREFETCH_OP
#define PNDE_PREFETCH_OP "prefetch.global.L2"
#endif
#ifndef PNDE_PREFETCH_AHEAD
#define PNDE_PREFETCH_AHEAD 1
#endif
template <class M>
__global__ void __launch_bounds__(128) smoother_kernel(const SmoothParams sp) {
  using SM = SmoothModel<M>;
  using SC = typename SM::SC;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC, SREC = SM::SREC, NF = SM::NF, DC = SM::DC;
  constexpr int DCOV = DC * (q + 1);  // dimension of one covariance factor

  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= sp.n) return;
  const long long n = sp.n;
  const int ns = sp.n_saved[tid];
  int status = 0;
  if (ns <= 0) {
    sp.status[tid] = 0;
    return;
  }
  // calibration (static diffusion models): every filtered covariance is scaled by the final global
  // diffusion (src/integrator_utils.jl:7-12); for fixedMV this is a per-dimension scale which the
  // Kronecker form carries OUTSIDE the (shared) factor.
  double gfin[ND];
PNDE_UNROLL
  for (int i = 0; i < ND; ++i) gfin[i] = sp.calibrate ? sp.final_diff[(long long)i * n + tid] : 1.0;
  double dimscale[d];
PNDE_UNROLL
  for (int a = 0; a < d; ++a) dimscale[a] = (sp.calibrate && !M::IS_EK1) ? (sp.is_mv ? gfin[a < ND ? a : 0] : gfin[0]) : 1.0;
  const double dense_cal = (sp.calibrate && M::IS_EK1) ? sqrt(gfin[0]) : 1.0;

  auto rec = [&](int slot) { return sp.hist + ((long long)slot * REC) * n + tid; };
  auto srec = [&](int slot) { return sp.smooth + ((long long)slot * SREC) * n + tid; };

  double ms[D];          // smoothed mean at i+1 (natural coordinates)
  double Ls[NF][SC::NP]; // smoothed factor(s) at i+1 (natural coordinates), packed lower (Kronecker models)
  // dense EK1 up to D = 8: the D x D scratch X / T' and the smoothed factor live in shared memory (no register
  // spills).  For D >= 10 that layout leaves 128 threads per SM; the plain path (everything in registers / local
  // memory, the whole L1 for the stack, 256 threads per SM) is faster there.
  constexpr bool USE_SMEM = SmoothModel<M>::USE_SMEM;
  extern __shared__ double sm_dyn[];
  constexpr int lst = SmemMat<DCOV>::st;  // == blockDim.x (launch_smooth_t / rtc_smooth)
  if (USE_SMEM && blockDim.x != lst) __trap();
  SmemMat<DCOV> Xs{sm_dyn + threadIdx.x};
  double* Lsv = sm_dyn + (size_t)DCOV * DCOV * lst + threadIdx.x;
  auto write = [&](int slot) {
    double* o = srec(slot);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) o[(long long)i * n] = ms[i];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f)
PNDE_UNROLL
      for (int i = 0; i < SC::NP; ++i)
        o[(long long)(D + f * SC::NP + i) * n] = USE_SMEM ? Lsv[i * lst] : Ls[f][i];
    if constexpr (!M::IS_EK1) {
PNDE_UNROLL
      for (int a = 0; a < d; ++a) o[(long long)(D + NF * SC::NP + a) * n] = dimscale[a];
    }
  };
  // x_smooth[i] = x_filt[i] as a triangular factor (last state; also used for the un-smoothed first)
  auto from_filtered = [&](int slot) {
    typename M::State st;
    M::load(st, rec(slot) + (long long)(1 + ND) * n, n);
PNDE_UNROLL
    for (int i = 0; i < D; ++i) ms[i] = st.m[i];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) {
      const Factor<DC, q>* F;
      if constexpr (M::IS_EK1) F = &st.F; else F = &st.F[f];
      double Y[SC::R][DCOV], Tt[DCOV][DCOV];
PNDE_UNROLL
      for (int c = 0; c < SC::R; ++c)
PNDE_UNROLL
        for (int i = 0; i < DCOV; ++i) {
          if (c < DC)
            Y[c][i] = F->W[c < DC ? c : 0][i] * dense_cal;
          else
            Y[c][i] = (i >= 2 * DC + (c - DC))
                          ? F->Lz[Factor<DC, q>::lz(c - DC >= 0 ? c - DC : 0, i - 2 * DC >= 0 ? i - 2 * DC : 0)] * dense_cal
                          : 0.0;
        }
PNDE_UNROLL
      for (int c = 0; c < DCOV; ++c)
PNDE_UNROLL
        for (int i = 0; i < DCOV; ++i) Tt[c][i] = 0.0;
      SC::template triangularize<SC::R>(Y, Tt, Ls[f], status);
      if constexpr (USE_SMEM) {
PNDE_UNROLL
        for (int i = 0; i < SC::NP; ++i) Lsv[i * lst] = Ls[0][i];
      }
    }
  };

  from_filtered(ns - 1);
  write(ns - 1);
  for (int i = ns - 2; i >= 1; --i) {
    const double* ri = rec(i);
    const double* rn = rec(i + 1);
    // the records are streamed once, backwards: pull the next one (i-1) towards L1/L2 while this step computes
    if (i > PNDE_PREFETCH_AHEAD) {
      const double* rp = rec(i - PNDE_PREFETCH_AHEAD);
PNDE_UNROLL
      for (int k = 0; k < REC; ++k) asm volatile(PNDE_PREFETCH_OP " [%0];" ::"l"(rp + (long long)k * n));
    }
    const double h = rn[0] - ri[0];
    // h == 0 (or a sliver, filter_kernel.cuh): the state is kept as it is (src/smoothing.jl:13-16); one copy of write() below
    if (!sliver_interval(h, ri[0], rn[0], ns, sp.calibrate)) {
    double Pk[q + 1], PIk[q + 1];
    precond_scales<q>(h, Pk, PIk);
    typename M::State st;
    M::load(st, ri + (long long)(1 + ND) * n, n);
    M::scale(st, Pk);  // x[i] = P * x[i]  (src/smoothing.jl:23)
    // diffusion of the interval t[i] -> t[i+1] is stored with state i+1 (src/integrator_utils.jl:44)
    double sig[NF];
PNDE_UNROLL
    for (int f = 0; f < NF; ++f) {
      const double g = sp.calibrate ? (M::IS_EK1 ? gfin[0] : 1.0) : rn[(long long)(1 + (NF > 1 ? f : 0)) * n];
      sig[f] = sqrt(g);
    }
    // delta = P m_next_smoothed - A P m_i ; mean replicas laid out per factor coordinate
    double mpred[D];
PNDE_UNROLL
    for (int k = 0; k < D; ++k) mpred[k] = st.m[k];
    apply_A<d, q>(mpred);
    if constexpr (!USE_SMEM) {
PNDE_UNROLL
      for (int f = 0; f < NF; ++f) {
        // scale smoothed factor at i+1 into P(h) coordinates
PNDE_UNROLL
        for (int r = 0; r < DCOV; ++r)
PNDE_UNROLL
          for (int c = 0; c <= r; ++c) Ls[f][SC::tri(r, c)] *= Pk[r / DC];
      }
    }
    if constexpr (M::IS_EK1 && !USE_SMEM) {
      st.F.scale_all(dense_cal);
      double delta[1][D];
PNDE_UNROLL
      for (int k = 0; k < D; ++k) delta[0][k] = fma(Pk[k / d], ms[k], -mpred[k]);
      SC::template step<1>(st.F, sig[0], sp.C, Ls[0], delta, status);
PNDE_UNROLL
      for (int k = 0; k < D; ++k) ms[k] = (st.m[k] + delta[0][k]) * PIk[k / d];
    } else if constexpr (M::IS_EK1) {
      st.F.scale_all(dense_cal);
      double delta[1][D];
PNDE_UNROLL
      for (int k = 0; k < D; ++k) delta[0][k] = fma(Pk[k / d], ms[k], -mpred[k]);
      double cols[SC::R][D];
      SC::cols_from_factor(st.F, cols);
      SC::template step_cols_smem<SC::R>(cols, sig[0], sp.C, Xs, Lsv, Pk, PIk, delta, status);
PNDE_UNROLL
      for (int k = 0; k < D; ++k) ms[k] = (st.m[k] + delta[0][k]) * PIk[k / d];
    } else if constexpr (NF == 1) {
      // Kronecker, shared factor: d mean replicas, replica a holds coordinates (k, a), k = 0..q.
      // With a static per-dimension calibration the scale cancels in G, so the shared factor is
      // smoothed uncalibrated and dimscale is applied at output.
      double delta[d][q + 1];
PNDE_UNROLL
      for (int a = 0; a < d; ++a)
PNDE_UNROLL
        for (int k = 0; k <= q; ++k) delta[a][k] = fma(Pk[k], ms[k * d + a], -mpred[k * d + a]);
      const double sg = sp.calibrate ? 1.0 : sig[0];
      SC::template step<d>(st.F[0], sg, sp.C, Ls[0], delta, status);
PNDE_UNROLL
      for (int a = 0; a < d; ++a)
PNDE_UNROLL
        for (int k = 0; k <= q; ++k) ms[k * d + a] = (st.m[k * d + a] + delta[a][k]) * PIk[k];
    } else {
      // dynamicMV: one factor per dimension, one replica each
PNDE_UNROLL
      for (int a = 0; a < d; ++a) {
        double delta[1][q + 1];
PNDE_UNROLL
        for (int k = 0; k <= q; ++k) delta[0][k] = fma(Pk[k], ms[k * d + a], -mpred[k * d + a]);
        SC::template step<1>(st.F[a < NF ? a : 0], sig[a < NF ? a : 0], sp.C, Ls[a < NF ? a : 0], delta, status);
PNDE_UNROLL
        for (int k = 0; k <= q; ++k) ms[k * d + a] = (st.m[k * d + a] + delta[0][k]) * PIk[k];
      }
    }
    if constexpr (!USE_SMEM) {
PNDE_UNROLL
      for (int f = 0; f < NF; ++f) {
PNDE_UNROLL
        for (int r = 0; r < DCOV; ++r) {
PNDE_UNROLL
          for (int c = 0; c <= r; ++c) Ls[f][SC::tri(r, c)] *= PIk[r / DC];
          if (!(Ls[f][SC::tri(r, r)] == Ls[f][SC::tri(r, r)])) status |= 1;
        }
      }
    }
PNDE_UNROLL
    for (int k = 0; k < D; ++k)
      if (!(ms[k] == ms[k])) status |= 1;
    }
    write(i);
  }
  if (ns >= 2) {
    // the first state is never smoothed (src/smoothing.jl:11: i runs down to 2)
    from_filtered(0);
    write(0);
  }
  sp.status[tid] = status;
}

}  // namespace pnde
