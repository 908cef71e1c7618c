// Lane-group RTS smoother for the dense EK1 model at D = d (q + 1) >= 10: FOUR adjacent lanes of a warp own one
// trajectory.  Same algorithm and the same arithmetic, operation for operation, as smoother_kernel / SmoothCov
// (smoother_kernel.cuh: stage-1 sweep, gain products without forming G, final triangularisation), but every matrix
// is split by COLUMNS (= state coordinates) over the lanes:
//
//   lane g = ga + 2 gk owns the coordinates (k, a) with a = 2 al + ga (dimension parity) and k = 2 kl + gk (derivative
//   block parity): CL = ceil((q + 1) / 2) d / 2 columns of the stage-1 stack [El | Er], of R-, X, Y, T' and the same
//   ROWS of the smoothed factor L^s.
//
// A Householder reflector is broadcast from the lane that owns its column (warp shuffles) and applied by every lane to
// its own columns; the forward substitutions with R- run row by row, each finished row of Z = R-^-T L^s broadcast once.
// One thread then holds ~1/4 of the ~540 doubles a D = 12 smoother step works on (the one-thread kernel spills 4-8 KB
// per thread there: 79 M steps/s at q = 5); X, R- and Y wait in shared memory between the stages.
// The interval loop is warp-uniform (full-mask shuffles, see wide_filter.cuh).
//
// Reference path: smooth_all! / smooth!  src/smoothing.jl:4-63.
#pragma once
#include "smoother_kernel.cuh"
#include "wide_filter.cuh"

namespace pnde {

#ifndef PNDE_WSMOOTH_BLOCK
#define PNDE_WSMOOTH_BLOCK 128
#endif

template <class VF_, int q_>
struct WideSmooth {
  using M = DenseEK1<VF_, q_>;
  using SC = SmoothCov<VF_::d, q_>;
  using Fac = Factor<VF_::d, q_>;
  static constexpr int d = M::d, q = q_, D = M::D, R = D - d, NZ = D - 2 * d, NP = SC::NP, G = 4;
  static_assert(d % 2 == 0, "lanes own dimensions of one parity");
  static constexpr int DL = d / 2, KL = (q + 2) / 2, CL = KL * DL;
  static constexpr int ST = PNDE_WSMOOTH_BLOCK;                 // stride of the [element][thread] shared-memory layout
  static constexpr int SM_X = 0, SM_R = D * CL, SM_Y = 2 * D * CL;  // element offsets of X, R-, Y
  static constexpr int SM_LEN = 2 * D * CL + R * CL;            // doubles per lane

  struct Lane {
    int ga, gk;
    // block, dimension and natural index j = k d + a of slot s; ok: k <= q (q even: the odd-block lanes have a phantom slot)
    __device__ __forceinline__ int k(int s) const { return 2 * (s / DL) + gk; }
    __device__ __forceinline__ int a(int s) const { return 2 * (s % DL) + ga; }
    __device__ __forceinline__ int j(int s) const { return k(s) * d + a(s); }
    __device__ __forceinline__ bool ok(int s) const { return k(s) <= q; }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(0xffffffffu, v, src, G); }
  };
  // largest natural index a slot can have (odd-block lane, last dimension): rows of L^s / columns of R- of slot s have
  // entries c <= jmax(s) at most -- everything beyond is never touched, so it costs no register
  __host__ __device__ static constexpr int jmax(int s) { return (2 * (s / DL) + 1) * d + d - 1 < D - 1 ? (2 * (s / DL) + 1) * d + d - 1 : D - 1; }
  __device__ __forceinline__ static constexpr int owner(int c) { return ((c % d) & 1) + 2 * ((c / d) & 1); }
  __device__ __forceinline__ static constexpr int slot_of(int c) { return ((c / d) >> 1) * DL + ((c % d) >> 1); }

  // Column r of the filtered factor S = [W | Lz], entries of dimension `a` in all blocks, from a history record
  // (layout of DenseEK1::store behind the mean: W[d][D], Lz packed)
  __device__ __forceinline__ static void load_factor_column(const double* fac, long long n, int r, int a, bool on,
                                                            double (&col)[q + 1]) {
#pragma unroll
    for (int kk = 0; kk <= q; ++kk) {
      double v = 0.0;
      if (r < d) {
        if (on) v = fac[(long long)(r * D + kk * d + a) * n];
      } else if (kk >= 2) {
        const int jj = r - d, il = (kk - 2) * d + a;
        if (on && il >= jj) v = fac[(long long)(d * D + Fac::lz(jj, il)) * n];
      }
      col[kk] = v;
    }
  }

  // Householder triangularisation of the stack [Y ; T'] (R + D rows, own columns) -> rows of the lower factor L.
  // Arithmetic of SmoothCov::triangularize_impl.
  __device__ __forceinline__ static void triangularize(const Lane& ln, double (&Y)[R][CL], double (&Tt)[D][CL],
                                                       double (&L)[CL][D], bool on, int& status) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const int oc = owner(c), psc = slot_of(c);
      double v[R + D];
#pragma unroll
      for (int i = c; i < R + D; ++i) v[i] = ln.bcast(i < R ? Y[i < R ? i : 0][psc] : Tt[i - R >= 0 ? i - R : 0][psc], oc);
      double n2 = 0.0;
#pragma unroll
      for (int i = c; i < R + D; ++i) n2 = fma(v[i], v[i], n2);
      const double pv = v[c];
      const bool nz = n2 > 0.0;
      const double rn = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double snrm = copysign(nrm, pv);
      const double v0 = pv + snrm;
      const double beta = nz ? fast_rcp(fma(fabs(pv), nrm, n2)) : 0.0;
      if (on && !(n2 == n2)) status |= 1;
#pragma unroll
      for (int s = 0; s < CL; ++s) {
        if (jmax(s) < c) continue;  // every column of this slot lies before the pivot (compile time)
        const bool behind = ln.ok(s) && ln.j(s) > c;
        const bool is_c = ln.ok(s) && ln.j(s) == c;
        const double prj = (c < R) ? Y[c < R ? c : 0][s] : Tt[c - R >= 0 ? c - R : 0][s];
        double w = v0 * prj;
#pragma unroll
        for (int i = c + 1; i < R + D; ++i) w = fma(v[i], i < R ? Y[i < R ? i : 0][s] : Tt[i - R >= 0 ? i - R : 0][s], w);
        const double sc = beta * w;
        const double rr = fma(-sc, v0, prj);
        const double sce = behind ? sc : 0.0;
#pragma unroll
        for (int i = c + 1; i < R + D; ++i) {
          if (i < R)
            Y[i < R ? i : 0][s] = fma(-sce, v[i], Y[i < R ? i : 0][s]);
          else
            Tt[i - R >= 0 ? i - R : 0][s] = fma(-sce, v[i], Tt[i - R >= 0 ? i - R : 0][s]);
        }
        L[s][c] = (on && is_c) ? -snrm : ((on && behind) ? rr : L[s][c]);
      }
    }
  }

  // smoothed record (layout of SmoothModel<DenseEK1>: mean[D], L packed lower by rows): the own rows
  __device__ __forceinline__ static void write_record(const Lane& ln, double* o, long long n, bool on,
                                                      const double (&ms)[CL], const double (&Ls)[CL][D]) {
    if (!on) return;
#pragma unroll
    for (int s = 0; s < CL; ++s) {
      if (!ln.ok(s)) continue;
      const int j = ln.j(s);
      o[(long long)j * n] = ms[s];
#pragma unroll
      for (int c = 0; c < D; ++c)
        if (c <= jmax(s) && c <= j) o[(long long)(D + j * (j + 1) / 2 + c) * n] = Ls[s][c];
    }
  }
  // x_smooth = x_filt as a triangular factor (last state; the un-smoothed first state).  r0: the record behind t and
  // the diffusion.
  __device__ __forceinline__ static void from_filtered(const Lane& ln, const double* r0, long long n, bool on,
                                                       double dense_cal, double (&ms)[CL], double (&Ls)[CL][D],
                                                       int& status) {
    double Y[R][CL], Tt[D][CL];
#pragma unroll
    for (int s = 0; s < CL; ++s) {
      const bool ld = on && ln.ok(s);
      const double v = ld ? r0[(long long)ln.j(s) * n] : 0.0;
      ms[s] = on ? v : ms[s];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        double col[q + 1];
        load_factor_column(r0 + (long long)D * n, n, r, 2 * al + ln.ga, on, col);
#pragma unroll
        for (int kl = 0; kl < KL; ++kl) {
          const double v = ln.gk ? (2 * kl + 1 <= q ? col[2 * kl + 1 <= q ? 2 * kl + 1 : 0] : 0.0) : col[2 * kl];
          Y[r][kl * DL + al] = v * dense_cal;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int s = 0; s < CL; ++s) Tt[c][s] = 0.0;
    triangularize(ln, Y, Tt, Ls, on, status);
  }
};

template <class VF, int q_>
__global__ void __launch_bounds__(PNDE_WSMOOTH_BLOCK, 1) wide_smoother_kernel(const SmoothParams sp) {
  using W = WideSmooth<VF, q_>;
  using M = typename W::M;
  using SC = typename W::SC;
  constexpr int d = W::d, q = W::q, D = W::D, R = W::R, G = W::G, CL = W::CL, DL = W::DL, KL = W::KL, ST = W::ST;
  constexpr int REC = M::REC, SREC = SmoothModel<M>::SREC;
  extern __shared__ double wsm[];
  if (blockDim.x != ST) __trap();
  double* sm = wsm + threadIdx.x;  // element e of this lane at sm[e * ST]
  const long long gth = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lid = gth / G;
  const bool exists = lid < sp.n;
  const long long tid = exists ? lid : sp.n - 1;
  const long long n = sp.n;
  typename W::Lane ln;
  {
    const int g = (int)(threadIdx.x % G);
    ln.ga = g & 1;
    ln.gk = g >> 1;
  }
  const int ns = exists ? sp.n_saved[tid] : 0;
  int status = 0;
  const double gfin = sp.calibrate ? sp.final_diff[tid] : 1.0;
  const double dense_cal = sp.calibrate ? sqrt(gfin) : 1.0;
  auto rec = [&](int slot) { return sp.hist + ((long long)slot * REC) * n + tid; };
  auto srec = [&](int slot) { return sp.smooth + ((long long)slot * SREC) * n + tid; };

  double ms[CL];     // smoothed mean at i + 1, own coordinates (natural)
  double Ls[CL][D];  // rows of the smoothed factor at i + 1 (natural coordinates); entries c <= j
#pragma unroll
  for (int s = 0; s < CL; ++s) {
    ms[s] = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c)
      if (c <= W::jmax(s)) Ls[s][c] = 0.0;
  }
  // the whole warp walks its trajectories backwards together (full-mask shuffles); a group whose trajectory is
  // shorter idles with its stores and loads switched off
  W::from_filtered(ln, rec(ns >= 1 ? ns - 1 : 0) + (long long)2 * n, n, ns >= 1, dense_cal, ms, Ls, status);
  W::write_record(ln, srec(ns >= 1 ? ns - 1 : 0), n, ns >= 1, ms, Ls);
  int i = ns - 2;
  while (__any_sync(0xffffffffu, i >= 1)) {
    const bool alive = i >= 1;
    const int ii = alive ? i : 0;
    const double* ri = rec(ii);
    const double* rn = rec(alive ? ii + 1 : 0);
    const double h = alive ? rn[0] - ri[0] : 1.0;
    // h == 0 (or a sliver, filter_kernel.cuh): the state is kept as it is (src/smoothing.jl:13-16)
    const bool work = alive && !sliver_interval(h, ri[0], rn[0], ns, sp.calibrate);
    double Pk[q + 1], PIk[q + 1];
    precond_scales<q>(work ? h : 1.0, Pk, PIk);
    const double g = sp.calibrate ? gfin : (alive ? rn[(long long)n] : 1.0);
    const double sig = sqrt(g);
    auto sL = [&](int kk, int kc) { return sig * sp.C.Lt[kk][kc]; };

    // ---- filtered mean (P coordinates) and its prediction for the own dimensions ----
    double mP[CL], mpred[CL];
#pragma unroll
    for (int al = 0; al < DL; ++al) {
      double col[q + 1];
#pragma unroll
      for (int kk = 0; kk <= q; ++kk) col[kk] = work ? ri[(long long)(2 + kk * d + 2 * al + ln.ga) * n] * Pk[kk] : 0.0;
      double pr[q + 1];
#pragma unroll
      for (int kk = 0; kk <= q; ++kk) pr[kk] = col[kk];
      apply_A<1, q>(pr);
#pragma unroll
      for (int kl = 0; kl < KL; ++kl) {
        const int k1 = 2 * kl + 1 <= q ? 2 * kl + 1 : 0;
        mP[kl * DL + al] = ln.gk ? (2 * kl + 1 <= q ? col[k1] : 0.0) : col[2 * kl];
        mpred[kl * DL + al] = ln.gk ? (2 * kl + 1 <= q ? pr[k1] : 0.0) : pr[2 * kl];
      }
    }
    // ---- stage 1: sweep over [sig Q_L' | 0 ; (A S)' | S'] ----
    double El[R][CL], Er[R][CL];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        double col[q + 1];
        W::load_factor_column(ri + (long long)(2 + D) * n, n, r, 2 * al + ln.ga, work, col);
#pragma unroll
        for (int kk = 0; kk <= q; ++kk) col[kk] = (col[kk] * Pk[kk]) * dense_cal;  // x = P x, then the calibration
        double pr[q + 1];
#pragma unroll
        for (int kk = 0; kk <= q; ++kk) pr[kk] = col[kk];
        apply_A<1, q>(pr);
#pragma unroll
        for (int kl = 0; kl < KL; ++kl) {
          const int k1 = 2 * kl + 1 <= q ? 2 * kl + 1 : 0;
          Er[r][kl * DL + al] = ln.gk ? (2 * kl + 1 <= q ? col[k1] : 0.0) : col[2 * kl];
          El[r][kl * DL + al] = ln.gk ? (2 * kl + 1 <= q ? pr[k1] : 0.0) : pr[2 * kl];
        }
      }
    }
    double rinvL[CL];
#pragma unroll
    for (int s = 0; s < CL; ++s) rinvL[s] = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const int kc = c / d, ac = c % d, oc = W::owner(c), psc = W::slot_of(c);
      double v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = ln.bcast(El[r][psc], oc);
      const double pv = sL(kc, kc);
      double n2 = pv * pv;
#pragma unroll
      for (int r = 0; r < R; ++r) n2 = fma(v[r], v[r], n2);
      const bool nz = n2 > 0.0;
      const double rn_ = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn_;
      const double v0 = pv + nrm;
      const double beta = nz ? fast_rcp(fma(pv, nrm, n2)) : 0.0;
      // left block: own columns behind c; R-[c][j] goes to shared memory
#pragma unroll
      for (int s = 0; s < CL; ++s) {
        if (W::jmax(s) < c) continue;  // in front of the pivot in both parities: R-[c][j] is never read
        const bool behind = ln.ok(s) && ln.j(s) > c;
        const bool is_c = ln.ok(s) && ln.j(s) == c;
        const bool pnz = ln.ok(s) && ln.a(s) == ac;
        // sig Ltilde[k][kc] for the block of this slot in this lane's parity
        const int kl = s / DL;
        const double lt = ln.gk ? (2 * kl + 1 <= q ? sp.C.Lt[2 * kl + 1 <= q ? 2 * kl + 1 : 0][kc] : 0.0) : sp.C.Lt[2 * kl][kc];
        const double prj = pnz ? sig * lt : 0.0;
        double w = v0 * prj;
#pragma unroll
        for (int r = 0; r < R; ++r) w = fma(v[r], El[r][s], w);
        const double sc = beta * w;
        const double rr = fma(-sc, v0, prj);
        const double sce = behind ? sc : 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) El[r][s] = fma(-sce, v[r], El[r][s]);
        sm[(W::SM_R + c * CL + s) * ST] = is_c ? -nrm : (behind ? rr : 0.0);
        rinvL[s] = is_c ? -rn_ : rinvL[s];
      }
      // right block: all own columns; X[c][j] goes to shared memory
#pragma unroll
      for (int s = 0; s < CL; ++s) {
        double w = v[0] * Er[0][s];
#pragma unroll
        for (int r = 1; r < R; ++r) w = fma(v[r], Er[r][s], w);
        const double sc = beta * w;
        sm[(W::SM_X + c * CL + s) * ST] = -sc * v0;
#pragma unroll
        for (int r = 0; r < R; ++r) Er[r][s] = fma(-sc, v[r], Er[r][s]);
      }
    }
    // Y = Er waits in shared memory for the last stage
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int s = 0; s < CL; ++s) sm[(W::SM_Y + r * CL + s) * ST] = Er[r][s];

    // ---- stage 2: y = R-^-T delta, Z = R-^-T (P L^s) row by row; delta <- X' y, T' = (X' Z)' accumulated on the fly ----
    double accY[CL], dn[CL], Tt[D][CL];
#pragma unroll
    for (int s = 0; s < CL; ++s) dn[s] = 0.0;
#pragma unroll
    for (int s = 0; s < CL; ++s) {
      const int kl = s / DL;
      const double pk = ln.gk ? Pk[2 * kl + 1 <= q ? 2 * kl + 1 : 0] : Pk[2 * kl];
      accY[s] = fma(pk, ms[s], -mpred[s]);  // delta = P m^s_{i+1} - A P m_i
#pragma unroll
      for (int c = 0; c < D; ++c)
        if (c <= W::jmax(s)) Ls[s][c] *= pk;  // L^s into P(h) coordinates
    }
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int s = 0; s < CL; ++s) Tt[c][s] = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r) {  // row r of R-^-T: coordinate r
      const int orr = W::owner(r), psr = W::slot_of(r);
      const double ri_ = rinvL[psr];
      const double yv = ln.bcast(accY[psr] * ri_, orr);
      double Zr[D];
#pragma unroll
      for (int c = 0; c <= r; ++c) Zr[c] = ln.bcast(Ls[psr][c <= W::jmax(psr) ? c : 0] * ri_, orr);
#pragma unroll
      for (int s = 0; s < CL; ++s) {
        const double x = sm[(W::SM_X + r * CL + s) * ST];
        dn[s] = fma(x, yv, dn[s]);
#pragma unroll
        for (int c = 0; c <= r; ++c) Tt[c][s] = (c == r) ? x * Zr[c] : fma(x, Zr[c], Tt[c][s]);
        if (W::jmax(s) <= r) continue;  // no row of this slot lies behind r
        const bool after = work && ln.ok(s) && ln.j(s) > r;  // an idle group leaves its L^s untouched (rm = 0)
        const double rm = after ? sm[(W::SM_R + r * CL + s) * ST] : 0.0;  // R-[r][j]
        accY[s] = fma(-rm, yv, accY[s]);
#pragma unroll
        for (int c = 0; c <= r; ++c) Ls[s][c] = fma(-rm, Zr[c], Ls[s][c]);
      }
    }
    // ---- stage 3: triangularise [Y ; T'] -> L^s_i ----
    double Y[R][CL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int s = 0; s < CL; ++s) Y[r][s] = sm[(W::SM_Y + r * CL + s) * ST];
    // (L^s of a working group is dead here -- it served as the accumulator of Z --, an idle group's is kept)
    W::triangularize(ln, Y, Tt, Ls, work, status);
    // back to natural coordinates; a group without work keeps its state
#pragma unroll
    for (int s = 0; s < CL; ++s) {
      const int kl = s / DL;
      const double pik = ln.gk ? PIk[2 * kl + 1 <= q ? 2 * kl + 1 : 0] : PIk[2 * kl];
      const double msn = (mP[s] + dn[s]) * pik;
      if (work) {  // (an idle group scaled its L^s by P = 1 and accumulated with R- = 0: it is unchanged)
        ms[s] = msn;
        if (ln.ok(s) && !(msn == msn)) status |= 1;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          if (c > W::jmax(s)) continue;
          Ls[s][c] *= pik;
          if (!(Ls[s][c] == Ls[s][c])) status |= 1;  // any NaN in the row (src/smoothing.jl:25,59 test the diagonal)
        }
      }
    }
    W::write_record(ln, srec(ii), n, alive, ms, Ls);
    --i;
  }
  // the first state is never smoothed (src/smoothing.jl:11: i runs down to 2)
  W::from_filtered(ln, rec(0) + (long long)2 * n, n, ns >= 2, dense_cal, ms, Ls, status);
  W::write_record(ln, srec(0), n, ns >= 2, ms, Ls);
  // the status flags of the four lanes are combined by the lead lane
  {
    int sall = status;
    sall |= __shfl_xor_sync(0xffffffffu, sall, 1, G);
    sall |= __shfl_xor_sync(0xffffffffu, sall, 2, G);
    if (exists && (threadIdx.x % G) == 0) sp.status[tid] = sall;
  }
}

}  // namespace pnde
