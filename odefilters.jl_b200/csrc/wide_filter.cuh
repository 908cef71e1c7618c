// Lane-group filter kernel for the dense EK1 model at D = d (q + 1) >= 10: G adjacent lanes of a warp own one
// trajectory.  Same algorithm and the same arithmetic, operation for operation, as DenseEK1 / cov_filter_step
// (cov_engine.cuh) -- one structured Householder triangularisation per attempted step in measurement-aligned
// coordinates -- but the D x D working matrix E is split by COLUMNS over the lanes:
//
//   lane g of a group owns the coordinates (k, a) of every derivative block k with dimension a = al * G + g,
//   i.e. CL = (q + 1) d / G mean entries, the same rows of the posterior factor S, and the columns
//   (y_a, x0_a, x2_a, ..., xq_a) of E.
//
// With that ownership the transition A = Atilde (x) I_d acts lane-locally; a Householder reflector is broadcast from
// the lane that owns its column with warp shuffles (<= D doubles per column) and applied by every lane to its own
// columns; the scalar chain (norm, rsqrt, beta, vector field, Jacobian, diffusion, error estimate, controller) is
// evaluated redundantly by all lanes of the group on identical inputs, so they stay in lockstep without any
// synchronisation.  A thread then holds D x CL instead of D x D doubles of E (72 instead of 144 at q = 5, d = 2),
// which is what does not fit one thread's 255 registers.
//
// Reference path: src/perform_step.jl:27-158, src/filtering.jl:22-48,79-91, src/diffusions.jl:72-80.
#pragma once
#include "filter_kernel.cuh"

namespace pnde {

// G adjacent lanes; every collective below is executed by the WHOLE warp in lockstep (full mask, sub-group width G)
template <int G>
struct LaneGroup {
  int g;  // this lane's index in its group
  __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(0xffffffffu, v, src, G); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

template <class VF_, int q_, int G_>
struct WideEK1 {
  using VF = VF_;
  using Base = DenseEK1<VF_, q_>;
  using Fac = Factor<VF_::d, q_>;
  static constexpr int d = VF::d, q = q_, D = d * (q + 1), ND = 1, G = G_;
  static_assert(d % G == 0, "lanes own whole dimensions");
  static constexpr int DL = d / G;         // dimensions per lane
  static constexpr int CL = (q + 1) * DL;  // coordinates per lane; local slot s = k * DL + al <-> (k, a = al * G + g)
  static constexpr int R = D - d;          // factor columns
  static constexpr int NZ = D - 2 * d;
  static constexpr bool IS_EK1 = true;
  static constexpr int REC = Base::REC;    // history records keep the single-thread layout (smoother, getters)

  // block-level structure of the posterior factor S = [W | Lz]: row block k, column r
  __host__ __device__ static constexpr bool nzb(int k, int r) { return r < d || (k >= 2 && (r - d) < (k - 1) * d); }
  __host__ __device__ static constexpr int count_state() {
    int c = CL;
    for (int k = 0; k <= q; ++k)
      for (int r = 0; r < R; ++r) c += nzb(k, r) ? DL : 0;
    return c;
  }
  static constexpr int STATE_LEN = count_state();  // doubles per LANE in the shared-memory stash
  static constexpr int SCR = D * R + q + 1;        // doubles per GROUP for final_cov

  struct State {
    double m[CL];
    double S[CL][R];  // rows of the factor for the own coordinates; entries with !nzb are never touched
  };

  __device__ __forceinline__ static void zero(State& s) {
#pragma unroll
    for (int i = 0; i < CL; ++i)
#pragma unroll
      for (int r = 0; r < R; ++r) s.S[i][r] = 0.0;
  }
  // full natural-order vector -> the own slice
  __device__ __forceinline__ static void take_mean(State& s, const LaneGroup<G>& gp, const double (&full)[D]) {
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        double v = full[k * d + al * G];
#pragma unroll
        for (int gg = 1; gg < G; ++gg) v = (gp.g == gg) ? full[k * d + al * G + gg] : v;
        s.m[k * DL + al] = v;
      }
  }
  __device__ __forceinline__ static void scale(State& s, const double (&sc)[q + 1]) {
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        s.m[k * DL + al] *= sc[k];
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (nzb(k, r)) s.S[k * DL + al][r] *= sc[k];
      }
  }
  // pre-step state in shared memory ([element][thread], conflict free), own layout
  __device__ __forceinline__ static void stash_store(const State& s, double* base, int stride) {
    int o = 0;
#pragma unroll
    for (int i = 0; i < CL; ++i) base[(o++) * stride] = s.m[i];
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al)
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (nzb(k, r)) base[(o++) * stride] = s.S[k * DL + al][r];
  }
  __device__ __forceinline__ static void stash_load(State& s, const double* base, int stride) {
    int o = 0;
#pragma unroll
    for (int i = 0; i < CL; ++i) s.m[i] = base[(o++) * stride];
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al)
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (nzb(k, r)) s.S[k * DL + al][r] = base[(o++) * stride];
  }
  // history record in the layout of DenseEK1::store: m[D], W[d][D], Lz packed; sc[k] scales block k on the way out
  __device__ __forceinline__ static void store(const State& s, const LaneGroup<G>& gp, const double (&sc)[q + 1],
                                               double* base, long long stride) {
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        const int a = al * G + gp.g;
        const int i = k * d + a;
        base[(long long)i * stride] = s.m[k * DL + al] * sc[k];
#pragma unroll
        for (int c = 0; c < d; ++c) base[(long long)(D + c * D + i) * stride] = s.S[k * DL + al][c] * sc[k];
        if (k >= 2) {
          const int il = (k - 2) * d + a;  // row among the triangular columns
#pragma unroll
          for (int j = 0; j < NZ; ++j)
            if (nzb(k, d + j) && j <= il)
              base[(long long)(D + d * D + Fac::lz(j, il)) * stride] = s.S[k * DL + al][d + j] * sc[k];
        }
      }
  }
  __device__ __forceinline__ static void write_mean(const State& s, const LaneGroup<G>& gp, const double (&sc)[q + 1],
                                                    double* mean, long long stride) {
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al) mean[(long long)(k * d + al * G + gp.g) * stride] = s.m[k * DL + al] * sc[k];
  }
  // Sigma = diag(sc) S S' diag(sc), packed lower.  The rows of S live in different lanes: they meet in shared memory
  // (scr: SCR = D * R + q + 1 doubles of this group) and the lanes split the D (D + 1) / 2 entries.  Once per trajectory.
  __device__ static void final_cov(const State& s, const LaneGroup<G>& gp, const double (&sc)[q + 1], double* scr,
                                   double* cov, long long stride, bool write) {
#pragma unroll
    for (int k = 0; k <= q; ++k)
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        const int i = k * d + al * G + gp.g;
#pragma unroll
        for (int r = 0; r < R; ++r) scr[i * R + r] = nzb(k, r) ? s.S[k * DL + al][r] : 0.0;
      }
#pragma unroll
    for (int k = 0; k <= q; ++k) scr[D * R + k] = sc[k];  // every lane writes the same values
    gp.sync();
    for (int e = gp.g; e < D * (D + 1) / 2; e += G) {
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      const int j = e - i * (i + 1) / 2;
      double acc = 0.0;
      for (int r = 0; r < R; ++r) acc = fma(scr[i * R + r], scr[j * R + r], acc);
      if (write) cov[(long long)e * stride] = acc * scr[D * R + i / d] * scr[D * R + j / d];
    }
    gp.sync();
  }

  __device__ __forceinline__ static void gather_block(const LaneGroup<G>& gp, const double (&loc)[CL], int k,
                                                      double (&out)[d]) {
#pragma unroll
    for (int a = 0; a < d; ++a) out[a] = gp.bcast(loc[k * DL + a / G], a % G);
  }

  // One attempted step in P(h) coordinates; every scalar output is identical in all lanes of the group.
  __device__ __forceinline__ static void step(State& s, const LaneGroup<G>& gp, const double* p, double pi0,
                                              double pi1, double ipi1, int diffusion, const IwpConsts& C,
                                              double (&u_new)[d], double (&err)[d], double (&local)[ND], double& quad,
                                              double& detS) {
    const int g = gp.g;
    apply_A<DL, q>(s.m);  // predict_mean!  (lane-local: A couples only the blocks of one dimension)
    double m0[d], m1[d];
    gather_block(gp, s.m, 0, m0);
    gather_block(gp, s.m, 1, m1);
    double uhat[d], fu[d], J[d][d], Jp[d][d], z[d];
#pragma unroll
    for (int i = 0; i < d; ++i) uhat[i] = pi0 * m0[i];
    VF::template f<double>(uhat, p, fu);
    VF::jac(uhat, p, J);
#pragma unroll
    for (int i = 0; i < d; ++i) {
      z[i] = fma(pi1, m1[i], -fu[i]);
#pragma unroll
      for (int j = 0; j < d; ++j) Jp[i][j] = pi0 * J[i][j];
    }
    double B[d][d];
#pragma unroll
    for (int i = 0; i < d; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < d; ++k) acc = fma(Jp[i][k], Jp[j][k], acc);
        acc *= C.Qt[0][0];
        acc = fma(-pi1 * C.Qt[0][1], Jp[i][j] + Jp[j][i], acc);
        if (i == j) acc = fma(pi1 * pi1, C.Qt[1][1], acc);
        B[i][j] = acc;
        B[j][i] = acc;
      }
    }
    double sig = 1.0;
    if (diffusion == DIFF_DYNAMIC) {
      double Lb[d][d], yb[d];
      double ss = 0.0;
#pragma unroll
      for (int j = 0; j < d; ++j) {
        double djj = B[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) djj = fma(-Lb[j][k], Lb[j][k], djj);
        const double il = (djj > 0.0) ? fast_rsqrt(djj) : 0.0;
        Lb[j][j] = djj * il;
#pragma unroll
        for (int i = j + 1; i < d; ++i) {
          double v = B[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) v = fma(-Lb[i][k], Lb[j][k], v);
          Lb[i][j] = v * il;
        }
        double yy = z[j];
#pragma unroll
        for (int k = 0; k < j; ++k) yy = fma(-Lb[j][k], yb[k], yy);
        yb[j] = yy * il;
        ss = fma(yb[j], yb[j], ss);
      }
      local[0] = ss * (1.0 / double(d));
      sig = (local[0] > 0.0) ? local[0] * fast_rsqrt(local[0]) : 0.0;
    }
    // rows of Jp for the own dimensions
    double JpO[DL][d];
#pragma unroll
    for (int al = 0; al < DL; ++al)
#pragma unroll
      for (int b = 0; b < d; ++b) {
        double v = Jp[al * G][b];
#pragma unroll
        for (int gg = 1; gg < G; ++gg) v = (g == gg) ? Jp[al * G + gg][b] : v;
        JpO[al][b] = v;
      }

    // ---- covariance: build E (own columns), sweep, read the update off R ----
    // sig * Ltilde[k][kk]: formed where it is used (one multiplication by a kernel-parameter constant; every entry is
    // used once per lane and step) instead of (q + 1)(q + 2) / 2 doubles that stay live across the whole sweep
    auto sL = [&](int k, int kk) { return sig * C.Lt[k][kk]; };

    double E[D][CL];  // local column slot ps = kk * DL + al: kk = 0 -> y_a, kk = 1 -> x0_a, kk >= 2 -> x_kk,a
    // rows 0..d-1: the prior rows of block 0
#pragma unroll
    for (int i = 0; i < d; ++i) {
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        const bool own_i = (al * G + g == i);
        double v = own_i ? pi1 * sL(1, 0) : 0.0;
        v = fma(-sL(0, 0), JpO[al][i], v);
        E[i][0 * DL + al] = v;
        E[i][1 * DL + al] = own_i ? sL(0, 0) : 0.0;
#pragma unroll
        for (int kk = 2; kk <= q; ++kk) E[i][kk * DL + al] = own_i ? sL(kk, 0) : 0.0;
      }
    }
    // rows d..D-1: (T A s)' for every factor column s
#pragma unroll
    for (int c = 0; c < R; ++c) {
      const int kb = (c < d) ? 0 : 2 + (c - d) / d;  // first structurally non-zero block of this column
      double w[CL];
#pragma unroll
      for (int k = 0; k <= q; ++k) {
#pragma unroll
        for (int al = 0; al < DL; ++al) {
          bool any = (k >= kb);
          double acc = any ? s.S[k * DL + al][c] : 0.0;
#pragma unroll
          for (int j = k + 1; j <= q; ++j) {
            if (j >= kb) {
              const double cf = inv_factorial(j - k);
              const double x = s.S[j * DL + al][c];
              if (!any) {
                acc = (j - k == 1) ? x : cf * x;
                any = true;
              } else {
                acc = (j - k == 1) ? acc + x : fma(cf, x, acc);
              }
            }
          }
          w[k * DL + al] = acc;
        }
      }
      double w0[d];
      gather_block(gp, w, 0, w0);
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        double y = pi1 * w[1 * DL + al];
#pragma unroll
        for (int bb = 0; bb < d; ++bb) y = fma(-JpO[al][bb], w0[bb], y);
        E[d + c][0 * DL + al] = y;
        E[d + c][1 * DL + al] = w[0 * DL + al];
#pragma unroll
        for (int kk = 2; kk <= q; ++kk) E[d + c][kk * DL + al] = w[kk * DL + al];
      }
    }

    double RtopL[d][CL], Rinv[d];
    // Householder sweep over the primed columns c = kc * d + ac (owner lane ac % G, its slot kc * DL + ac / G)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const int first = (c < d) ? 0 : (c < 2 * d ? c - d + 1 : d);
      const int kc = c / d, ac = c % d, oc = ac % G, psc = kc * DL + ac / G;
      double v[D];
#pragma unroll
      for (int i = first; i < D; ++i) v[i] = gp.bcast(E[i][psc], oc);
      double pv;
      if (c < d)
        pv = pi1 * sL(1, 1);
      else if (c < 2 * d)
        pv = gp.bcast(E[c - d][psc], oc);
      else
        pv = sL(kc, kc);
      double n2a = pv * pv, n2b = v[first] * v[first];
#pragma unroll
      for (int i = first + 1; i < D; ++i) {
        if ((i - first) & 1)
          n2a = fma(v[i], v[i], n2a);
        else
          n2b = fma(v[i], v[i], n2b);
      }
      const double n2 = n2a + n2b;
      const bool nzcol = n2 > 0.0;
      const double rn = nzcol ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double snrm = copysign(nrm, pv);
      const double v0 = pv + snrm;
      const double beta = nzcol ? fast_rcp(fma(fabs(pv), nrm, n2)) : 0.0;
      if (c < d) Rinv[c] = -copysign(rn, pv);
      // own columns j = kk * d + a behind c
#pragma unroll
      for (int kk = kc; kk <= q; ++kk) {
#pragma unroll
        for (int al = 0; al < DL; ++al) {
          const int ps = kk * DL + al;
          if (kk == kc && al * G + G - 1 < ac) {  // this slot is in front of the pivot in every lane: R[c][j] = 0
            if (c < d) {
              RtopL[c][ps] = 0.0;
            } else if (c < 2 * d) {
              s.S[0 * DL + al][c - d] = 0.0;  // (kk == kc == 1)
            } else if (nzb(kk, c - d)) {
              s.S[kk * DL + al][c - d] = 0.0;
            }
            continue;
          }
          // j > c ?  (decided at compile time except inside the pivot's own block)
          const bool behind = (kk > kc) || (al > ac / G) || (al == ac / G && g > oc);
          const bool is_c = (kk == kc) && (al == ac / G) && (g == oc);
          const bool same_dim = (al == ac / G) && (g == oc);  // a == ac
          bool pnz;
          double prj = 0.0;
          if (c < d) {
            pnz = (kk >= 2) && same_dim;
            if (kk >= 2) prj = sL(kk, 1);
          } else if (c < 2 * d) {
            pnz = true;
            prj = E[c - d][ps];
          } else {
            pnz = same_dim;
            prj = sL(kk, kc);
          }
          // The pivot-row entry is selected to zero where the prior row has none: w then starts from v0 * 0 = 0 and
          // fma(a, b, 0) == a * b, fma(-s, v0, 0) == -s * v0 -- bit for bit the two code paths of cov_filter_step,
          // without evaluating both.  Likewise a lane that has no column behind the pivot in this slot applies the
          // reflector with s = 0 (E is unchanged exactly): no divergent branch inside the sweep, so ptxas schedules
          // the whole column step as one block and overlaps the next column's norm chain with these updates.
          const bool never = (c < d) && (kk < 2);  // no prior-row entry in any lane (compile time)
          const double prj_e = pnz ? prj : 0.0;
          double w = never ? v[first] * E[first][ps] : fma(v[first], E[first][ps], v0 * prj_e);
#pragma unroll
          for (int i = first + 1; i < D; ++i) w = fma(v[i], E[i][ps], w);
          const double sc = beta * w;
          double rr = never ? -sc * v0 : fma(-sc, v0, prj_e);  // R[c][j]
          const double sce = behind ? sc : 0.0;
#pragma unroll
          for (int i = first; i < D; ++i) E[i][ps] = fma(-sce, v[i], E[i][ps]);
          rr = is_c ? -snrm : (behind ? rr : 0.0);
          // consume the finished entry of R
          if (c < d) {
            RtopL[c][ps] = rr;
          } else if (c < 2 * d) {
            // column a_w = c - d of W: x0 entries (primed block 1) are the block-0 rows, blocks >= 2 as they are
            if (kk == 1) s.S[0 * DL + al][c - d] = rr;
            if (kk >= 2) s.S[kk * DL + al][c - d] = rr;
          } else {
            if (nzb(kk, c - d)) s.S[kk * DL + al][c - d] = rr;
          }
        }
      }
      if (c == d - 1) {
        // the first d rows of R are complete: innovation and mean update now, so that Rtop, z, y are dead for the
        // rest of the sweep (same operations as after the sweep in the one-thread kernel)
        // ---- innovation: S_z = G G', G = Rtop[:, :d]' lower triangular; y = G^-1 z ----
        double Ry[d][d];  // Rtop[b][a], a, b < d (the y columns), gathered
#pragma unroll
        for (int b = 0; b < d; ++b) {
          double row[CL];
#pragma unroll
          for (int i = 0; i < CL; ++i) row[i] = RtopL[b][i];
          gather_block(gp, row, 0, Ry[b]);
        }
        double y[d];
        double yy2 = 0.0, dets = 1.0;
#pragma unroll
        for (int a = 0; a < d; ++a) {
          double acc = z[a];
#pragma unroll
          for (int b = 0; b < a; ++b) acc = fma(-Ry[b][a], y[b], acc);
          y[a] = acc * Rinv[a];
          yy2 = fma(y[a], y[a], yy2);
          dets *= fabs(Ry[a][a]);
        }
        quad = yy2;
        detS = dets;
        if (diffusion != DIFF_DYNAMIC) local[0] = yy2 * (1.0 / double(d));
        // mean update (lane-local), block 1 from the linearised vector field
#pragma unroll
        for (int al = 0; al < DL; ++al) {
          double acc = s.m[0 * DL + al];
#pragma unroll
          for (int a = 0; a < d; ++a) acc = fma(-RtopL[a][1 * DL + al], y[a], acc);
          s.m[0 * DL + al] = acc;
#pragma unroll
          for (int kk = 2; kk <= q; ++kk) {
            double acc2 = s.m[kk * DL + al];
#pragma unroll
            for (int a = 0; a < d; ++a) acc2 = fma(-RtopL[a][kk * DL + al], y[a], acc2);
            s.m[kk * DL + al] = acc2;
          }
        }
        double m0new[d];
        gather_block(gp, s.m, 0, m0new);
#pragma unroll
        for (int al = 0; al < DL; ++al) {
          double fo = fu[al * G];
#pragma unroll
          for (int gg = 1; gg < G; ++gg) fo = (g == gg) ? fu[al * G + gg] : fo;
          double acc = fo;
#pragma unroll
          for (int bb = 0; bb < d; ++bb) acc = fma(JpO[al][bb], m0new[bb] - m0[bb], acc);
          s.m[1 * DL + al] = acc * ipi1;
        }
#pragma unroll
        for (int i = 0; i < d; ++i) {
          u_new[i] = pi0 * m0new[i];
          err[i] = sqrt(local[0] * B[i][i]);
        }
      }
    }
    // block 1 of the posterior factor is slaved to block 0: x_1 = (Jp x_0) / pi1  (H S+ = 0)
#pragma unroll
    for (int a = 0; a < d; ++a) {
      double col0[CL], W0[d];
#pragma unroll
      for (int i = 0; i < CL; ++i) col0[i] = s.S[i][a];
      gather_block(gp, col0, 0, W0);
#pragma unroll
      for (int al = 0; al < DL; ++al) {
        double vv = 0.0;
#pragma unroll
        for (int bb = 0; bb < d; ++bb) vv = fma(JpO[al][bb], W0[bb], vv);
        s.S[1 * DL + al][a] = vv * ipi1;
      }
    }

  }
};

// ---------------------------------------------------------------------------------------------
// The kernel: the control flow of filter_kernel (filter_kernel.cuh), G lanes per trajectory.
//
// The time loop is WARP-UNIFORM: every lane iterates until the slowest trajectory of its warp is done, a finished
// group keeps executing the loop body with its bookkeeping switched off (`alive`) -- in SIMT lockstep that costs
// nothing, the warp runs until its slowest trajectory anyway.  This is what lets every shuffle use the full-warp mask:
// a partial-mask __shfl_sync compiles to WARPSYNC + a convergence-barrier region per shuffle (measured on the first
// version of this kernel: 324 WARPSYNC / 388 BSSY-BSYNC pairs per step, and slower than the one-thread kernel it was
// meant to replace); with the full mask it is a bare SHFL.IDX.
// The accepted state also lives in shared memory ([element][thread]): a rejected step and a finished group both
// simply reload it.
// ---------------------------------------------------------------------------------------------
#ifndef PNDE_WIDE_BLOCK
#define PNDE_WIDE_BLOCK 128
#endif
template <class M, bool ADAPTIVE>
__global__ void __launch_bounds__(PNDE_WIDE_BLOCK, 1) wide_filter_kernel(const FilterParams prm) {
  using VF = typename M::VF;
  constexpr int d = M::d, q = M::q, D = M::D, ND = M::ND, REC = M::REC, G = M::G;
  const long long gth = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lid_ = gth / G;
  const bool exists = lid_ < prm.count;  // lanes past the end of the ensemble shadow its last trajectory, silently
  LaneGroup<G> gp;
  gp.g = (int)(threadIdx.x % G);
  const bool lead = gp.g == 0;
  const long long tid = prm.first + (exists ? lid_ : prm.count - 1);
  const long long n = prm.n;
  const CtrlParams& K = prm.K;
  const int diffusion = prm.diffusion;
  const bool is_static = (diffusion == DIFF_FIXED || diffusion == DIFF_FIXED_MAP);

  double p[VF::np], u0[d];
#pragma unroll
  for (int i = 0; i < VF::np; ++i) p[i] = prm.p[(long long)i * n + tid];
#pragma unroll
  for (int i = 0; i < d; ++i) u0[i] = prm.u0[(long long)i * n + tid];

  // shared memory: [STATE_LEN][blockDim] accepted state (ADAPTIVE only), then the [blockDim / G][SCR] scratch of
  // final_cov
  extern __shared__ double wsm[];
  typename M::State st;
  {
    double full[D];
    taylor_init<VF, q>(u0, p, full);
    M::take_mean(st, gp, full);
  }
  M::zero(st);

  double t = K.t0;
  int iter = 0, nacc = 0, nrej = 0, nfe = 0, ret = RET_SUCCESS, nsaved = 0;
  double gsaved[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) gsaved[i] = 1.0;
  double uprev[d];
#pragma unroll
  for (int i = 0; i < d; ++i) uprev[i] = u0[i];
  double ll_quad = 0.0, ll_log = 0.0, ll_mant = 1.0;
  long long ll_exp = 0;
  int ll_n = 0;
  double one[q + 1];
#pragma unroll
  for (int k = 0; k <= q; ++k) one[k] = 1.0;
  bool stopped = !exists;  // left the loop through a `break` of the one-thread kernel (ret != success), or shadow lane

  auto save = [&](const typename M::State& sv, const double (&sc)[q + 1], double tt, const double (&g)[ND]) {
    if (nsaved >= prm.max_saved) {
      ret = RET_HISTORY_FULL;  // adaptive runs keep stepping without saving: naccept is the capacity the run needs
      if (!ADAPTIVE) stopped = true;
      return;
    }
    double* base = prm.hist + ((long long)nsaved * REC) * n + tid;
    if (lead) {
      base[0] = tt;
#pragma unroll
      for (int i = 0; i < ND; ++i) base[(long long)(1 + i) * n] = g[i];
    }
    M::store(sv, gp, sc, base + (long long)(1 + ND) * n, n);
    ++nsaved;
  };
  if (prm.save_mode != SAVE_FINAL && exists) save(st, one, t, gsaved);

  double dt;
  if (ADAPTIVE) {
    if (K.dt > 0.0) {
      dt = K.dt;
    } else {
      dt = initdt<VF, q>(u0, p, K);
      nfe += 2;
    }
  } else {
    dt = K.dt;
  }
  const double lqold0 = ctrl_state_init(K);
  double dtpropose = dt, lqold = lqold0, q11 = 1.0;
  bool accepted_prev = true;
  double hcur = -1.0;
  double Pk[q + 1], PIk[q + 1];
#pragma unroll
  for (int k = 0; k <= q; ++k) Pk[k] = PIk[k] = 1.0;
  if (ADAPTIVE) M::stash_store(st, wsm + threadIdx.x, blockDim.x);

  while (__any_sync(0xffffffffu, !stopped && t < K.t1)) {
    bool alive = !stopped && t < K.t1;  // this group really takes a step in this iteration
    // ---- loopheader! ----
    double dtn = dt;
    if (iter > 0) {
      if (accepted_prev)
        dtn = dtpropose;
      else
        dtn = dt / fmin(1.0 / K.qmin, q11 / K.gamma);
    }
    if (alive) {
      ++iter;
      if (iter > K.maxiters) {
        ret = RET_MAXITERS;
        stopped = true;
        alive = false;
      }
    }
    if (ADAPTIVE) {
      dtn = fmin(dtn, K.dtmax);
      dtn = fmax(dtn, K.dtmin);
      dtn = fmin(dtn, K.t1 - t);
    } else {
      dtn = fmin(K.dt, K.t1 - t);
    }
    if (alive) {
      dt = dtn;
      if (dt != dt) {
        ret = RET_DTNAN;
        stopped = true;
        alive = false;
      } else if (ADAPTIVE && iter > 1 && !accepted_prev && fabs(dt) <= fabs(K.dtmin)) {
        ret = RET_DTMIN;
        stopped = true;
        alive = false;
      }
    }
    // a group that is not alive runs the step on a harmless step size and throws the result away
    const double h = alive ? dt : (K.t1 - K.t0);
    // ---- perform_step! ----
    if (ADAPTIVE) {
      precond_scales<q>(h, Pk, PIk);
      M::scale(st, Pk);
    } else if (h != hcur) {
      double Pn[q + 1], PIn[q + 1], sc[q + 1];
      precond_scales<q>(h, Pn, PIn);
#pragma unroll
      for (int k = 0; k <= q; ++k) {
        sc[k] = Pn[k] * PIk[k];
        Pk[k] = Pn[k];
        PIk[k] = PIn[k];
      }
      M::scale(st, sc);
      hcur = h;
    }
    double unew[d], err[d], local[ND], quad, detS;
#pragma unroll
    for (int i = 0; i < ND; ++i) local[i] = 1.0;
    M::step(st, gp, p, PIk[0], PIk[1], Pk[1], diffusion, prm.C, unew, err, local, quad, detS);
    double gcur[ND];
    if (diffusion == DIFF_DYNAMIC) {
      gcur[0] = local[0];
    } else if (diffusion == DIFF_FIXED) {
      gcur[0] = (nacc == 0) ? local[0] : gsaved[0] + (local[0] - gsaved[0]) / double(nacc);
    } else {
      const double Nn = double(nacc + 1), al = 0.5, be = 0.5;
      if (nacc == 0) {
        gcur[0] = (be + 0.5 * local[0]) / (al + Nn * d / 2.0 + 1.0);
      } else {
        const double res_prev = (gsaved[0] * (al + (Nn - 1.0) * d / 2.0 + 1.0) - be) * 2.0;
        gcur[0] = (be + 0.5 * (res_prev + local[0])) / (al + Nn * d / 2.0 + 1.0);
      }
    }
    double EEst = 0.0;
    bool finite = true;
    if (ADAPTIVE) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < d; ++i) {
        const double r = dt * err[i] / (K.abstol + fmax(fabs(uprev[i]), fabs(unew[i])) * K.reltol);
        acc = fma(r, r, acc);
      }
      EEst = sqrt(acc / double(d));
    }
#pragma unroll
    for (int i = 0; i < d; ++i) finite = finite && (fabs(unew[i]) <= 1.79769313486231570e308);
    const bool commit = alive && (!ADAPTIVE || (EEst < 1.0));
    const bool accept = !ADAPTIVE || (EEst <= 1.0);
    if (ADAPTIVE) {
      if (commit) {
        M::scale(st, PIk);
        M::stash_store(st, wsm + threadIdx.x, blockDim.x);
      } else {
        M::stash_load(st, wsm + threadIdx.x, blockDim.x);  // rejected, or not alive: back to the accepted state
      }
    }
    if (alive) {
      ++nfe;
#pragma unroll
      for (int i = 0; i < d; ++i) uprev[i] = unew[i];  // integ.u .= u_filt, even when rejected (:86)
      if (commit) {
        ll_quad += quad;
        ++ll_n;
        if (detS > 1e-290 && detS < 1e290) {
          const long long bits = __double_as_longlong(detS);
          ll_exp += ((bits >> 52) & 0x7ff) - 1023;
          ll_mant *= __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3ff0000000000000LL);
          if (ll_mant > 1e250) {
            ll_log += log(ll_mant);
            ll_mant = 1.0;
          }
        } else {
          ll_log += log(detS);
        }
      }
      if (!finite) {
        ret = RET_NONFINITE;
        stopped = true;
      } else {
        // ---- loopfooter! ----
        const double ttmp = t + dt;
        if (ADAPTIVE) {
          double lE;
          double qc = controller_factor(EEst, K, lqold, lE);
          if (accept) {
            ++nacc;
            if (K.qsteady_min <= qc && qc <= K.qsteady_max) qc = 1.0;
            lqold = ctrl_state_accept(EEst, lE, lqold0, K);
            const double dtnew = dt / qc;
            t = (fabs(ttmp - K.t1) < 10.0 * ulp_of(fmax(t, K.t1))) ? K.t1 : ttmp;
            dtpropose = fmax(K.dtmin, fmin(K.dtmax, dtnew));
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
#pragma unroll
          for (int i = 0; i < ND; ++i) gsaved[i] = gcur[i];
          const bool want = (prm.save_mode == SAVE_EVERY) ||
                            (prm.save_mode == SAVE_STRIDE && (nacc % prm.save_stride == 0 || !(t < K.t1)));
          if (want) {
            if (ADAPTIVE)
              save(st, one, t, gsaved);
            else
              save(st, PIk, t, gsaved);
          }
        }
      }
    }
  }

  // ---- outputs (the warp is converged here) ----
  double sc[q + 1];
#pragma unroll
  for (int k = 0; k <= q; ++k) sc[k] = (ADAPTIVE || hcur < 0.0) ? 1.0 : PIk[k];
  if (prm.mean && exists) M::write_mean(st, gp, sc, prm.mean + tid, n);
  double ll = -0.5 * (ll_quad + 2.0 * (ll_log + log(ll_mant) + double(ll_exp) * 0.6931471805599453) +
                      double(ll_n) * double(d) * 1.8378770664093453);
  double cal = 1.0;
  if (is_static && nacc > 0) {
    cal = gsaved[0];
    ll = nan("");
  }
  if (prm.cov) {
    double sc2[q + 1];
    const double gq = sqrt(cal);
#pragma unroll
    for (int k = 0; k <= q; ++k) sc2[k] = sc[k] * gq;
    double* scr = wsm + (ADAPTIVE ? (size_t)M::STATE_LEN * blockDim.x : 0) + (size_t)(threadIdx.x / G) * M::SCR;
    M::final_cov(st, gp, sc2, scr, prm.cov + tid, n, exists);
  }
  if (lead && exists) {
    if (prm.final_diff) prm.final_diff[tid] = gsaved[0];
    if (prm.t_final) prm.t_final[tid] = t;
    if (prm.loglik) prm.loglik[tid] = ll;
    prm.retcode[tid] = ret;
    prm.naccept[tid] = nacc;
    prm.nreject[tid] = nrej;
    prm.nf[tid] = nfe;
    prm.njacs[tid] = nfe - ((ADAPTIVE && !(K.dt > 0.0)) ? 2 : 0);
    prm.n_saved[tid] = nsaved;
  }
}

}  // namespace pnde
