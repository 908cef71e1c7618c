// EK1 with the full covariance at large D = d(q+1) (BASELINE config 4, Lorenz-96 d = 1024, D = 4096):
// the same measurement-aligned Householder triangularisation as cov_engine.cuh, but as a BLOCKED
// (compact-WY) QR whose trailing updates are FP64 tensor-core GEMMs (mma.sync m8n8k4 -> SASS DMMA).
//
// Reference path: perform_step! src/perform_step.jl:27-93 with alg isa EK1 (predict_cov!
// src/filtering.jl:33-48, measure! :95-132, DynamicDiffusion src/diffusions.jl:72-80, update!
// src/filtering.jl:79-91) -- the reference does all of it with dense LAPACK on D x D matrices.
//
// Working matrices (row major, leading dimension ld = D):
//   E [D][D]  dense rows of the stack in primed coordinates x' = (y, x0, x2, .., xq):
//             rows 0..d-1  = the prior rows of block 0 (the only prior rows that are not sparse pivots),
//             rows d..D-1  = (T A s)' for the D-d columns s of the posterior factor.
//   R [D][D]  upper triangular result; rows 0..d-1 carry innovation factor + gain, rows d.. are the
//             new posterior factor.
//   The other prior rows are sparse "pivot rows" generated on the fly (prior_pivot): reflector c is
//   [pivot(c,c) ; E[:,c]], so V = [diag(v0) ; Vb] and the compact-WY T needs only Vb'Vb.
// One panel of NB = 32 columns: (1) cluster panel factorisation (8 CTAs, entries in registers, pivot column in
// shared memory, warp-shuffle + DSMEM reductions), preceded inside the same kernel by the previous panel's block
// reflector applied to these 32 columns (look-ahead), (2) W = Vb' E_trail (DMMA, split-K atomics),
// (3) W2 = T'(W + pivots), R rows, (4) E_trail -= Vb W2 (DMMA).  (2)-(4) cover the columns behind the next panel
// and run on the main stream while the next panel is factored on a high-priority auxiliary stream.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "cov_engine.cuh"

namespace pnde {
namespace big {

namespace cg = cooperative_groups;

constexpr int NB = 32;        // panel width
#ifndef PNDE_PANEL_CLUSTER
#define PNDE_PANEL_CLUSTER 8
#endif
constexpr int NCLUSTER = PNDE_PANEL_CLUSTER;  // CTAs per panel cluster (16 needs the non-portable opt-in)
constexpr int PANEL_THREADS = 256;
constexpr int CPW = NB / (PANEL_THREADS / 32);  // panel columns per warp
static_assert(CPW == 4, "the panel kernel loads and reduces its columns four at a time");

struct Geometry {
  int d, q, D;
};

// scalars that live on the device for the whole solve (no host round trips inside a step)
struct Scalars {
  double sigma;        // sqrt of the diffusion used in the predict of this step
  double local, global_saved;
  double quad, logdet;
  double ll_quad, ll_logdet;
  double pi0, pi1, ipi1, h;
  int nacc, nonfinite, ll_n, pad;
};

// entry (c, j) of the sparse prior pivot rows, WITHOUT the sigma factor (see cov_engine.cuh)
__device__ __forceinline__ double prior_pivot(int c, int j, int d, double pi1, const IwpConsts& C, int mode) {
  if (mode == 0) return 0.0;  // plain QR (zero pivots)
  if (c < d) {
    if (j == c) return pi1 * C.Lt[1][1];
    if (j >= 2 * d && (j % d) == c) return C.Lt[j / d][1];
    return 0.0;
  }
  if (c < 2 * d) return 0.0;
  const int k = c / d, a = c % d;
  if ((j % d) == a && j / d >= k) return C.Lt[j / d][k];
  return 0.0;
}

// the same with (block, dimension) = (index / d, index % d) of both indices supplied by the caller
__device__ __forceinline__ double prior_pivot_split(int c, int cd, int ca, int j, int jd, int ja, int d, double pi1,
                                                    const IwpConsts& C, int mode) {
  (void)d;
  if (mode == 0) return 0.0;
  if (cd == 0) {
    if (j == c) return pi1 * C.Lt[1][1];
    if (jd >= 2 && ja == c) return C.Lt[jd][1];
    return 0.0;
  }
  if (cd == 1) return 0.0;
  if (ja == ca && jd >= cd) return C.Lt[jd][cd];
  return 0.0;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// (1) panel factorisation: columns [c0, c0+NB) of E (nrows x .), cluster of NCLUSTER CTAs
// outputs: E panel <- Vb;  R[c0+k][c0+j] (k <= j < NB);  v0[NB], T[NB][NB] (compact WY)
// ---------------------------------------------------------------------------------------------
// Explicit 128-bit shared-window accesses for the hot loops of the panel kernel (with cluster launch the compiler
// addresses `extern __shared__` through the cluster window and may rebuild the window base in front of an access).
__device__ __forceinline__ void lds128(unsigned addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts128(unsigned addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

constexpr int panel_rs(int rpt) { return 32 * rpt + 2; }  // +2: the 32 columns start in different banks
inline size_t panel_smem_bytes(int rpt) { return ((size_t)NB * panel_rs(rpt) + 2 * NB + 3 * NB * NB) * sizeof(double); }

// Sum over the warp of 4 values per lane with 6 shuffles: lanes 8c .. 8c+7 return the total of v[c].
__device__ __forceinline__ double warp_reduce4(const double (&v)[4], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
  const double u0 = (b4 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, b4 ? v[0] : v[2], 16);
  const double u1 = (b4 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, b4 ? v[1] : v[3], 16);
  double t = (b3 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, b3 ? u0 : u1, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 4);
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

struct PanelArgs {
  double* E;
  double* R;
  int ld, nrows, c0, d, mode;
  const Scalars* sc;
  double* v0;    // [NB]
  double* T;     // [NB][NB]
  // previous panel (columns [cprev, cprev+NB), cprev < 0: none): its reflectors are applied to this panel's
  // columns inside the kernel before the factorisation starts (the "narrow" trailing update of a look-ahead QR)
  int cprev;
  const double* v0p;
  const double* Tp;
  IwpConsts C;
};

// Mapping: 8 warps, warp w owns the CPW = 4 panel columns [4w, 4w+4); the lanes of a warp are rows: lane l holds the
// row pairs (64 p + 2 l, 64 p + 2 l + 1), p < RPT/2, of its warp's columns in registers (creg) for the whole kernel.
// Shared memory holds a column-major copy slabT[k][r] (a lane's row pair of one column is one 128-bit access):
// the previous panel's reflectors during the fused update, afterwards the panel itself, of which only the current
// pivot column has to be up to date (its owner warp publishes it right after its last update).  A column lives in
// one warp, so inner products are warp-shuffle reductions (no block reduction), and every pivot value read from
// shared memory feeds 4 columns.
template <int RPT>
__global__ void __cluster_dims__(NCLUSTER, 1, 1) __launch_bounds__(PANEL_THREADS, 1) panel_kernel(const PanelArgs a) {
  static_assert(RPT % 2 == 0, "rows per lane are processed in pairs");
  extern __shared__ double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int rows_per = a.nrows / NCLUSTER;
  const int row0 = rank * rows_per;
  constexpr int RS = panel_rs(RPT);  // row stride of slabT (rows >= rows_per are zero padding)
  constexpr int NP = RPT / 2;
  double* slab = smem;                       // [NB][RS]
  double* gl = slab + (size_t)NB * RS;       // [2][NB] this CTA's inner products of the current column (double buffered)
  double* gram = gl + 2 * NB;                // [NB][NB] Gram of Vb (cluster total, rank 0)
  double* wloc = gram + NB * NB;             // [NB][NB] this CTA's share of a block product
  double* wful = wloc + NB * NB;             // [NB][NB] cluster total, then W2
  __shared__ double s_v0[NB], s_beta[NB], s_T[NB][NB + 1], s_coef[NB];
  __shared__ int s_cd[NB], s_ca[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wc0 = warp * CPW;  // first column of this warp
  const double sigma = (a.mode == 0) ? 1.0 : a.sc->sigma;
  const double pi1 = a.sc->pi1;
  const bool fused = a.cprev >= 0;
  // byte address (shared window) of this lane's first row pair in column 0 of slabT
  const unsigned lbase = (unsigned)__cvta_generic_to_shared(slab) + (unsigned)(2 * lane) * 8u;

  auto stage = [&](int cfirst) {  // slabT <- columns [cfirst, cfirst + NB) of this CTA's rows of E (coalesced reads)
    for (int idx = tid; idx < 32 * RPT * NB; idx += PANEL_THREADS) {
      const int r = idx / NB, k = idx % NB;
      slab[k * RS + r] = (r < rows_per) ? a.E[(size_t)(row0 + r) * a.ld + cfirst + k] : 0.0;
    }
  };
  // this thread's entries straight from global memory (4 consecutive doubles per row: one full sector per lane)
  // while the block stages what the sweep reads from shared memory first: Vp (fused) or the panel itself
  double creg[CPW][RPT];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = 64 * p + 2 * lane + h;
      if (r < rows_per) {
        const double2* src = reinterpret_cast<const double2*>(a.E + (size_t)(row0 + r) * a.ld + a.c0 + wc0);
        const double2 v01 = src[0], v23 = src[1];
        creg[0][2 * p + h] = v01.x;
        creg[1][2 * p + h] = v01.y;
        creg[2][2 * p + h] = v23.x;
        creg[3][2 * p + h] = v23.y;
      } else {
        creg[0][2 * p + h] = creg[1][2 * p + h] = creg[2][2 * p + h] = creg[3][2 * p + h] = 0.0;
      }
    }
  stage(fused ? a.cprev : a.c0);
  if (tid < NB) {  // (block, dimension) of the panel's columns for the sparse prior pivots
    s_cd[tid] = (a.c0 + tid) / a.d;
    s_ca[tid] = (a.c0 + tid) % a.d;
  }
  if (fused) {
    for (int e = tid; e < NB * NB; e += PANEL_THREADS) s_T[e / NB][e % NB] = a.Tp[e];
    if (tid < NB) s_v0[tid] = a.v0p[tid];
  }
  __syncthreads();

  // wloc[k][wc0 + c] = sum over this CTA's rows of slabT[k][r] * creg[c][r]   (k = 0..NB-1)
  auto block_product = [&]() {
    for (int k = 0; k < NB; ++k) {
      double v[CPW] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        double x, y;
        lds128(lbase + (unsigned)(k * RS + 64 * p) * 8u, x, y);
#pragma unroll
        for (int j = 0; j < CPW; ++j) v[j] = fma(y, creg[j][2 * p + 1], fma(x, creg[j][2 * p], v[j]));
      }
      const double t = warp_reduce4(v, lane);
      if ((lane & 7) == 0) wloc[k * NB + wc0 + (lane >> 3)] = t;
    }
    __syncthreads();
  };

  if (fused) {
    // ---- apply the previous panel's block reflector to these NB columns (C in creg, Vp in slabT) ----
    //   W = Vp' C (cluster-wide sum);  W2 = T' (W + diag(v0) pivots);  R rows;  C -= Vp W2
    block_product();
    cluster.sync();
    constexpr int EPT = NB * NB / PANEL_THREADS;  // entries of an NB x NB block per thread
    double w2r[EPT], pivr[EPT];
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      const int e = tid + m * PANEL_THREADS, k = e / NB, c = e % NB;
      double w = 0.0;
#pragma unroll
      for (int rk = 0; rk < NCLUSTER; ++rk) w += cluster.map_shared_rank(wloc, rk)[e];
      pivr[m] = sigma * prior_pivot(a.cprev + k, a.c0 + c, a.d, pi1, a.C, a.mode);
      wful[e] = fma(s_v0[k], pivr[m], w);
    }
    __syncthreads();
    {
      // W2[k][c] = sum_{mm <= k} T[mm][k] (W + ...)[mm][c]  (T' is lower triangular).  A thread's EPT entries share
      // the column c (PANEL_THREADS % NB == 0): one pass over mm feeds all of them
      const int c = tid % NB, kg = tid / NB;
#pragma unroll
      for (int m = 0; m < EPT; ++m) w2r[m] = 0.0;
      for (int mm = 0; mm < NB; ++mm) {
        const double w = wful[mm * NB + c];
#pragma unroll
        for (int m = 0; m < EPT; ++m) {
          const int k = kg + m * (PANEL_THREADS / NB);
          if (mm <= k) w2r[m] = fma(s_T[mm][k], w, w2r[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int k = kg + m * (PANEL_THREADS / NB);
        if (rank == 0) a.R[(size_t)(a.cprev + k) * a.ld + a.c0 + c] = fma(-s_v0[k], w2r[m], pivr[m]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < EPT; ++m) wful[tid + m * PANEL_THREADS] = w2r[m];
    __syncthreads();
    for (int k = 0; k < NB; ++k) {
      double w2[CPW];
#pragma unroll
      for (int j = 0; j < CPW; ++j) w2[j] = wful[k * NB + wc0 + j];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        double x, y;
        lds128(lbase + (unsigned)(k * RS + 64 * p) * 8u, x, y);
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          creg[j][2 * p] = fma(-x, w2[j], creg[j][2 * p]);
          creg[j][2 * p + 1] = fma(-y, w2[j], creg[j][2 * p + 1]);
        }
      }
    }
    __syncthreads();  // every warp is done with Vp: slabT <- C
#pragma unroll
    for (int j = 0; j < CPW; ++j)
#pragma unroll
      for (int p = 0; p < NP; ++p) sts128(lbase + (unsigned)((wc0 + j) * RS + 64 * p) * 8u, creg[j][2 * p], creg[j][2 * p + 1]);
    __syncthreads();
  }
  // ---- factorisation.  Per column: inner products with the pivot column (warp reductions) -> one value per column
  // published to the cluster -> warp 0 gathers the ranks through DSMEM, computes the reflector scalars s_j -> update.
  const int jd = (a.c0 + lane) / a.d, ja = (a.c0 + lane) % a.d;  // used by warp 0 (lane == panel column)
  for (int k = 0; k < NB; ++k) {
    double* glk = gl + (k & 1) * NB;
    if (wc0 + CPW - 1 >= k) {
      double v[CPW] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        double x, y;
        lds128(lbase + (unsigned)(k * RS + 64 * p) * 8u, x, y);
#pragma unroll
        for (int j = 0; j < CPW; ++j) v[j] = fma(y, creg[j][2 * p + 1], fma(x, creg[j][2 * p], v[j]));
      }
      const double t = warp_reduce4(v, lane);
      if ((lane & 7) == 0) glk[wc0 + (lane >> 3)] = t;
    }
    cluster.sync();
    if (tid < NB) {
      double g = 0.0;
#pragma unroll
      for (int rk = 0; rk < NCLUSTER; ++rk) g += cluster.map_shared_rank(glk, rk)[tid];
      const double gk = __shfl_sync(0xffffffffu, g, k);
      const int c = a.c0 + k;
      const int cd = s_cd[k], ca = s_ca[k];
      const double pv = sigma * prior_pivot_split(c, cd, ca, c, cd, ca, a.d, pi1, a.C, a.mode);
      const double n2 = fma(pv, pv, gk);
      const bool nz = n2 > 0.0;
      const double rn = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double v0 = pv + nrm;
      const double beta = nz ? fast_rcp(fma(pv, nrm, n2)) : 0.0;
      double sj = 0.0;
      if (tid > k) {
        const double prj = sigma * prior_pivot_split(c, cd, ca, a.c0 + tid, jd, ja, a.d, pi1, a.C, a.mode);
        sj = beta * fma(v0, prj, g);
        if (rank == 0) a.R[(size_t)c * a.ld + a.c0 + tid] = fma(-sj, v0, prj);
      }
      s_coef[tid] = sj;
      if (tid == 0) {
        s_v0[k] = v0;
        s_beta[k] = beta;
        if (rank == 0) a.R[(size_t)c * a.ld + c] = -nrm;
      }
    }
    __syncthreads();
    if (wc0 + CPW - 1 > k) {
      double sj[CPW];
#pragma unroll
      for (int j = 0; j < CPW; ++j) sj[j] = s_coef[wc0 + j];  // zero for columns <= k
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        double x, y;
        lds128(lbase + (unsigned)(k * RS + 64 * p) * 8u, x, y);
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          creg[j][2 * p] = fma(-sj[j], x, creg[j][2 * p]);
          creg[j][2 * p + 1] = fma(-sj[j], y, creg[j][2 * p + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        if (wc0 + j == k + 1) {  // the next pivot column: publish it for the other warps
#pragma unroll
          for (int p = 0; p < NP; ++p) sts128(lbase + (unsigned)((k + 1) * RS + 64 * p) * 8u, creg[j][2 * p], creg[j][2 * p + 1]);
        }
      }
    }
    __syncthreads();
  }
  // every column of slabT is in its final state (= Vb): back to E, then the Gram of Vb as one more block product
  for (int idx = tid; idx < rows_per * NB; idx += PANEL_THREADS) {
    const int r = idx / NB, k = idx % NB;
    a.E[(size_t)(row0 + r) * a.ld + a.c0 + k] = slab[k * RS + r];
  }
  block_product();
  cluster.sync();
  if (rank == 0) {
    for (int e = tid; e < NB * NB; e += PANEL_THREADS) {
      double sg = 0.0;
#pragma unroll
      for (int rk = 0; rk < NCLUSTER; ++rk) sg += cluster.map_shared_rank(wloc, rk)[e];
      gram[e] = sg;
    }
    __syncthreads();
    // T = U^-1 with U = striu(Vb'Vb) + diag(1/beta) (compact WY).  From T U = I, row i of T only depends on itself:
    //   T[i][i] = beta_i,  T[i][k] = -beta_k sum_{m=i}^{k-1} T[i][m] G[m][k]   -- one thread per row, no barriers
    // (the row lives in registers: fully unrolled, entries left of the diagonal are zero and drop out by themselves)
    if (tid < NB) {
      const int i = tid;
      double Ti[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < k; ++m) acc = fma(Ti[m], gram[m * NB + k], acc);  // gram[m][k]: one address per warp
        Ti[k] = (k < i) ? 0.0 : (k == i ? s_beta[k] : -s_beta[k] * acc);
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) s_T[i][k] = Ti[k];
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += PANEL_THREADS) a.T[e] = s_T[e / NB][e % NB];
    if (tid < NB) a.v0[tid] = s_v0[tid];
  }
  cluster.sync();  // keep every CTA's shared memory alive until rank 0 has read the Gram pieces
}

// ---------------------------------------------------------------------------------------------
// (2) C[m][n] += sum_i A[i][m] * B[i][n]   (A: K x M, B: K x N row major; M multiple of 32)
// CTA tile 32 (m) x 64 (n), K split over blockIdx.z, DMMA, results added with FP64 atomics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) atb_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B,
                                                  int ldb, double* __restrict__ Cm, int ldc, int N, int K, int kchunk) {
  constexpr int LDA_S = 36, LDB_S = 68;
  __shared__ double As[32 * LDA_S], Bs[32 * LDB_S];
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[4][2][2];  // [m tile][n tile within the warp's 16 columns][2]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // software pipeline: the next 32-row slice travels from global memory into registers while the tensor cores
  // work on the current one (8 + 16 doubles per thread)
  double ra[8], rb[16];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = threadIdx.x + i * 128, r = idx / 32, c = idx % 32;
      ra[i] = (k0 + r < kend) ? A[(size_t)(k0 + r) * lda + m0 + c] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int idx = threadIdx.x + i * 128, r = idx / 64, c = idx % 64;
      rb[i] = (k0 + r < kend && n0 + c < N) ? B[(size_t)(k0 + r) * ldb + n0 + c] : 0.0;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = threadIdx.x + i * 128;
      As[(idx / 32) * LDA_S + idx % 32] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int idx = threadIdx.x + i * 128;
      Bs[(idx / 64) * LDB_S + idx % 64] = rb[i];
    }
    __syncthreads();
    if (k0 + 32 < kend) fetch(k0 + 32);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      double bf[2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) bf[nt] = Bs[(kk * 4 + (lane & 3)) * LDB_S + warp * 16 + nt * 8 + (lane >> 2)];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const double af = As[(kk * 4 + (lane & 3)) * LDA_S + mt * 8 + (lane >> 2)];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], af, bf[nt]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int m = m0 + mt * 8 + (lane >> 2);
      const int n = n0 + warp * 16 + nt * 8 + 2 * (lane & 3);
      if (n < N) atomicAdd(&Cm[(size_t)m * ldc + n], acc[mt][nt][0]);
      if (n + 1 < N) atomicAdd(&Cm[(size_t)m * ldc + n + 1], acc[mt][nt][1]);
    }
}

// ---------------------------------------------------------------------------------------------
// (4) Cm[i][n] -= sum_k A[i][k] * B[k][n]   (A: M x 32, B: 32 x N; rank-32 update, DMMA)
// CTA tile 64 (i) x 64 (n)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) rank_update_kernel(const double* __restrict__ A, int lda,
                                                             const double* __restrict__ B, int ldb,
                                                             double* __restrict__ Cm, int ldc, int M, int N) {
  // 8 warps, warp w owns rows [8w, 8w+8) of the 64 x 64 tile: 16 accumulators per thread keep the register count low
  // enough for 3-4 CTAs (24-32 warps) per SM -- the kernel is bound by the latency of its global loads
  constexpr int LDA_S = 36, LDB_S = 68;
  __shared__ double As[64 * LDA_S], Bs[32 * LDB_S];
  const int i0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {
    const int r = idx / 32, c = idx % 32;
    As[r * LDA_S + c] = (i0 + r < M) ? A[(size_t)(i0 + r) * lda + c] : 0.0;
  }
  for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
    const int r = idx / 64, c = idx % 64;
    Bs[r * LDB_S + c] = (n0 + c < N) ? B[(size_t)r * ldb + n0 + c] : 0.0;
  }
  // the C fragment is read up front (acc starts as C, the product is subtracted through -A), so that all
  // global loads of the tile are in flight together
  double acc[8][2];
  const int i = i0 + warp * 8 + (lane >> 2);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int n = n0 + nt * 8 + 2 * (lane & 3);
    acc[nt][0] = (i < M && n < N) ? Cm[(size_t)i * ldc + n] : 0.0;
    acc[nt][1] = (i < M && n + 1 < N) ? Cm[(size_t)i * ldc + n + 1] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const double af = -As[(warp * 8 + (lane >> 2)) * LDA_S + kk * 4 + (lane & 3)];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const double bf = Bs[(kk * 4 + (lane & 3)) * LDB_S + nt * 8 + (lane >> 2)];
      dmma(acc[nt][0], acc[nt][1], af, bf);
    }
  }
  if (i < M) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int n = n0 + nt * 8 + 2 * (lane & 3);
      if (n < N) Cm[(size_t)i * ldc + n] = acc[nt][0];
      if (n + 1 < N) Cm[(size_t)i * ldc + n + 1] = acc[nt][1];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (3) W2 = T' (W + diag(v0) pivots);  R[c0+k][j] = pivot(c0+k, j) - v0[k] W2[k][j];  W <- W2
// one thread per trailing column j
// ---------------------------------------------------------------------------------------------
struct W2Args {
  double* W;   // [NB][ldw], column j of the trailing block at W[k][j - (c0+NB)]
  int ldw;
  double* R;
  int ld, c0, ncols, d, mode;
  int jt0, jt1;  // trailing-column range [jt0, jt1) handled by this launch (jt = j - (c0 + NB))
  const double* v0;
  const double* T;
  const Scalars* sc;
  IwpConsts C;
};

__global__ void __launch_bounds__(256) w2_kernel(const W2Args a) {
  // block = 8 trailing columns x 32 panel rows: thread (k, column) computes W2[k][j]
  __shared__ double sT[NB][NB + 1], sv0[NB], sw[8][NB + 1];
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) sT[e / NB][e % NB] = a.T[e];
  if (threadIdx.x < NB) sv0[threadIdx.x] = a.v0[threadIdx.x];
  __syncthreads();
  const int k = threadIdx.x % NB, cj = threadIdx.x / NB;
  const int jt = a.jt0 + blockIdx.x * 8 + cj;
  const int j = a.c0 + NB + jt;
  const bool valid = (jt < a.jt1) && (j < a.ncols);
  const double sigma = (a.mode == 0) ? 1.0 : a.sc->sigma;
  const double pi1 = a.sc->pi1;
  double piv = 0.0;
  if (valid) {
    piv = sigma * prior_pivot(a.c0 + k, j, a.d, pi1, a.C, a.mode);
    sw[cj][k] = fma(sv0[k], piv, a.W[(size_t)k * a.ldw + jt]);
  }
  __syncthreads();
  if (!valid) return;
  double acc = 0.0;
  for (int m = 0; m <= k; ++m) acc = fma(sT[m][k], sw[cj][m], acc);  // T' is lower triangular
  a.W[(size_t)k * a.ldw + jt] = acc;
  a.R[(size_t)(a.c0 + k) * a.ld + j] = fma(-sv0[k], acc, piv);
}

// Blocked QR driver: E (nrows x ncols, ld) -> R rows [0, ncols) (upper triangular part written).
// nrows % NCLUSTER == 0 with nrows / NCLUSTER <= 640, ncols % NB == 0.  work: W [NB][ld], v0 [2][NB], T [2][NB*NB].
// Look-ahead: the panel kernel of panel j+1 (8 SMs, stream `aux`) applies panel j's reflectors to its own NB columns
// and overlaps the update of the remaining trailing columns with panel j's reflectors ("wide" update, main stream).
struct QrWork {
  double* W;
  double* v0;  // [2][NB]
  double* T;   // [2][NB*NB]
  cudaStream_t aux;
  cudaEvent_t ev_narrow[2], ev_panel[2];
};

inline cudaError_t blocked_qr(double* E, double* R, int ld, int nrows, int ncols, int d, int mode, const Scalars* sc,
                              const IwpConsts& C, const QrWork& wk, cudaStream_t s, long long* launches) {
  const int rows_per = nrows / NCLUSTER;
  const int rpt = rows_per <= 8 * 32 ? 8 : (rows_per <= 16 * 32 ? 16 : 20);
  if (rows_per > 20 * 32) return cudaErrorInvalidValue;
  const size_t smem = panel_smem_bytes(rpt);
  void (*kern)(const PanelArgs) = rpt == 8 ? panel_kernel<8> : (rpt == 16 ? panel_kernel<16> : panel_kernel<20>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (NCLUSTER > 8 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) return e;
  auto launch_panel = [&](int c0, int buf, cudaStream_t st, int cprev) {
    PanelArgs pa;
    pa.cprev = cprev;
    pa.v0p = wk.v0 + (1 - buf) * NB;
    pa.Tp = wk.T + (1 - buf) * NB * NB;
    pa.E = E;
    pa.R = R;
    pa.ld = ld;
    pa.nrows = nrows;
    pa.c0 = c0;
    pa.d = d;
    pa.mode = mode;
    pa.sc = sc;
    pa.v0 = wk.v0 + buf * NB;
    pa.T = wk.T + buf * NB * NB;
    pa.C = C;
    kern<<<NCLUSTER, PANEL_THREADS, smem, st>>>(pa);
    ++*launches;
  };
  // trailing update of the columns jt in [jt0, jt1) (relative to c0 + NB) with the reflectors of panel c0
  auto update = [&](int c0, int buf, int jt0, int jt1) -> cudaError_t {
    const int nc = jt1 - jt0;
    if (nc <= 0) return cudaSuccess;
    cudaError_t e2 = cudaMemset2DAsync(wk.W + jt0, (size_t)ld * sizeof(double), 0, (size_t)nc * sizeof(double), NB, s);
    if (e2 != cudaSuccess) return e2;
    static const int kchunk = getenv("PNDE_ATB_KCHUNK") ? atoi(getenv("PNDE_ATB_KCHUNK")) : 256;
    dim3 g1((nc + 63) / 64, 1, (nrows + kchunk - 1) / kchunk);
    atb_kernel<<<g1, 128, 0, s>>>(E + c0, ld, E + c0 + NB + jt0, ld, wk.W + jt0, ld, nc, nrows, kchunk);
    W2Args wa;
    wa.W = wk.W;
    wa.ldw = ld;
    wa.R = R;
    wa.ld = ld;
    wa.c0 = c0;
    wa.ncols = ncols;
    wa.d = d;
    wa.mode = mode;
    wa.jt0 = jt0;
    wa.jt1 = jt1;
    wa.v0 = wk.v0 + buf * NB;
    wa.T = wk.T + buf * NB * NB;
    wa.sc = sc;
    wa.C = C;
    w2_kernel<<<(nc + 7) / 8, 256, 0, s>>>(wa);
    dim3 g2((nc + 63) / 64, (nrows + 63) / 64);
    rank_update_kernel<<<g2, 256, 0, s>>>(E + c0, ld, wk.W + jt0, ld, E + c0 + NB + jt0, ld, nrows, nc);
    *launches += 3;
    return cudaSuccess;
  };
  // Look-ahead: the panel chain runs on `aux` (panel j+1 applies panel j's reflectors to its own NB columns
  // itself), the update of the remaining trailing columns with panel j's reflectors ("wide") on `s`:
  //   panel(j+1) needs panel(j) [stream order] and wide(j-1);   wide(j) needs panel(j) and wide(j-1) [stream order]
  if ((e = cudaEventRecord(wk.ev_narrow[0], s)) != cudaSuccess) return e;  // everything that built E
  if ((e = cudaStreamWaitEvent(wk.aux, wk.ev_narrow[0], 0)) != cudaSuccess) return e;
  launch_panel(0, 0, wk.aux, -1);
  if ((e = cudaEventRecord(wk.ev_panel[0], wk.aux)) != cudaSuccess) return e;
  int last = 0;
  for (int c0 = 0, j = 0; c0 < ncols; c0 += NB, ++j) {
    const int ntrail = ncols - c0 - NB;
    if (ntrail <= 0) break;
    const int par = j & 1;
    if (j >= 1 && (e = cudaStreamWaitEvent(wk.aux, wk.ev_narrow[(j - 1) & 1], 0)) != cudaSuccess) return e;
    launch_panel(c0 + NB, 1 - par, wk.aux, c0);
    if ((e = cudaEventRecord(wk.ev_panel[1 - par], wk.aux)) != cudaSuccess) return e;
    last = 1 - par;
    if ((e = cudaStreamWaitEvent(s, wk.ev_panel[par], 0)) != cudaSuccess) return e;
    e = update(c0, par, NB, ntrail);  // wide: everything behind the next panel
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(wk.ev_narrow[par], s)) != cudaSuccess) return e;  // "wide(j) done"
  }
  if ((e = cudaStreamWaitEvent(s, wk.ev_panel[last], 0)) != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace big
}  // namespace pnde
