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
// One panel of NB = 32 columns: (1) cluster panel factorisation (8 CTAs, slab in shared memory,
// DSMEM reductions), (2) W = Vb' E_trail (DMMA, split-K atomics), (3) W2 = T'(W + pivots), R rows,
// (4) E_trail -= Vb W2 (DMMA).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "cov_engine.cuh"

namespace pnde {
namespace big {

namespace cg = cooperative_groups;

constexpr int NB = 32;        // panel width
constexpr int NCLUSTER = 8;   // CTAs per panel cluster
constexpr int PANEL_THREADS = 1024;

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

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// (1) panel factorisation: columns [c0, c0+NB) of E (nrows x .), cluster of NCLUSTER CTAs
// outputs: E panel <- Vb;  R[c0+k][c0+j] (k <= j < NB);  v0[NB], T[NB][NB] (compact WY)
// ---------------------------------------------------------------------------------------------
struct PanelArgs {
  double* E;
  double* R;
  int ld, nrows, c0, d, mode;
  const Scalars* sc;
  double* v0;    // [NB]
  double* T;     // [NB][NB]
  IwpConsts C;
};

__global__ void __cluster_dims__(NCLUSTER, 1, 1) __launch_bounds__(PANEL_THREADS) panel_kernel(const PanelArgs a) {
  extern __shared__ double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int rows_per = a.nrows / NCLUSTER;
  const int row0 = rank * rows_per;
  constexpr int LDS = NB + 1;
  double* slab = smem;                         // [rows_per][LDS]
  double* part = slab + (size_t)rows_per * LDS;      // [NSTRIPE][NB] partial sums + [2][NB] column sums
  double* gram = part + (PANEL_THREADS / NB + 2) * NB;  // [NB][NB] Gram of Vb (local, then total in rank 0)
  __shared__ double s_v0[NB], s_beta[NB], s_T[NB][NB];
  const int tid = threadIdx.x;
  const double sigma = (a.mode == 0) ? 1.0 : a.sc->sigma;
  const double pi1 = a.sc->pi1;

  for (int idx = tid; idx < rows_per * NB; idx += PANEL_THREADS) {
    const int r = idx / NB, k = idx % NB;
    slab[r * LDS + k] = a.E[(size_t)(row0 + r) * a.ld + a.c0 + k];
  }
  __syncthreads();
  // Thread (col, stripe): lanes of a warp are the NB panel columns, warp w owns row stripe w.  Per column:
  // partial inner products -> block reduction -> one value per column published to the cluster ->
  // warp 0 gathers the 8 ranks through DSMEM, computes the reflector scalars and broadcasts s_j.
  const int col = tid % NB, stripe = tid / NB;
  constexpr int NSTRIPE = PANEL_THREADS / NB;
  double* gl = part + NSTRIPE * NB;  // [2][NB]
  __shared__ double s_coef[NB];
  for (int k = 0; k < NB; ++k) {
    double acc = 0.0;
    if (col >= k)
      for (int r = stripe; r < rows_per; r += NSTRIPE) acc = fma(slab[r * LDS + k], slab[r * LDS + col], acc);
    part[stripe * NB + col] = acc;
    __syncthreads();
    double* glk = gl + (k & 1) * NB;
    if (tid < NB) {
      double sm_ = 0.0;
#pragma unroll
      for (int st = 0; st < NSTRIPE; ++st) sm_ += part[st * NB + tid];
      glk[tid] = sm_;
    }
    cluster.sync();
    if (tid < NB) {
      double g = 0.0;
#pragma unroll
      for (int rk = 0; rk < NCLUSTER; ++rk) g += cluster.map_shared_rank(glk, rk)[tid];
      const double gk = __shfl_sync(0xffffffffu, g, k);
      const int c = a.c0 + k;
      const double pv = sigma * prior_pivot(c, c, a.d, pi1, a.C, a.mode);
      const double n2 = fma(pv, pv, gk);
      const bool nz = n2 > 0.0;
      const double rn = nz ? fast_rsqrt(n2) : 0.0;
      const double nrm = n2 * rn;
      const double v0 = pv + nrm;
      const double beta = nz ? fast_rcp(fma(pv, nrm, n2)) : 0.0;
      double sj = 0.0;
      if (tid > k) {
        const double prj = sigma * prior_pivot(c, a.c0 + tid, a.d, pi1, a.C, a.mode);
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
    if (col > k) {
      const double sj = s_coef[col];
      for (int r = stripe; r < rows_per; r += NSTRIPE) slab[r * LDS + col] = fma(-sj, slab[r * LDS + k], slab[r * LDS + col]);
    }
    __syncwarp();
  }
  __syncthreads();
  // write Vb back, local Gram of Vb
  for (int idx = tid; idx < rows_per * NB; idx += PANEL_THREADS) {
    const int r = idx / NB, k = idx % NB;
    a.E[(size_t)(row0 + r) * a.ld + a.c0 + k] = slab[r * LDS + k];
  }
  for (int e = tid; e < NB * NB; e += PANEL_THREADS) {
    const int i = e / NB, j = e % NB;
    double acc = 0.0;
    if (i <= j)
      for (int r = 0; r < rows_per; ++r) acc = fma(slab[r * LDS + i], slab[r * LDS + j], acc);
    gram[e] = acc;
  }
  cluster.sync();
  if (rank == 0) {
    for (int e = tid; e < NB * NB; e += PANEL_THREADS) {
      double s = gram[e];
      for (int rk = 1; rk < NCLUSTER; ++rk) s += cluster.map_shared_rank(gram, rk)[e];
      gram[e] = s;  // only entries i <= j are meaningful
    }
    __syncthreads();
    // T (upper triangular): T[k][k] = beta_k, T[0:k,k] = -beta_k T[0:k,0:k] (Vb[:,0:k]' Vb[:,k])
    if (tid < NB) {
      for (int j = 0; j < NB; ++j) s_T[tid][j] = 0.0;
    }
    __syncthreads();
    for (int k = 0; k < NB; ++k) {
      if (tid < k) {
        double acc = 0.0;
        for (int m = tid; m < k; ++m) acc = fma(s_T[tid][m], gram[m * NB + k], acc);
        s_T[tid][k] = -s_beta[k] * acc;
      }
      if (tid == k) s_T[k][k] = s_beta[k];
      __syncthreads();
    }
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
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 128) {
      const int r = idx / 32, c = idx % 32;
      As[r * LDA_S + c] = (k0 + r < kend) ? A[(size_t)(k0 + r) * lda + m0 + c] : 0.0;
    }
    for (int idx = threadIdx.x; idx < 32 * 64; idx += 128) {
      const int r = idx / 64, c = idx % 64;
      Bs[r * LDB_S + c] = (k0 + r < kend && n0 + c < N) ? B[(size_t)(k0 + r) * ldb + n0 + c] : 0.0;
    }
    __syncthreads();
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
__global__ void __launch_bounds__(128) rank_update_kernel(const double* __restrict__ A, int lda,
                                                          const double* __restrict__ B, int ldb,
                                                          double* __restrict__ Cm, int ldc, int M, int N) {
  constexpr int LDA_S = 36, LDB_S = 68;
  __shared__ double As[64 * LDA_S], Bs[32 * LDB_S];
  const int i0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < 64 * 32; idx += 128) {
    const int r = idx / 32, c = idx % 32;
    As[r * LDA_S + c] = (i0 + r < M) ? A[(size_t)(i0 + r) * lda + c] : 0.0;
  }
  for (int idx = threadIdx.x; idx < 32 * 64; idx += 128) {
    const int r = idx / 64, c = idx % 64;
    Bs[r * LDB_S + c] = (n0 + c < N) ? B[(size_t)r * ldb + n0 + c] : 0.0;
  }
  // the C fragment is read up front (acc starts as C, the product is subtracted through -A), so that all
  // global loads of the tile are in flight together
  double acc[2][8][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int i = i0 + warp * 16 + mt * 8 + (lane >> 2);
      const int n = n0 + nt * 8 + 2 * (lane & 3);
      acc[mt][nt][0] = (i < M && n < N) ? Cm[(size_t)i * ldc + n] : 0.0;
      acc[mt][nt][1] = (i < M && n + 1 < N) ? Cm[(size_t)i * ldc + n + 1] : 0.0;
    }
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    double af[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) af[mt] = -As[(warp * 16 + mt * 8 + (lane >> 2)) * LDA_S + kk * 4 + (lane & 3)];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const double bf = Bs[(kk * 4 + (lane & 3)) * LDB_S + nt * 8 + (lane >> 2)];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int i = i0 + warp * 16 + mt * 8 + (lane >> 2);
      const int n = n0 + nt * 8 + 2 * (lane & 3);
      if (i < M) {
        if (n < N) Cm[(size_t)i * ldc + n] = acc[mt][nt][0];
        if (n + 1 < N) Cm[(size_t)i * ldc + n + 1] = acc[mt][nt][1];
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
// nrows % (8 * NCLUSTER) == 0, ncols % NB == 0.  work: W [NB][ld], v0 [2][NB], T [2][NB*NB].
// Look-ahead: after panel j, the NB columns of panel j+1 are updated first ("narrow" update) so that the
// factorisation of panel j+1 (8 SMs, stream `aux`) overlaps the update of the remaining trailing columns
// with panel j's reflectors ("wide" update, main stream).
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
  const size_t smem = ((size_t)rows_per * (NB + 1) + (PANEL_THREADS / NB + 2) * NB + NB * NB) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  auto launch_panel = [&](int c0, int buf, cudaStream_t st) {
    PanelArgs pa;
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
    panel_kernel<<<NCLUSTER, PANEL_THREADS, smem, st>>>(pa);
    ++*launches;
  };
  // trailing update of the columns jt in [jt0, jt1) (relative to c0 + NB) with the reflectors of panel c0
  auto update = [&](int c0, int buf, int jt0, int jt1) -> cudaError_t {
    const int nc = jt1 - jt0;
    if (nc <= 0) return cudaSuccess;
    cudaError_t e2 = cudaMemset2DAsync(wk.W + jt0, (size_t)ld * sizeof(double), 0, (size_t)nc * sizeof(double), NB, s);
    if (e2 != cudaSuccess) return e2;
    static const int kchunk = getenv("PNDE_ATB_KCHUNK") ? atoi(getenv("PNDE_ATB_KCHUNK")) : 64;
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
    rank_update_kernel<<<g2, 128, 0, s>>>(E + c0, ld, wk.W + jt0, ld, E + c0 + NB + jt0, ld, nrows, nc);
    *launches += 3;
    return cudaSuccess;
  };
  launch_panel(0, 0, s);
  for (int c0 = 0, j = 0; c0 < ncols; c0 += NB, ++j) {
    const int ntrail = ncols - c0 - NB;
    if (ntrail <= 0) break;
    const int par = j & 1;
    const int nn = ntrail < NB ? ntrail : NB;
    e = update(c0, par, 0, nn);  // narrow: the next panel's columns
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(wk.ev_narrow[par], s)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(wk.aux, wk.ev_narrow[par], 0)) != cudaSuccess) return e;
    launch_panel(c0 + NB, 1 - par, wk.aux);
    if ((e = cudaEventRecord(wk.ev_panel[1 - par], wk.aux)) != cudaSuccess) return e;
    e = update(c0, par, nn, ntrail);  // wide: everything behind it, overlapping the panel factorisation
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(s, wk.ev_panel[1 - par], 0)) != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace big
}  // namespace pnde
