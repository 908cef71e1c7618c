// Per-model dispatch table: every (vector field, order, algorithm) combination is a separate
// template instantiation; the C-ABI layer (pnde_api.cu) looks its launchers up at pnde_create time.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

#include "filter_kernel.cuh"

namespace pnde {

struct ConvertParams {
  long long n;           // trajectories in the ensemble
  long long traj_begin, traj_end;
  long long max_saved;
  const int* n_saved;        // [n]
  const long long* offsets;  // [traj_end - traj_begin + 1] (device)
  const double* hist;        // filtered records
  const double* smooth;      // smoothed records (or nullptr)
  const double* final_diff;  // [ND][n]
  int which;                 // 0 filtered, 1 smoothed
  int calibrate;             // static diffusion model: scale by the final global diffusion
  int is_mv;
  int marginals;             // 1: write u / cov_u only
  double* t;                 // [total]
  double* mean;              // [total][D] or [total][d]
  double* cov;               // [total][D(D+1)/2] or [total][d(d+1)/2]
  double* diffusion;         // [total][nd_out]
  int nd_out;
  double* sqrt;              // [total][D][D] row-major square root S, Sigma = S S' (SRMatrix.squareroot), or null
};

struct SmoothParams {
  long long n;
  long long max_saved;
  const int* n_saved;
  const double* hist;
  double* smooth;            // [max_saved][SREC][n]
  const double* final_diff;  // [ND][n]
  int calibrate;
  int is_mv;
  int* status;               // [n] non-zero: non-finite / negative-variance flag (src/smoothing.jl:25,59)
  int flags;                 // FLAG_ONE_THREAD: the one-thread kernel also where the lane-group smoother exists
  IwpConsts C;
};

struct SampleParams;
struct DenseParams;
struct StepParams;

#ifndef __CUDACC_RTC__
// `self` lets run-time compiled models (rtc_model.cu) carry their CUfunction handles
struct ModelOps {
  int d, q, D, nd, rec, srec, np;
  int sample_len;  // doubles of sampler scratch per (trajectory, interval): SamplePrep<M>::LEN
  bool ek1;
  cudaError_t (*launch_filter)(const ModelOps* self, const FilterParams&, bool adaptive, cudaStream_t);
  cudaError_t (*launch_convert)(const ModelOps* self, const ConvertParams&, cudaStream_t);
  cudaError_t (*launch_smooth)(const ModelOps* self, const SmoothParams&, cudaStream_t);
  cudaError_t (*launch_sample)(const ModelOps* self, const SampleParams&, cudaStream_t);
  cudaError_t (*launch_dense)(const ModelOps* self, const DenseParams&, cudaStream_t);
  cudaError_t (*launch_step)(const ModelOps* self, const StepParams&, cudaStream_t);  // null: not built (NVRTC models)
};

// defined in inst_*.cu
const ModelOps* ops_fhn_readme(int alg, int q, bool mvdyn);
const ModelOps* ops_fhn_lib(int alg, int q, bool mvdyn);
const ModelOps* ops_lotka_volterra(int alg, int q, bool mvdyn);
const ModelOps* ops_vanderpol(int alg, int q, bool mvdyn);
const ModelOps* ops_linear2(int alg, int q, bool mvdyn);
const ModelOps* ops_logistic(int alg, int q, bool mvdyn);
const ModelOps* ops_linear1(int alg, int q, bool mvdyn);
#endif  // !__CUDACC_RTC__

}  // namespace pnde
