// Interface between pnde_api.cu and the large-D EK1 path (big_dense.cu).
#pragma once
#include <cuda_runtime.h>

#include "filter_kernel.cuh"

namespace pnde {
namespace big {

struct BigRunArgs {
  long long n;       // trajectories (processed one after the other)
  int d, q, diffusion;
  int adaptive;      // 1: PI-controlled steps (the controller runs on the host, one 8-byte read-back per attempted step)
  const double* u0;  // [d][n] device
  const double* p;   // [1][n] device (forcing F)
  double* mean;      // [D][n]
  double* cov;       // [D(D+1)/2][n] or nullptr
  double* t_final;
  double* loglik;
  double* final_diff;
  int* retcode;
  int* naccept;
  int* nreject;
  int* nf;
  int* njacs;
  int* n_saved;
  void* work;        // big_work_bytes(d, q) bytes of device memory
  // history of the solution marginals (a saved full state would be a (D-d) x D factor: 100 MB at D = 4096):
  // record = [t, diffusion of the interval, u[d], diag(Sigma_u)[d]], layout [slot][2 + 2d][n]; null: final state only
  double* hist;
  long long max_saved;
  int save_mode, save_stride;
  IwpConsts C;
  CtrlParams K;
};

size_t big_work_bytes(int d, int q, bool adaptive = false);
cudaError_t big_run(const BigRunArgs& args, cudaStream_t stream, long long* launches);

}  // namespace big
}  // namespace pnde
