// Iterated extended Kalman smoother (reference: src/ieks.jl:53-61 + the IEKS branch of
// measure!, src/perform_step.jl:111-113): every iterate is an EK1 solve whose Jacobian is evaluated at the
// PREVIOUS iterate's dense output sol(t + dt).mu; the vector field itself is still evaluated at the predicted
// mean.  On the device this is the ordinary filter kernel with a linearisation-point policy that calls the
// dense-output routine of post_kernels.cuh on the previous iterate's history once per attempted step.
// The kernels are compiled on demand (rtc_model.cu) so that the statically built catalogue stays as it is.
#pragma once
#include "post_kernels.cuh"

namespace pnde {

struct DenseLin {
  static constexpr bool enabled = true;
  // returns false on the first iterate (no previous solution): the step then linearises at the predicted mean
  template <class M>
  __device__ __noinline__ static bool point(const FilterParams& prm, long long tr, double tval, double (&u)[M::d]) {
    if (!prm.lin.hist) return false;
    DenseParams dp;
    dp.n = prm.n;
    dp.traj_begin = 0;
    dp.traj_end = prm.n;
    dp.max_saved = prm.lin.max_saved;
    dp.n_t = 1;
    dp.n_saved = prm.lin.n_saved;
    dp.hist = prm.lin.hist;
    dp.smooth = prm.lin.smooth;
    dp.final_diff = prm.lin.final_diff;
    dp.calibrate = prm.lin.calibrate;
    dp.is_mv = prm.lin.is_mv;
    dp.smoothed = 1;  // IEKS asserts smooth = true (src/ieks.jl:38)
    dp.tq = nullptr;
    dp.mean = nullptr;
    dp.cov = nullptr;
    dp.C = prm.C;
    double mean[M::D], cov[M::D * (M::D + 1) / 2];
    dense_state<M>(dp, tr, tval, mean, cov);
PNDE_UNROLL
    for (int i = 0; i < M::d; ++i) u[i] = mean[i];  // SolProj * posterior(t)  (src/solution.jl:211-214)
    return true;
  }
};

}  // namespace pnde
