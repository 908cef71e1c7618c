// Shared by the inst_*.cu translation units: launchers + the history conversion kernel.
#pragma once
#include <stdlib.h>
#include "model_ops.cuh"
#include "convert_kernel.cuh"
#include "post_kernels.cuh"
#include "dense_sample.cuh"
#include "smoother_kernel.cuh"
#include "wide_filter.cuh"
#include "wide_smoother.cuh"
#include "step_kernel.cuh"

namespace pnde {

template <class M>
cudaError_t launch_filter_t(const ModelOps*, const FilterParams& prm, bool adaptive, cudaStream_t s) {
  // (CTA size does not change the cost of a partial last wave -- 125 k trajectories = 3.3 waves of 148 SMs x 256
  // resident threads run at 0.918 of the 1e6 rate with 128-, 64- and 32-thread CTAs alike, profiles/
  // r2_tail_waves_cta_size.jsonl; PNDE_FILTER_BLOCK_RT is the knob that measurement used)
  int block = PNDE_FILTER_BLOCK;
  if (const char* e = getenv("PNDE_FILTER_BLOCK_RT")) block = atoi(e);
  const long long grid = (prm.count + block - 1) / block;
  if (adaptive) {
    // shared-memory stash of the pre-step state (read back on rejection)
    const size_t smem = (size_t)M::STATE_LEN * block * sizeof(double);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(filter_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    filter_kernel<M, true><<<(unsigned)grid, block, smem, s>>>(prm);
  } else
    filter_kernel<M, false><<<(unsigned)grid, block, 0, s>>>(prm);
  return cudaGetLastError();
}

// Dense EK1 at D >= 10: G = 2 lanes per trajectory (wide_filter.cuh) unless the handle asks for one thread per trajectory
// (PNDE_FLAG_ONE_THREAD: A/B measurements and the bitwise-equality test of the two kernels).
template <class VF, int Q>
cudaError_t launch_filter_wide_t(const ModelOps* self, const FilterParams& prm, bool adaptive, cudaStream_t s) {
  using W = WideEK1<VF, Q, 2>;
  if (prm.flags & FLAG_ONE_THREAD) return launch_filter_t<DenseEK1<VF, Q>>(self, prm, adaptive, s);
  const int block = PNDE_WIDE_BLOCK;
  const long long grid = (prm.count * W::G + block - 1) / block;
  const size_t scr = (size_t)(block / W::G) * W::SCR * sizeof(double);
  const size_t smem = scr + (adaptive ? (size_t)W::STATE_LEN * block * sizeof(double) : 0);
  cudaError_t e = adaptive ? cudaFuncSetAttribute(wide_filter_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(wide_filter_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (adaptive)
    wide_filter_kernel<W, true><<<(unsigned)grid, block, smem, s>>>(prm);
  else
    wide_filter_kernel<W, false><<<(unsigned)grid, block, smem, s>>>(prm);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_convert_t(const ModelOps*, const ConvertParams& c, cudaStream_t s) {
  const int block = 128;
  const long long total = (c.traj_end - c.traj_begin) * c.max_saved;
  if (total <= 0) return cudaSuccess;
  const long long grid = (total + block - 1) / block;
  convert_kernel<M><<<(unsigned)grid, block, 0, s>>>(c);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_smooth_t(const ModelOps*, const SmoothParams& sp, cudaStream_t s) {
  // dense EK1: D*D + D(D+1)/2 doubles of shared memory per thread (X / T' scratch + smoothed factor)
  const int block = PNDE_SMOOTH_BLOCK;  // compile-time stride of the shared-memory scratch
  const size_t smem =
      SmoothModel<M>::USE_SMEM ? (size_t)(M::D * M::D + M::D * (M::D + 1) / 2) * block * sizeof(double) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(smoother_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const long long grid = (sp.n + block - 1) / block;
  smoother_kernel<M><<<(unsigned)grid, block, smem, s>>>(sp);
  return cudaGetLastError();
}

// Dense EK1 at D >= 10: four lanes per trajectory (wide_smoother.cuh) unless PNDE_FLAG_ONE_THREAD
template <class VF, int Q>
cudaError_t launch_smooth_wide_t(const ModelOps* self, const SmoothParams& sp, cudaStream_t s) {
  using W = WideSmooth<VF, Q>;
  if (sp.flags & FLAG_ONE_THREAD) return launch_smooth_t<DenseEK1<VF, Q>>(self, sp, s);
  const int block = W::ST;
  const size_t smem = (size_t)W::SM_LEN * block * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(wide_smoother_kernel<VF, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long grid = (sp.n * W::G + block - 1) / block;
  wide_smoother_kernel<VF, Q><<<(unsigned)grid, block, smem, s>>>(sp);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_sample_t(const ModelOps*, const SampleParams& sp, cudaStream_t s) {
  const int block = 128;
  const long long ntr = sp.traj_end - sp.traj_begin;
  if (ntr <= 0 || sp.n_samples <= 0) return cudaSuccess;
  if (sp.tq) {  // dense_sample: the caller's time grid (dense_sample.cuh)
    if (sp.n_t < 1) return cudaSuccess;
    if (sp.n_t > 1) {
      const long long prep = ntr * (sp.n_t - 1);
      dense_sample_prep_kernel<M><<<(unsigned)((prep + block - 1) / block), block, 0, s>>>(sp);
    }
    const long long total = ntr * sp.n_samples;
    dense_sample_draw_kernel<M><<<(unsigned)((total + block - 1) / block), block, 0, s>>>(sp);
    return cudaGetLastError();
  }
  if (sp.max_saved > 1) {
    const long long prep = ntr * (sp.max_saved - 1);
    sample_prep_kernel<M><<<(unsigned)((prep + block - 1) / block), block, 0, s>>>(sp);
  }
  const long long total = ntr * sp.n_samples;
  sample_draw_kernel<M><<<(unsigned)((total + block - 1) / block), block, 0, s>>>(sp);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_dense_t(const ModelOps*, const DenseParams& dp, cudaStream_t s) {
  const int block = 128;
  const long long total = (dp.traj_end - dp.traj_begin) * dp.n_t;
  if (total <= 0) return cudaSuccess;
  dense_kernel<M><<<(unsigned)((total + block - 1) / block), block, 0, s>>>(dp);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_step_t(const ModelOps*, const StepParams& sp, cudaStream_t s) {
  const int block = 64;
  if (sp.n <= 0) return cudaSuccess;
  step_kernel<M><<<(unsigned)((sp.n + block - 1) / block), block, 0, s>>>(sp);
  return cudaGetLastError();
}

template <class M>
const ModelOps* make_ops() {
  static const ModelOps ops = {M::d,
                               M::q,
                               M::D,
                               M::ND,
                               M::REC,
                               SmoothModel<M>::SREC,
                               M::VF::np,
                               SamplePrep<M>::LEN,
                               M::IS_EK1,
                               &launch_filter_t<M>,
                               &launch_convert_t<M>,
                               &launch_smooth_t<M>,
                               &launch_sample_t<M>,
                               &launch_dense_t<M>,
                               &launch_step_t<M>};
  return &ops;
}

// D = 8 (q = 3, d = 2): experiment switch PNDE_WIDE_SMOOTH=1 runs the four-lane smoother instead of the one-thread one
template <class VF, int Q>
cudaError_t launch_smooth_try_wide_t(const ModelOps* self, const SmoothParams& sp, cudaStream_t s) {
  static const bool wide = getenv("PNDE_WIDE_SMOOTH") && atoi(getenv("PNDE_WIDE_SMOOTH")) != 0;
  if (wide) return launch_smooth_wide_t<VF, Q>(self, sp, s);
  return launch_smooth_t<DenseEK1<VF, Q>>(self, sp, s);
}

template <class VF, int Q>
const ModelOps* make_ops_ek1() {
  using M = DenseEK1<VF, Q>;
  if constexpr (M::D == 8 && VF::d == 2) {
    static const ModelOps ops = {M::d, M::q, M::D, M::ND, M::REC, SmoothModel<M>::SREC, M::VF::np, SamplePrep<M>::LEN, true,
                                 &launch_filter_t<M>, &launch_convert_t<M>, &launch_smooth_try_wide_t<VF, Q>,
                                 &launch_sample_t<M>, &launch_dense_t<M>, &launch_step_t<M>};
    return &ops;
  } else if constexpr (M::D >= 10 && VF::d % 2 == 0) {
    static const ModelOps ops = {M::d, M::q, M::D, M::ND, M::REC, SmoothModel<M>::SREC, M::VF::np, SamplePrep<M>::LEN, true,
                                 &launch_filter_wide_t<VF, Q>, &launch_convert_t<M>, &launch_smooth_wide_t<VF, Q>,
                                 &launch_sample_t<M>, &launch_dense_t<M>, &launch_step_t<M>};
    return &ops;
  } else {
    return make_ops<M>();
  }
}

// q = 1..QINST for every algorithm
#define PNDE_QCASE(VF, Q)                                                     \
  case Q:                                                                     \
    if (alg == 1) return make_ops_ek1<VF, Q>();                               \
    return mvdyn ? make_ops<KronEK0<VF, Q, true>>() : make_ops<KronEK0<VF, Q, false>>();

#define PNDE_DEFINE_OPS(NAME, VF)                             \
  const ModelOps* NAME(int alg, int q, bool mvdyn) {          \
    switch (q) {                                              \
      PNDE_QCASE(VF, 1)                                       \
      PNDE_QCASE(VF, 2)                                       \
      PNDE_QCASE(VF, 3)                                       \
      PNDE_QCASE(VF, 4)                                       \
      PNDE_QCASE(VF, 5)                                       \
      default:                                                \
        return nullptr;                                       \
    }                                                         \
  }

}  // namespace pnde
