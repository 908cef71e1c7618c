// C-ABI layer of libpnde.so (see include/pnde.h).  Host logic only: configuration checks, IWP
// constants (src/priors.jl:7-59), device buffers, kernel dispatch, result marshalling.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pnde.h"
#include "model_ops.cuh"
#include "big_dense_api.h"
#include "lorenz96_kernel.cuh"
#include "post_kernels.cuh"
#include "rtc_model.h"
#include "step_kernel.cuh"

using namespace pnde;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t want) {
    if (want <= bytes && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    if (want == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// Qtilde and its Cholesky factor for the once-per-solve constants (src/priors.jl:29-55, d = 1).
bool build_iwp(int q, IwpConsts& C) {
  memset(&C, 0, sizeof(C));
  double fact[2 * QMAX + 2];
  fact[0] = 1.0;
  for (int i = 1; i < 2 * QMAX + 2; ++i) fact[i] = fact[i - 1] * i;
  for (int r = 0; r <= q; ++r)
    for (int c = 0; c <= q; ++c) C.Qt[r][c] = 1.0 / (double(2 * q + 1 - r - c) * fact[q - r] * fact[q - c]);
  for (int j = 0; j <= q; ++j) {
    double s = C.Qt[j][j];
    for (int k = 0; k < j; ++k) s -= C.Lt[j][k] * C.Lt[j][k];
    if (!(s > 0.0)) return false;
    C.Lt[j][j] = sqrt(s);
    for (int i = j + 1; i <= q; ++i) {
      double t = C.Qt[i][j];
      for (int k = 0; k < j; ++k) t -= C.Lt[i][k] * C.Lt[j][k];
      C.Lt[i][j] = t / C.Lt[j][j];
    }
  }
  return true;
}

const ModelOps* find_ops(int vf, int alg, int q, bool mvdyn) {
  switch (vf) {
    case PNDE_VF_FHN_README: return ops_fhn_readme(alg, q, mvdyn);
    case PNDE_VF_FHN_LIB: return ops_fhn_lib(alg, q, mvdyn);
    case PNDE_VF_LOTKA_VOLTERRA: return ops_lotka_volterra(alg, q, mvdyn);
    case PNDE_VF_VANDERPOL: return ops_vanderpol(alg, q, mvdyn);
    case PNDE_VF_LINEAR2: return ops_linear2(alg, q, mvdyn);
    case PNDE_VF_LOGISTIC: return ops_logistic(alg, q, mvdyn);
    case PNDE_VF_LINEAR1: return ops_linear1(alg, q, mvdyn);
    default: return nullptr;
  }
}

}  // namespace

struct pnde_handle {
  pnde_config cfg;
  const ModelOps* ops = nullptr;  // nullptr for the CTA-per-trajectory Lorenz-96 path
  bool owns_ops = false;           // run-time compiled model (pnde_create_custom)
  bool lorenz = false;
  bool ieks = false;       // PNDE_ALG_IEKS: cfg.alg is stored as EK1, the filter kernels carry the DenseLin policy
  bool have_prev = false;  // ... and prev_* hold the previous iterate (linearisation points of the next pnde_run)
  long long prev_max_saved = 0;
  int d = 0, D = 0, np = 0, nd = 1, ncov = 0;  // dimensions (from ops, or from cfg for Lorenz-96)
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;  // pnde_solve_ensemble_to_host: device-to-host copies of finished slices
  cudaStream_t stream2 = nullptr;      // ... and the second compute stream of its slices
  cudaEvent_t slice_ev[3] = {nullptr, nullptr, nullptr};
  IwpConsts C;
  long long n = 0;
  long long max_saved = 0;
  bool ran = false, smoothed = false;
  double filter_ms = 0.0, smooth_ms = 0.0;
  long long launches = 0;
  DevBuf u0, p, mean, cov, t_final, loglik, final_diff, retcode, naccept, nreject, nf, njacs, n_saved, hist, smooth,
      sstatus, scratch_off, scratch_out, sample_scratch, bigwork, prev_hist, prev_smooth, prev_n_saved, prev_final_diff;
  std::string err;
  // multi-device handle (cfg.n_devices > 1): one single-device child per GPU, trajectories in contiguous shards
  // [kid_lo[k], kid_lo[k+1]); the parent owns no device memory.  SURVEY 8(e): no exchange step, results are written
  // straight into disjoint slices of the caller's arrays.
  std::vector<pnde_handle*> kids;
  std::vector<long long> kid_lo;
  long long global_lo = 0;  // child: index of its first trajectory in the parent's ensemble (keys the sampler)
  bool multi() const { return !kids.empty(); }

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return PNDE_ERR_CUDA;
  }
};

#define CK(call, what)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return h->cuda_fail(e__, what); \
  } while (0)

extern "C" {

int pnde_default_config(pnde_config* cfg, int32_t alg, int32_t order, int32_t vf_kind) {
  if (!cfg) return PNDE_ERR_ARG;
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = PNDE_ABI_VERSION;
  cfg->alg = alg;
  cfg->order = order;
  cfg->vf_kind = vf_kind;
  cfg->d = 0;
  cfg->diffusion = PNDE_DIFF_DYNAMIC;
  cfg->smooth = 0;
  cfg->adaptive = 1;
  cfg->save_mode = PNDE_SAVE_FINAL;
  cfg->save_stride = 1;
  cfg->device = -1;
  cfg->abstol = 1e-6;
  cfg->reltol = 1e-3;
  cfg->dt = 0.0;
  cfg->t0 = 0.0;
  cfg->t1 = 1.0;
  cfg->qmin = 1.0 / 5.0;
  cfg->qmax = 10.0;
  cfg->gamma = 9.0 / 10.0;
  cfg->qsteady_min = 1.0;
  cfg->qsteady_max = 1.0;
  cfg->qoldinit = 1e-4;
  cfg->beta1 = 0.0;  // <= 0: 7/(10(q+1)), src/alg_utils.jl:24
  cfg->beta2 = 0.0;  // <= 0: 2/(5(q+1)),  src/alg_utils.jl:23
  cfg->dtmin = 0.0;
  cfg->dtmax = 0.0;  // <= 0: t1 - t0
  cfg->maxiters = 100000;
  cfg->max_saved = 0;
  cfg->n_devices = 0;
  cfg->flags = 0;
  return PNDE_OK;
}

struct CustomVf {
  int d, np;
  const char* f_body;
  const char* jac_body;
};

static int create_impl(const pnde_config* cfg, pnde_handle** out, const CustomVf* custom, bool as_ieks = false) {
  CustomVf catalogue_rtc = {0, 0, nullptr, nullptr};
  if (!cfg || !out) {
    g_create_error = "null argument";
    return PNDE_ERR_ARG;
  }
  *out = nullptr;
  if (cfg->abi_version != PNDE_ABI_VERSION) {
    g_create_error = "abi_version mismatch";
    return PNDE_ERR_ARG;
  }
  if (cfg->alg != PNDE_ALG_EK0 && cfg->alg != PNDE_ALG_EK1 && cfg->alg != PNDE_ALG_IEKS) {
    g_create_error = "alg must be PNDE_ALG_EK0, PNDE_ALG_EK1 or PNDE_ALG_IEKS";
    return PNDE_ERR_ARG;
  }
  const bool ieks = cfg->alg == PNDE_ALG_IEKS;
  if (ieks) {
    if (!cfg->smooth || cfg->save_mode != PNDE_SAVE_EVERY) {
      g_create_error = "IEKS requires smooth = 1 and save_mode = PNDE_SAVE_EVERY (src/ieks.jl:38-40)";
      return PNDE_ERR_ARG;
    }
    if (cfg->vf_kind == PNDE_VF_LORENZ96) {
      g_create_error = "IEKS is not built for the large-d Lorenz-96 paths";
      return PNDE_ERR_UNSUPPORTED;
    }
    pnde_config c2 = *cfg;  // everything below sees an EK1
    c2.alg = PNDE_ALG_EK1;
    return create_impl(&c2, out, custom, true);
  }
  if (cfg->order < 1 || cfg->order > QMAX) {
    g_create_error = "order must be in 1..7";
    return PNDE_ERR_UNSUPPORTED;
  }
  if (cfg->diffusion < 0 || cfg->diffusion > 4) {
    g_create_error = "unknown diffusion model";
    return PNDE_ERR_ARG;
  }
  const bool mv = (cfg->diffusion == PNDE_DIFF_DYNAMIC_MV || cfg->diffusion == PNDE_DIFF_FIXED_MV);
  if (mv && cfg->alg != PNDE_ALG_EK0) {
    // the reference asserts this (src/diffusions.jl:97-101,128,134)
    g_create_error = "MV diffusion models are EK0-only (src/diffusions.jl:97)";
    return PNDE_ERR_ARG;
  }
  if (!cfg->adaptive && !(cfg->dt > 0.0)) {
    g_create_error = "Fixed timestep methods require a choice of dt";  // test/errors.jl:16-20
    return PNDE_ERR_ARG;
  }
  if (!(cfg->t1 > cfg->t0)) {
    g_create_error = "tspan must satisfy t1 > t0";
    return PNDE_ERR_ARG;
  }
  if (cfg->smooth && cfg->save_mode != PNDE_SAVE_EVERY) {
    g_create_error = "smooth requires save_mode = PNDE_SAVE_EVERY (src/perform_step.jl:3)";
    return PNDE_ERR_ARG;
  }
  if (cfg->save_mode == PNDE_SAVE_STRIDE && cfg->save_stride < 1) {
    g_create_error = "save_stride must be >= 1";
    return PNDE_ERR_ARG;
  }
  const bool lorenz = (cfg->vf_kind == PNDE_VF_LORENZ96);
  const ModelOps* ops = nullptr;
  if (lorenz) {
    if (cfg->alg == PNDE_ALG_EK1) {
      // large-D dense path (blocked Householder QR with FP64 tensor-core updates, big_dense.cu)
      if (cfg->d < 32 || cfg->d % 32 != 0 || cfg->d * (cfg->order + 1) > 5120) {
        g_create_error = "Lorenz-96 EK1 (dense, D >= 64): d must be a multiple of 32 with d (q+1) <= 5120";
        return PNDE_ERR_UNSUPPORTED;
      }
    } else if (cfg->d < 4 || cfg->d > 8 * LORENZ_THREADS) {
      g_create_error = "Lorenz-96: d must be in 4..2048";
      return PNDE_ERR_ARG;
    } else if (cfg->order > 5) {
      g_create_error = "Lorenz-96 EK0: orders 1..5 are built";
      return PNDE_ERR_UNSUPPORTED;
    }
    if (mv) {
      g_create_error = "Lorenz-96: MV diffusion models are not built for the large-d path";
      return PNDE_ERR_UNSUPPORTED;
    }
    if (cfg->alg == PNDE_ALG_EK1 && cfg->smooth) {
      g_create_error = "Lorenz-96 EK1 (dense): smoothing is not built (the saved history holds the solution marginals only)";
      return PNDE_ERR_UNSUPPORTED;
    }
  } else if (custom) {
    // EK1 carries a dense D x (D - d) factor per thread: d <= 8.  EK0's covariance is the (q+1) x q Kronecker factor
    // whatever d is; only the mean grows (d (q + 1) doubles per thread): d <= 16.
    // Up to D = d (q+1) = 16 (EK1) / 64 (EK0) the kernels are unrolled into registers like the catalogue models; beyond
    // that they are compiled with their loops rolled and their arrays in local memory (PNDE_ROLLED, cov_engine.cuh):
    // any dimension works, at a fraction of the speed.  EK1 carries D x D arrays per thread: D <= 96.
    const int Dc = custom->d * (cfg->order + 1);
    const int Dmax = (cfg->alg == PNDE_ALG_EK0) ? 1024 : 96;
    if (custom->d < 1 || custom->d > 128 || Dc > Dmax || custom->np < 0 || custom->np > 64) {
      g_create_error = "custom vector field: d must be in 1..128 with d (q+1) <= 96 (EK1) / 1024 (EK0), n_params in 0..64";
      return PNDE_ERR_ARG;
    }
  } else {
    // PNDE_FLAG_REFERENCE_QUIRKS (b) lives only in run-time compiled kernels (filter_kernel.cuh): take that route
    const bool quirk_rtc = (cfg->flags & PNDE_FLAG_REFERENCE_QUIRKS) && cfg->diffusion == PNDE_DIFF_FIXED;
    ops = (as_ieks || quirk_rtc) ? nullptr : find_ops(cfg->vf_kind, cfg->alg, cfg->order, cfg->diffusion == PNDE_DIFF_DYNAMIC_MV);
    if (!ops) {
      // orders 6 and 7 of the catalogue, and the IEKS flavour of every order, are not instantiated at build
      // time: compile them on demand (NVRTC)
      static const struct { int kind, d, np; const char* name; } cat[] = {
          {PNDE_VF_FHN_README, 2, 3, "@catalogue:VfFhnReadme"}, {PNDE_VF_FHN_LIB, 2, 4, "@catalogue:VfFhnLib"},
          {PNDE_VF_LOTKA_VOLTERRA, 2, 4, "@catalogue:VfLotkaVolterra"}, {PNDE_VF_VANDERPOL, 2, 1, "@catalogue:VfVanDerPol"},
          {PNDE_VF_LINEAR2, 2, 2, "@catalogue:VfLinear2"}, {PNDE_VF_LOGISTIC, 1, 1, "@catalogue:VfLogistic"},
          {PNDE_VF_LINEAR1, 1, 1, "@catalogue:VfLinear1"}};
      for (const auto& e : cat)
        if (e.kind == cfg->vf_kind) {
          catalogue_rtc = {e.d, e.np, e.name, nullptr};
          custom = &catalogue_rtc;
        }
      if (!custom) {
        g_create_error = "no built-in kernel for this (vf_kind, alg, order)";
        return PNDE_ERR_UNSUPPORTED;
      }
    } else if (cfg->d != 0 && cfg->d != ops->d) {
      g_create_error = "d does not match the vector field";
      return PNDE_ERR_ARG;
    }
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
    return PNDE_ERR_CUDA;
  }
  int dev = cfg->device;
  if (dev < 0) {
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
      g_create_error = std::string("cudaGetDevice: ") + cudaGetErrorString(e);
      return PNDE_ERR_CUDA;
    }
  }
  if (dev >= ndev) {
    g_create_error = "device ordinal out of range";
    return PNDE_ERR_ARG;
  }
  bool owns = false;
  if (custom) {
    if (cudaSetDevice(dev) != cudaSuccess) {
      g_create_error = "cudaSetDevice failed";
      return PNDE_ERR_CUDA;
    }
    std::string rerr;
    ops = rtc_build(cfg->alg, cfg->order, cfg->diffusion == PNDE_DIFF_DYNAMIC_MV, custom->d, custom->np, custom->f_body,
                    custom->jac_body, rerr, as_ieks, cfg->adaptive ? 1 : 0,
                    (cfg->flags & PNDE_FLAG_REFERENCE_QUIRKS) && cfg->diffusion == PNDE_DIFF_FIXED,
                    !(cfg->flags & PNDE_FLAG_ONE_THREAD));
    if (!ops) {
      g_create_error = rerr;
      return PNDE_ERR_ARG;
    }
    owns = true;
    // adaptive kernels park the pre-step state in shared memory (STATE_LEN x 128 doubles per CTA): refuse what
    // cannot be launched now instead of failing at the first pnde_run with an opaque "invalid value"
    const bool rolled = rtc_rolled(cfg->alg, custom->d, cfg->order) && strncmp(custom->f_body ? custom->f_body : "", "@catalogue:", 11) != 0;
    const bool lanes = !rolled && cfg->alg == PNDE_ALG_EK1 && !as_ieks && ops->D >= 10 && ops->d % 2 == 0 && !(cfg->flags & PNDE_FLAG_ONE_THREAD) &&
                       !((cfg->flags & PNDE_FLAG_REFERENCE_QUIRKS) && cfg->diffusion == PNDE_DIFF_FIXED);
    const size_t stash = (size_t)(ops->rec - 1 - ops->nd) * 128 * sizeof(double) / (lanes ? 2 : 1) + (lanes ? (size_t)64 * (ops->D * (ops->D - ops->d) + cfg->order + 1) * 8 : 0);
    if (cfg->adaptive && !rolled && stash > 227 * 1024) {
      rtc_destroy(ops);
      g_create_error = "adaptive steps with d = " + std::to_string(custom->d) + ", order = " + std::to_string(cfg->order) +
                       " need " + std::to_string(stash) + " B of shared memory per CTA for the pre-step state (limit 232448): "
                       "use a lower order or fixed steps";
      return PNDE_ERR_UNSUPPORTED;
    }
  }
  pnde_handle* h = new pnde_handle();
  h->cfg = *cfg;
  h->ops = ops;
  h->owns_ops = owns;
  h->lorenz = lorenz;
  h->ieks = as_ieks;
  if (lorenz) {
    h->d = cfg->d;
    h->D = cfg->d * (cfg->order + 1);
    h->np = 1;
    h->nd = 1;
    h->ncov = (cfg->order + 1) * (cfg->order + 2) / 2;  // Kronecker factor Ctilde
    if (cfg->alg == PNDE_ALG_EK1) h->ncov = h->D * (h->D + 1) / 2;
  } else {
    h->d = ops->d;
    h->D = ops->D;
    h->np = ops->np;
    h->nd = ops->nd;
    h->ncov = ops->D * (ops->D + 1) / 2;
  }
  h->cfg.d = h->d;
  h->device = dev;
  if (!build_iwp(cfg->order, h->C)) {
    if (owns) rtc_destroy(ops);
    delete h;
    g_create_error = "IWP process-noise Cholesky failed";
    return PNDE_ERR_ARG;
  }
  const int q = cfg->order;
  if (!(h->cfg.beta2 > 0.0)) h->cfg.beta2 = 2.0 / (5.0 * (q + 1));
  if (!(h->cfg.beta1 > 0.0)) h->cfg.beta1 = 7.0 / (10.0 * (q + 1));
  if (!(h->cfg.dtmax > 0.0)) h->cfg.dtmax = cfg->t1 - cfg->t0;
  if (h->cfg.maxiters <= 0) h->cfg.maxiters = 100000;
  e = cudaSetDevice(dev);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
  if (e != cudaSuccess) {
    g_create_error = std::string("stream/event creation: ") + cudaGetErrorString(e);
    for (int i = 0; i < 4; ++i)
      if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (owns) rtc_destroy(ops);
    delete h;
    return PNDE_ERR_CUDA;
  }
  *out = h;
  return PNDE_OK;
}

// cfg.n_devices > 1: one child handle per listed device under a parent that owns no device memory
static int create_any(const pnde_config* cfg, pnde_handle** out, const CustomVf* custom) {
  if (!cfg || !out || cfg->abi_version != PNDE_ABI_VERSION || cfg->n_devices <= 1) {
    if (cfg && out && cfg->abi_version == PNDE_ABI_VERSION && cfg->n_devices == 1) {
      pnde_config c1 = *cfg;
      c1.device = cfg->device_list[0];
      c1.n_devices = 0;
      return create_impl(&c1, out, custom);
    }
    return create_impl(cfg, out, custom);
  }
  *out = nullptr;
  if (cfg->n_devices > PNDE_MAX_DEVICES) {
    g_create_error = "n_devices exceeds PNDE_MAX_DEVICES";
    return PNDE_ERR_ARG;
  }
  for (int i = 0; i < cfg->n_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (cfg->device_list[i] == cfg->device_list[j]) {
        g_create_error = "device_list holds a device twice";
        return PNDE_ERR_ARG;
      }
  pnde_handle* parent = new pnde_handle();
  for (int i = 0; i < cfg->n_devices; ++i) {
    pnde_config ck = *cfg;
    ck.device = cfg->device_list[i];
    ck.n_devices = 0;
    pnde_handle* kid = nullptr;
    const int rc = create_impl(&ck, &kid, custom);
    if (rc != PNDE_OK) {
      for (pnde_handle* k : parent->kids) pnde_destroy(k);
      delete parent;
      return rc;
    }
    parent->kids.push_back(kid);
  }
  const pnde_handle* k0 = parent->kids[0];
  parent->cfg = k0->cfg;
  parent->cfg.n_devices = cfg->n_devices;
  parent->d = k0->d;
  parent->D = k0->D;
  parent->np = k0->np;
  parent->nd = k0->nd;
  parent->ncov = k0->ncov;
  parent->lorenz = k0->lorenz;
  parent->ieks = k0->ieks;
  parent->ops = k0->ops;  // dimension helpers only; never launched through the parent
  parent->device = k0->device;
  *out = parent;
  return PNDE_OK;
}

int pnde_create(const pnde_config* cfg, pnde_handle** out) { return create_any(cfg, out, nullptr); }

int pnde_create_custom(const pnde_config* cfg, int32_t d, int32_t n_params, const char* f_body, const char* jac_body,
                       pnde_handle** out) {
  CustomVf c = {d, n_params, f_body, jac_body};
  if (cfg && cfg->vf_kind != PNDE_VF_CUSTOM) {
    g_create_error = "pnde_create_custom: cfg.vf_kind must be PNDE_VF_CUSTOM";
    return PNDE_ERR_ARG;
  }
  return create_any(cfg, out, &c);
}

int pnde_check_custom(int32_t alg, int32_t order, int32_t diffusion, int32_t d, int32_t n_params, const char* f_body,
                      const char* jac_body, char* log, int64_t log_len) {
  std::string err;
  const bool ieks = alg == PNDE_ALG_IEKS;
  const bool ok = rtc_check(ieks ? PNDE_ALG_EK1 : alg, order, diffusion == PNDE_DIFF_DYNAMIC_MV, d, n_params, f_body, jac_body,
                            err, ieks);
  if (log && log_len > 0) {
    strncpy(log, err.c_str(), (size_t)log_len - 1);
    log[log_len - 1] = 0;
  }
  return ok ? PNDE_OK : PNDE_ERR_ARG;
}

int pnde_destroy(pnde_handle* h) {
  if (!h) return PNDE_OK;
  if (h->multi()) {
    for (pnde_handle* k : h->kids) pnde_destroy(k);
    delete h;
    return PNDE_OK;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = {&h->u0,      &h->p,      &h->mean,  &h->cov,     &h->t_final, &h->loglik,
                    &h->final_diff, &h->retcode, &h->naccept, &h->nreject, &h->nf,      &h->njacs,
                    &h->n_saved, &h->hist,   &h->smooth, &h->sstatus, &h->scratch_off, &h->scratch_out, &h->sample_scratch,
                    &h->bigwork, &h->prev_hist, &h->prev_smooth, &h->prev_n_saved, &h->prev_final_diff};
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < 4; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) {
    cudaStreamDestroy(h->copy_stream);
    cudaStreamDestroy(h->stream2);
    for (int i = 0; i < 3; ++i) cudaEventDestroy(h->slice_ev[i]);
  }
  if (h->owns_ops) rtc_destroy(h->ops);
  delete h;
  return PNDE_OK;
}

const char* pnde_last_error(const pnde_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t pnde_state_dim(const pnde_handle* h) { return h ? h->D : 0; }
int64_t pnde_n_params(const pnde_handle* h) { return h ? h->np : 0; }
// doubles per saved state: catalogue / NVRTC models from their ops; large-d Kronecker path: t, diffusion, factor, mean
static int lorenz_rec(const pnde_handle* h) {
  const int q = h->cfg.order, nz = q - 1;
  return 2 + (q + 1) + (nz > 0 ? nz * (nz + 1) / 2 : 0) + h->D;
}
static int lorenz_srec(const pnde_handle* h) { return (h->cfg.order + 1) * (h->cfg.order + 2) / 2 + 1 + h->D; }
int64_t pnde_record_len(const pnde_handle* h) {
  if (!h) return 0;
  if (h->ops) return h->ops->rec;
  if (h->lorenz && h->cfg.alg == PNDE_ALG_EK1) return 2 + 2 * h->d;  // large-D dense path: t, diffusion, u, diag(Sigma_u)
  return h->lorenz ? lorenz_rec(h) : 0;
}
int64_t pnde_cov_len(const pnde_handle* h) { return h ? h->ncov : 0; }

static long long derive_max_saved(const pnde_handle* h) {
  const pnde_config& c = h->cfg;
  if (c.save_mode == PNDE_SAVE_FINAL) return 0;
  if (c.max_saved > 0) return c.max_saved;
  if (c.adaptive) return 0;  // must be given
  // fixed step: number of steps of the t += dt loop (+ possible sliver step) + initial state
  long long steps = (long long)ceil((c.t1 - c.t0) / c.dt) + 2;
  if (c.save_mode == PNDE_SAVE_STRIDE) steps = steps / c.save_stride + 3;
  return steps + 1;
}

}  // extern "C"  (templates below)

// Runs fn(kid index) for every child that holds trajectories, one host thread per device, and returns the first
// failure (its text is copied into the parent's error slot).
template <class Fn>
static int for_each_kid(pnde_handle* h, Fn fn) {
  const size_t nk = h->kids.size();
  std::vector<int> rc(nk, PNDE_OK);
  std::vector<std::thread> th;
  for (size_t k = 0; k < nk; ++k) {
    if (h->kid_lo.size() == nk + 1 && h->kid_lo[k + 1] <= h->kid_lo[k]) continue;  // empty shard
    th.emplace_back([&, k] { rc[k] = fn((int)k); });
  }
  for (auto& t : th) t.join();
  for (size_t k = 0; k < nk; ++k)
    if (rc[k] != PNDE_OK) {
      h->err = "device " + std::to_string(h->kids[k]->device) + ": " + h->kids[k]->err;
      return rc[k];
    }
  return PNDE_OK;
}

extern "C" {

// contiguous block partition, remainder spread over the first shards (|shard sizes| differ by at most one)
static void partition(pnde_handle* h, long long n) {
  const long long nk = (long long)h->kids.size();
  h->kid_lo.assign((size_t)nk + 1, 0);
  for (long long k = 0; k < nk; ++k) {
    h->kid_lo[(size_t)k + 1] = h->kid_lo[(size_t)k] + n / nk + (k < n % nk ? 1 : 0);
    h->kids[(size_t)k]->global_lo = h->kid_lo[(size_t)k];
  }
}

// src_pitch: elements between consecutive rows of the caller's [rows][n_total] arrays (== n_traj for a whole ensemble)
static int upload_impl(pnde_handle* h, int64_t n_traj, const double* u0, const double* p, int64_t src_pitch) {
  if (!h) return PNDE_ERR_ARG;
  if (n_traj <= 0 || !u0 || (!p && h->np > 0)) return h->fail(PNDE_ERR_ARG, "pnde_upload: bad arguments");
  if (h->multi()) {
    partition(h, n_traj);
    h->n = n_traj;
    h->ran = h->smoothed = false;
    return for_each_kid(h, [&](int k) {
      const long long lo = h->kid_lo[(size_t)k], cnt = h->kid_lo[(size_t)k + 1] - lo;
      return upload_impl(h->kids[(size_t)k], cnt, u0 + lo, p ? p + lo : nullptr, src_pitch);
    });
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  struct { int d, np, D, nd, rec; } dims = {h->d, h->np, h->D, h->nd, (int)pnde_record_len(h)};
  const auto* o = &dims;
  const size_t n = (size_t)n_traj;
  const long long ms = derive_max_saved(h);
  if (h->cfg.save_mode != PNDE_SAVE_FINAL && ms <= 0)
    return h->fail(PNDE_ERR_ARG, "adaptive runs that save history need cfg.max_saved > 0");
  h->max_saved = ms;
  CK(h->u0.ensure(n * o->d * 8), "alloc u0");
  CK(h->p.ensure(n * (o->np > 0 ? o->np : 1) * 8), "alloc p");
  CK(h->mean.ensure(n * o->D * 8), "alloc mean");
  CK(h->cov.ensure(n * (size_t)h->ncov * 8), "alloc cov");
  CK(h->t_final.ensure(n * 8), "alloc t_final");
  CK(h->loglik.ensure(n * 8), "alloc loglik");
  CK(h->final_diff.ensure(n * o->nd * 8), "alloc final_diff");
  CK(h->retcode.ensure(n * 4), "alloc retcode");
  CK(h->naccept.ensure(n * 4), "alloc naccept");
  CK(h->nreject.ensure(n * 4), "alloc nreject");
  CK(h->nf.ensure(n * 4), "alloc nf");
  CK(h->njacs.ensure(n * 4), "alloc njacs");
  CK(h->n_saved.ensure(n * 4), "alloc n_saved");
  if (ms > 0) {
    cudaError_t e = h->hist.ensure(n * (size_t)ms * o->rec * 8);
    if (e != cudaSuccess) {
      h->err = std::string("history allocation (") + std::to_string(n * (size_t)ms * o->rec * 8) +
               " bytes): " + cudaGetErrorString(e);
      cudaGetLastError();
      return PNDE_ERR_ALLOC;
    }
  }
  CK(cudaMemcpy2DAsync(h->u0.p, n * 8, u0, (size_t)src_pitch * 8, n * 8, (size_t)o->d, cudaMemcpyHostToDevice, h->stream), "H2D u0");
  if (o->np > 0)
    CK(cudaMemcpy2DAsync(h->p.p, n * 8, p, (size_t)src_pitch * 8, n * 8, (size_t)o->np, cudaMemcpyHostToDevice, h->stream), "H2D p");
  // with page-locked inputs the copies above are truly asynchronous: wait, so that the caller may reuse or free
  // u0 / p as soon as this call returns (the kernel cannot start before they have arrived anyway)
  CK(cudaStreamSynchronize(h->stream), "H2D synchronize");
  h->n = n_traj;
  h->ran = false;
  h->smoothed = false;
  h->have_prev = false;
  return PNDE_OK;
}

int pnde_upload(pnde_handle* h, int64_t n_traj, const double* u0, const double* p) {
  return upload_impl(h, n_traj, u0, p, n_traj);
}

static void fill_filter_params(pnde_handle* h, FilterParams& fp) {
  const pnde_config& c = h->cfg;
  memset(&fp, 0, sizeof(fp));
  fp.n = h->n;
  fp.first = 0;
  fp.count = h->n;
  fp.u0 = h->u0.as<double>();
  fp.p = h->p.as<double>();
  fp.mean = h->mean.as<double>();
  fp.cov = h->cov.as<double>();
  fp.t_final = h->t_final.as<double>();
  fp.loglik = h->loglik.as<double>();
  fp.final_diff = h->final_diff.as<double>();
  fp.retcode = h->retcode.as<int>();
  fp.naccept = h->naccept.as<int>();
  fp.nreject = h->nreject.as<int>();
  fp.nf = h->nf.as<int>();
  fp.njacs = h->njacs.as<int>();
  fp.n_saved = h->n_saved.as<int>();
  fp.hist = h->hist.as<double>();
  fp.max_saved = h->max_saved;
  fp.save_mode = c.save_mode;
  fp.save_stride = c.save_stride > 0 ? c.save_stride : 1;
  fp.diffusion = c.diffusion;
  fp.flags = c.flags;
  fp.C = h->C;
  fp.K.abstol = c.abstol;
  fp.K.reltol = c.reltol;
  fp.K.dt = c.dt;
  fp.K.t0 = c.t0;
  fp.K.t1 = c.t1;
  fp.K.qmin = c.qmin;
  fp.K.qmax = c.qmax;
  fp.K.gamma = c.gamma;
  fp.K.qsteady_min = c.qsteady_min;
  fp.K.qsteady_max = c.qsteady_max;
  fp.K.qoldinit = c.qoldinit;
  fp.K.beta1 = c.beta1;
  fp.K.beta2 = c.beta2;
  fp.K.dtmin = c.dtmin;
  fp.K.dtmax = c.dtmax;
  fp.K.maxiters = c.maxiters;
}

int pnde_run(pnde_handle* h) {
  if (!h) return PNDE_ERR_ARG;
  if (h->n <= 0) return h->fail(PNDE_ERR_STATE, "pnde_run: nothing uploaded");
  if (h->multi()) {
    // asynchronous launches: one thread suffices, every device works on its own stream
    for (size_t k = 0; k < h->kids.size(); ++k) {
      if (h->kid_lo[k + 1] <= h->kid_lo[k]) continue;
      const int rc = pnde_run(h->kids[k]);
      if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
    }
    h->ran = true;
    h->smoothed = false;
    return PNDE_OK;
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const pnde_config& c = h->cfg;
  if (h->ieks && h->ran) {
    // next IEKS iterate (src/ieks.jl:57-59): the solution so far becomes the linearisation trajectory
    if (!h->smoothed)
      return h->fail(PNDE_ERR_STATE, "IEKS: pnde_smooth the previous iterate before the next pnde_run (src/ieks.jl:38)");
    std::swap(h->hist, h->prev_hist);
    std::swap(h->smooth, h->prev_smooth);
    std::swap(h->n_saved, h->prev_n_saved);
    std::swap(h->final_diff, h->prev_final_diff);
    h->prev_max_saved = h->max_saved;
    CK(h->hist.ensure((size_t)h->n * (size_t)h->max_saved * h->ops->rec * 8), "alloc history (IEKS iterate)");
    CK(h->n_saved.ensure((size_t)h->n * 4), "alloc n_saved");
    CK(h->final_diff.ensure((size_t)h->n * h->nd * 8), "alloc final_diff");
    h->have_prev = true;
  }
  FilterParams fp;
  fill_filter_params(h, fp);
  if (h->ieks && h->have_prev) {
    fp.lin.hist = h->prev_hist.as<double>();
    fp.lin.smooth = h->prev_smooth.as<double>();
    fp.lin.n_saved = h->prev_n_saved.as<int>();
    fp.lin.final_diff = h->prev_final_diff.as<double>();
    fp.lin.max_saved = h->prev_max_saved;
    const int df = c.diffusion;
    fp.lin.calibrate = (df == PNDE_DIFF_FIXED || df == PNDE_DIFF_FIXED_MAP || df == PNDE_DIFF_FIXED_MV);
    fp.lin.is_mv = 0;
  }
  CK(cudaEventRecord(h->ev[0], h->stream), "event record");
  if (h->lorenz && c.alg == PNDE_ALG_EK1) {
    cudaError_t e = h->bigwork.ensure(big::big_work_bytes(h->d, c.order, c.adaptive != 0));
    if (e != cudaSuccess) {
      h->err = std::string("work-space allocation for the dense large-D path: ") + cudaGetErrorString(e);
      cudaGetLastError();
      return PNDE_ERR_ALLOC;
    }
    big::BigRunArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.n = fp.n;
    ba.d = h->d;
    ba.q = c.order;
    ba.diffusion = c.diffusion;
    ba.adaptive = c.adaptive;
    ba.u0 = fp.u0;
    ba.p = fp.p;
    ba.mean = fp.mean;
    ba.cov = fp.cov;
    ba.t_final = fp.t_final;
    ba.loglik = fp.loglik;
    ba.final_diff = fp.final_diff;
    ba.retcode = fp.retcode;
    ba.naccept = fp.naccept;
    ba.nreject = fp.nreject;
    ba.nf = fp.nf;
    ba.njacs = fp.njacs;
    ba.n_saved = fp.n_saved;
    ba.work = h->bigwork.p;
    ba.hist = (c.save_mode != PNDE_SAVE_FINAL) ? h->hist.as<double>() : nullptr;
    ba.max_saved = h->max_saved;
    ba.save_mode = c.save_mode;
    ba.save_stride = c.save_stride;
    ba.C = fp.C;
    ba.K = fp.K;
    long long nl = 0;
    CK(big::big_run(ba, h->stream, &nl), "dense large-D EK1 run");
    CK(cudaEventRecord(h->ev[1], h->stream), "event record");
    h->launches = nl;
    h->ran = true;
    h->smoothed = false;
    h->smooth_ms = 0.0;
    return PNDE_OK;
  } else if (h->lorenz) {
    LorenzParams lp;
    memset(&lp, 0, sizeof(lp));
    lp.n = fp.n;
    lp.d = h->d;
    lp.u0 = fp.u0;
    lp.p = fp.p;
    lp.mean = fp.mean;
    lp.cov = fp.cov;
    lp.t_final = fp.t_final;
    lp.loglik = fp.loglik;
    lp.final_diff = fp.final_diff;
    lp.retcode = fp.retcode;
    lp.naccept = fp.naccept;
    lp.nreject = fp.nreject;
    lp.nf = fp.nf;
    lp.n_saved = fp.n_saved;
    lp.hist = fp.hist;
    lp.max_saved = fp.max_saved;
    lp.save_mode = fp.save_mode;
    lp.save_stride = fp.save_stride;
    lp.diffusion = fp.diffusion;
    lp.adaptive = c.adaptive;
    lp.C = fp.C;
    lp.K = fp.K;
    CK(cudaMemsetAsync(h->njacs.p, 0, (size_t)h->n * 4, h->stream), "memset njacs");
    CK(launch_lorenz(c.order, lp, h->stream), "lorenz96 kernel launch");
  } else {
    CK(h->ops->launch_filter(h->ops, fp, c.adaptive != 0, h->stream), "filter kernel launch");
  }
  CK(cudaEventRecord(h->ev[1], h->stream), "event record");
  h->launches = 1;
  h->ran = true;
  h->smoothed = false;
  h->smooth_ms = 0.0;
  return PNDE_OK;
}

int pnde_smooth(pnde_handle* h) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "pnde_smooth: run the filter first");
  if (h->cfg.save_mode != PNDE_SAVE_EVERY)
    return h->fail(PNDE_ERR_STATE, "pnde_smooth needs save_mode = PNDE_SAVE_EVERY");
  if (h->multi()) {
    for (size_t k = 0; k < h->kids.size(); ++k) {
      if (h->kid_lo[k + 1] <= h->kid_lo[k]) continue;
      const int rc = pnde_smooth(h->kids[k]);
      if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
    }
    h->smoothed = true;
    return PNDE_OK;
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  if (h->lorenz) {
    if (h->cfg.alg != PNDE_ALG_EK0) return h->fail(PNDE_ERR_UNSUPPORTED, "smoothing is not built for the dense large-D EK1 path");
    cudaError_t e2 = h->smooth.ensure((size_t)h->n * (size_t)h->max_saved * lorenz_srec(h) * 8);
    if (e2 != cudaSuccess) {
      h->err = std::string("smoothed-history allocation: ") + cudaGetErrorString(e2);
      cudaGetLastError();
      return PNDE_ERR_ALLOC;
    }
    CK(h->sstatus.ensure((size_t)h->n * 4), "alloc smoother status");
    LorenzSmoothParams ls;
    memset(&ls, 0, sizeof(ls));
    ls.n = h->n;
    ls.d = h->d;
    ls.max_saved = h->max_saved;
    ls.n_saved = h->n_saved.as<int>();
    ls.hist = h->hist.as<double>();
    ls.smooth = h->smooth.as<double>();
    ls.final_diff = h->final_diff.as<double>();
    const int dfl = h->cfg.diffusion;
    ls.calibrate = (dfl == PNDE_DIFF_FIXED || dfl == PNDE_DIFF_FIXED_MAP);
    ls.status = h->sstatus.as<int>();
    ls.C = h->C;
    CK(cudaEventRecord(h->ev[2], h->stream), "event record");
    CK(launch_lorenz_smooth(h->cfg.order, ls, h->stream), "lorenz smoother launch");
    CK(cudaEventRecord(h->ev[3], h->stream), "event record");
    h->launches += 1;
    h->smoothed = true;
    return PNDE_OK;
  }
  const ModelOps* o = h->ops;
  cudaError_t e = h->smooth.ensure((size_t)h->n * (size_t)h->max_saved * o->srec * 8);
  if (e != cudaSuccess) {
    h->err = std::string("smoothed-history allocation: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return PNDE_ERR_ALLOC;
  }
  CK(h->sstatus.ensure((size_t)h->n * 4), "alloc smoother status");
  SmoothParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.n = h->n;
  sp.max_saved = h->max_saved;
  sp.n_saved = h->n_saved.as<int>();
  sp.hist = h->hist.as<double>();
  sp.smooth = h->smooth.as<double>();
  sp.final_diff = h->final_diff.as<double>();
  const int df = h->cfg.diffusion;
  sp.calibrate = (df == PNDE_DIFF_FIXED || df == PNDE_DIFF_FIXED_MAP || df == PNDE_DIFF_FIXED_MV);
  sp.is_mv = (df == PNDE_DIFF_DYNAMIC_MV || df == PNDE_DIFF_FIXED_MV);
  sp.status = h->sstatus.as<int>();
  sp.flags = h->cfg.flags;
  sp.C = h->C;
  CK(cudaEventRecord(h->ev[2], h->stream), "event record");
  CK(o->launch_smooth(o, sp, h->stream), "smoother kernel launch");
  CK(cudaEventRecord(h->ev[3], h->stream), "event record");
  h->launches += 1;
  h->smoothed = true;
  return PNDE_OK;
}

int pnde_synchronize(pnde_handle* h) {
  if (!h) return PNDE_ERR_ARG;
  if (h->multi()) {
    for (pnde_handle* k : h->kids) {
      const int rc = pnde_synchronize(k);
      if (rc != PNDE_OK) return h->fail(rc, k->err);
    }
    return PNDE_OK;
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_last_run_ms(pnde_handle* h, double* filter_ms, double* smooth_ms) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->multi()) {  // the devices work concurrently: max over shards
    double f = 0.0, s = 0.0;
    for (size_t k = 0; k < h->kids.size(); ++k) {
      if (h->kid_lo[k + 1] <= h->kid_lo[k]) continue;
      double fk = 0.0, sk = 0.0;
      const int rc = pnde_last_run_ms(h->kids[k], &fk, &sk);
      if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
      f = std::max(f, fk);
      s = std::max(s, sk);
    }
    if (filter_ms) *filter_ms = f;
    if (smooth_ms) *smooth_ms = s;
    return PNDE_OK;
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]), "event elapsed");
  h->filter_ms = ms;
  if (h->smoothed) {
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]), "event elapsed");
    h->smooth_ms = ms;
  }
  if (filter_ms) *filter_ms = h->filter_ms;
  if (smooth_ms) *smooth_ms = h->smooth_ms;
  return PNDE_OK;
}

int64_t pnde_last_launch_count(const pnde_handle* h) {
  if (!h) return 0;
  if (!h->multi()) return h->launches;
  int64_t tot = 0;
  for (size_t k = 0; k < h->kids.size(); ++k)
    if (h->kid_lo.size() > k + 1 && h->kid_lo[k + 1] > h->kid_lo[k]) tot += h->kids[k]->launches;
  return tot;
}

static int solve_ensemble_impl(pnde_handle* h, int64_t n_traj, const double* u0, const double* p, int64_t src_pitch) {
  if (h && h->multi()) {
    if (n_traj <= 0 || !u0 || (!p && h->np > 0)) return h->fail(PNDE_ERR_ARG, "pnde_solve_ensemble: bad arguments");
    partition(h, n_traj);
    h->n = n_traj;
    const int rc = for_each_kid(h, [&](int k) {
      const long long lo = h->kid_lo[(size_t)k], cnt = h->kid_lo[(size_t)k + 1] - lo;
      return solve_ensemble_impl(h->kids[(size_t)k], cnt, u0 + lo, p ? p + lo : nullptr, src_pitch);
    });
    h->ran = (rc == PNDE_OK);
    h->smoothed = h->ran && h->cfg.smooth;
    return rc;
  }
  int rc = upload_impl(h, n_traj, u0, p, src_pitch);
  if (rc != PNDE_OK) return rc;
  const int iterations = h->ieks ? (h->cfg.ieks_iterations > 0 ? h->cfg.ieks_iterations : 10) : 1;
  for (int it = 0; it < iterations; ++it) {
    rc = pnde_run(h);
    if (rc != PNDE_OK) return rc;
    if (h->cfg.smooth) {
      rc = pnde_smooth(h);
      if (rc != PNDE_OK) return rc;
    }
  }
  return pnde_synchronize(h);
}

int pnde_solve_ensemble(pnde_handle* h, int64_t n_traj, const double* u0, const double* p) {
  return solve_ensemble_impl(h, n_traj, u0, p, n_traj);
}

static int get_final_impl(pnde_handle* h, double* mean, double* cov, double* t_final, double* loglik, int64_t dst_pitch);

// Pipelined variant of solve + get_final for the thread-per-trajectory models: the ensemble is cut into slices;
// the device-to-host copy of slice k (auxiliary stream) overlaps the filter kernel of slice k+1.
// pitch: elements between consecutive rows of the caller's input AND output arrays (the total ensemble size)
static int solve_to_host_impl(pnde_handle* h, int64_t n_traj, const double* u0, const double* p, double* mean,
                              double* cov, double* t_final, double* loglik, int64_t pitch_n) {
  if (!h) return PNDE_ERR_ARG;
  if (h->multi()) {
    if (n_traj <= 0 || !u0 || (!p && h->np > 0)) return h->fail(PNDE_ERR_ARG, "pnde_solve_ensemble_to_host: bad arguments");
    partition(h, n_traj);
    h->n = n_traj;
    const int rc = for_each_kid(h, [&](int k) {
      const long long lo = h->kid_lo[(size_t)k], cnt = h->kid_lo[(size_t)k + 1] - lo;
      return solve_to_host_impl(h->kids[(size_t)k], cnt, u0 + lo, p ? p + lo : nullptr, mean ? mean + lo : nullptr,
                                cov ? cov + lo : nullptr, t_final ? t_final + lo : nullptr,
                                loglik ? loglik + lo : nullptr, pitch_n);
    });
    h->ran = (rc == PNDE_OK);
    h->smoothed = h->ran && h->cfg.smooth;
    return rc;
  }
  if (!h->ops || h->cfg.smooth || h->cfg.save_mode != PNDE_SAVE_FINAL) {
    // models without slices (Lorenz-96 paths) or runs that keep history: plain sequence
    int rc = solve_ensemble_impl(h, n_traj, u0, p, pitch_n);
    if (rc != PNDE_OK) return rc;
    return get_final_impl(h, mean, cov, t_final, loglik, pitch_n);
  }
  int rc = upload_impl(h, n_traj, u0, p, pitch_n);
  if (rc != PNDE_OK) return rc;
  if (!h->copy_stream) {
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), "copy stream");
    CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking), "second compute stream");
    for (int i = 0; i < 3; ++i) CK(cudaEventCreateWithFlags(&h->slice_ev[i], cudaEventDisableTiming), "slice event");
  }
  FilterParams fp;
  fill_filter_params(h, fp);
  const long long n = h->n;
  const int nslices = n >= 65536 ? 8 : 1;
  const long long per = ((n + nslices - 1) / nslices + 127) / 128 * 128;
  CK(cudaEventRecord(h->ev[0], h->stream), "event record");
  // slices alternate between two compute streams so that the tail wave of one slice is filled by the next
  CK(cudaEventRecord(h->slice_ev[2], h->stream), "upload event");
  CK(cudaStreamWaitEvent(h->stream2, h->slice_ev[2], 0), "upload wait");
  long long launches = 0;
  for (int k = 0; k < nslices; ++k) {
    const long long lo = k * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    cudaStream_t cs = (k & 1) ? h->stream2 : h->stream;
    fp.first = lo;
    fp.count = hi - lo;
    CK(h->ops->launch_filter(h->ops, fp, h->cfg.adaptive != 0, cs), "filter kernel launch");
    ++launches;
    CK(cudaEventRecord(h->slice_ev[k & 1], cs), "slice event record");
    CK(cudaStreamWaitEvent(h->copy_stream, h->slice_ev[k & 1], 0), "slice wait");
    const size_t w = (size_t)(hi - lo) * 8, pitch = (size_t)n * 8, dpitch = (size_t)pitch_n * 8;
    if (mean)
      CK(cudaMemcpy2DAsync(mean + lo, dpitch, h->mean.as<double>() + lo, pitch, w, h->D, cudaMemcpyDeviceToHost, h->copy_stream), "D2H mean");
    if (cov)
      CK(cudaMemcpy2DAsync(cov + lo, dpitch, h->cov.as<double>() + lo, pitch, w, h->ncov, cudaMemcpyDeviceToHost, h->copy_stream), "D2H cov");
    if (t_final) CK(cudaMemcpyAsync(t_final + lo, h->t_final.as<double>() + lo, w, cudaMemcpyDeviceToHost, h->copy_stream), "D2H t");
    if (loglik) CK(cudaMemcpyAsync(loglik + lo, h->loglik.as<double>() + lo, w, cudaMemcpyDeviceToHost, h->copy_stream), "D2H loglik");
  }
  CK(cudaEventRecord(h->slice_ev[2], h->stream2), "join event");
  CK(cudaStreamWaitEvent(h->stream, h->slice_ev[2], 0), "join wait");
  CK(cudaEventRecord(h->ev[1], h->stream), "event record");
  h->launches = launches;
  h->ran = true;
  h->smoothed = false;
  h->smooth_ms = 0.0;
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  CK(cudaStreamSynchronize(h->copy_stream), "copy stream synchronize");
  return PNDE_OK;
}

int pnde_solve_ensemble_to_host(pnde_handle* h, int64_t n_traj, const double* u0, const double* p, double* mean,
                                double* cov, double* t_final, double* loglik) {
  return solve_to_host_impl(h, n_traj, u0, p, mean, cov, t_final, loglik, n_traj);
}

static int fetch_i32(pnde_handle* h, const DevBuf& b, std::vector<int>& out) {
  out.resize((size_t)h->n);
  CK(cudaMemcpyAsync(out.data(), b.p, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream), "D2H counts");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_query_sizes(pnde_handle* h, int64_t* n_saved_total, int64_t* max_saved) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->multi()) {
    int64_t tot = 0, mx = 0;
    for (size_t k = 0; k < h->kids.size(); ++k) {
      if (h->kid_lo[k + 1] <= h->kid_lo[k]) continue;
      int64_t tk = 0, mk = 0;
      const int rc = pnde_query_sizes(h->kids[k], &tk, &mk);
      if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
      tot += tk;
      mx = std::max(mx, mk);
    }
    if (n_saved_total) *n_saved_total = tot;
    if (max_saved) *max_saved = mx;
    return PNDE_OK;
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  std::vector<int> ns;
  int rc = fetch_i32(h, h->n_saved, ns);
  if (rc != PNDE_OK) return rc;
  long long tot = 0;
  int mx = 0;
  for (int v : ns) {
    tot += v;
    if (v > mx) mx = v;
  }
  if (n_saved_total) *n_saved_total = tot;
  if (max_saved) *max_saved = mx;
  return PNDE_OK;
}

int pnde_get_counts(pnde_handle* h, int64_t* naccept, int64_t* nreject, int64_t* nf, int64_t* njacs, int32_t* retcode,
                    int64_t* n_saved) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->multi())
    return for_each_kid(h, [&](int k) {
      const long long lo = h->kid_lo[(size_t)k];
      return pnde_get_counts(h->kids[(size_t)k], naccept ? naccept + lo : nullptr, nreject ? nreject + lo : nullptr,
                             nf ? nf + lo : nullptr, njacs ? njacs + lo : nullptr, retcode ? retcode + lo : nullptr,
                             n_saved ? n_saved + lo : nullptr);
    });
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  std::vector<int> tmp;
  struct {
    int64_t* dst;
    const DevBuf* src;
  } items[] = {{naccept, &h->naccept}, {nreject, &h->nreject}, {nf, &h->nf}, {njacs, &h->njacs}, {n_saved, &h->n_saved}};
  for (auto& it : items) {
    if (!it.dst) continue;
    int rc = fetch_i32(h, *it.src, tmp);
    if (rc != PNDE_OK) return rc;
    for (long long i = 0; i < h->n; ++i) it.dst[i] = tmp[(size_t)i];
  }
  if (retcode) {
    CK(cudaMemcpyAsync(retcode, h->retcode.p, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream), "D2H retcode");
    CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  }
  return PNDE_OK;
}

static int get_final_impl(pnde_handle* h, double* mean, double* cov, double* t_final, double* loglik, int64_t dst_pitch) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->multi())
    return for_each_kid(h, [&](int k) {
      const long long lo = h->kid_lo[(size_t)k];
      return get_final_impl(h->kids[(size_t)k], mean ? mean + lo : nullptr, cov ? cov + lo : nullptr,
                            t_final ? t_final + lo : nullptr, loglik ? loglik + lo : nullptr, dst_pitch);
    });
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const size_t n = (size_t)h->n, dp = (size_t)dst_pitch * 8;
  if (mean) CK(cudaMemcpy2DAsync(mean, dp, h->mean.p, n * 8, n * 8, (size_t)h->D, cudaMemcpyDeviceToHost, h->stream), "D2H mean");
  if (cov) CK(cudaMemcpy2DAsync(cov, dp, h->cov.p, n * 8, n * 8, (size_t)h->ncov, cudaMemcpyDeviceToHost, h->stream), "D2H cov");
  if (t_final) CK(cudaMemcpyAsync(t_final, h->t_final.p, n * 8, cudaMemcpyDeviceToHost, h->stream), "D2H t");
  if (loglik) CK(cudaMemcpyAsync(loglik, h->loglik.p, n * 8, cudaMemcpyDeviceToHost, h->stream), "D2H loglik");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_get_final(pnde_handle* h, double* mean, double* cov, double* t_final, double* loglik) {
  return get_final_impl(h, mean, cov, t_final, loglik, h ? h->n : 0);
}

}  // extern "C"  (template below)

// Multi-device form of the CSR getters: the trajectory range is cut at the shard boundaries, every child writes its
// part of the flat outputs (trajectory order is shard order), offsets are rebased.  `call(kid, lo, hi, offs, base)`
// runs the child getter for its local range with output pointers advanced by `base` saved states.
template <class Fn>
static int multi_csr(pnde_handle* h, int64_t tb, int64_t te, int64_t* offsets, Fn call) {
  if (tb < 0 || te > h->n || tb >= te || !offsets) return h->fail(PNDE_ERR_ARG, "bad trajectory range");
  int64_t base = 0;
  offsets[0] = 0;
  std::vector<int64_t> offs;
  for (size_t k = 0; k < h->kids.size(); ++k) {
    const int64_t lo = std::max<int64_t>(tb, h->kid_lo[k]), hi = std::min<int64_t>(te, h->kid_lo[k + 1]);
    if (lo >= hi) continue;
    offs.assign((size_t)(hi - lo) + 1, 0);
    const int rc = call(h->kids[k], lo - h->kid_lo[k], hi - h->kid_lo[k], offs.data(), base);
    if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
    for (int64_t i = 0; i <= hi - lo; ++i) offsets[lo - tb + i] = base + offs[(size_t)i];
    base += offs[(size_t)(hi - lo)];
  }
  return PNDE_OK;
}

// History of the large-D dense path (records [t, diffusion, u[d], diag(Sigma_u)[d]], layout [slot][2 + 2d][n]) ->
// the caller's CSR arrays; static diffusion models: calibrated by the final global diffusion like every other path
// (src/integrator_utils.jl:4-18).  One thread per (saved state, dimension).
static __global__ void __launch_bounds__(256) bigdense_convert_kernel(const double* hist, const double* final_diff,
                                                                       const int* naccept, const long long* offsets,
                                                                       long long n, long long tb, long long te, int d,
                                                                       int calibrate, double* t, double* u, double* var) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = offsets[te - tb];
  if (idx >= total * d) return;
  const long long g = idx / d;
  const int b = (int)(idx % d);
  long long lo = 0, hi = te - tb;  // trajectory of state g: last i with offsets[i] <= g
  while (hi - lo > 1) {
    const long long mid = (lo + hi) / 2;
    if (offsets[mid] <= g) lo = mid; else hi = mid;
  }
  const long long tr = tb + lo, slot = g - offsets[lo];
  const int REC = 2 + 2 * d;
  const double* r = hist + (slot * REC) * n + tr;
  const double cal = (calibrate && naccept[tr] > 0) ? final_diff[tr] : 1.0;
  if (b == 0) t[g] = r[0];
  u[g * d + b] = r[(long long)(2 + b) * n];
  var[g * d + b] = cal * r[(long long)(2 + d + b) * n];
}

extern "C" {

static int get_history_impl(pnde_handle* h, int32_t which, int64_t tb, int64_t te, int64_t* offsets, double* t,
                            double* mean, double* cov, double* diffusion, bool marginals, double* sqrt_out = nullptr) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->cfg.save_mode == PNDE_SAVE_FINAL) return h->fail(PNDE_ERR_STATE, "no history was saved (save_mode = final)");
  if (h->multi()) {
    const bool kron = h->lorenz && h->cfg.alg == PNDE_ALG_EK0, bigd = h->lorenz && h->cfg.alg == PNDE_ALG_EK1;
    const int64_t DM = marginals ? h->d : h->D;
    const int64_t NC = kron ? (marginals ? 1 : h->ncov) : ((bigd && marginals) ? h->d : DM * (DM + 1) / 2);
    const int df0 = h->cfg.diffusion;
    const int64_t ndo = (df0 == PNDE_DIFF_DYNAMIC_MV || df0 == PNDE_DIFF_FIXED_MV) ? h->d : 1;
    return multi_csr(h, tb, te, offsets, [&](pnde_handle* kid, int64_t lo, int64_t hi, int64_t* offs, int64_t base) {
      return get_history_impl(kid, which, lo, hi, offs, t ? t + base : nullptr, mean ? mean + base * DM : nullptr,
                              cov ? cov + base * NC : nullptr, diffusion ? diffusion + base * ndo : nullptr, marginals,
                              sqrt_out ? sqrt_out + base * h->D * h->D : nullptr);
    });
  }
  if (which == PNDE_HIST_SMOOTHED && !h->smoothed) return h->fail(PNDE_ERR_STATE, "history has not been smoothed");
  if (which != PNDE_HIST_FILTERED && which != PNDE_HIST_SMOOTHED) return h->fail(PNDE_ERR_ARG, "bad 'which'");
  if (tb < 0 || te > h->n || tb >= te || !offsets) return h->fail(PNDE_ERR_ARG, "bad trajectory range");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const ModelOps* o = h->ops;
  const bool bigdense = !o && h->lorenz && h->cfg.alg == PNDE_ALG_EK1;
  if (bigdense && (!marginals || which != PNDE_HIST_FILTERED))
    return h->fail(PNDE_ERR_UNSUPPORTED, "the large-D dense path saves the solution marginals only: use pnde_get_marginals(PNDE_HIST_FILTERED)");
  if (!o && !h->lorenz) return h->fail(PNDE_ERR_UNSUPPORTED, "no history on this path");
  if (!o && sqrt_out) return h->fail(PNDE_ERR_UNSUPPORTED, "the large-d path returns Ctilde, not D x D factors");
  std::vector<int> ns;
  int rc = fetch_i32(h, h->n_saved, ns);
  if (rc != PNDE_OK) return rc;
  const long long ntr = te - tb;
  std::vector<long long> off((size_t)ntr + 1);
  off[0] = 0;
  for (long long i = 0; i < ntr; ++i) off[(size_t)i + 1] = off[(size_t)i] + ns[(size_t)(tb + i)];
  const long long total = off[(size_t)ntr];
  for (long long i = 0; i <= ntr; ++i) offsets[i] = off[(size_t)i];
  if (bigdense) {
    // large-D dense path: u [total][d], cov_u = diag(Sigma_u) [total][d] (the d x d matrix is never formed per state)
    const size_t dd = (size_t)h->d, nml = (size_t)total * dd;
    CK(h->scratch_off.ensure(((size_t)ntr + 1) * 8), "alloc offsets");
    CK(h->scratch_out.ensure(((size_t)total + 2 * nml) * 8 + 64), "alloc history staging");
    CK(cudaMemcpyAsync(h->scratch_off.p, off.data(), ((size_t)ntr + 1) * 8, cudaMemcpyHostToDevice, h->stream), "H2D offsets");
    double* lt = h->scratch_out.as<double>();
    const int dfl = h->cfg.diffusion;
    int calibrate = (dfl == PNDE_DIFF_FIXED || dfl == PNDE_DIFF_FIXED_MAP);
    if (h->cfg.flags & PNDE_FLAG_REFERENCE_QUIRKS) calibrate = 0;  // quirk (a): sol.pu of an un-smoothed solve
    if (total > 0) {
      const long long work = total * (long long)dd;
      bigdense_convert_kernel<<<(unsigned)((work + 255) / 256), 256, 0, h->stream>>>(
          h->hist.as<double>(), h->final_diff.as<double>(), h->naccept.as<int>(), h->scratch_off.as<long long>(), h->n, tb, te,
          h->d, calibrate, lt, lt + total, lt + total + nml);
      CK(cudaGetLastError(), "marginal history converter launch");
    }
    if (t) CK(cudaMemcpyAsync(t, lt, (size_t)total * 8, cudaMemcpyDeviceToHost, h->stream), "D2H t");
    if (mean) CK(cudaMemcpyAsync(mean, lt + total, nml * 8, cudaMemcpyDeviceToHost, h->stream), "D2H u");
    if (cov) CK(cudaMemcpyAsync(cov, lt + total + nml, nml * 8, cudaMemcpyDeviceToHost, h->stream), "D2H var");
    CK(cudaStreamSynchronize(h->stream), "stream synchronize");
    return PNDE_OK;
  }
  if (!o) {
    // large-d Kronecker path: mean [total][D] (marginals: [total][d]), cov = packed Ctilde [total][(q+1)(q+2)/2]
    // (marginals: ONE entry per state, Ctilde[0][0]: Sigma_u = Ctilde[0][0] I_d), diffusion [total]
    const size_t DMl = marginals ? (size_t)h->d : (size_t)h->D, NCl = marginals ? 1 : (size_t)h->ncov;
    const size_t nml = (size_t)total * DMl, ncl = (size_t)total * NCl;
    CK(h->scratch_off.ensure(((size_t)ntr + 1) * 8), "alloc offsets");
    CK(h->scratch_out.ensure(((size_t)total * 2 + nml + ncl) * 8 + 64), "alloc history staging");
    CK(cudaMemcpyAsync(h->scratch_off.p, off.data(), ((size_t)ntr + 1) * 8, cudaMemcpyHostToDevice, h->stream), "H2D offsets");
    double* lt = h->scratch_out.as<double>();
    LorenzConvertParams lc;
    memset(&lc, 0, sizeof(lc));
    lc.n = h->n;
    lc.traj_begin = tb;
    lc.traj_end = te;
    lc.max_saved = h->max_saved;
    lc.d = h->d;
    lc.n_saved = h->n_saved.as<int>();
    lc.offsets = h->scratch_off.as<long long>();
    lc.hist = h->hist.as<double>();
    lc.smooth = h->smooth.as<double>();
    lc.final_diff = h->final_diff.as<double>();
    lc.which = which;
    const int dfl = h->cfg.diffusion;
    lc.calibrate = (dfl == PNDE_DIFF_FIXED || dfl == PNDE_DIFF_FIXED_MAP);
    if (marginals && which == PNDE_HIST_FILTERED && !h->cfg.smooth && (h->cfg.flags & PNDE_FLAG_REFERENCE_QUIRKS)) lc.calibrate = 0;
    lc.marginals = marginals ? 1 : 0;
    lc.t = lt;
    lc.diffusion = lt + total;
    lc.mean = lt + 2 * total;
    lc.cov = lc.mean + nml;
    CK(launch_lorenz_convert(h->cfg.order, lc, h->stream), "lorenz convert launch");
    if (t) CK(cudaMemcpyAsync(t, lc.t, (size_t)total * 8, cudaMemcpyDeviceToHost, h->stream), "D2H t");
    if (mean) CK(cudaMemcpyAsync(mean, lc.mean, nml * 8, cudaMemcpyDeviceToHost, h->stream), "D2H mean");
    if (cov) CK(cudaMemcpyAsync(cov, lc.cov, ncl * 8, cudaMemcpyDeviceToHost, h->stream), "D2H cov");
    if (diffusion) CK(cudaMemcpyAsync(diffusion, lc.diffusion, (size_t)total * 8, cudaMemcpyDeviceToHost, h->stream), "D2H diffusion");
    CK(cudaStreamSynchronize(h->stream), "stream synchronize");
    return PNDE_OK;
  }
  const int df = h->cfg.diffusion;
  const bool is_mv = (df == PNDE_DIFF_DYNAMIC_MV || df == PNDE_DIFF_FIXED_MV);
  const int nd_out = is_mv ? o->d : 1;
  const int DM = marginals ? o->d : o->D;
  const size_t nm = (size_t)total * DM, nc = (size_t)total * (size_t)(DM * (DM + 1) / 2), ndif = (size_t)total * nd_out;
  const size_t nsq = sqrt_out ? (size_t)total * o->D * o->D : 0;
  CK(h->scratch_off.ensure(((size_t)ntr + 1) * 8), "alloc offsets");
  CK(h->scratch_out.ensure(((size_t)total + nm + nc + ndif + nsq) * 8 + 64), "alloc history staging");
  CK(cudaMemcpyAsync(h->scratch_off.p, off.data(), ((size_t)ntr + 1) * 8, cudaMemcpyHostToDevice, h->stream), "H2D offsets");
  double* dt_ = h->scratch_out.as<double>();
  double* dmean = dt_ + total;
  double* dcov = dmean + nm;
  double* ddif = dcov + nc;
  ConvertParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.n = h->n;
  cp.traj_begin = tb;
  cp.traj_end = te;
  cp.max_saved = h->max_saved;
  cp.n_saved = h->n_saved.as<int>();
  cp.offsets = h->scratch_off.as<long long>();
  cp.hist = h->hist.as<double>();
  cp.smooth = h->smooth.as<double>();
  cp.final_diff = h->final_diff.as<double>();
  cp.which = which;
  cp.calibrate = (df == PNDE_DIFF_FIXED || df == PNDE_DIFF_FIXED_MAP || df == PNDE_DIFF_FIXED_MV);
  // PNDE_FLAG_REFERENCE_QUIRKS (a): sol.pu of an un-smoothed solve keeps the uncalibrated covariances
  if (marginals && which == PNDE_HIST_FILTERED && !h->cfg.smooth && (h->cfg.flags & PNDE_FLAG_REFERENCE_QUIRKS))
    cp.calibrate = 0;
  cp.is_mv = is_mv;
  cp.marginals = marginals ? 1 : 0;
  cp.t = dt_;
  cp.mean = dmean;
  cp.cov = dcov;
  cp.diffusion = ddif;
  cp.nd_out = nd_out;
  cp.sqrt = sqrt_out ? ddif + ndif : nullptr;
  if (sqrt_out) {  // only the factor was asked for
    cp.mean = nullptr;
    cp.cov = nullptr;
    cp.diffusion = nullptr;
  }
  CK(o->launch_convert(o, cp, h->stream), "convert kernel launch");
  if (sqrt_out) CK(cudaMemcpyAsync(sqrt_out, cp.sqrt, nsq * 8, cudaMemcpyDeviceToHost, h->stream), "D2H sqrt");
  if (t) CK(cudaMemcpyAsync(t, dt_, (size_t)total * 8, cudaMemcpyDeviceToHost, h->stream), "D2H t");
  if (mean) CK(cudaMemcpyAsync(mean, dmean, nm * 8, cudaMemcpyDeviceToHost, h->stream), "D2H mean");
  if (cov) CK(cudaMemcpyAsync(cov, dcov, nc * 8, cudaMemcpyDeviceToHost, h->stream), "D2H cov");
  if (diffusion) CK(cudaMemcpyAsync(diffusion, ddif, ndif * 8, cudaMemcpyDeviceToHost, h->stream), "D2H diffusion");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_get_history(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end, int64_t* offsets, double* t,
                     double* mean, double* cov, double* diffusion) {
  return get_history_impl(h, which, traj_begin, traj_end, offsets, t, mean, cov, diffusion, false);
}

int pnde_get_history_sqrt(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end, int64_t* offsets,
                          double* sqrt_out) {
  if (!sqrt_out) return h ? h->fail(PNDE_ERR_ARG, "pnde_get_history_sqrt: null output") : PNDE_ERR_ARG;
  return get_history_impl(h, which, traj_begin, traj_end, offsets, nullptr, nullptr, nullptr, nullptr, false, sqrt_out);
}

int pnde_get_marginals(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end, int64_t* offsets, double* t,
                       double* u, double* cov_u) {
  return get_history_impl(h, which, traj_begin, traj_end, offsets, t, u, cov_u, nullptr, true);
}

int pnde_sample(pnde_handle* h, int64_t tb, int64_t te, int32_t n_samples, uint64_t seed, int64_t* offsets, double* t,
                double* samples) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->cfg.save_mode != PNDE_SAVE_EVERY) return h->fail(PNDE_ERR_STATE, "pnde_sample needs save_mode = PNDE_SAVE_EVERY");
  if (tb < 0 || te > h->n || tb >= te || !offsets || n_samples < 1) return h->fail(PNDE_ERR_ARG, "bad arguments");
  if (h->lorenz) return h->fail(PNDE_ERR_UNSUPPORTED, "sampling is not built for the large-d paths");
  if (h->multi()) {
    // the generator is keyed by the trajectory index: children are handed their global offset so that the draws do
    // not depend on how the ensemble was sharded
    return multi_csr(h, tb, te, offsets, [&](pnde_handle* kid, int64_t lo, int64_t hi, int64_t* offs, int64_t base) {
      return pnde_sample(kid, lo, hi, n_samples, seed, offs, t ? t + base : nullptr,
                         samples ? samples + base * n_samples * h->D : nullptr);
    });
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const ModelOps* o = h->ops;
  std::vector<int> ns;
  int rc = fetch_i32(h, h->n_saved, ns);
  if (rc != PNDE_OK) return rc;
  const long long ntr = te - tb;
  std::vector<long long> off((size_t)ntr + 1);
  off[0] = 0;
  for (long long i = 0; i < ntr; ++i) off[(size_t)i + 1] = off[(size_t)i] + ns[(size_t)(tb + i)];
  const long long total = off[(size_t)ntr];
  for (long long i = 0; i <= ntr; ++i) offsets[i] = off[(size_t)i];
  const size_t nout = (size_t)total * n_samples * o->D;
  CK(h->scratch_off.ensure(((size_t)ntr + 1) * 8), "alloc offsets");
  CK(h->scratch_out.ensure(((size_t)total + nout) * 8 + 64), "alloc sample staging");
  CK(cudaMemcpyAsync(h->scratch_off.p, off.data(), ((size_t)ntr + 1) * 8, cudaMemcpyHostToDevice, h->stream), "H2D offsets");
  const int df = h->cfg.diffusion;
  double* dt_ = h->scratch_out.as<double>();
  double* dout = dt_ + total;
  // time stamps through the converter (marginals mode with null outputs writes only t)
  ConvertParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.n = h->n;
  cp.traj_begin = tb;
  cp.traj_end = te;
  cp.max_saved = h->max_saved;
  cp.n_saved = h->n_saved.as<int>();
  cp.offsets = h->scratch_off.as<long long>();
  cp.hist = h->hist.as<double>();
  cp.final_diff = h->final_diff.as<double>();
  cp.which = 0;
  cp.marginals = 1;
  cp.t = dt_;
  cp.nd_out = 1;
  CK(o->launch_convert(o, cp, h->stream), "convert kernel launch");
  // the per-interval backward kernels (stage-1 sweep, once per trajectory and interval) live in a scratch block;
  // large ranges are processed in chunks of trajectories so that the block stays below 2 GiB
  int mxs = 0;
  for (long long i = tb; i < te; ++i) mxs = std::max(mxs, ns[(size_t)i]);
  const size_t per_traj = (size_t)std::max(mxs - 1, 1) * (size_t)o->sample_len * 8;
  const long long chunk = std::max<long long>(1, std::min<long long>(ntr, (long long)(((size_t)2 << 30) / per_traj)));
  CK(h->sample_scratch.ensure((size_t)chunk * per_traj), "alloc sampler scratch");
  SampleParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.n = h->n;
  sp.n_saved = h->n_saved.as<int>();
  sp.hist = h->hist.as<double>();
  sp.final_diff = h->final_diff.as<double>();
  sp.calibrate = (df == PNDE_DIFF_FIXED || df == PNDE_DIFF_FIXED_MAP || df == PNDE_DIFF_FIXED_MV);
  sp.is_mv = (df == PNDE_DIFF_DYNAMIC_MV || df == PNDE_DIFF_FIXED_MV);
  sp.n_samples = n_samples;
  sp.seed = seed;
  sp.key_offset = h->global_lo;
  sp.scratch = h->sample_scratch.as<double>();
  sp.C = h->C;
  for (long long c0 = tb; c0 < te; c0 += chunk) {
    sp.traj_begin = c0;
    sp.traj_end = std::min<long long>(te, c0 + chunk);
    sp.max_saved = std::max(mxs, 1);  // stride of the scratch: intervals per trajectory + 1
    sp.offsets = h->scratch_off.as<long long>() + (c0 - tb);
    sp.out = dout;
    CK(o->launch_sample(o, sp, h->stream), "sample kernel launch");
    h->launches += 2;
  }
  if (t) CK(cudaMemcpyAsync(t, dt_, (size_t)total * 8, cudaMemcpyDeviceToHost, h->stream), "D2H t");
  if (samples) CK(cudaMemcpyAsync(samples, dout, nout * 8, cudaMemcpyDeviceToHost, h->stream), "D2H samples");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_dense_sample(pnde_handle* h, int64_t tb, int64_t te, int64_t n_t, const double* tq, int32_t n_samples,
                      uint64_t seed, double* samples) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->cfg.save_mode != PNDE_SAVE_EVERY) return h->fail(PNDE_ERR_STATE, "pnde_dense_sample needs save_mode = PNDE_SAVE_EVERY");
  if (tb < 0 || te > h->n || tb >= te || n_t < 1 || !tq || n_samples < 1 || !samples) return h->fail(PNDE_ERR_ARG, "bad arguments");
  for (int64_t k = 0; k + 1 < n_t; ++k)
    if (!(tq[k + 1] >= tq[k])) return h->fail(PNDE_ERR_ARG, "pnde_dense_sample: the time grid must be non-decreasing");
  if (h->lorenz) return h->fail(PNDE_ERR_UNSUPPORTED, "sampling is not built for the large-d paths");
  if (h->multi()) {
    const size_t per = (size_t)n_t * n_samples * h->D;
    for (size_t k = 0; k < h->kids.size(); ++k) {
      const int64_t lo = std::max<int64_t>(tb, h->kid_lo[k]), hi = std::min<int64_t>(te, h->kid_lo[k + 1]);
      if (lo >= hi) continue;
      const int rc = pnde_dense_sample(h->kids[k], lo - h->kid_lo[k], hi - h->kid_lo[k], n_t, tq, n_samples, seed,
                                       samples + (size_t)(lo - tb) * per);
      if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
    }
    return PNDE_OK;
  }
  if (!h->ops) return h->fail(PNDE_ERR_UNSUPPORTED, "pnde_dense_sample: not built for this path");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const ModelOps* o = h->ops;
  // scratch record of one grid interval: DenseSamplePrep<M>::LEN = 2 D + NF (NP + DCOV + 2 DCOV^2)
  const size_t Dq = (size_t)o->q + 1;
  const size_t dcov = o->ek1 ? (size_t)o->D : Dq, npk = dcov * (dcov + 1) / 2;
  const size_t nf = o->ek1 ? 1 : ((size_t)o->srec - (size_t)o->D - (size_t)o->d) / npk;
  const size_t len = 2 * (size_t)o->D + nf * (npk + dcov + 2 * dcov * dcov);
  const long long ntr = te - tb;
  const size_t per_traj = (size_t)std::max<int64_t>(n_t - 1, 1) * len * 8;
  const long long chunk = std::max<long long>(1, std::min<long long>(ntr, (long long)(((size_t)2 << 30) / per_traj)));
  CK(h->sample_scratch.ensure((size_t)chunk * per_traj), "alloc sampler scratch");
  const size_t nout = (size_t)ntr * n_t * n_samples * o->D;
  CK(h->scratch_out.ensure(((size_t)n_t + nout) * 8 + 64), "alloc sample staging");
  double* dtq = h->scratch_out.as<double>();
  double* dout = dtq + n_t;
  CK(cudaMemcpyAsync(dtq, tq, (size_t)n_t * 8, cudaMemcpyHostToDevice, h->stream), "H2D time grid");
  const int df = h->cfg.diffusion;
  SampleParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.n = h->n;
  sp.n_saved = h->n_saved.as<int>();
  sp.hist = h->hist.as<double>();
  sp.final_diff = h->final_diff.as<double>();
  sp.calibrate = (df == PNDE_DIFF_FIXED || df == PNDE_DIFF_FIXED_MAP || df == PNDE_DIFF_FIXED_MV);
  sp.is_mv = (df == PNDE_DIFF_DYNAMIC_MV || df == PNDE_DIFF_FIXED_MV);
  sp.n_samples = n_samples;
  sp.seed = seed;
  sp.key_offset = h->global_lo;
  sp.scratch = h->sample_scratch.as<double>();
  sp.max_saved = h->max_saved;
  sp.tq = dtq;
  sp.n_t = n_t;
  sp.C = h->C;
  for (long long c0 = tb; c0 < te; c0 += chunk) {
    sp.traj_begin = c0;
    sp.traj_end = std::min<long long>(te, c0 + chunk);
    sp.out = dout + (size_t)(c0 - tb) * n_t * n_samples * o->D;
    CK(o->launch_sample(o, sp, h->stream), "dense sample kernel launch");
    h->launches += 2;
  }
  CK(cudaMemcpyAsync(samples, dout, nout * 8, cudaMemcpyDeviceToHost, h->stream), "D2H samples");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_step_from_state(pnde_handle* h, int64_t n, const double* mean, const double* sqrt_in, const double* t,
                         const double* dt, const double* p, const double* uprev, double* mean_out, double* cov_out,
                         double* sigma2, double* eest, double* u_out, double* quad_logdet, int32_t* status) {
  if (!h) return PNDE_ERR_ARG;
  if (n <= 0 || !mean || !sqrt_in || !dt || (!p && h->np > 0) || !status)
    return h->fail(PNDE_ERR_ARG, "pnde_step_from_state: bad arguments");
  if (h->multi()) {
    h->kid_lo.assign(h->kids.size() + 1, 0);  // a stateless call: run it on the first device
    return pnde_step_from_state(h->kids[0], n, mean, sqrt_in, t, dt, p, uprev, mean_out, cov_out, sigma2, eest, u_out,
                                quad_logdet, status);
  }
  if (!h->ops || !h->ops->launch_step)
    return h->fail(PNDE_ERR_UNSUPPORTED, "pnde_step_from_state is built for the catalogue models (thread-per-trajectory paths)");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const ModelOps* o = h->ops;
  const size_t N = (size_t)n, D = (size_t)o->D, NC = D * (D + 1) / 2, npar = (size_t)(o->np > 0 ? o->np : 1);
  // one staging block: inputs then outputs
  const size_t in_len = D + D * D + 2 + npar + (size_t)o->d, out_len = D + NC + (size_t)o->nd + 1 + (size_t)o->d + 2;
  CK(h->scratch_out.ensure((in_len + out_len) * N * 8 + N * 4 + 64), "alloc step staging");
  double* b = h->scratch_out.as<double>();
  double *dmean = b, *dsq = dmean + D * N, *dt_ = dsq + D * D * N, *ddt = dt_ + N, *dp = ddt + N, *dup = dp + npar * N;
  double *omean = dup + (size_t)o->d * N, *ocov = omean + D * N, *osig = ocov + NC * N, *oee = osig + (size_t)o->nd * N,
         *ou = oee + N, *oql = ou + (size_t)o->d * N;
  int* ost = reinterpret_cast<int*>(oql + 2 * N);
  auto up = [&](double* dst, const double* src, size_t len) {
    return cudaMemcpyAsync(dst, src, len * 8, cudaMemcpyHostToDevice, h->stream);
  };
  CK(up(dmean, mean, D * N), "H2D mean");
  CK(up(dsq, sqrt_in, D * D * N), "H2D sqrt");
  if (t) CK(up(dt_, t, N), "H2D t");
  CK(up(ddt, dt, N), "H2D dt");
  if (o->np > 0) CK(up(dp, p, npar * N), "H2D p");
  if (uprev) CK(up(dup, uprev, (size_t)o->d * N), "H2D uprev");
  StepParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.n = n;
  sp.mean = dmean;
  sp.sqrt = dsq;
  sp.t = dt_;
  sp.dt = ddt;
  sp.p = dp;
  sp.uprev = uprev ? dup : nullptr;
  sp.mean_out = omean;
  sp.cov_out = ocov;
  sp.sigma2 = osig;
  sp.eest = oee;
  sp.u_out = ou;
  sp.quad_logdet = oql;
  sp.status = ost;
  sp.diffusion = h->cfg.diffusion;
  sp.abstol = h->cfg.abstol;
  sp.reltol = h->cfg.reltol;
  sp.C = h->C;
  CK(o->launch_step(o, sp, h->stream), "step kernel launch");
  auto down = [&](void* dst, const void* src, size_t bytes) {
    return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
  };
  CK(down(mean_out, omean, D * N * 8), "D2H mean");
  CK(down(cov_out, ocov, NC * N * 8), "D2H cov");
  CK(down(sigma2, osig, (size_t)o->nd * N * 8), "D2H sigma2");
  CK(down(eest, oee, N * 8), "D2H eest");
  CK(down(u_out, ou, (size_t)o->d * N * 8), "D2H u");
  CK(down(quad_logdet, oql, 2 * N * 8), "D2H quad/logdet");
  CK(down(status, ost, N * 4), "D2H status");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

int pnde_eval_dense(pnde_handle* h, int32_t which, int64_t tb, int64_t te, int64_t n_t, const double* t, double* mean,
                    double* cov) {
  if (!h) return PNDE_ERR_ARG;
  if (!h->ran) return h->fail(PNDE_ERR_STATE, "nothing has run");
  if (h->cfg.save_mode != PNDE_SAVE_EVERY) return h->fail(PNDE_ERR_STATE, "dense output needs save_mode = PNDE_SAVE_EVERY");
  if (which == PNDE_HIST_SMOOTHED && !h->smoothed) return h->fail(PNDE_ERR_STATE, "history has not been smoothed");
  if (which != PNDE_HIST_FILTERED && which != PNDE_HIST_SMOOTHED) return h->fail(PNDE_ERR_ARG, "bad 'which'");
  if (tb < 0 || te > h->n || tb >= te || n_t < 1 || !t) return h->fail(PNDE_ERR_ARG, "bad arguments");
  if (h->lorenz) return h->fail(PNDE_ERR_UNSUPPORTED, "dense output is not built for the large-d paths");
  if (h->multi()) {
    const int64_t NC = (int64_t)h->D * (h->D + 1) / 2;
    for (size_t k = 0; k < h->kids.size(); ++k) {
      const int64_t lo = std::max<int64_t>(tb, h->kid_lo[k]), hi = std::min<int64_t>(te, h->kid_lo[k + 1]);
      if (lo >= hi) continue;
      const int64_t o = (lo - tb) * n_t;
      const int rc = pnde_eval_dense(h->kids[k], which, lo - h->kid_lo[k], hi - h->kid_lo[k], n_t, t,
                                     mean ? mean + o * h->D : nullptr, cov ? cov + o * NC : nullptr);
      if (rc != PNDE_OK) return h->fail(rc, h->kids[k]->err);
    }
    return PNDE_OK;
  }
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const ModelOps* o = h->ops;
  const size_t ntr = (size_t)(te - tb), NC = (size_t)(o->D * (o->D + 1) / 2);
  const size_t nm = ntr * (size_t)n_t * o->D, nc = ntr * (size_t)n_t * NC;
  CK(h->scratch_out.ensure(((size_t)n_t + nm + nc) * 8 + 64), "alloc dense staging");
  double* dtq = h->scratch_out.as<double>();
  double* dmean = dtq + n_t;
  double* dcov = dmean + nm;
  CK(cudaMemcpyAsync(dtq, t, (size_t)n_t * 8, cudaMemcpyHostToDevice, h->stream), "H2D query times");
  const int df = h->cfg.diffusion;
  DenseParams dp;
  memset(&dp, 0, sizeof(dp));
  dp.n = h->n;
  dp.traj_begin = tb;
  dp.traj_end = te;
  dp.max_saved = h->max_saved;
  dp.n_t = n_t;
  dp.n_saved = h->n_saved.as<int>();
  dp.hist = h->hist.as<double>();
  dp.smooth = h->smooth.as<double>();
  dp.final_diff = h->final_diff.as<double>();
  dp.calibrate = (df == PNDE_DIFF_FIXED || df == PNDE_DIFF_FIXED_MAP || df == PNDE_DIFF_FIXED_MV);
  dp.is_mv = (df == PNDE_DIFF_DYNAMIC_MV || df == PNDE_DIFF_FIXED_MV);
  dp.smoothed = (which == PNDE_HIST_SMOOTHED);
  dp.tq = dtq;
  dp.mean = dmean;
  dp.cov = dcov;
  dp.C = h->C;
  CK(o->launch_dense(o, dp, h->stream), "dense kernel launch");
  if (mean) CK(cudaMemcpyAsync(mean, dmean, nm * 8, cudaMemcpyDeviceToHost, h->stream), "D2H mean");
  if (cov) CK(cudaMemcpyAsync(cov, dcov, nc * 8, cudaMemcpyDeviceToHost, h->stream), "D2H cov");
  CK(cudaStreamSynchronize(h->stream), "stream synchronize");
  return PNDE_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Roofline denominators (bench.py): FP64 FMA peak and HBM copy bandwidth of the bound device.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0,
         x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b);
    x1 = fma(x1, a, b);
    x2 = fma(x2, a, b);
    x3 = fma(x3, a, b);
    x4 = fma(x4, a, b);
    x5 = fma(x5, a, b);
    x6 = fma(x6, a, b);
    x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

}  // namespace

extern "C" {

int pnde_measure_fp64_peak(int32_t device, double* tflops) {
  if (!tflops) return PNDE_ERR_ARG;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return PNDE_ERR_CUDA;
  cudaDeviceProp prop;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return PNDE_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return PNDE_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess || best == 0.0) return PNDE_ERR_CUDA;
  *tflops = best;
  return PNDE_OK;
}

int pnde_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return PNDE_ERR_ARG;
  *ptr = nullptr;
  const cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_create_error = std::string("pnde_host_alloc: ") + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? PNDE_ERR_ALLOC : PNDE_ERR_CUDA;
  }
  return PNDE_OK;
}

int pnde_host_free(void* ptr) {
  if (!ptr) return PNDE_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? PNDE_OK : PNDE_ERR_CUDA;
}

int pnde_measure_hbm_copy(int32_t device, double* gbs) {
  if (!gbs) return PNDE_ERR_ARG;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return PNDE_ERR_CUDA;
  cudaDeviceProp prop;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return PNDE_ERR_CUDA;
  const size_t bytes = (size_t)1 << 30;
  double2 *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
    if (a) cudaFree(a);
    return PNDE_ERR_CUDA;
  }
  cudaMemset(a, 1, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    copy_kernel<<<prop.multiProcessorCount * 16, 256>>>(a, b, bytes / sizeof(double2));
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = 2.0 * bytes / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  if (cudaGetLastError() != cudaSuccess || best == 0.0) return PNDE_ERR_CUDA;
  *gbs = best;
  return PNDE_OK;
}

}  // extern "C"
