/*
 * pnde.h -- C ABI of libpnde.so, the B200-native ODE-filter hot path.
 *
 * Drop-in boundary for ProbNumDiffEq.jl v0.1.5 (nathanaelbosch/ODEFilters.jl).  The
 * reference has no FFI: its hot path is reached through OrdinaryDiffEq's plugin hooks
 *   alg_cache      src/caches.jl:42-43        -> pnde_create
 *   initialize!    src/perform_step.jl:2-12   -> inside pnde_run (Taylor-mode init on device)
 *   perform_step!  src/perform_step.jl:27-93  -> inside pnde_run (the persistent filter kernel)
 *   savevalues!    src/integrator_utils.jl:33-48 -> history buffers, pnde_get_history
 *   postamble!     src/integrator_utils.jl:2-30  -> static-diffusion rescale + pnde_smooth
 *   build_solution src/solution.jl:45-80      -> pnde_get_final / pnde_get_history / pnde_get_counts
 * and through the external OrdinaryDiffEq loop (solve!, loopheader!, loopfooter!, PI controller,
 * initdt), which a per-step ccall cannot keep on a GPU; the boundary therefore sits one level
 * up: Julia's DiffEqBase.__solve(prob, alg::AbstractEK; ...) / __solve(::EnsembleProblem, ...)
 * marshal to structure-of-arrays and ccall the functions below once per ensemble.
 * INTEGRATION.md shows the Julia ccall stubs.
 *
 * Conventions: plain C, no exceptions; all reals are double, all sizes int64_t.  Return code 0 =
 * ok, <0 = argument / CUDA error (text via pnde_last_error).  Per-trajectory numerical failures are
 * DATA (the retcode array), never a failed call.  The caller allocates every output buffer; the
 * library owns device memory inside the handle.  A handle is bound to one CUDA device -- or, with
 * cfg.n_devices > 1, to several, over which it shards the ensemble -- and is not re-entrant; distinct handles
 * may be used from distinct host threads / processes.
 * There is no CPU fallback: every entry point that computes fails with PNDE_ERR_CUDA when no
 * device is present.
 *
 * Array layout: structure-of-arrays with the TRAJECTORY INDEX FASTEST, e.g. u0[c * n_traj + i]
 * is component c of trajectory i.
 */
#ifndef PNDE_H
#define PNDE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNDE_ABI_VERSION 2 /* 2: n_devices / flags / device_list appended to pnde_config */
#define PNDE_MAX_DEVICES 16

/* return codes */
#define PNDE_OK 0
#define PNDE_ERR_ARG (-1)
#define PNDE_ERR_CUDA (-2)
#define PNDE_ERR_UNSUPPORTED (-3)
#define PNDE_ERR_STATE (-4)
#define PNDE_ERR_ALLOC (-5)

/* algorithms: src/algorithms.jl:23-51 */
#define PNDE_ALG_EK0 0
#define PNDE_ALG_EK1 1
/* Iterated extended Kalman smoother, src/ieks.jl:2-61: an EK1 solve repeated `ieks_iterations` times, the Jacobian
 * of every iterate evaluated at the dense output sol(t + dt) of the previous one (src/perform_step.jl:111-113).
 * Needs smooth = 1 and PNDE_SAVE_EVERY; its kernels are compiled at create time through NVRTC. */
#define PNDE_ALG_IEKS 2

/* diffusion models: src/caches.jl:89-96 / src/diffusions.jl */
#define PNDE_DIFF_DYNAMIC 0
#define PNDE_DIFF_FIXED 1
#define PNDE_DIFF_FIXED_MAP 2
#define PNDE_DIFF_DYNAMIC_MV 3
#define PNDE_DIFF_FIXED_MV 4

/* built-in vector-field catalogue (Julia closures cannot run on the device) */
#define PNDE_VF_FHN_README 0     /* README.md:36-40, p = (a,b,c) */
#define PNDE_VF_FHN_LIB 1        /* prob_ode_fitzhughnagumo, p = (a,b,tauinv,l) */
#define PNDE_VF_LOTKA_VOLTERRA 2 /* prob_ode_lotkavoltera, p = (a,b,c,d) */
#define PNDE_VF_VANDERPOL 3      /* prob_ode_vanstiff ordering u=(y,x), p = (mu) */
#define PNDE_VF_LINEAR2 4        /* du_i = p_i u_i, d = 2 (test/state_init.jl:15) */
#define PNDE_VF_LOGISTIC 5       /* du = p u (1-u), d = 1 (test/specific_problems.jl:62) */
#define PNDE_VF_LORENZ96 6       /* d given in the config (4..2048), p = (F).  EK0: one CTA per trajectory, Kronecker-
                                    factored covariance Sigma = Ctilde (x) I_d, history / smoothing available (the
                                    getters then return the packed Ctilde where other paths return Sigma, and
                                    pnde_get_marginals ONE covariance entry per state, Ctilde[0][0]); EK1: dense
                                    D = d (q+1) <= 5120 (BASELINE config 4), fixed or adaptive steps; with save_mode !=
                                    PNDE_SAVE_FINAL the history holds the solution marginals of every saved state (a full
                                    state would be a 100 MB factor at D = 4096): pnde_get_marginals(PNDE_HIST_FILTERED)
                                    returns u [total][d] and cov_u = diag(Sigma_u) [total][d]; no smoothing */
#define PNDE_VF_LINEAR1 7        /* du = p u, d = 1 (test/convergence.jl:10) */
#define PNDE_VF_CUSTOM 100       /* user source, compiled at run time: pnde_create_custom */

/* what is written to the device-side history */
#define PNDE_SAVE_FINAL 0  /* final state only */
#define PNDE_SAVE_EVERY 1  /* every accepted step (what the reference does, integrator_utils.jl:43-45) */
#define PNDE_SAVE_STRIDE 2 /* every save_stride-th accepted step and the last one */

/* per-trajectory return codes (data) */
#define PNDE_RET_SUCCESS 0
#define PNDE_RET_MAXITERS 1
#define PNDE_RET_DTNAN 2
#define PNDE_RET_NONFINITE 3
#define PNDE_RET_HISTORY_FULL 4 /* more accepted steps than cfg.max_saved slots.  Adaptive thread-per-trajectory runs keep
                                    stepping without saving: naccept + 1 is the capacity a retry needs */
#define PNDE_RET_DTMIN 5
#define PNDE_RET_ZERO_RESIDUAL 6 /* only with PNDE_FLAG_REFERENCE_QUIRKS: FixedDiffusion met an exactly zero residual,
                                    where the reference throws (src/diffusions.jl:18-20) */

/* pnde_config.flags */
#define PNDE_FLAG_REFERENCE_QUIRKS 1 /* reproduce two accidents of the reference instead of the intended behaviour:
   (a) static diffusion model + smooth = 0: sol.pu keeps the UNcalibrated covariances, because savevalues!
       (src/integrator_utils.jl:43-45) fills it before postamble! rescales x_filt (:4-18) -- pnde_get_marginals(
       PNDE_HIST_FILTERED) then returns uncalibrated cov_u (pnde_get_history stays calibrated, like sol.x_filt);
   (b) FixedDiffusion with z == 0 exactly: the reference returns one value where two are destructured and throws
       (src/diffusions.jl:18-20) -- the trajectory ends there with PNDE_RET_ZERO_RESIDUAL instead of sigma^2 = 0.
       The test is not in the ahead-of-time kernels (it costs 2-3 % there): a handle with this flag and the fixed
       diffusion model gets its kernels compiled at create time through NVRTC (1-3 s). */

#define PNDE_FLAG_ONE_THREAD 2 /* dense EK1 with d (q+1) >= 10 normally runs with two lanes of a warp per trajectory
   (wide_filter.cuh); this flag selects the one-thread-per-trajectory kernel instead (A/B measurements; the two
   kernels perform the same operations in the same order and return identical results) */

/* which states pnde_get_history returns */
#define PNDE_HIST_FILTERED 0
#define PNDE_HIST_SMOOTHED 1

typedef struct pnde_config {
  int32_t abi_version; /* PNDE_ABI_VERSION */
  int32_t alg;         /* PNDE_ALG_* */
  int32_t order;       /* q, src/algorithms.jl:25; 1..7 (1..5 built ahead of time, 6..7 compiled on demand via NVRTC) */
  int32_t d;           /* ODE dimension (fixed by the vector field except Lorenz-96) */
  int32_t vf_kind;     /* PNDE_VF_* */
  int32_t diffusion;   /* PNDE_DIFF_* */
  int32_t smooth;      /* run the RTS pass in pnde_solve_ensemble (needs PNDE_SAVE_EVERY) */
  int32_t adaptive;    /* 1: PI-controlled steps, 0: fixed dt */
  int32_t save_mode;   /* PNDE_SAVE_* */
  int32_t save_stride; /* for PNDE_SAVE_STRIDE */
  int32_t device;      /* CUDA device ordinal, -1 = current device */
  int32_t ieks_iterations; /* PNDE_ALG_IEKS: re-solves per pnde_solve_ensemble (<= 0: 10, src/ieks.jl:54) */
  double abstol, reltol; /* defaults 1e-6 / 1e-3 (OrdinaryDiffEq) */
  double dt;             /* fixed step, or initial step when adaptive (<= 0: Hairer initdt) */
  double t0, t1;
  /* PI controller, OrdinaryDiffEq defaults + src/alg_utils.jl:23-24; beta <= 0 selects the default */
  double qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit, beta1, beta2, dtmin, dtmax;
  int64_t maxiters;  /* attempted steps per trajectory (default 100000) */
  int64_t max_saved; /* history slots per trajectory incl. the initial state; 0 = derive (fixed step) */
  /* --- ABI version 2 --- */
  int32_t n_devices; /* 0 or 1: the single device `device`.  > 1: the ensemble is sharded over device_list[0..n_devices):
                        contiguous blocks of trajectories, one host thread + stream per GPU inside every call, results
                        written straight into disjoint slices of the caller's arrays (EnsembleProblem over
                        the GPUs of one box, no exchange step).  Every entry point then acts on the whole ensemble. */
  int32_t flags;     /* PNDE_FLAG_* */
  int32_t device_list[PNDE_MAX_DEVICES];
} pnde_config;

typedef struct pnde_handle pnde_handle;

/* Fill cfg with the reference defaults for (alg, order): EK*(order=q), dynamic diffusion,
 * smooth=false, adaptive=true, abstol=1e-6, reltol=1e-3, PI controller of src/alg_utils.jl:13-24. */
int pnde_default_config(pnde_config* cfg, int32_t alg, int32_t order, int32_t vf_kind);

/* alg_cache (src/caches.jl:42-114): validates the configuration, builds the IWP constants
 * (src/priors.jl:7-59), binds the device. */
int pnde_create(const pnde_config* cfg, pnde_handle** out);
/* Any autonomous user ODE.  D = d (q+1) <= 16 (EK1) / 64 (EK0): unrolled into registers like the catalogue; larger, up
 * to D <= 96 (EK1) / 1024 (EK0), d <= 128: the same kernels with rolled loops and arrays in local memory (slow, general).
 * The vector field and its Jacobian are given as CUDA C++ statement lists and
 * compiled at run time (NVRTC) into the same kernels as the catalogue.  f_body assigns du[i] from u[] and p[]
 * and must be generic in the scalar type T (it is also evaluated on truncated Taylor series for the exact
 * initial state, src/state_initialization.jl:15-42): + - * / exp log sin cos sqrt are available.  jac_body
 * assigns J[i][j] = d f_i / d u_j in doubles (what src/jacobian.jl:6-22 obtains from ModelingToolkit; entries it does
 * not assign are zero); it may be NULL for EK0.  cfg->vf_kind must be PNDE_VF_CUSTOM.  Example (Lotka-Volterra):
 *   f_body   "du[0] = p[0]*u[0] - p[1]*u[0]*u[1]; du[1] = -p[2]*u[1] + p[3]*u[0]*u[1];"
 *   jac_body "J[0][0] = p[0]-p[1]*u[1]; J[0][1] = -p[1]*u[0]; J[1][0] = p[3]*u[1]; J[1][1] = -p[2]+p[3]*u[0];" */
int pnde_create_custom(const pnde_config* cfg, int32_t d, int32_t n_params, const char* f_body,
                       const char* jac_body, pnde_handle** out);
/* Compile-only validation of such a source (needs libnvrtc but no GPU); the compiler log goes to `log`. */
int pnde_check_custom(int32_t alg, int32_t order, int32_t diffusion, int32_t d, int32_t n_params,
                      const char* f_body, const char* jac_body, char* log, int64_t log_len);
int pnde_destroy(pnde_handle* h);
const char* pnde_last_error(const pnde_handle* h); /* h may be NULL: last create error */

/* Dimension helpers. */
int64_t pnde_state_dim(const pnde_handle* h);    /* D = d (q+1) */
int64_t pnde_n_params(const pnde_handle* h);     /* parameters per trajectory */
int64_t pnde_record_len(const pnde_handle* h);   /* doubles per saved state in the device history */
int64_t pnde_cov_len(const pnde_handle* h);      /* rows of pnde_get_final's cov: D(D+1)/2, or (q+1)(q+2)/2 for
                                                    the Lorenz-96 Kronecker path (Sigma = Ctilde (x) I_d) */

/* One call = one EnsembleProblem solve: host buffers in, results kept on the device.
 * u0: [d][n_traj], p: [n_params][n_traj].  Equivalent to pnde_upload + pnde_run (+ pnde_smooth);
 * PNDE_ALG_IEKS: pnde_upload + ieks_iterations x (pnde_run + pnde_smooth)  (solve_ieks, src/ieks.jl:53-61). */
int pnde_solve_ensemble(pnde_handle* h, int64_t n_traj, const double* u0, const double* p);

/* pnde_solve_ensemble + pnde_get_final in one call, pipelined: the ensemble is processed in slices and the
 * device-to-host copy of a finished slice overlaps the kernel of the next one (final-state runs of the
 * thread-per-trajectory models; other configurations fall back to the plain sequence).  Output layout as in
 * pnde_get_final; any output pointer may be NULL. */
int pnde_solve_ensemble_to_host(pnde_handle* h, int64_t n_traj, const double* u0, const double* p, double* mean,
                                double* cov, double* t_final, double* loglik);

/* Split form, so that inputs can stay resident in HBM across runs.  pnde_upload returns after the copies have
 * completed: u0 / p may be reused or freed immediately. */
int pnde_upload(pnde_handle* h, int64_t n_traj, const double* u0, const double* p);
int pnde_run(pnde_handle* h);         /* initialize! + the whole solve! loop, asynchronous.  PNDE_ALG_IEKS: every
                                         pnde_run after the first one since pnde_upload linearises at the solution
                                         of the previous pnde_run + pnde_smooth pair (which it keeps on the device) */
int pnde_synchronize(pnde_handle* h); /* wait for the handle's stream */
/* device time (ms, CUDA events on the handle's stream) of the last pnde_run / pnde_smooth */
int pnde_last_run_ms(pnde_handle* h, double* filter_ms, double* smooth_ms);
/* number of kernels the last pnde_run / pnde_smooth launched */
int64_t pnde_last_launch_count(const pnde_handle* h);

/* postamble! (src/integrator_utils.jl:2-30): RTS smoother over the saved history
 * (src/smoothing.jl:4-63).  Requires PNDE_SAVE_EVERY. */
int pnde_smooth(pnde_handle* h);

/* Sizes for the getters: total saved states over all trajectories and the per-trajectory maximum. */
int pnde_query_sizes(pnde_handle* h, int64_t* n_saved_total, int64_t* max_saved);

/* destats / retcode.  Any pointer may be NULL.  Arrays of n_traj. */
int pnde_get_counts(pnde_handle* h, int64_t* naccept, int64_t* nreject, int64_t* nf, int64_t* njacs,
                    int32_t* retcode, int64_t* n_saved);

/* Final filtering state: mean [D][n], cov packed lower triangle by rows [D(D+1)/2][n]
 * (entry (i,j), j<=i, at i(i+1)/2+j), t_final [n], log-likelihood [n] (NaN for static diffusion
 * models, src/integrator_utils.jl:6).  Any pointer may be NULL. */
int pnde_get_final(pnde_handle* h, double* mean, double* cov, double* t_final, double* loglik);

/* History of trajectory range [traj_begin, traj_end) in CSR form: offsets[i - traj_begin] ..
 * offsets[i - traj_begin + 1] index the saved states of trajectory i in the flat outputs.
 *   t         [total]           time stamps (sol.t)
 *   mean      [total][D]        state means (x_filt.mu / x_smooth.mu)
 *   cov       [total][D(D+1)/2] packed lower covariance (Sigma = S S')
 *   diffusion [total][nd]       sol.diffusions: entry k of a trajectory belongs to the interval
 *                               ending at saved state k (entry 0 is unused); nd = 1, or d for MV models
 * offsets has traj_end - traj_begin + 1 entries and is an OUTPUT (query sizes first). */
int pnde_get_history(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end,
                     int64_t* offsets, double* t, double* mean, double* cov, double* diffusion);

/* The square roots of the same states (SRMatrix.squareroot, src/squarerootmatrix.jl:10-16; consumed e.g. by
 * src/solution_sampling.jl:6-12): sqrt [total][D][D] row-major with Sigma = S S'.  Filtered states return the
 * rank-(D - d) factor the kernels carry, padded with d zero columns; smoothed states the lower-triangular factor of
 * the smoother.  No factorisation happens on the way out: these ARE the device-side representations. */
int pnde_get_history_sqrt(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end, int64_t* offsets,
                          double* sqrt);

/* sol.pu (src/integrator_utils.jl:45): marginals of the solution block, same CSR layout:
 * u [total][d], cov_u [total][d(d+1)/2]  (Lorenz-96 EK0: [total][1]; Lorenz-96 EK1: [total][d], see PNDE_VF_LORENZ96). */
int pnde_get_marginals(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end,
                       int64_t* offsets, double* t, double* u, double* cov_u);

/* sample_states (src/solution_sampling.jl:24-62): n_samples backward-sampled paths per trajectory of
 * [traj_begin, traj_end), drawn with an in-kernel counter-based generator (Philox4x32-10 keyed by seed;
 * reproducible for a given (seed, trajectory, sample, state)).  Needs PNDE_SAVE_EVERY.  Same CSR layout as
 * pnde_get_history: t [total], samples [total][n_samples][D] (sample() of the reference is the first d
 * components).  offsets is an output. */
int pnde_sample(pnde_handle* h, int64_t traj_begin, int64_t traj_end, int32_t n_samples, uint64_t seed,
                int64_t* offsets, double* t, double* samples);

/* dense_sample_states (src/solution_sampling.jl:63-74): `n_samples` joint draws from the smoothing posterior on the
 * caller's non-decreasing time grid tq[0..n_t) instead of the solver's own grid (the reference: 1000 equidistant points
 * over the solution's time span): backward sampling through the filtering posterior extrapolated to every grid point.
 * samples [traj_end - traj_begin][n_t][n_samples][D]; dense_sample (:75-79) is its first d entries of each state.
 * Needs save_mode = PNDE_SAVE_EVERY. */
int pnde_dense_sample(pnde_handle* h, int64_t traj_begin, int64_t traj_end, int64_t n_t, const double* tq,
                      int32_t n_samples, uint64_t seed, double* samples);

/* perform_step! (src/perform_step.jl:27-93) applied once to n caller-supplied states: the entry a step!/callback
 * driver needs, and the teacher-forced unit of the parity protocol.  Stateless with respect to the
 * handle's ensemble; uses the handle's algorithm, order, diffusion model and tolerances.
 *   in : mean [D][n], sqrt [D*D][n] (entry (i,j) of S at (i*D + j)*n + traj, Sigma = S S'; must be a state the filter
 *        can be in: zero or a filter posterior -- anything else yields status 1), t [n] (may be NULL: autonomous
 *        fields), dt [n], p [n_params][n], uprev [d][n] (integ.u before the step, for EEst; NULL: EEst = 0)
 *   out: mean_out [D][n], cov_out [D(D+1)/2][n] packed lower (x_filt), sigma2 [nd][n] (local diffusion),
 *        eest [n] (src/perform_step.jl:84), u_out [d][n], quad_logdet [2][n] (z' S^-1 z, log det S of :66),
 *        status [n] (0 ok, 1 not representable, 2 non-finite).  Any output except status may be NULL. */
int pnde_step_from_state(pnde_handle* h, int64_t n, const double* mean, const double* sqrt, const double* t,
                         const double* dt, const double* p, const double* uprev, double* mean_out, double* cov_out,
                         double* sigma2, double* eest, double* u_out, double* quad_logdet, int32_t* status);

/* Dense output sol(t) (GaussianODEFilterPosterior, src/solution.jl:165-215) of every trajectory in
 * [traj_begin, traj_end) at the n_t query times t[]: predict from the left filtered neighbour and, when
 * which == PNDE_HIST_SMOOTHED, smooth against the right smoothed neighbour.  mean [ntr][n_t][D],
 * cov [ntr][n_t][D(D+1)/2] packed lower.  Queries before t0 yield NaN (the reference throws). */
int pnde_eval_dense(pnde_handle* h, int32_t which, int64_t traj_begin, int64_t traj_end, int64_t n_t,
                    const double* t, double* mean, double* cov);

/* Page-locked host memory for the output getters: results copied into such buffers move at the full host-link
 * rate (the copies into pageable memory are staged by the driver).  Julia: unsafe_wrap(Array, ptr, dims). */
int pnde_host_alloc(void** ptr, int64_t bytes);
int pnde_host_free(void* ptr);

/* Device micro-benchmarks used as roofline denominators by bench.py (not part of the path). */
int pnde_measure_fp64_peak(int32_t device, double* tflops);
int pnde_measure_hbm_copy(int32_t device, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* PNDE_H */
