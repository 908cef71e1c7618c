"""Host-side interface (see package docstring).  Mirrors the reference's user API:

    prob = ODEProblem("fhn_readme", u0=[-1.0, 1.0], tspan=(0.0, 20.0), p=(0.2, 0.2, 3.0))
    sol = solve(prob, EK0(order=1), abstol=1e-1, reltol=1e-2)          # README.md:47
    eprob = EnsembleProblem(prob, p=P)                                   # P: [N, n_params]
    esol = solve(eprob, EK1(order=3), EnsembleB200(), trajectories=N, adaptive=False, dt=0.01)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


# --------------------------------------------------------------------------------------------
# Algorithm / problem types (src/algorithms.jl:23-51; DiffEqBase ODEProblem / EnsembleProblem)
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class _EK:
    order: int = 3
    diffusionmodel: str = "dynamic"
    smooth: bool = True
    prior: str = "ibm"
    kind: int = L.ALG_EK1
    iterations: int = 0  # IEKS only (set by solve_ieks)

    def __post_init__(self):
        if self.prior != "ibm":
            raise ValueError("Only the ibm prior is implemented so far")  # src/caches.jl:69
        if self.diffusionmodel not in L.DIFFUSIONS:
            raise ValueError(f"diffusionmodel must be one of {list(L.DIFFUSIONS)}")


def EK0(order: int = 3, diffusionmodel: str = "dynamic", smooth: bool = True, prior: str = "ibm") -> _EK:
    """Gaussian ODE filtering with zeroth order extended Kalman filtering (src/algorithms.jl:23-28)."""
    return _EK(order, diffusionmodel, smooth, prior, L.ALG_EK0)


def EK1(order: int = 3, diffusionmodel: str = "dynamic", smooth: bool = True, prior: str = "ibm") -> _EK:
    """Gaussian ODE filtering with first order extended Kalman filtering (src/algorithms.jl:46-51)."""
    return _EK(order, diffusionmodel, smooth, prior, L.ALG_EK1)


def IEKS(order: int = 1, diffusionmodel: str = "dynamic", prior: str = "ibm") -> _EK:
    """Gaussian ODE filtering with iterated extended Kalman smoothing (src/ieks.jl:10-41): use ``solve_ieks``.
    ``smooth`` is always on.  Every iterate is an EK1 solve linearised at the previous iterate's dense output."""
    return _EK(order, diffusionmodel, True, prior, L.ALG_IEKS)


@dataclass(frozen=True)
class CustomVectorField:
    """A user ODE for the run-time compiled path (pnde_create_custom): CUDA C++ statement lists for f and its
    Jacobian, the role ModelingToolkit-generated code plays on the Julia side (src/jacobian.jl:6-22).

        lv = CustomVectorField(d=2, n_params=4,
                               f="du[0] = p[0]*u[0] - p[1]*u[0]*u[1]; du[1] = -p[2]*u[1] + p[3]*u[0]*u[1];",
                               jac="J[0][0] = p[0]-p[1]*u[1]; J[0][1] = -p[1]*u[0]; J[1][0] = p[3]*u[1]; J[1][1] = -p[2]+p[3]*u[0];")
    """

    d: int
    n_params: int
    f: str
    jac: Optional[str] = None

    def check(self, alg: "_EK") -> str:
        """Compile-only validation (no GPU needed); returns the compiler log, raises on errors."""
        lib = L.load()
        log = C.create_string_buffer(1 << 16)
        rc = lib.pnde_check_custom(alg.kind, alg.order, L.DIFFUSIONS[alg.diffusionmodel], self.d, self.n_params,
                                   self.f.encode(), self.jac.encode() if self.jac else None, log, len(log))
        if rc != 0:
            raise ValueError(log.value.decode())
        return log.value.decode()


@dataclass
class ODEProblem:
    """ODEProblem(f, u0, tspan, p) with ``f`` a name from the built-in catalogue (include/pnde.h) or a
    CustomVectorField."""

    f: object
    u0: Sequence[float]
    tspan: Sequence[float]
    p: Sequence[float] = ()

    def __post_init__(self):
        self.custom = self.f if isinstance(self.f, CustomVectorField) else None
        if self.custom is not None:
            self.f = "custom"
        if self.f not in L.VF_KINDS:
            raise ValueError(f"unknown vector field {self.f!r}; catalogue: {sorted(L.VF_KINDS)}")
        u0 = np.asarray(self.u0, dtype=np.float64)
        if u0.ndim != 1:
            # src/caches.jl:46-49
            raise ValueError("Problems which are not scalar- or vector-valued (e.g. u0 is a scalar or a matrix) "
                             "are currently not supported")
        d, npar = L.VF_DIMS.get(self.f, (u0.shape[0], 1))  # lorenz96: d = len(u0), p = (F,)
        if self.custom is not None:
            d, npar = self.custom.d, self.custom.n_params
        if u0.shape[0] != d:
            raise ValueError(f"{self.f} has dimension {d}")
        p = np.atleast_1d(np.asarray(self.p, dtype=np.float64))
        if npar == 0 and p.size == 0:
            p = np.zeros(0)
        if p.shape[0] != npar:
            raise ValueError(f"{self.f} takes {npar} parameters")
        self.u0, self.p = u0, p


@dataclass
class EnsembleProblem:
    """EnsembleProblem(prob; prob_func): the per-trajectory remake is given as arrays.

    ``u0``: [N, d] or None (broadcast prob.u0); ``p``: [N, n_params] or None (broadcast prob.p).
    """

    prob: ODEProblem
    u0: Optional[np.ndarray] = None
    p: Optional[np.ndarray] = None


class EnsembleB200:
    """Ensemble algorithm: all trajectories in one device launch (stands where EnsembleThreads does).

    ``devices``: CUDA device ordinals to shard the ensemble over inside the ONE library call (contiguous blocks, one
    host thread + stream per GPU, results written into disjoint slices of the output arrays).  None: the
    current device.  ``EnsembleB200(devices="all")`` uses every visible GPU."""

    def __init__(self, devices=None):
        self.devices = devices


def device_count() -> int:
    """Number of visible CUDA devices (cudaGetDeviceCount through the runtime the library links)."""
    import ctypes.util

    for name in ("libcudart.so.12", "libcudart.so", ctypes.util.find_library("cudart")):
        if not name:
            continue
        try:
            rt = C.CDLL(name)
        except OSError:
            continue
        n = C.c_int(0)
        return int(n.value) if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0
    return 0


def shard_range(n: int, rank: int, world: int):
    """Contiguous block partition [lo, hi) of n trajectories: remainder to the last rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi


def shard_strided(n: int, rank: int, world: int) -> np.ndarray:
    """Interleaved assignment rank, rank + world, ...: for adaptive ensembles whose cost varies smoothly with
    the trajectory index, every rank then sees the same mix of cheap and expensive trajectories."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return np.arange(rank, n, world)


# --------------------------------------------------------------------------------------------
# Result containers (src/squarerootmatrix.jl, GaussianDistributions.Gaussian, src/solution.jl:8-24)
# --------------------------------------------------------------------------------------------
class SRMatrix:
    """PSD matrix with a square-root factor (src/squarerootmatrix.jl:9-16).  ``mat`` and ``squareroot`` both come from
    the device (pnde_get_history / pnde_get_history_sqrt: the factor the kernels carry, S S' = mat; like the
    reference's, it is not unique).  Only an SRMatrix built from a bare matrix derives a factor on the host."""

    def __init__(self, mat: np.ndarray, squareroot: Optional[np.ndarray] = None):
        self.mat = mat
        self._sr = squareroot

    @property
    def squareroot(self) -> np.ndarray:
        if self._sr is None:
            w, v = np.linalg.eigh((self.mat + self.mat.T) / 2)
            self._sr = v * np.sqrt(np.clip(w, 0.0, None))
        return self._sr

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.mat, dtype=dtype)

    @property
    def shape(self):
        return self.mat.shape


@dataclass
class Gaussian:
    mu: np.ndarray
    Sigma: SRMatrix


class _PinnedBlock:
    def __init__(self, nbytes: int):
        self.lib = L.load()
        self.ptr = C.c_void_p()
        rc = self.lib.pnde_host_alloc(C.byref(self.ptr), int(nbytes))
        if rc != 0:
            raise MemoryError(f"pnde_host_alloc({nbytes}) failed ({rc}): {self.lib.pnde_last_error(None).decode()}")

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.pnde_host_free(self.ptr)
            self.ptr = None


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array in page-locked host memory (pnde_host_alloc): device-to-host copies into it run at the full
    host-link rate instead of being staged through the driver's bounce buffer."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    blk = _PinnedBlock(max(nbytes, 8))
    buf = (C.c_char * max(nbytes, 8)).from_address(blk.ptr.value)
    buf._pnde_block = blk  # the array's base keeps the allocation alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def _unpack_lower(packed: np.ndarray, D: int) -> np.ndarray:
    """[..., D(D+1)/2] packed lower triangle (rows) -> [..., D, D] symmetric."""
    out = np.zeros(packed.shape[:-1] + (D, D))
    il = np.tril_indices(D)
    out[..., il[0], il[1]] = packed
    out[..., il[1], il[0]] = packed
    return out


class _GaussianList:
    """StructArray{Gaussian} over SoA buffers (src/solution.jl:60-64): ``.mu`` [N, D], ``.Sigma`` [N, D, D]."""

    def __init__(self, mu: np.ndarray, cov: np.ndarray, sqrt: Optional[np.ndarray] = None):
        self.mu, self.Sigma, self.sqrt = mu, cov, sqrt  # sqrt: [N, rows, D] factors from the device, or None

    def __len__(self):
        return self.mu.shape[0]

    def __getitem__(self, i) -> Gaussian:
        return Gaussian(self.mu[i], SRMatrix(self.Sigma[i], None if self.sqrt is None else self.sqrt[i]))


def mean(x):
    """mean(::SRGaussianList) = s.mu (src/ProbNumDiffEq.jl:64); also for a single Gaussian."""
    return x.mu


def var(x):
    """var = diag(Sigma), per state for a list (src/ProbNumDiffEq.jl:62,65)."""
    S = x.Sigma.mat if isinstance(x, Gaussian) else x.Sigma
    return np.diagonal(S, axis1=-2, axis2=-1).copy()


def std(x):
    """std = sqrt.(diag(Sigma)) (src/ProbNumDiffEq.jl:63,66)."""
    return np.sqrt(np.maximum(var(x), 0.0))


@dataclass
class ProbODESolution:
    """Fields of the reference's ProbODESolution (src/solution.jl:8-24)."""

    t: np.ndarray
    u: np.ndarray
    pu: _GaussianList
    x_filt: _GaussianList
    x_smooth: Optional[_GaussianList]
    diffusions: np.ndarray
    log_likelihood: float
    destats: dict
    retcode: str
    prob: ODEProblem = None
    alg: _EK = None
    _solver: "FilterSolver" = None
    _index: int = 0

    def __len__(self):
        return len(self.t)

    def __call__(self, t):
        """Dense output sol(t) = SolProj * posterior(t) (src/solution.jl:211-215); evaluated on the device."""
        scalar = np.ndim(t) == 0
        tq = np.atleast_1d(np.asarray(t, dtype=np.float64))
        which = L.HIST_SMOOTHED if self.x_smooth is not None else L.HIST_FILTERED
        mean, cov = self._solver.dense(which, self._index, self._index + 1, tq)
        d = self._solver.d
        out = _GaussianList(mean[0][:, :d].copy(), cov[0][:, :d, :d].copy())
        return out[0] if scalar else out

    def posterior(self, t):
        """Full-state posterior at t (GaussianODEFilterPosterior call, src/solution.jl:165-210)."""
        tq = np.atleast_1d(np.asarray(t, dtype=np.float64))
        which = L.HIST_SMOOTHED if self.x_smooth is not None else L.HIST_FILTERED
        mean, cov = self._solver.dense(which, self._index, self._index + 1, tq)
        return _GaussianList(mean[0], cov[0])

    def sample_states(self, n: int = 1, seed: int = 0) -> np.ndarray:
        """sample_states(sol, n) (src/solution_sampling.jl:15-18): [len(sol), D, n]."""
        if self.x_smooth is None:
            raise ValueError("sampling not implemented for non-smoothed posteriors")  # :16
        _, _, smp = self._solver.sample(self._index, self._index + 1, n, seed)
        return np.ascontiguousarray(np.transpose(smp, (0, 2, 1)))

    def sample(self, n: int = 1, seed: int = 0) -> np.ndarray:
        """sample(sol, n) (src/solution_sampling.jl:19-23): [len(sol), d, n]."""
        return self.sample_states(n, seed)[:, : self._solver.d, :]

    def dense_sample_states(self, n: int = 1, seed: int = 0, n_times: int = 1000):
        """dense_sample_states(sol, n) (src/solution_sampling.jl:63-74): (samples [n_times, D, n], times) on
        range(sol.t[1], sol.t[end], length = 1000)."""
        if self.x_smooth is None:
            raise ValueError("sampling not implemented for non-smoothed posteriors")  # :64
        times = np.linspace(self.t[0], self.t[-1], n_times)
        smp = self._solver.dense_sample(self._index, self._index + 1, times, n, seed)[0]
        return np.ascontiguousarray(np.transpose(smp, (0, 2, 1))), times

    def dense_sample(self, n: int = 1, seed: int = 0, n_times: int = 1000):
        """dense_sample(sol, n) (src/solution_sampling.jl:75-79): (samples [n_times, d, n], times)."""
        smp, times = self.dense_sample_states(n, seed, n_times)
        return smp[:, : self._solver.d, :], times


@dataclass
class EnsembleSolution:
    """EnsembleSolution over the device results: final states for every trajectory plus lazily
    fetched per-trajectory histories."""

    solver: "FilterSolver"
    n: int
    mean: np.ndarray        # [N, D] final filtering mean
    cov: np.ndarray         # [N, D(D+1)/2] packed
    t_final: np.ndarray
    log_likelihood: np.ndarray
    destats: dict
    retcode: np.ndarray
    converged: bool = True
    _cache: dict = field(default_factory=dict)
    _slot: Optional[np.ndarray] = None  # device slot of trajectory i when the ensemble was reordered (balance_by)

    def __len__(self):
        return self.n

    @property
    def u(self) -> np.ndarray:
        return self.mean[:, : self.solver.d]

    def __getitem__(self, i: int) -> ProbODESolution:
        return self.solver.solution(int(self._slot[i]) if self._slot is not None else i)


# --------------------------------------------------------------------------------------------
# Low-level handle wrapper
# --------------------------------------------------------------------------------------------
class FilterSolver:
    """One pnde_handle: alg_cache + the device buffers of one ensemble (include/pnde.h)."""

    def __init__(self, prob: ODEProblem, alg: _EK, *, abstol=1e-6, reltol=1e-3, adaptive=True, dt=None,
                 save_everystep=True, save_stride=None, smooth=None, maxiters=100000, max_saved=0, device=-1,
                 devices=None, reference_quirks=False, one_thread=False,
                 dtmin=0.0, dtmax=None, qmin=None, qmax=None, gamma=None, beta1=None, beta2=None):
        if not adaptive and dt is None:
            raise ValueError("Fixed timestep methods require a choice of dt")  # test/errors.jl:16-20
        self.lib = L.load()
        self.prob, self.alg = prob, alg
        cfg = L.PndeConfig()
        self.lib.pnde_default_config(C.byref(cfg), alg.kind, alg.order, L.VF_KINDS[prob.f])
        cfg.diffusion = L.DIFFUSIONS[alg.diffusionmodel]
        smooth = alg.smooth if smooth is None else smooth
        cfg.smooth = 1 if (smooth and save_everystep) else 0
        cfg.adaptive = 1 if adaptive else 0
        cfg.save_mode = L.SAVE_EVERY if save_everystep else (L.SAVE_STRIDE if save_stride else L.SAVE_FINAL)
        cfg.save_stride = int(save_stride or 1)
        cfg.device = device
        if isinstance(devices, str) and devices == "all":
            devices = list(range(device_count()))
        if devices is not None:
            devices = [int(x) for x in devices]
            if not 1 <= len(devices) <= L.MAX_DEVICES:
                raise ValueError(f"devices must list 1..{L.MAX_DEVICES} CUDA ordinals")
            cfg.n_devices = len(devices)
            for i, dv in enumerate(devices):
                cfg.device_list[i] = dv
        cfg.flags = (L.FLAG_REFERENCE_QUIRKS if reference_quirks else 0) | (L.FLAG_ONE_THREAD if one_thread else 0)
        self.devices = devices
        cfg.ieks_iterations = int(alg.iterations)
        cfg.abstol, cfg.reltol = float(abstol), float(reltol)
        cfg.dt = float(dt) if dt is not None else 0.0
        cfg.t0, cfg.t1 = float(prob.tspan[0]), float(prob.tspan[1])
        cfg.maxiters = int(maxiters)
        cfg.max_saved = int(max_saved)
        cfg.dtmin = float(dtmin)
        cfg.d = len(prob.u0)
        for name, val in (("dtmax", dtmax), ("qmin", qmin), ("qmax", qmax), ("gamma", gamma), ("beta1", beta1),
                          ("beta2", beta2)):
            if val is not None:
                setattr(cfg, name, float(val))
        self.cfg = cfg
        self._h = C.c_void_p()
        if prob.custom is not None:
            cv = prob.custom
            rc = self.lib.pnde_create_custom(C.byref(cfg), cv.d, cv.n_params, cv.f.encode(),
                                             cv.jac.encode() if cv.jac else None, C.byref(self._h))
        else:
            rc = self.lib.pnde_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"pnde_create failed ({rc}): {self.lib.pnde_last_error(None).decode()}")
        self.d, self.npar = L.VF_DIMS.get(prob.f, (len(prob.u0), 1))
        if prob.custom is not None:
            self.d, self.npar = prob.custom.d, prob.custom.n_params
        self.D = int(self.lib.pnde_state_dim(self._h))
        self.ncov = int(self.lib.pnde_cov_len(self._h))
        # Lorenz-96 EK0: covariance returned as the Kronecker factor Ctilde (EK1 returns the full matrix)
        self.kron = prob.f == "lorenz96" and alg.kind == L.ALG_EK0
        # Lorenz-96 EK1 (large-D dense path): the saved history holds the solution marginals only (u, diag Sigma_u)
        self.bigdense = prob.f == "lorenz96" and alg.kind == L.ALG_EK1
        self.n = 0
        self.is_mv = alg.diffusionmodel in ("dynamicMV", "fixedMV")

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.pnde_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.pnde_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- data movement / execution ---------------------------------------------------------
    @staticmethod
    def _soa(a: np.ndarray) -> np.ndarray:
        """[N, k] -> C-contiguous [k, N] (trajectory index fastest)."""
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)

    def upload(self, u0: np.ndarray, p: np.ndarray, soa: bool = False):
        u0s = u0 if soa else self._soa(u0)
        ps = p if soa else self._soa(p)
        n = u0s.shape[1]
        assert u0s.shape == (self.d, n) and ps.shape == (self.npar, n)
        self._keep = (u0s, ps)
        self._check(self.lib.pnde_upload(self._h, n, u0s.ctypes.data, ps.ctypes.data), "pnde_upload")
        self.n = n

    def run(self, sync: bool = True):
        self._check(self.lib.pnde_run(self._h), "pnde_run")
        if sync:
            self.synchronize()

    def smooth(self, sync: bool = True):
        self._check(self.lib.pnde_smooth(self._h), "pnde_smooth")
        if sync:
            self.synchronize()

    def synchronize(self):
        self._check(self.lib.pnde_synchronize(self._h), "pnde_synchronize")

    def solve_ensemble(self, u0: np.ndarray, p: np.ndarray, soa: bool = False):
        u0s = u0 if soa else self._soa(u0)
        ps = p if soa else self._soa(p)
        n = u0s.shape[1]
        self._keep = (u0s, ps)
        self._check(self.lib.pnde_solve_ensemble(self._h, n, u0s.ctypes.data, ps.ctypes.data), "pnde_solve_ensemble")
        self.n = n

    def solve_to_host(self, u0: np.ndarray, p: np.ndarray, soa: bool = False):
        """pnde_solve_ensemble_to_host: pipelined solve + final-state download; returns final()'s tuple."""
        u0s = u0 if soa else self._soa(u0)
        ps = p if soa else self._soa(p)
        n, D = u0s.shape[1], self.D
        self._keep = (u0s, ps)
        mean, cov, tf, ll = np.empty((D, n)), np.empty((self.ncov, n)), np.empty(n), np.empty(n)
        self._check(self.lib.pnde_solve_ensemble_to_host(self._h, n, u0s.ctypes.data, ps.ctypes.data, mean.ctypes.data,
                                                         cov.ctypes.data, tf.ctypes.data, ll.ctypes.data),
                    "pnde_solve_ensemble_to_host")
        self.n = n
        return mean.T.copy(), cov.T.copy(), tf, ll

    def last_run_ms(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.pnde_last_run_ms(self._h, C.byref(a), C.byref(b)), "pnde_last_run_ms")
        return a.value, b.value

    def launch_count(self) -> int:
        return int(self.lib.pnde_last_launch_count(self._h))

    # --- results --------------------------------------------------------------------------
    def counts(self) -> dict:
        n = self.n
        arrs = {k: np.zeros(n, dtype=np.int64) for k in ("naccept", "nreject", "nf", "njacs", "n_saved")}
        ret = np.zeros(n, dtype=np.int32)
        self._check(self.lib.pnde_get_counts(self._h, arrs["naccept"].ctypes.data, arrs["nreject"].ctypes.data,
                                             arrs["nf"].ctypes.data, arrs["njacs"].ctypes.data, ret.ctypes.data,
                                             arrs["n_saved"].ctypes.data), "pnde_get_counts")
        arrs["retcode"] = ret
        return arrs

    def final(self):
        """(mean [N, D], cov packed [N, D(D+1)/2], t_final [N], loglik [N])."""
        n, D = self.n, self.D
        mean = np.empty((D, n))
        cov = np.empty((self.ncov, n))
        tf = np.empty(n)
        ll = np.empty(n)
        self._check(self.lib.pnde_get_final(self._h, mean.ctypes.data, cov.ctypes.data, tf.ctypes.data,
                                            ll.ctypes.data), "pnde_get_final")
        return mean.T.copy(), cov.T.copy(), tf, ll

    def final_u(self):
        """Only the solution block of the final mean and the final time (small D2H read)."""
        n, D = self.n, self.D
        mean = np.empty((D, n))
        tf = np.empty(n)
        self._check(self.lib.pnde_get_final(self._h, mean.ctypes.data, None, tf.ctypes.data, None), "pnde_get_final")
        return mean[: self.d].T.copy(), tf

    def history(self, which: int, lo: int, hi: int, marginals: bool = False, pinned: bool = False, out=None):
        """CSR history of trajectories [lo, hi): (offsets, t, mean, cov_packed, diffusion).
        pinned: allocate the outputs in page-locked memory (large reads: several times the pageable copy rate);
        out = (t, mean, cov) reuses caller buffers (e.g. from pinned_empty) that are at least as large."""
        empty = pinned_empty if pinned else np.empty
        if out is not None:
            bufs = iter(out)
            def empty(shape, _b=bufs):  # noqa: E306
                shape = (shape,) if np.isscalar(shape) else tuple(shape)
                a = next(_b).reshape(-1)
                return a[: int(np.prod(shape, dtype=np.int64))].reshape(shape)
        tot, mx = C.c_int64(), C.c_int64()
        self._check(self.lib.pnde_query_sizes(self._h, C.byref(tot), C.byref(mx)), "pnde_query_sizes")
        cnt = self.counts()["n_saved"][lo:hi]
        total = int(cnt.sum())
        DM = self.d if marginals else self.D
        offsets = np.zeros(hi - lo + 1, dtype=np.int64)
        t = empty(total)
        mean = empty((total, DM))
        # large-d Kronecker path: the packed Ctilde (Sigma = Ctilde (x) I_d); marginals: one entry, Ctilde[0][0]
        # large-D dense path: marginals only, cov_u = diag(Sigma_u), d entries per state
        cov = empty((total, (1 if marginals else self.ncov) if self.kron else
                     (self.d if (self.bigdense and marginals) else DM * (DM + 1) // 2)))
        if marginals:
            self._check(self.lib.pnde_get_marginals(self._h, which, lo, hi, offsets.ctypes.data, t.ctypes.data,
                                                    mean.ctypes.data, cov.ctypes.data), "pnde_get_marginals")
            return offsets, t, mean, cov, None
        nd = self.d if self.is_mv else 1
        diff = np.empty((total, nd))
        self._check(self.lib.pnde_get_history(self._h, which, lo, hi, offsets.ctypes.data, t.ctypes.data,
                                              mean.ctypes.data, cov.ctypes.data, diff.ctypes.data), "pnde_get_history")
        return offsets, t, mean, cov, diff

    def history_sqrt(self, which: int, lo: int, hi: int):
        """pnde_get_history_sqrt: (offsets, S [total, D, D]) with Sigma = S S' (SRMatrix.squareroot) -- the factors the
        kernels carry, no factorisation on the way out."""
        total = int(self.counts()["n_saved"][lo:hi].sum())
        offsets = np.zeros(hi - lo + 1, dtype=np.int64)
        S = np.empty((total, self.D, self.D))
        self._check(self.lib.pnde_get_history_sqrt(self._h, which, lo, hi, offsets.ctypes.data, S.ctypes.data),
                    "pnde_get_history_sqrt")
        return offsets, S

    def step_from_state(self, mean, sqrt, dt, p, t=None, uprev=None) -> dict:
        """pnde_step_from_state: perform_step! (src/perform_step.jl:27-93) once from n given states.
        mean [n, D], sqrt [n, D, D] (Sigma = S S'), dt [n], p [n, n_params], uprev [n, d] (for EEst)."""
        mean = np.asarray(mean, dtype=np.float64)
        n, D, d = mean.shape[0], self.D, self.d
        soa = lambda a, k: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(n, k).T)  # noqa: E731
        m_, s_, dt_, p_ = soa(mean, D), soa(sqrt, D * D), soa(dt, 1), soa(p, max(self.npar, 1))
        t_ = soa(t, 1) if t is not None else None
        up_ = soa(uprev, d) if uprev is not None else None
        nd = d if (self.alg.kind == L.ALG_EK0) else 1
        out = dict(mean=np.empty((D, n)), cov=np.empty((D * (D + 1) // 2, n)), sigma2=np.empty((nd, n)), eest=np.empty(n),
                   u=np.empty((d, n)), ql=np.empty((2, n)), status=np.zeros(n, dtype=np.int32))
        ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
        self._check(self.lib.pnde_step_from_state(self._h, n, ptr(m_), ptr(s_), ptr(t_), ptr(dt_), ptr(p_), ptr(up_),
                                                  ptr(out["mean"]), ptr(out["cov"]), ptr(out["sigma2"]), ptr(out["eest"]),
                                                  ptr(out["u"]), ptr(out["ql"]), ptr(out["status"])),
                    "pnde_step_from_state")
        return dict(mean=out["mean"].T.copy(), cov=_unpack_lower(out["cov"].T.copy(), D), sigma2=out["sigma2"].T.copy(),
                    eest=out["eest"], u=out["u"].T.copy(), quad=out["ql"][0].copy(), logdet=out["ql"][1].copy(),
                    status=out["status"])

    def sample(self, lo: int, hi: int, n_samples: int, seed: int = 0):
        """pnde_sample for trajectories [lo, hi): (offsets, t [total], samples [total, n_samples, D])."""
        cnt = self.counts()["n_saved"][lo:hi]
        total = int(cnt.sum())
        offsets = np.zeros(hi - lo + 1, dtype=np.int64)
        t = np.empty(total)
        smp = np.empty((total, n_samples, self.D))
        self._check(self.lib.pnde_sample(self._h, lo, hi, n_samples, seed, offsets.ctypes.data, t.ctypes.data,
                                         smp.ctypes.data), "pnde_sample")
        return offsets, t, smp

    def dense_sample(self, lo: int, hi: int, tq: np.ndarray, n_samples: int, seed: int = 0) -> np.ndarray:
        """pnde_dense_sample for trajectories [lo, hi): samples [ntr, n_t, n_samples, D] on the time grid tq."""
        tq = np.ascontiguousarray(tq, dtype=np.float64)
        smp = np.empty((hi - lo, len(tq), n_samples, self.D))
        self._check(self.lib.pnde_dense_sample(self._h, lo, hi, len(tq), tq.ctypes.data, n_samples, seed, smp.ctypes.data),
                    "pnde_dense_sample")
        return smp

    def dense(self, which: int, lo: int, hi: int, tq: np.ndarray):
        """pnde_eval_dense: (mean [ntr, n_t, D], cov [ntr, n_t, D, D]) of trajectories [lo, hi) at times tq."""
        tq = np.ascontiguousarray(tq, dtype=np.float64)
        D = self.D
        mean = np.empty((hi - lo, len(tq), D))
        cov = np.empty((hi - lo, len(tq), D * (D + 1) // 2))
        self._check(self.lib.pnde_eval_dense(self._h, which, lo, hi, len(tq), tq.ctypes.data, mean.ctypes.data,
                                             cov.ctypes.data), "pnde_eval_dense")
        return mean, _unpack_lower(cov, D)

    def solution(self, i: int, counts: Optional[dict] = None, final=None) -> ProbODESolution:
        """build_solution for trajectory i (src/solution.jl:45-80, src/integrator_utils.jl:20-26)."""
        counts = counts or self.counts()
        D, d = self.D, self.d
        smoothed = bool(self.cfg.smooth)
        if self.cfg.save_mode == L.SAVE_FINAL:
            mean, cov, tf, ll = final or self.final()
            if self.kron:
                # Sigma = Ctilde (x) I_d: the full D x D matrix is never formed; marginals use Ctilde[0, 0]
                ct = _unpack_lower(cov[i:i + 1], self.alg.order + 1)
                pu = _GaussianList(mean[i:i + 1, :d].copy(), ct[:, 0, 0][:, None, None] * np.eye(d)[None])
                return ProbODESolution(
                    t=tf[i:i + 1], u=pu.mu, pu=pu, x_filt=_GaussianList(mean[i:i + 1], ct), x_smooth=None,
                    diffusions=np.zeros(0), log_likelihood=float(ll[i]),
                    destats={k: int(counts[k][i]) for k in ("naccept", "nreject", "nf", "njacs")},
                    retcode=L.RETCODES.get(int(counts["retcode"][i]), "Failure"), prob=self.prob, alg=self.alg,
                    _solver=self, _index=i)
            xf = _GaussianList(mean[i:i + 1], _unpack_lower(cov[i:i + 1], D))
            t = tf[i:i + 1]
            diffs = np.zeros((0, 1))
            xs = None
            llv = ll[i]
        else:
            if self.kron:
                return self._kron_solution(i, counts, final)
            if self.bigdense:
                return self._bigdense_solution(i, counts, final)
            _, t, mean, cov, diffs = self.history(L.HIST_FILTERED, i, i + 1)
            xf = _GaussianList(mean, _unpack_lower(cov, D), self.history_sqrt(L.HIST_FILTERED, i, i + 1)[1])
            xs = None
            if smoothed:
                _, _, ms, cs, _ = self.history(L.HIST_SMOOTHED, i, i + 1)
                xs = _GaussianList(ms, _unpack_lower(cs, D), self.history_sqrt(L.HIST_SMOOTHED, i, i + 1)[1])
            diffs = diffs[1:]  # entry 0 belongs to no interval
            llv = (final or self.final())[3][i]
        src = xs if xs is not None else xf
        # sol.pu = SolProj * x: the factor of the marginal is E0 S, d x D (src/squarerootmatrix.jl:38-39)
        pu = _GaussianList(src.mu[:, :d].copy(), src.Sigma[:, :d, :d].copy(),
                           None if src.sqrt is None else src.sqrt[:, :d, :].copy())
        return ProbODESolution(
            t=t, u=pu.mu, pu=pu, x_filt=xf, x_smooth=xs,
            diffusions=diffs if self.is_mv else diffs[:, 0],
            log_likelihood=float(llv),
            destats={k: int(counts[k][i]) for k in ("naccept", "nreject", "nf", "njacs")},
            retcode=L.RETCODES.get(int(counts["retcode"][i]), "Failure"), prob=self.prob, alg=self.alg,
            _solver=self, _index=i)


    def _bigdense_solution(self, i: int, counts: dict, final=None) -> ProbODESolution:
        """Large-D dense EK1 path with history: a saved full state would be a (D-d) x D factor (100 MB at D = 4096), so
        the history holds sol.t, sol.u and the marginal variances diag(Sigma_u) [N, d]; x_filt is the final state."""
        _, t, u, var, _ = self.history(L.HIST_FILTERED, i, i + 1, marginals=True)
        mean, cov, _, ll = final or self.final()
        xf = _GaussianList(mean[i:i + 1], _unpack_lower(cov[i:i + 1], self.D))
        pu = _GaussianList(u, var)
        return ProbODESolution(
            t=t, u=pu.mu, pu=pu, x_filt=xf, x_smooth=None, diffusions=np.zeros(0), log_likelihood=float(ll[i]),
            destats={k: int(counts[k][i]) for k in ("naccept", "nreject", "nf", "njacs")},
            retcode=L.RETCODES.get(int(counts["retcode"][i]), "Failure"), prob=self.prob, alg=self.alg, _solver=self, _index=i)

    def _kron_solution(self, i: int, counts: dict, final=None) -> ProbODESolution:
        """Large-d Kronecker path with history: Sigma = Ctilde (x) I_d is never expanded.  x_filt / x_smooth carry the
        means [N, D] and Ctilde [N, q+1, q+1]; sol.pu the means [N, d] and the marginal variance Ctilde[0, 0] [N]."""
        q1 = self.alg.order + 1
        _, t, mean, cov, diffs = self.history(L.HIST_FILTERED, i, i + 1)
        xf = _GaussianList(mean, _unpack_lower(cov, q1))
        xs = None
        if self.cfg.smooth:
            _, _, ms, cs, _ = self.history(L.HIST_SMOOTHED, i, i + 1)
            xs = _GaussianList(ms, _unpack_lower(cs, q1))
        src = xs if xs is not None else xf
        pu = _GaussianList(src.mu[:, : self.d].copy(), src.Sigma[:, 0, 0].copy())
        return ProbODESolution(
            t=t, u=pu.mu, pu=pu, x_filt=xf, x_smooth=xs, diffusions=diffs[1:, 0],
            log_likelihood=float((final or self.final())[3][i]),
            destats={k: int(counts[k][i]) for k in ("naccept", "nreject", "nf", "njacs")},
            retcode=L.RETCODES.get(int(counts["retcode"][i]), "Failure"), prob=self.prob, alg=self.alg, _solver=self, _index=i)


# --------------------------------------------------------------------------------------------
# solve
# --------------------------------------------------------------------------------------------
def _ensemble_arrays(eprob: EnsembleProblem, trajectories: Optional[int]):
    prob = eprob.prob
    n = trajectories
    for a in (eprob.u0, eprob.p):
        if a is not None:
            n = len(a) if n is None else n
    if n is None:
        raise ValueError("trajectories must be given")
    u0 = np.broadcast_to(prob.u0, (n, len(prob.u0))) if eprob.u0 is None else np.asarray(eprob.u0, dtype=np.float64)[:n]
    p = np.broadcast_to(prob.p, (n, len(prob.p))) if eprob.p is None else np.asarray(eprob.p, dtype=np.float64)[:n]
    return u0, p, n


def solve_ieks(prob, alg: _EK, *args, iterations: int = 10, **kwargs):
    """solve_ieks(prob, IEKS(...); iterations=10, kwargs...) (src/ieks.jl:44-61): a fixed number of re-solves, no
    stopping criterion; all iterates run on the device inside one ``pnde_solve_ensemble`` call."""
    if alg.kind != L.ALG_IEKS:
        raise ValueError("solve_ieks needs an IEKS algorithm")
    from dataclasses import replace
    return solve(prob, replace(alg, iterations=int(iterations)), *args, **kwargs)


def solve(prob, alg: _EK, ensemblealg: Optional[EnsembleB200] = None, *, trajectories: Optional[int] = None,
          abstol=1e-6, reltol=1e-3, adaptive=True, dt=None, dense=None, save_everystep=None, save_stride=None,
          maxiters=100000, max_saved=0, device=-1, devices=None, balance_by=None, **ctrl):
    """solve(prob, EK0/EK1(order=q); abstol, reltol, adaptive, dt) -> ProbODESolution, or
    solve(EnsembleProblem, alg, EnsembleB200(); trajectories=N, ...) -> EnsembleSolution.

    ``dense`` must equal ``alg.smooth`` (the reference asserts this, src/perform_step.jl:3); it defaults to it.
    ``balance_by`` (ensembles): a per-trajectory cost key, e.g. the stiffness parameter.  The ensemble is processed in
    the order of that key so that the 32 trajectories of a warp take similar numbers of adaptive steps (SURVEY 8e;
    -16 % on BASELINE config 3); results are returned in the caller's order.
    """
    if dense is not None and bool(dense) != bool(alg.smooth):
        raise ValueError("`dense` and `smooth` should have the same value! ")
    ensemble = isinstance(prob, EnsembleProblem)
    if devices is None and ensemblealg is not None:
        devices = getattr(ensemblealg, "devices", None)
    if save_everystep is None:
        save_everystep = not ensemble
    perm = None
    if ensemble:
        u0, p, n = _ensemble_arrays(prob, trajectories)
        base = prob.prob
        if balance_by is not None:
            key = np.asarray(balance_by, dtype=np.float64)
            if key.shape != (n,):
                raise ValueError("balance_by must hold one key per trajectory")
            perm = np.argsort(key, kind="stable")
            u0, p = u0[perm], p[perm]
    else:
        base = prob
        u0, p, n = base.u0[None, :], base.p[None, :], 1
    user_cap = max_saved
    if save_everystep and adaptive and not max_saved:
        max_saved = 1024 if ensemble else 8192
    for attempt in range(8):
        solver = FilterSolver(base, alg, abstol=abstol, reltol=reltol, adaptive=adaptive, dt=dt,
                              save_everystep=save_everystep, save_stride=save_stride, maxiters=maxiters,
                              max_saved=max_saved, device=device, devices=devices, **ctrl)
        solver.solve_ensemble(u0, p)
        counts = solver.counts()
        full = counts["retcode"] == 4
        if full.any() and not user_cap and max_saved < maxiters + 1:
            # history full: the kernels kept stepping without saving, so naccept is what the run needs -- ONE retry
            # with the exact capacity (the large-d path stops at a full history: grow geometrically there)
            need = int(counts["naccept"][full].max()) + 1
            solver.close()
            max_saved = min(max(need, max_saved * 4 if need <= max_saved else need), maxiters + 1)
            continue
        break
    if not ensemble:
        return solver.solution(0, counts)
    mean, cov, tf, ll = solver.final()
    slot = None
    if perm is not None:  # back to the caller's order
        slot = np.empty(n, dtype=np.int64)
        slot[perm] = np.arange(n)
        mean, cov, tf, ll = mean[slot], cov[slot], tf[slot], ll[slot]
        counts = {k: v[slot] for k, v in counts.items()}
    return EnsembleSolution(solver=solver, n=n, mean=mean, cov=cov, t_final=tf, log_likelihood=ll,
                            destats={k: counts[k] for k in ("naccept", "nreject", "nf", "njacs")},
                            retcode=counts["retcode"], converged=bool((counts["retcode"] == 0).all()), _slot=slot)
