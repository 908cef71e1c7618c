"""ctypes wrapper of oracle/pnde_ref.c (TEST INFRASTRUCTURE / CPU baseline; see that file's header)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpnde_ref.so")
VF = {"fhn_readme": 0, "fhn_lib": 1, "lotka_volterra": 2, "vanderpol": 3}


class RefConfig(C.Structure):
    _fields_ = [("alg", C.c_int), ("order", C.c_int), ("vf", C.c_int), ("adaptive", C.c_int),
                ("abstol", C.c_double), ("reltol", C.c_double), ("dt", C.c_double), ("t0", C.c_double),
                ("t1", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double), ("gamma", C.c_double),
                ("qsteady_min", C.c_double), ("qsteady_max", C.c_double), ("qoldinit", C.c_double),
                ("dtmin", C.c_double), ("dtmax", C.c_double), ("maxiters", C.c_int64)]


_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) and build:
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
        _lib.pnde_ref_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(load().pnde_ref_max_threads())


def solve_ensemble(vf, alg, order, u0, p, tspan, *, adaptive=True, dt=0.0, abstol=1e-6, reltol=1e-3, nthreads=0,
                   want_cov=True):
    """u0: [N, 2], p: [N, np].  Returns dict(mean [N, D], cov [N, D, D], t, loglik, naccept, nreject, nf,
    chol_fail, retcode)."""
    lib = load()
    u0 = np.ascontiguousarray(np.asarray(u0, dtype=float).T)
    p = np.ascontiguousarray(np.asarray(p, dtype=float).T)
    n = u0.shape[1]
    D = 2 * (order + 1)
    cfg = RefConfig(alg=1 if alg == "EK1" else 0, order=order, vf=VF[vf], adaptive=int(adaptive), abstol=abstol,
                    reltol=reltol, dt=dt or 0.0, t0=tspan[0], t1=tspan[1], qmin=0.2, qmax=10.0, gamma=0.9,
                    qsteady_min=1.0, qsteady_max=1.0, qoldinit=1e-4, dtmin=0.0, dtmax=0.0, maxiters=100000)
    mean = np.empty((D, n))
    cov = np.empty((D * D, n)) if want_cov else None
    t = np.empty(n)
    ll = np.empty(n)
    counts = np.empty((4, n), dtype=np.int64)
    ret = np.empty(n, dtype=np.int32)
    rc = lib.pnde_ref_solve_ensemble(C.byref(cfg), C.c_int64(n), C.c_void_p(u0.ctypes.data), C.c_void_p(p.ctypes.data),
                                     C.c_void_p(mean.ctypes.data), C.c_void_p(cov.ctypes.data if want_cov else None),
                                     C.c_void_p(t.ctypes.data), C.c_void_p(ll.ctypes.data),
                                     C.c_void_p(counts.ctypes.data), C.c_void_p(ret.ctypes.data), C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"pnde_ref_solve_ensemble failed: {rc}")
    return dict(mean=mean.T.copy(), cov=(cov.T.reshape(n, D, D).copy() if want_cov else None), t=t, loglik=ll,
                naccept=counts[0], nreject=counts[1], nf=counts[2], chol_fail=counts[3], retcode=ret)
