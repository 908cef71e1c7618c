"""EK0 in Kronecker form, numpy -- TEST INFRASTRUCTURE.

The reference (and oracle/pnde_oracle.py) carry the dense D x D covariance, which at Lorenz-96
d = 1024 (D = 4096) is far too slow for a checker.  For EK0 with a scalar diffusion the covariance is
exactly C (x) I_d, so this restates the same recursion on the (q+1) x (q+1) factor.
It is validated against the dense oracle at small d (tests/test_oracle_golden.py) and then used as the
checker for the large-d GPU path.
"""
import math

import numpy as np

from pnde_oracle import Controller, _eps, ibm


def lorenz96_f(u, F):
    return (np.roll(u, -1) - np.roll(u, 2)) * np.roll(u, 1) - u + F


def lorenz96_jets(u0, F, q):
    """Time-Taylor coefficients c_0..c_q of the solution through u0 (src/state_initialization.jl:15-42)."""
    c = [np.asarray(u0, dtype=float)]
    for k in range(q):
        acc = np.zeros_like(c[0])
        for a in range(k + 1):
            acc += (np.roll(c[a], -1) - np.roll(c[a], 2)) * np.roll(c[k - a], 1)
        acc -= c[k]
        if k == 0:
            acc += F
        c.append(acc / (k + 1))
    return [math.factorial(k) * c[k] for k in range(q + 1)]


def solve_ek0_kron(f, derivs, tspan, q, *, adaptive=True, dt=None, abstol=1e-6, reltol=1e-3, diffusion="dynamic",
                   ctrl=None):
    """derivs: [u0, u'(t0), ..., u^(q)(t0)] (each length d).  Returns dict with t, M (final mean (q+1, d)),
    C (final (q+1)x(q+1) covariance factor product), naccept, nreject, nf, us (list of u per accepted step)."""
    ctrl = ctrl or Controller()
    At, Q = ibm(1, q)
    Qt = Q.mat
    d = len(derivs[0])
    M = np.array(derivs, dtype=float)  # (q+1, d), natural coordinates
    C = np.zeros((q + 1, q + 1))
    beta2, beta1 = 2 / (5 * (q + 1)), 7 / (10 * (q + 1))
    t0, t1 = map(float, tspan)
    dtmax = t1 - t0
    nf = 0

    def rms(x):
        return math.sqrt(float(np.sum(x * x)) / len(x))

    if adaptive and dt is None:  # Hairer initdt
        u0, f0 = M[0], f(M[0])
        sk = abstol + np.abs(u0) * reltol
        d0, d1 = rms(u0 / sk), rms(f0 / sk)
        dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else (d0 / d1) / 100
        dt0 = min(dt0, dtmax)
        f1 = f(u0 + dt0 * f0)
        nf += 2
        d2 = rms((f1 - f0) / sk) / dt0
        mx = max(d1, d2)
        dt1 = max(1e-6, dt0 * 1e-3) if mx <= 1e-15 else 10.0 ** (-(2 + math.log10(mx)) / (q + 1))
        cur = min(100 * dt0, dt1, dtmax)
    else:
        cur = float(dt)
    dt_user = cur
    t, qold, q11, it, acc_prev, dtpropose = t0, ctrl.qoldinit, 1.0, 0, None, cur
    nacc = nrej = 0
    uprev = M[0].copy()
    ts, us, diffs = [t0], [M[0].copy()], []
    gsaved = 1.0
    while t < t1:
        if it > 0:
            cur = dtpropose if acc_prev else cur / min(1 / ctrl.qmin, q11 / ctrl.gamma)
        it += 1
        if adaptive:
            cur = min(min(cur, dtmax), t1 - t)
        else:
            cur = min(dt_user, t1 - t)
        h = cur
        P = np.array([h ** (k - q - 0.5) for k in range(q + 1)])
        PI = 1 / P
        Mb = P[:, None] * M
        Cb = P[:, None] * C * P[None, :]
        Mp = At @ Mb
        uhat = PI[0] * Mp[0]
        fu = f(uhat)
        nf += 1
        z = PI[1] * Mp[1] - fu
        B = PI[1] ** 2 * Qt[1, 1]
        if diffusion == "dynamic":
            local = float(z @ z) / (d * B)
            Cp = At @ Cb @ At.T + local * Qt
        else:
            Cp = At @ Cb @ At.T + Qt
        hvec = np.zeros(q + 1)
        hvec[1] = PI[1]
        s = float(hvec @ Cp @ hvec)
        k = Cp @ hvec / s
        Mn = Mp - np.outer(k, z)
        IKH = np.eye(q + 1) - np.outer(k, hvec)
        Cn = IKH @ Cp @ IKH.T
        if diffusion != "dynamic":
            local = float(z @ z) / s / d
        if diffusion == "dynamic":
            gcur = local
        elif diffusion == "fixed":
            gcur = local if nacc == 0 else gsaved + (local - gsaved) / nacc
        else:
            raise ValueError(diffusion)
        unew = PI[0] * Mn[0]
        EEst = 0.0
        if adaptive:
            err = math.sqrt(local * B)
            r = h * err / (abstol + np.maximum(np.abs(uprev), np.abs(unew)) * reltol)
            EEst = rms(r)
        uprev = unew.copy()
        if (not adaptive) or EEst < 1.0:
            M = PI[:, None] * Mn
            C = PI[:, None] * Cn * PI[None, :]
        ttmp = t + h
        if adaptive:
            if EEst == 0:
                qc = 1 / ctrl.qmax
            else:
                q11 = EEst ** beta1
                qc = q11 / qold ** beta2
                qc = max(1 / ctrl.qmax, min(1 / ctrl.qmin, qc / ctrl.gamma))
            accept = EEst <= 1.0
            if accept:
                nacc += 1
                if ctrl.qsteady_min <= qc <= ctrl.qsteady_max:
                    qc = 1.0
                qold = max(EEst, ctrl.qoldinit)
                t = t1 if abs(ttmp - t1) < 10 * _eps(max(t, t1)) else ttmp
                dtpropose = max(ctrl.dtmin, min(dtmax, h / qc))
            else:
                nrej += 1
        else:
            accept = True
            nacc += 1
            t = t1 if abs(ttmp - t1) < 10 * _eps(max(t, t1)) else ttmp
            dtpropose = h
        acc_prev = accept
        if accept:
            gsaved = gcur
            ts.append(t)
            us.append(M[0].copy())
            diffs.append(gcur)
    if diffusion != "dynamic" and nacc > 0:
        C = C * gsaved
    return dict(t=np.array(ts), M=M, C=C, naccept=nacc, nreject=nrej, nf=nf, us=np.array(us), diffusions=np.array(diffs))
