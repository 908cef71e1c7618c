#!/usr/bin/env python
"""mpmath arbiter (TEST INFRASTRUCTURE): runs the oracle's recursion in 60-digit arithmetic to tell which
FP64 implementation is closer to the exact recursion when two of them disagree beyond 1e-10.

    python oracle/arbiter_mpmath.py fhn      # fixed steps, 2000 steps, a stiff-ish draw of config 2
    python oracle/arbiter_mpmath.py vdp      # adaptive, Van der Pol mu=1e3, EK1(order=5) (config 3 centre)
"""
import os
import sys
import time

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pnde_oracle as O  # noqa: E402
import pnde_ref as R  # noqa: E402

mp.mp.dps = 60


def mpf_list(xs):
    return [mp.mpf(float(x)) for x in xs]


def fhn():
    p = [0.1808504754815618, 0.2644266438544973, 3.4351888806647247]
    prob = O.Problem(O.CATALOGUE["fhn_readme"], mpf_list([-1, 1]), (0.0, 20.0), mpf_list(p))
    t0 = time.time()
    sm = O.solve_ivp(prob, O.EK1(order=3, smooth=False), adaptive=False, dt=0.01, dtype=object)
    exact = np.array([float(x) for x in sm.x_filt[-1].mu[:2]])
    so = O.solve_ivp(O.Problem(O.CATALOGUE["fhn_readme"], [-1.0, 1.0], (0.0, 20.0), p), O.EK1(order=3, smooth=False),
                     adaptive=False, dt=0.01)
    rc = R.solve_ensemble("fhn_readme", "EK1", 3, [[-1.0, 1.0]], [p], (0.0, 20.0), adaptive=False, dt=0.01, want_cov=False)
    sc = np.abs(exact).max()
    print(f"exact (60 digits) u(20) = {exact}  [{time.time() - t0:.0f} s]")
    print("numpy oracle, chol-first :", np.abs(so.x_filt[-1].mu[:2] - exact).max() / sc)
    print("C restatement, chol-first:", np.abs(rc["mean"][0][:2] - exact).max() / sc)


def vdp():
    prob = O.Problem(O.CATALOGUE["vanderpol"], mpf_list([0.0, 3.0 ** 0.5]), (0.0, 1.0), mpf_list([1e3]))
    t0 = time.time()
    sm = O.solve_ivp(prob, O.EK1(order=5, smooth=False), dtype=object)
    so = O.solve_ivp(O.Problem(O.CATALOGUE["vanderpol"], [0.0, 3.0 ** 0.5], (0.0, 1.0), [1e3]), O.EK1(order=5, smooth=False))
    print(f"60-digit recursion : naccept {sm.naccept} nreject {sm.nreject}  [{time.time() - t0:.0f} s]")
    print(f"FP64 oracle (chol-first, {so.stats.get('chol_fail', 0)} Cholesky failures): naccept {so.naccept} nreject {so.nreject}")
    print("u(1) 60-digit:", [float(x) for x in sm.u[-1]], " FP64 oracle:", list(so.u[-1]))


if __name__ == "__main__":
    {"fhn": fhn, "vdp": vdp}[sys.argv[1] if len(sys.argv) > 1 else "fhn"]()
