#!/usr/bin/env python
"""Regenerates tests/golden/arbiter_config3.npz: the first N draws of BASELINE config 3 (Van der Pol, mu ~
logU(5e2, 2e3), EK1(order=5), adaptive, default tolerances) run through the oracle's recursion in 60-digit mpmath
arithmetic -- the EXACT recursion of the reference's algorithm, against which FP64 implementations that disagree with
each other on accept/reject decisions are judged (oracle/arbiter_mpmath.py).

    python tests/golden/make_arbiter_config3.py [N]       (about 30 s per trajectory and core)
"""
import multiprocessing as mpc
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(args):
    import mpmath as mp

    import pnde_oracle as O

    mp.mp.dps = 60
    u0, mu = args
    prob = O.Problem(O.CATALOGUE["vanderpol"], [mp.mpf(float(x)) for x in u0], (0.0, 1.0), [mp.mpf(float(mu))])
    s = O.solve_ivp(prob, O.EK1(order=5, smooth=False), dtype=object)
    return s.naccept, s.nreject, [float(x) for x in s.u[-1]]


if __name__ == "__main__":
    from ensembles import config3_inputs

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    u0, p = config3_inputs(10_000)  # the draws of the 1e4 test ensemble; trajectory i uses draw i
    with mpc.Pool(min(n, os.cpu_count() or 1)) as pool:
        res = pool.map(one, [(u0[i], p[i, 0]) for i in range(n)])
    np.savez(os.path.join(HERE, "arbiter_config3.npz"), index=np.arange(n), u0=u0[:n], mu=p[:n, 0],
             naccept=np.array([r[0] for r in res]), nreject=np.array([r[1] for r in res]),
             u1=np.array([r[2] for r in res]))
    print("exact recursion:", [(r[0], r[1]) for r in res])
