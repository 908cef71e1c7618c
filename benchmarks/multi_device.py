#!/usr/bin/env python
"""ONE process, ONE C-ABI call, N GPUs (cfg.n_devices): BASELINE configs[1] (1e6 FHN trajectories in total,
EK1(order=3), 2000 fixed steps) through pnde_solve_ensemble_to_host on devices [0, N) with pinned host buffers, against
the same call on one device.  Prints one JSON line per device count: wall time of the call, steps/s, bitwise equality.

    python benchmarks/multi_device.py [--n 1000000] [--devices 1 2 4 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import odefilters_b200 as B  # noqa: E402
from ensembles import config2_inputs  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--devices", type=int, nargs="*", default=None)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    have = B.api.device_count()
    counts = a.devices or [k for k in (1, 2, 4, 8) if k <= have]
    n = a.n
    u0, p = config2_inputs(n)
    u0s, ps = B.pinned_empty((2, n)), B.pinned_empty((3, n))
    u0s[:], ps[:] = u0.T, p.T
    prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 20.0), (0.2, 0.2, 3.0))
    ref = None
    for k in counts:
        s = B.FilterSolver(prob, B.EK1(order=3, smooth=False), adaptive=False, dt=0.01, save_everystep=False,
                           devices=list(range(k)))
        out = [B.pinned_empty((8, n)), B.pinned_empty((36, n)), B.pinned_empty(n), B.pinned_empty(n)]

        def call():
            s._check(s.lib.pnde_solve_ensemble_to_host(s._h, n, u0s.ctypes.data, ps.ctypes.data, *[o.ctypes.data for o in out]),
                     "pnde_solve_ensemble_to_host")

        call()
        call()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            call()
        dt = (time.perf_counter() - t0) / a.reps
        s.n = n
        dev_ms = s.last_run_ms()[0]
        if ref is None:
            ref = [o.copy() for o in out]
        same = all(np.array_equal(x, y) for x, y in zip(out, ref))
        print(json.dumps({"devices": k, "trajectories_total": n, "call_ms": 1e3 * dt, "device_ms_max_over_gpus": dev_ms,
                          "steps_per_s_e2e": n * 2000 / dt, "bitwise_equal_to_first": bool(same),
                          "h2d_bytes": 40 * n, "d2h_bytes": 368 * n}), flush=True)
        s.close()
