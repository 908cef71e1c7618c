#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 1, 3 and 5 (bench.py covers config 2, the headline).

    python benchmarks/configs.py [--n3 100000] [--n5 250000]

Prints one JSON line per config with device times from the library's CUDA events.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odefilters_b200 as B  # noqa: E402

SEED = 20260118


def config1():
    prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 20.0), (0.2, 0.2, 3.0))
    t0 = time.perf_counter()
    sol = B.solve(prob, B.EK0(order=1), abstol=1e-1, reltol=1e-2)
    dt = time.perf_counter() - t0
    return {"config": 1, "what": "FHN README single solve, EK0(order=1), adaptive, smoothed", "naccept": sol.destats["naccept"],
            "nreject": sol.destats["nreject"], "u_end": sol.u[-1].tolist(), "wall_s_incl_alloc": dt}


def config3(n):
    rng = np.random.default_rng(SEED)
    mu = np.exp(rng.uniform(np.log(5e2), np.log(2e3), n))
    u0 = np.stack([np.zeros(n), np.sqrt(3.0) * (1 + 0.01 * rng.standard_normal(n))], axis=1)
    prob = B.ODEProblem("vanderpol", [0.0, np.sqrt(3.0)], (0.0, 1.0), (1e3,))
    s = B.FilterSolver(prob, B.EK1(order=5, smooth=False), save_everystep=False)
    s.upload(u0, mu[:, None])
    for _ in range(2):
        s.run()
    ms = s.last_run_ms()[0]
    c = s.counts()
    steps = int(c["naccept"].sum() + c["nreject"].sum())
    return {"config": 3, "what": "Van der Pol mu~logU(5e2,2e3), EK1(order=5), adaptive, masked per-trajectory control",
            "n": n, "attempted_steps": steps, "accepted": int(c["naccept"].sum()), "rejected": int(c["nreject"].sum()),
            "naccept_minmax": [int(c["naccept"].min()), int(c["naccept"].max())], "ms": ms,
            "steps_per_s": steps / (ms * 1e-3), "all_success": bool((c["retcode"] == 0).all())}


def config5(n):
    rng = np.random.default_rng(SEED)
    p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
    u0 = np.ones((n, 2))
    prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
    s = B.FilterSolver(prob, B.EK1(order=3, smooth=True), adaptive=False, dt=0.05, save_everystep=True)
    s.upload(u0, p)
    for _ in range(2):
        s.run()
        s.smooth()
    fms, sms = s.last_run_ms()
    c = s.counts()
    steps = int(c["naccept"].sum())
    rec = int(s.lib.pnde_record_len(s._h))
    return {"config": 5, "what": "Lotka-Volterra, EK1(order=3), dt=0.05 on (0,10): filter (every step saved) + RTS smoother",
            "n": n, "filter_steps": steps, "filter_ms": fms, "smooth_ms": sms,
            "filter_steps_per_s": steps / (fms * 1e-3), "smoother_steps_per_s": (steps - n) / (sms * 1e-3),
            "filter_hist_GBps": steps * rec * 8 / (fms * 1e-3) / 1e9, "record_bytes": rec * 8,
            "all_success": bool((c["retcode"] == 0).all())}


def config4(d=1024):
    rng = np.random.default_rng(SEED)
    u0 = 8.0 + 0.01 * rng.standard_normal(d)
    prob = B.ODEProblem("lorenz96", u0, (0.0, 1.0), (8.0,))
    s = B.FilterSolver(prob, B.EK0(order=3, smooth=False), adaptive=False, dt=1e-3, save_everystep=False)
    s.upload(u0[None, :], np.array([[8.0]]))
    for _ in range(3):
        s.run()
    ms = s.last_run_ms()[0]
    c = s.counts()
    steps = int(c["naccept"].sum())
    return {"config": 4, "what": f"Lorenz-96 d={d}, EK0(order=3) Kronecker covariance, fixed dt=1e-3, one CTA",
            "steps": steps, "ms": ms, "us_per_step": 1e3 * ms / steps, "all_success": bool((c["retcode"] == 0).all()),
            "note": "latency bound (one sequential chain, 3 block syncs per step)"}


def config2_ek0(n=1_000_000):
    """Not a BASELINE config: the headline workload with EK0 (Kronecker covariance) for comparison."""
    rng = np.random.default_rng(SEED)
    p = np.stack([rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 0.3, n), rng.uniform(2.0, 4.0, n)], axis=1)
    prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 20.0), (0.2, 0.2, 3.0))
    s = B.FilterSolver(prob, B.EK0(order=3, smooth=False), adaptive=False, dt=0.01, save_everystep=False)
    s.upload(np.tile([-1.0, 1.0], (n, 1)), p)
    for _ in range(3):
        s.run()
    ms = s.last_run_ms()[0]
    steps = int(s.counts()["naccept"].sum())
    return {"config": "2-EK0", "what": "FHN sweep, EK0(order=3) Kronecker, dt=0.01, 2000 steps, final state", "n": n,
            "ms": ms, "steps_per_s": steps / (ms * 1e-3)}


def config4_ek1(d=1024, nsteps=20):
    rng = np.random.default_rng(SEED)
    u0 = 8.0 + 0.01 * rng.standard_normal(d)
    dt = 1e-3
    prob = B.ODEProblem("lorenz96", u0, (0.0, nsteps * dt), (8.0,))
    s = B.FilterSolver(prob, B.EK1(order=3, smooth=False), adaptive=False, dt=dt, save_everystep=False)
    s.upload(u0[None, :], np.array([[8.0]]))
    for _ in range(2):
        s.run()
    ms = s.last_run_ms()[0]
    c = s.counts()
    D = 4 * d
    flop = 2.0 * D * D * D + 2.0 * (2 * d) * d * d  # blocked QR of the D x D dense rows + the diffusion QR
    return {"config": 4, "what": f"Lorenz-96 d={d}, EK1(order=3) dense D={D}, fixed dt=1e-3, blocked Householder QR with "
            "DMMA trailing updates", "steps": nsteps, "ms": ms, "ms_per_step": ms / nsteps,
            "ms_includes": "final covariance product S'S (one more 2 D^2 (D-d) flop GEMM)",
            "algorithmic_tflops": flop * nsteps / (ms * 1e-3) / 1e12, "launches": s.launch_count(),
            "all_success": bool((c["retcode"] == 0).all())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n3", type=int, default=100000)
    ap.add_argument("--n5", type=int, default=250000)
    a = ap.parse_args()
    for fn, arg in ((config1, None), (config2_ek0, 1_000_000), (config3, a.n3), (config4, 1024), (config4_ek1, 1024), (config5, a.n5)):
        out = fn() if arg is None else fn(arg)
        print(json.dumps(out), flush=True)
