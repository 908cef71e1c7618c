"""Smoother throughput by order (LV, EK1, fixed dt = 0.05 on (0, 10): 200 steps), for DESIGN section 4."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(20260118)
p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
for kind in ("EK1", "EK0"):
    for q in (1, 2, 3, 4, 5):
        alg = (B.EK1 if kind == "EK1" else B.EK0)(order=q, smooth=True)
        s = B.FilterSolver(prob, alg, adaptive=False, dt=0.05, save_everystep=True)
        s.upload(np.ones((n, 2)), p)
        for _ in range(2):
            s.run(); s.smooth()
        f, sm = s.last_run_ms()
        print(f"{kind} q={q}: filter {f:8.3f} ms  smoother {sm:8.3f} ms  -> {n * 200 / sm / 1e3:9.1f} M smoother steps/s, {n * 200 / f / 1e3:9.1f} M filter steps/s")
        s.close()
