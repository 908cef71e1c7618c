"""Adaptive FHN ensemble, EK1(order=3) / EK0(order=3), default tolerances (per-trajectory step control)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(20260118)
p = np.stack([rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 0.3, n), rng.uniform(2, 4, n)], axis=1)
prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 20.0), (0.2, 0.2, 3.0))
for alg in (B.EK1(order=3, smooth=False), B.EK0(order=3, smooth=False), B.EK1(order=2, smooth=False)):
    s = B.FilterSolver(prob, alg, save_everystep=False)
    s.upload(np.tile([-1.0, 1.0], (n, 1)), p)
    for _ in range(2):
        s.run()
    c = s.counts()
    att = int(c["naccept"].sum() + c["nreject"].sum())
    print(f"kind={alg.kind} q={alg.order}: {s.last_run_ms()[0]:8.3f} ms, attempted {att}, {att / s.last_run_ms()[0] / 1e6:8.2f} G steps/s, checksum {c['naccept'].sum()} {c['nreject'].sum()}")
    s.close()
