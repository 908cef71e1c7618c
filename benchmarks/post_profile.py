"""Post-processing kernels at reduced size (for ncu): sampling (prepare + draw), dense output, history conversion.
LV, EK1(order=q), fixed dt = 0.05 on (0, 10), smoothed."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(20260118)
p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
s = B.FilterSolver(prob, B.EK1(order=q, smooth=True), adaptive=False, dt=0.05, save_everystep=True)
s.solve_ensemble(np.ones((n, 2)), p)
import time
t0 = time.perf_counter(); s.sample(0, n, 64, seed=1); t1 = time.perf_counter()
s.dense(1, 0, n, np.linspace(0.01, 9.99, 64)); t2 = time.perf_counter()
print(f"q={q} n={n}: 64 samples x 201 states per trajectory {1e3 * (t1 - t0):.1f} ms (incl. D2H), "
      f"64 dense evaluations per trajectory {1e3 * (t2 - t1):.1f} ms (incl. D2H)")
