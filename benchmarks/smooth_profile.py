"""Config 5 at reduced size (for ncu): LV, EK1(order=3), dt=0.05 on (0,10), filter with history + smoother."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
vf = sys.argv[2] if len(sys.argv) > 2 else "lotka_volterra"
rng = np.random.default_rng(20260118)
if vf == "lotka_volterra":
    p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
    prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
else:  # experiment builds only carry the fhn_readme kernels; the smoother's cost does not depend on the field
    p = np.stack([rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 0.3, n), rng.uniform(2, 4, n)], axis=1)
    prob = B.ODEProblem("fhn_readme", [1.0, 1.0], (0.0, 10.0), (0.2, 0.2, 3.0))
s = B.FilterSolver(prob, B.EK1(order=3, smooth=True), adaptive=False, dt=0.05, save_everystep=True)
s.upload(np.ones((n, 2)), p)
for _ in range(2):
    s.run(); s.smooth()
print("filter ms, smooth ms", s.last_run_ms(), "steps", n * 200)
