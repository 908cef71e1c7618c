"""q = 5 smoother at reduced size (for ncu): LV, EK1(order=5), dt = 0.05 on (0, 10)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(20260118)
p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
prob = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 10.0), (1.5, 1.0, 3.0, 1.0))
s = B.FilterSolver(prob, B.EK1(order=q, smooth=True), adaptive=False, dt=0.05, save_everystep=True,
                   one_thread="--one-thread" in sys.argv)
s.upload(np.ones((n, 2)), p)
for _ in range(2):
    s.run(); s.smooth()
f, sm = s.last_run_ms()
print(f"q={q} n={n}: filter {f:.3f} ms, smoother {sm:.3f} ms -> {n * 200 / sm / 1e3:.1f} M smoother steps/s")
