"""Config 3 at reduced size (for ncu): Van der Pol, EK1(order=5), adaptive."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 50000
rng = np.random.default_rng(20260118)
mu = np.exp(rng.uniform(np.log(5e2), np.log(2e3), n))
if "--sort" in sys.argv:  # neighbours in a warp get similar stiffness => similar step counts
    mu = np.sort(mu)
u0 = np.stack([np.zeros(n), np.sqrt(3.0) * (1 + 0.01 * rng.standard_normal(n))], axis=1)
prob = B.ODEProblem("vanderpol", [0.0, np.sqrt(3.0)], (0.0, 1.0), (1e3,))
s = B.FilterSolver(prob, B.EK1(order=5, smooth=False), save_everystep=False, one_thread="--one-thread" in sys.argv)
s.upload(u0, mu[:, None])
for _ in range(2):
    s.run()
c = s.counts()
print("ms", s.last_run_ms()[0], "attempted", int(c["naccept"].sum() + c["nreject"].sum()))
