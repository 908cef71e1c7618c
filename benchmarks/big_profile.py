"""A few steps of the large-D EK1 path (for ncu launch lists): big_profile.py [d] [nsteps]."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(20260118)
u0 = 8.0 + 0.01 * rng.standard_normal(d)
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prob = B.ODEProblem("lorenz96", u0, (0.0, nsteps * 1e-3), (8.0,))
s = B.FilterSolver(prob, B.EK1(order=3, smooth=False), adaptive=False, dt=1e-3, save_everystep=False)
s.upload(u0[None, :], np.array([[8.0]]))
s.run()
print("ms", s.last_run_ms()[0], "launches", s.launch_count(), "ms/step", s.last_run_ms()[0] / nsteps)
