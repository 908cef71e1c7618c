"""Probe: device time of the sliced solve (pnde_solve_ensemble_to_host) vs the single launch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import odefilters_b200 as B
n = 1_000_000
rng = np.random.default_rng(1)
p = np.ascontiguousarray(np.stack([rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 0.3, n), rng.uniform(2, 4, n)], axis=0))
u0 = np.ascontiguousarray(np.stack([np.full(n, -1.0), np.full(n, 1.0)], axis=0))
pin = lambda a: torch.from_numpy(a).pin_memory()
u0p, pp = pin(u0), pin(p)
mean, cov, tf, ll = (torch.empty(s, dtype=torch.float64).pin_memory() for s in ((8, n), (36, n), (n,), (n,)))
prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 20.0), (0.2, 0.2, 3.0))
s = B.FilterSolver(prob, B.EK1(order=3, smooth=False), adaptive=False, dt=0.01, save_everystep=False)
lib, h = s.lib, s._h
for it in range(4):
    t0 = time.perf_counter()
    rc = lib.pnde_solve_ensemble_to_host(h, n, u0p.data_ptr(), pp.data_ptr(), mean.data_ptr(), cov.data_ptr(), tf.data_ptr(), ll.data_ptr())
    t1 = time.perf_counter()
    s.n = n
    print("sliced: wall %.2f ms, device (first launch .. last kernel) %.2f ms" % (1e3 * (t1 - t0), s.last_run_ms()[0]))
for it in range(3):
    t0 = time.perf_counter()
    lib.pnde_upload(h, n, u0p.data_ptr(), pp.data_ptr()); ta = time.perf_counter()
    lib.pnde_run(h); lib.pnde_synchronize(h); tb = time.perf_counter()
    lib.pnde_get_final(h, mean.data_ptr(), cov.data_ptr(), tf.data_ptr(), ll.data_ptr()); tc = time.perf_counter()
    print("plain: upload %.2f run %.2f (device %.2f) fetch %.2f ms" % (1e3 * (ta - t0), 1e3 * (tb - ta), s.last_run_ms()[0], 1e3 * (tc - tb)))
