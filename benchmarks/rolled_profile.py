"""Throughput of the general-(d, q) fallback (rolled, local-memory NVRTC builds): ring ODE, ensemble of n trajectories."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import odefilters_b200 as B


def ring(d):
    f = "; ".join(f"du[{i}] = -p[0]*u[{i}] + p[1]*u[{(i + 1) % d}]*u[{(i + d - 1) % d}]" for i in range(d)) + ";"
    j = "; ".join(f"J[{i}][{i}] = -p[0]; J[{i}][{(i + 1) % d}] = p[1]*u[{(i + d - 1) % d}]; "
                  f"J[{i}][{(i + d - 1) % d}] = p[1]*u[{(i + 1) % d}]" for i in range(d)) + ";"
    return B.CustomVectorField(d=d, n_params=2, f=f, jac=j)


for d, kind, q, n in ((24, "EK1", 3, 4096), (12, "EK1", 2, 16384), (40, "EK0", 3, 16384), (128, "EK0", 2, 4096)):
    rng = np.random.default_rng(1)
    u0 = 1.0 + 0.3 * rng.standard_normal((n, d))
    p = np.tile([0.7, 0.4], (n, 1))
    alg = (B.EK1 if kind == "EK1" else B.EK0)(order=q, smooth=False)
    prob = B.ODEProblem(ring(d), u0[0], (0.0, 1.0), p[0])
    t0 = time.time()
    s = B.FilterSolver(prob, alg, adaptive=False, dt=0.02, save_everystep=False)
    tc = time.time() - t0
    s.upload(u0, p)
    s.run()
    s.run()
    ms = s.last_run_ms()[0]
    print(json.dumps({"d": d, "alg": kind, "q": q, "D": d * (q + 1), "trajectories": n, "steps": 50, "compile_s": round(tc, 1),
                      "ms": round(ms, 2), "steps_per_s": round(n * 50 / ms * 1e3)}), flush=True)
