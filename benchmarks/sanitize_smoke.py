"""Touches every kernel once on tiny inputs (meant to run under compute-sanitizer --tool memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odefilters_b200 as B

def run(prob, alg, **kw):
    sol = B.solve(prob, alg, **kw)
    if alg.smooth:
        sol(np.linspace(prob.tspan[0], prob.tspan[1], 7)); sol.sample(3, seed=1)
    return sol

lv = B.ODEProblem("lotka_volterra", [1.0, 1.0], (0.0, 0.5), (1.5, 1.0, 3.0, 1.0))
for alg in (B.EK1(order=3), B.EK0(order=2), B.EK0(order=2, diffusionmodel="dynamicMV"), B.EK1(order=2, diffusionmodel="fixed"),
            B.EK1(order=5, smooth=False)):
    run(lv, alg)
    run(lv, alg, adaptive=False, dt=0.05)
P = np.tile([1.5, 1.0, 3.0, 1.0], (130, 1))
es = B.solve(B.EnsembleProblem(lv, p=P), B.EK1(order=3, smooth=False), B.EnsembleB200(), adaptive=False, dt=0.05)
s = es.solver
s2 = B.FilterSolver(lv, B.EK1(order=3, smooth=True), adaptive=False, dt=0.05)
s2.solve_ensemble(np.ones((5, 2)), P[:5])
off, t, u, cu, _ = s2.history(1, 1, 4, marginals=True)
assert u.shape[1] == 2 and off[-1] == len(t)
cv = B.CustomVectorField(2, 1, "du[0] = u[1]; du[1] = -p[0]*sin(u[0]);", "J[0][0]=0.0; J[0][1]=1.0; J[1][0]=-p[0]*cos(u[0]); J[1][1]=0.0;")
run(B.ODEProblem(cv, [1.0, 0.0], (0.0, 0.5), (2.0,)), B.EK1(order=2))
u0 = 8.0 + 0.01 * np.random.default_rng(0).standard_normal(64)
B.solve(B.ODEProblem("lorenz96", u0, (0.0, 0.02), (8.0,)), B.EK0(order=3, smooth=False), save_everystep=False)
B.solve(B.ODEProblem("lorenz96", u0, (0.0, 0.02), (8.0,)), B.EK0(order=2, smooth=False), adaptive=False, dt=0.01, save_everystep=False)
B.solve(B.ODEProblem("lorenz96", u0, (0.0, 0.02), (8.0,)), B.EK1(order=1, smooth=False), adaptive=False, dt=0.01, save_everystep=False)
print("sanitize smoke ok")
