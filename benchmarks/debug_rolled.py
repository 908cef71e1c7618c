"""Debug helper: rolled vs unrolled build of the same small user ODE (dynamicMV smoother)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import odefilters_b200 as B

f = "du[0] = u[0] - u[0]*u[0]*u[0]/3.0 - u[1] + p[3]; du[1] = p[2]*(u[0] + p[0] - p[1]*u[1]);"
j = "J[0][0] = 1.0 - u[0]*u[0]; J[0][1] = -1.0; J[1][0] = p[2]; J[1][1] = -p[2]*p[1];"
u0, p, tspan = [-1.0, 1.0], [0.2, 0.2, 3.0, 0.5], (0.0, 2.0)
np.set_printoptions(linewidth=200, precision=6)
for q, diff in ((3, "dynamicMV"), (2, "dynamicMV"), (1, "dynamicMV"), (3, "fixedMV")):
    res = []
    for rolled in (False, True):
        if rolled:
            os.environ["PNDE_FORCE_ROLLED"] = "1"
        else:
            os.environ.pop("PNDE_FORCE_ROLLED", None)
        cv = B.CustomVectorField(d=2, n_params=4, f=f, jac=j)
        sg = B.solve(B.ODEProblem(cv, u0, tspan, p), B.EK0(order=q, diffusionmodel=diff, smooth=True), adaptive=False, dt=0.25)
        res.append(sg)
    a, b = res
    print(q, diff, "filt mean diff", np.abs(a.x_filt.mu - b.x_filt.mu).max(), "filt cov diff", np.abs(a.x_filt.Sigma - b.x_filt.Sigma).max())
    print(" smooth mean diff per state", np.abs(a.x_smooth.mu - b.x_smooth.mu).max(axis=1))
    print(" smooth cov diff per state", np.abs(a.x_smooth.Sigma - b.x_smooth.Sigma).max(axis=(1, 2)))
    k = len(a.t) - 2
    print(" unrolled smooth mean[k]", a.x_smooth.mu[k], "\n rolled               ", b.x_smooth.mu[k])
    print(" unrolled diag cov[k]", np.diag(a.x_smooth.Sigma[k]), "\n rolled            ", np.diag(b.x_smooth.Sigma[k]))
