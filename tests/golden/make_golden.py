#!/usr/bin/env python
"""Regenerates tests/golden/*.npz / reference_literals.json.

The reference is a Julia package and Julia is not installed in the build image, so the fixtures cannot come
from running it.  They are:
  * reference_literals.json -- the literal numbers the reference's OWN tests hold for this path, copied with
    their file:line (the only golden values the reference has);
  * oracle_*.npz -- trajectories produced by oracle/pnde_oracle.py (the numpy restatement, pinned to the
    literals above by tests/test_oracle_golden.py) for the BASELINE configs at small ensemble sizes.
Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pnde_oracle as O  # noqa: E402

SEED = 20260118


def literals():
    return {
        "test/specific_problems.jl:147-148": {
            "what": "ForwardDiff.gradient of norm(sol.u[end]) w.r.t. p, prob_ode_fitzhughnagumo, EK1(order=3), adaptive",
            "value": [0.026680212891877435, -0.028019989130281753, 0.3169977494388167, 0.6749351039218744]},
        "test/specific_problems.jl:154-155": {
            "what": "same, w.r.t. u0 (NOT reproduced by the oracle: TaylorSeries x Dual initialisation; unpinned)",
            "value": [0.6500925873857853, -0.004812245513746423]},
        "test/priors.jl:50-59": {
            "what": "preconditioned ibm(1,2): A and Q",
            "A": [[1, 1, 0.5], [0, 1, 1], [0, 0, 1]],
            "Q": [[1 / 20, 1 / 8, 1 / 6], [1 / 8, 1 / 3, 1 / 2], [1 / 6, 1 / 2, 1]]},
    }


def pack(sol, d):
    return dict(t=np.array(sol.t), mean=np.array([g.mu for g in sol.x_filt]),
                cov_u=np.array([g.Sigma.mat[:d, :d] for g in sol.x_filt]),
                diffusions=np.array([np.atleast_1d(x)[0] for x in sol.diffusions]),
                counts=np.array([sol.naccept, sol.nreject, sol.nf]), loglik=np.array(sol.log_likelihood),
                smooth_mean=np.array([g.mu for g in sol.x_smooth]) if sol.x_smooth is not None else np.zeros(0),
                smooth_cov_u=np.array([g.Sigma.mat[:d, :d] for g in sol.x_smooth]) if sol.x_smooth is not None else np.zeros(0))


def main():
    json.dump(literals(), open(os.path.join(HERE, "reference_literals.json"), "w"), indent=1)
    rng = np.random.default_rng(SEED)
    # config 1: README example (README.md:36-47)
    s = O.solve_ivp(O.Problem(O.CATALOGUE["fhn_readme"], [-1.0, 1.0], (0.0, 20.0), [0.2, 0.2, 3.0]), O.EK0(order=1),
                    abstol=1e-1, reltol=1e-2)
    np.savez_compressed(os.path.join(HERE, "oracle_config1_fhn_readme_ek0q1.npz"), **pack(s, 2))
    # config 2 (small): 4 draws of the parameter sweep, EK1(3), dt = 0.01, first 200 steps
    P = np.stack([rng.uniform(0.1, 0.3, 4), rng.uniform(0.1, 0.3, 4), rng.uniform(2, 4, 4)], axis=1)
    for i in range(4):
        s = O.solve_ivp(O.Problem(O.CATALOGUE["fhn_readme"], [-1.0, 1.0], (0.0, 2.0), list(P[i])),
                        O.EK1(order=3, smooth=False), adaptive=False, dt=0.01)
        np.savez_compressed(os.path.join(HERE, f"oracle_config2_fhn_ek1q3_traj{i}.npz"), p=P[i], **pack(s, 2))
    # config 3 (small): Van der Pol mu = 1e3, EK1(5), adaptive
    s = O.solve_ivp(O.Problem(O.CATALOGUE["vanderpol"], [0.0, 3.0 ** 0.5], (0.0, 1.0), [1e3]), O.EK1(order=5, smooth=False))
    np.savez_compressed(os.path.join(HERE, "oracle_config3_vdp_ek1q5.npz"), chol_fail=np.array(s.stats.get("chol_fail", 0)),
                        **pack(s, 2))
    # config 5 (small): Lotka-Volterra, EK1(3), dt = 0.05 on (0,10), filter + smoother
    s = O.solve_ivp(O.Problem(O.CATALOGUE["lotka_volterra"], [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0]),
                    O.EK1(order=3, smooth=True), adaptive=False, dt=0.05)
    np.savez_compressed(os.path.join(HERE, "oracle_config5_lv_ek1q3_smooth.npz"), **pack(s, 2))


if __name__ == "__main__":
    main()
