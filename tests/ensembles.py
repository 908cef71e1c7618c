"""Seeded synthetic ensembles of the BASELINE configs and the ensemble-scale parity statistic
(SURVEY 8c protocol (iii)): the fraction of trajectories whose (naccept, nreject) equals the reference arithmetic's
(oracle/pnde_ref.c, the C restatement of the reference's dense algorithm) and the error of u(t1).

Shared by tests/test_gpu_parity.py and benchmarks/parity_ensemble.py.  TEST INFRASTRUCTURE."""
import numpy as np

SEED = 20260118


def config2_inputs(n, seed=SEED):
    """FHN sweep: a, b ~ U(0.1, 0.3), c ~ U(2, 4), u0 = (-1, 1)."""
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 0.3, n), rng.uniform(2.0, 4.0, n)], axis=1)
    return np.tile([-1.0, 1.0], (n, 1)), p


def config3_inputs(n, seed=SEED):
    """Van der Pol: mu ~ logU(5e2, 2e3), u0 = (0, sqrt 3) (1 + 0.01 N(0, 1))."""
    rng = np.random.default_rng(seed)
    mu = np.exp(rng.uniform(np.log(5e2), np.log(2e3), n))
    u0 = np.stack([np.zeros(n), np.sqrt(3.0) * (1 + 0.01 * rng.standard_normal(n))], axis=1)
    return u0, mu[:, None]


def config5_inputs(n, seed=SEED):
    """Lotka-Volterra: p = (1.5, 1, 3, 1) (1 + 0.1 U(-1, 1)), u0 = (1, 1)."""
    rng = np.random.default_rng(seed)
    p = np.array([1.5, 1.0, 3.0, 1.0]) * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
    return np.ones((n, 2)), p


ENSEMBLES = {
    # name: (vector field, order, tspan, inputs)
    "config3_vdp_ek1q5": ("vanderpol", 5, (0.0, 1.0), config3_inputs),
    "fhn_adaptive_ek1q3": ("fhn_readme", 3, (0.0, 20.0), config2_inputs),
}


def count_parity(name, n, abstol=1e-6, reltol=1e-3, device=-1):
    """Run ensemble `name` (n trajectories, adaptive EK1) on the GPU through the C ABI and on the host through the C
    restatement of the reference; returns the statistics dict and the raw arrays."""
    import odefilters_b200 as B
    import pnde_ref as R

    vf, q, tspan, make = ENSEMBLES[name]
    u0, p = make(n)
    prob = B.ODEProblem(vf, u0[0], tspan, p[0])
    s = B.FilterSolver(prob, B.EK1(order=q, smooth=False), abstol=abstol, reltol=reltol, save_everystep=False,
                       device=device)
    s.solve_ensemble(u0, p)
    cg = s.counts()
    mg = s.final()[0]
    s.close()
    ref = R.solve_ensemble(vf, "EK1", q, u0, p, tspan, abstol=abstol, reltol=reltol, want_cov=False)
    same = (cg["naccept"] == ref["naccept"]) & (cg["nreject"] == ref["nreject"])
    d = u0.shape[1]
    scale = np.maximum(np.abs(ref["mean"][:, :d]).max(axis=1), 1e-300)
    relu = np.abs(mg[:, :d] - ref["mean"][:, :d]).max(axis=1) / scale
    stats = {
        "ensemble": name, "n": int(n), "frac_identical_counts": float(same.mean()),
        "n_differ": int((~same).sum()),
        "max_abs_dnaccept": int(np.abs(cg["naccept"] - ref["naccept"]).max()),
        "max_abs_dnreject": int(np.abs(cg["nreject"] - ref["nreject"]).max()),
        "mean_naccept_gpu": float(cg["naccept"].mean()), "mean_naccept_ref": float(ref["naccept"].mean()),
        "mean_nreject_gpu": float(cg["nreject"].mean()), "mean_nreject_ref": float(ref["nreject"].mean()),
        "max_rel_u_identical": float(relu[same].max()) if same.any() else None,
        "max_rel_u_all": float(relu.max()),
        "median_rel_u": float(np.median(relu)),
        "all_success_gpu": bool((cg["retcode"] == 0).all()), "all_success_ref": bool((ref["retcode"] == 0).all()),
        "ref_chol_failures": int(ref["chol_fail"].sum()),
    }
    return stats, dict(gpu_counts=cg, gpu_mean=mg, ref=ref, same=same, relu=relu, u0=u0, p=p)
