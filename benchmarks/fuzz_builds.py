"""Randomised differential run between BUILDS of the same model: the ahead-of-time catalogue kernels, the run-time
compiled (NVRTC, unrolled) kernels and the rolled general-(d, q) fallback (PNDE_FORCE_ROLLED=1) of Lotka-Volterra.
Filter, smoother, dense output and sampling must agree to rounding; prints the worst relative difference per case."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import odefilters_b200 as B

F = "du[0] = p[0]*u[0] - p[1]*u[0]*u[1]; du[1] = -p[2]*u[1] + p[3]*u[0]*u[1];"
J = "J[0][0] = p[0]-p[1]*u[1]; J[0][1] = -p[1]*u[0]; J[1][0] = p[3]*u[1]; J[1][1] = -p[2]+p[3]*u[0];"


def outputs(prob, alg, kw):
    s = B.solve(prob, alg, **kw)
    tq = np.linspace(0.05, 0.95 * prob.tspan[1], 5)
    return dict(t=np.asarray(s.t), u=s.u, fm=s.x_filt.mu, fc=s.x_filt.Sigma, sm=s.x_smooth.mu, sc=s.x_smooth.Sigma,
                dm=s(tq).mu, smp=s.sample(2, seed=3), ds=s.dense_sample(2, seed=4, n_times=23)[0], counts=s.destats)


def rel(a, b):
    sc = np.max(np.abs(a)) + 1e-300
    return float(np.max(np.abs(a - b)) / sc)


def run(seed=0, ncase=12, verbose=True):
    rng = np.random.default_rng(seed)
    bad = []
    for case in range(ncase):
        kind = rng.choice(["EK0", "EK1"])
        q = int(rng.integers(1, 5))
        diffusion = str(rng.choice(["dynamic", "fixed", "fixedMAP"] + (["dynamicMV", "fixedMV"] if kind == "EK0" else [])))
        adaptive = bool(rng.integers(0, 2))
        T = float(rng.choice([1.0, 2.0]))
        kw = dict(abstol=1e-6, reltol=float(rng.choice([1e-3, 1e-4]))) if adaptive else dict(adaptive=False, dt=float(rng.choice([0.05, 0.02])))
        alg = (B.EK0 if kind == "EK0" else B.EK1)(order=q, diffusionmodel=diffusion, smooth=True)
        u0, p = [1.0, 1.0], (1.5, 1.0, 3.0, 1.0)
        ref = outputs(B.ODEProblem("lotka_volterra", u0, (0.0, T), p), alg, kw)
        tag = dict(case=case, alg=kind, q=q, diffusion=diffusion, T=T, **kw)
        for build in ("nvrtc", "rolled"):
            if build == "rolled":
                os.environ["PNDE_FORCE_ROLLED"] = "1"
            else:
                os.environ.pop("PNDE_FORCE_ROLLED", None)
            out = outputs(B.ODEProblem(B.CustomVectorField(d=2, n_params=4, f=F, jac=J), u0, (0.0, T), p), alg, kw)
            os.environ.pop("PNDE_FORCE_ROLLED", None)
            if out["counts"] != ref["counts"] or len(out["t"]) != len(ref["t"]):
                tag[build] = "counts differ"
                bad.append(dict(tag))
                continue
            worst = max(rel(ref[k], out[k]) for k in ("t", "u", "fm", "sm", "dm", "smp", "ds"))
            sd = np.sqrt(np.abs(np.diagonal(ref["sc"], axis1=1, axis2=2)).max(axis=0))
            sd = np.maximum(sd, 1e-12 * sd.max())
            wc = max(float(np.max(np.abs(ref[k] - out[k]) / np.outer(sd_, sd_))) for k, sd_ in (("sc", sd), ("fc", np.maximum(np.sqrt(np.abs(np.diagonal(ref["fc"], axis1=1, axis2=2)).max(axis=0)), 1e-12 * sd.max()))))
            tag[build] = dict(worst_rel=worst, worst_cov=wc)
            if worst > 1e-6 or wc > 1e-5:
                bad.append(dict(tag))
        if verbose:
            print(json.dumps(tag), flush=True)
    return bad, ncase


if __name__ == "__main__":
    bad, ncase = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 12)
    print("DISAGREE", len(bad), "of", ncase)
    for b in bad:
        print("  ", json.dumps(b))
