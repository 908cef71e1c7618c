"""CPU tests: pin the oracle against every golden vector / known-answer test the reference holds
for this path, and the C restatement against the oracle."""
import math

import numpy as np
import pytest

import pnde_oracle as O


# ---- test/priors.jl:25-59 ------------------------------------------------------------------
def test_vanilla_ibm_literal_matrices():
    h, s = 0.37, 0.81
    A, Q = O.vanilla_ibm(2, 2, h, s ** 2)
    AH = np.array([[1, 0, h, 0, h ** 2 / 2, 0], [0, 1, 0, h, 0, h ** 2 / 2], [0, 0, 1, 0, h, 0], [0, 0, 0, 1, 0, h],
                   [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]], dtype=float)
    QH = s ** 2 * np.array([
        [h ** 5 / 20, 0, h ** 4 / 8, 0, h ** 3 / 6, 0], [0, h ** 5 / 20, 0, h ** 4 / 8, 0, h ** 3 / 6],
        [h ** 4 / 8, 0, h ** 3 / 3, 0, h ** 2 / 2, 0], [0, h ** 4 / 8, 0, h ** 3 / 3, 0, h ** 2 / 2],
        [h ** 3 / 6, 0, h ** 2 / 2, 0, h, 0], [0, h ** 3 / 6, 0, h ** 2 / 2, 0, h]])
    assert np.allclose(A, AH, rtol=1e-14) and np.allclose(Q, QH, rtol=1e-14)


def test_preconditioned_ibm_literal_matrices():
    A, Q = O.ibm(1, 2)
    assert np.allclose(A, [[1, 1, 0.5], [0, 1, 1], [0, 0, 1]], rtol=1e-15)
    assert np.allclose(Q.mat, [[1 / 20, 1 / 8, 1 / 6], [1 / 8, 1 / 3, 1 / 2], [1 / 6, 1 / 2, 1]], rtol=1e-14)


@pytest.mark.parametrize("q", [1, 2, 3, 4, 5])
def test_prior_dim(q):  # test/priors.jl:64-74
    A, Q = O.ibm(2, q)
    assert A.shape == (2 * (q + 1),) * 2 and Q.mat.shape == A.shape


# ---- test/preconditioning.jl:12-39 ---------------------------------------------------------
def test_preconditioner_identity():
    h, s, d, q = 0.043, 0.7, 2, 3
    Ah, Qh = O.vanilla_ibm(d, q, h, s ** 2)
    Ap, Qp = O.ibm(d, q)
    P = np.diag(O.preconditioner_diag(d, q, h))
    PI = np.linalg.inv(P)
    assert np.allclose(Qp.mat * s ** 2, P @ Qh @ P.T, rtol=1e-10)
    assert np.allclose(Ap, P @ Ah @ PI, rtol=1e-12)
    assert np.linalg.cond(Qh) > np.linalg.cond(Qp.mat * s ** 2) ** 2


# ---- test/filtering.jl ---------------------------------------------------------------------
def _rand_setup(seed):
    rng = np.random.default_rng(seed)
    d = 5
    m = rng.random(d)
    Lp = np.tril(rng.random((d, d)))
    A = rng.random((d, d))
    LQ = np.tril(rng.random((d, d)))
    return rng, d, m, Lp, A, LQ


def test_predict_dense_and_sr():
    _, d, m, Lp, A, LQ = _rand_setup(0)
    P, Q = Lp @ Lp.T, LQ @ LQ.T
    m_p, P_p = A @ m, A @ P @ A.T + Q
    out = O.predict(O.Gaussian(m, P), A, Q)
    assert np.array_equal(out.mu, m_p) and np.array_equal(out.Sigma, P_p)
    out = O.predict(O.Gaussian(m, O.SRMatrix(Lp)), A, O.SRMatrix(LQ))
    assert np.array_equal(out.mu, m_p) and np.allclose(out.Sigma.mat, P_p, rtol=1e-12)


def test_update_vs_textbook():
    rng, d, m_p, Lp, _, _ = _rand_setup(1)
    P_p = Lp @ Lp.T
    H = rng.random((3, d))
    z, S = H @ m_p, H @ P_p @ H.T
    K = P_p @ H.T @ np.linalg.inv(S)
    m, P = m_p + K @ (0 - z), P_p - K @ S @ K.T
    out = O.update(O.Gaussian(m_p, P_p), O.Gaussian(z, S), H)
    assert np.allclose(out.mu, m, rtol=1e-13) and np.allclose(out.Sigma, P, atol=1e-12)
    out = O.update(O.Gaussian(m_p, O.SRMatrix(Lp)), O.Gaussian(z, S), H)
    assert np.allclose(out.Sigma.mat, P, atol=1e-12)


def test_sr_smooth_vs_textbook_rts():
    rng, d, m, Lp, A, LQ = _rand_setup(2)
    m_s, Ls = rng.random(d), np.tril(rng.random((d, d)))
    P, P_s, Q = Lp @ Lp.T, Ls @ Ls.T, LQ @ LQ.T
    m_p, P_p = A @ m, A @ P @ A.T + Q
    G = P @ A.T @ np.linalg.inv(P_p)
    ms, Ps = m + G @ (m_s - m_p), P + G @ (P_s - P_p) @ G.T
    out, _ = O.smooth(O.Gaussian(m, O.SRMatrix(Lp)), O.Gaussian(m_s, O.SRMatrix(Ls)), A, O.SRMatrix(LQ))
    assert np.allclose(out.mu, ms, rtol=1e-10) and np.allclose(out.Sigma.mat, Ps, atol=1e-9)


# ---- test/state_init.jl:21-45 --------------------------------------------------------------
def test_taylor_mode_initial_derivatives():
    a, b, q = 1.1, -0.5, 6
    u0 = [0.1, 1.0]
    dfs = O.get_derivatives(u0, O.CATALOGUE["linear2"], [a, b], 0.0, q)
    truth = [[a ** k * u0[0], b ** k * u0[1]] for k in range(1, q + 1)]
    assert np.allclose(np.array(dfs), np.array(truth), rtol=1e-13)


def test_initial_state_is_exact():  # test/solution.jl:38-41
    prob = O.Problem(O.CATALOGUE["lotka_volterra"], [1.0, 1.0], (0.0, 0.1), [1.5, 1.0, 3.0, 1.0])
    sol = O.solve_ivp(prob, O.EK1(order=3, smooth=False))
    assert np.array_equal(sol.pu[0].mu, [1.0, 1.0]) and not sol.pu[0].Sigma.mat.any()
    assert len(sol.t) == sol.naccept + 1 and sol.t[0] == 0.0 and sol.t[-1] == 0.1  # test/solution.jl:20-28


# ---- test/specific_problems.jl:141-148: the only end-to-end numeric golden vector -----------
def test_golden_parameter_gradient():
    p0 = [0.7, 0.8, 1 / 12.5, 0.5]
    pd = [O.Dual(p0[i], np.eye(4)[i]) for i in range(4)]
    prob = O.Problem(O.CATALOGUE["fhn_lib"], [1.0, 1.0], (0.0, 1.0), pd)
    sol = O.solve_ivp(prob, O.EK1(order=3), dtype=object)
    ue = sol.u[-1]
    nrm = (ue[0] * ue[0] + ue[1] * ue[1]).sqrt()
    golden = np.array([0.026680212891877435, -0.028019989130281753, 0.3169977494388167, 0.6749351039218744])
    # Julia's `≈` is rtol = sqrt(eps) ~ 1.5e-8
    assert np.max(np.abs(nrm.p - golden) / np.abs(golden)) < 1.5e-8
    assert sol.naccept == 9 and sol.nreject == 0


# ---- accuracy gates: test/correctness.jl, test/smoothing.jl, test/convergence.jl ------------
def _rk4_ref(f, u0, p, t1, n=20000):
    u = np.array(u0, dtype=float)
    h = t1 / n
    F = lambda x: np.array(f(list(x), p, 0.0))
    for _ in range(n):
        k1 = F(u); k2 = F(u + h / 2 * k1); k3 = F(u + h / 2 * k2); k4 = F(u + h * k3)
        u = u + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return u


@pytest.mark.parametrize("kind", ["EK0", "EK1"])
@pytest.mark.parametrize("diffusion", ["fixed", "dynamic", "fixedMAP", "fixedMV", "dynamicMV"])
def test_constant_step_accuracy(kind, diffusion):  # test/correctness.jl:15-39
    if kind == "EK1" and diffusion.endswith("MV"):
        pytest.skip("MV diffusions are EK0-only")
    vf = O.CATALOGUE["lotka_volterra"]
    p = [1.5, 1.0, 3.0, 1.0]
    sol = O.solve_ivp(O.Problem(vf, [1.0, 1.0], (0.0, 1.0), p), O.Alg(kind, 3, diffusion, True), adaptive=False, dt=5e-3)
    truth = _rk4_ref(vf.f, [1.0, 1.0], p, 1.0)
    assert np.allclose(sol.u[-1], truth, rtol=1e-5)
    if diffusion.startswith("fixed"):
        assert math.isnan(sol.log_likelihood)


def test_smooth_vs_nonsmooth():  # test/smoothing.jl:24-48
    vf = O.CATALOGUE["lotka_volterra"]
    p = [1.5, 1.0, 3.0, 1.0]
    prob = O.Problem(vf, [1.0, 1.0], (0.0, 1.0), p)
    s1 = O.solve_ivp(prob, O.EK0(order=3, smooth=False), adaptive=False, dt=1e-2)
    s2 = O.solve_ivp(prob, O.EK0(order=3, smooth=True), adaptive=False, dt=1e-2)
    assert np.allclose(s1.t, s2.t) and np.array_equal(s1.u[-1], s2.u[-1]) and not np.array_equal(s1.u[-2], s2.u[-2])


@pytest.mark.parametrize("q", [1, 2, 3])
def test_convergence_order(q):  # test/convergence.jl:17-30 (Float64 instead of BigFloat)
    vf = O.CATALOGUE["linear1"]
    errs = []
    for dt in (1 / 16, 1 / 32):
        sol = O.solve_ivp(O.Problem(vf, [0.5], (0.0, 1.0), [1.01]), O.EK0(order=q, smooth=False), adaptive=False, dt=dt)
        errs.append(abs(sol.u[-1][0] - 0.5 * math.exp(1.01)))
    assert abs(math.log2(errs[0] / errs[1]) - (q + 1)) < 0.35


def test_fixed_step_requires_dt():  # test/errors.jl:16-20
    with pytest.raises(ValueError):
        O.solve_ivp(O.Problem(O.CATALOGUE["lotka_volterra"], [1.0, 1.0], (0.0, 1.0), [1.5, 1, 3, 1]), O.EK0(),
                    adaptive=False)


def test_fixed_step_grid_hazard():  # t += dt accumulation and the 10-ulp snap
    vf = O.CATALOGUE["lotka_volterra"]
    s = O.solve_ivp(O.Problem(vf, [1.0, 1.0], (0.0, 1.0), [1.5, 1, 3, 1]), O.EK0(order=1, smooth=False), adaptive=False,
                    dt=5e-3)
    assert s.naccept == 200 and s.t[-1] == 1.0


# ---- the C restatement (CPU baseline) against the oracle -------------------------------------
@pytest.mark.parametrize("vf,p,u0", [("fhn_readme", [0.2, 0.2, 3.0], [-1.0, 1.0]),
                                    ("lotka_volterra", [1.5, 1.0, 3.0, 1.0], [1.0, 1.0])])
@pytest.mark.parametrize("kind,q", [("EK1", 1), ("EK1", 3), ("EK0", 2)])
def test_c_restatement_fixed(vf, p, u0, kind, q):
    import pnde_ref as R

    so = O.solve_ivp(O.Problem(O.CATALOGUE[vf], u0, (0.0, 1.0), p), O.Alg(kind, q, "dynamic", False), adaptive=False,
                     dt=0.01)
    r = R.solve_ensemble(vf, kind, q, [u0], [p], (0.0, 1.0), adaptive=False, dt=0.01)
    ref = so.x_filt[-1]
    assert np.max(np.abs(r["mean"][0][:2] - ref.mu[:2])) < 1e-11
    tol = {1: 1e-11, 2: 1e-7, 3: 1e-5}[q]
    assert np.max(np.abs(r["cov"][0] - ref.Sigma.mat)) / np.max(np.abs(ref.Sigma.mat)) < tol
    assert r["naccept"][0] == so.naccept and r["nf"][0] == so.nf


def test_c_restatement_adaptive_counts():
    import pnde_ref as R

    so = O.solve_ivp(O.Problem(O.CATALOGUE["fhn_lib"], [1.0, 1.0], (0.0, 1.0), [0.7, 0.8, 1 / 12.5, 0.5]),
                     O.EK1(order=3, smooth=False))
    r = R.solve_ensemble("fhn_lib", "EK1", 3, [[1.0, 1.0]], [[0.7, 0.8, 1 / 12.5, 0.5]], (0.0, 1.0))
    assert (r["naccept"][0], r["nreject"][0], r["nf"][0]) == (so.naccept, so.nreject, so.nf) == (7, 0, 9)
    assert np.allclose(r["mean"][0][:2], so.x_filt[-1].mu[:2], rtol=1e-9)


# ---- the Kronecker EK0 model (checker of the large-d GPU path) against the dense oracle ---------
@pytest.mark.parametrize("adaptive", [False, True])
@pytest.mark.parametrize("diffusion", ["dynamic", "fixed"])
def test_kronecker_model_equals_dense_oracle(adaptive, diffusion):
    import kron_model as KM

    d, F, q = 8, 8.0, 3
    u0 = F + 0.01 * np.random.default_rng(0).standard_normal(d)
    kw = dict(adaptive=False, dt=0.01) if not adaptive else dict()
    so = O.solve_ivp(O.Problem(O.lorenz96(d), list(u0), (0.0, 0.5), [F]), O.Alg("EK0", q, diffusion, False), **kw)
    km = KM.solve_ek0_kron(lambda u: KM.lorenz96_f(u, F), KM.lorenz96_jets(u0, F, q), (0.0, 0.5), q,
                           diffusion=diffusion, **kw)
    assert (so.naccept, so.nreject, so.nf) == (km["naccept"], km["nreject"], km["nf"])
    ref = so.x_filt[-1]
    assert np.max(np.abs(km["M"].ravel() - ref.mu)) / np.max(np.abs(ref.mu)) < 1e-10
    assert np.max(np.abs(np.kron(km["C"], np.eye(d)) - ref.Sigma.mat)) / np.max(np.abs(ref.Sigma.mat)) < 1e-9


# ---- committed fixtures (tests/golden/): the oracle must keep reproducing them ------------------------
def test_oracle_reproduces_committed_fixtures():
    import json
    import os

    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    lit = json.load(open(os.path.join(gdir, "reference_literals.json")))
    A, Q = O.ibm(1, 2)
    assert np.allclose(A, lit["test/priors.jl:50-59"]["A"]) and np.allclose(Q.mat, lit["test/priors.jl:50-59"]["Q"])
    g = np.load(os.path.join(gdir, "oracle_config1_fhn_readme_ek0q1.npz"))
    s = O.solve_ivp(O.Problem(O.CATALOGUE["fhn_readme"], [-1.0, 1.0], (0.0, 20.0), [0.2, 0.2, 3.0]), O.EK0(order=1),
                    abstol=1e-1, reltol=1e-2)
    assert [s.naccept, s.nreject, s.nf] == list(g["counts"]) and np.allclose(np.array(s.t), g["t"], rtol=1e-12)
    g = np.load(os.path.join(gdir, "oracle_config5_lv_ek1q3_smooth.npz"))
    s = O.solve_ivp(O.Problem(O.CATALOGUE["lotka_volterra"], [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0]),
                    O.EK1(order=3, smooth=True), adaptive=False, dt=0.05)
    assert np.allclose(np.array([x.mu for x in s.x_smooth]), g["smooth_mean"], rtol=1e-9, atol=1e-12)


def test_ieks_oracle_first_iterate_is_ek1_and_iterates_contract():
    """src/ieks.jl:53-61 + src/perform_step.jl:111-113: iterate 1 has no linearisation trajectory and equals an EK1
    solve; on a fixed grid the iterates then contract towards the MAP estimate."""
    prob = O.Problem(O.CATALOGUE["fhn_lib"], [1.0, 1.0], (0.0, 5.0), [0.7, 0.8, 1 / 12.5, 0.5])
    kw = dict(adaptive=False, dt=0.05)
    s1 = O.solve_ieks(prob, O.IEKS(order=2), iterations=1, **kw)
    se = O.solve_ivp(prob, O.EK1(order=2, smooth=True), **kw)
    assert np.array_equal(np.array(s1.u), np.array(se.u))
    its = [np.array(O.solve_ieks(prob, O.IEKS(order=2), iterations=k, **kw).u) for k in (1, 2, 3)]
    d12, d23 = np.abs(its[1] - its[0]).max(), np.abs(its[2] - its[1]).max()
    assert 0 < d23 < d12
    # the reference's own test (test/ieks.jl:10-13: runs and returns a solution), with fewer iterates
    s = O.solve_ieks(O.Problem(O.CATALOGUE["fhn_lib"], [1.0, 1.0], (0.0, 10.0), [0.7, 0.8, 1 / 12.5, 0.5]),
                     O.IEKS(order=4, diffusionmodel="fixed"), iterations=3)
    assert s.retcode == "Success" and s.t[-1] == 10.0 and s.smoothed


def test_dense_sample_law_on_the_solver_grid_is_the_smoother():
    """oracle.dense_sample_law (the marginal law of src/solution_sampling.jl:63-74's draws) pinned on the one case with
    a known answer: on the solver's own grid the backward recursion is smooth_all (src/smoothing.jl:4-28)."""
    vf = O.CATALOGUE["lotka_volterra"]
    so = O.solve_ivp(O.Problem(vf, [1.0, 1.0], (0.0, 2.0), [1.5, 1.0, 3.0, 1.0]), O.Alg("EK1", 2, "dynamic", True),
                     adaptive=False, dt=0.1)
    law = O.dense_sample_law(so, np.array(so.t))
    for i in range(1, len(so.t)):
        assert np.allclose(law[i].mu, so.x_smooth[i].mu, rtol=1e-10, atol=1e-12)
        assert np.allclose(law[i].Sigma.mat, so.x_smooth[i].Sigma.mat, rtol=1e-8, atol=1e-16)
