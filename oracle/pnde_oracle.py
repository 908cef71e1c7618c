"""CPU oracle: a numpy restatement of ProbNumDiffEq v0.1.5's ODE-filter hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package, the C-ABI
library, the CUDA kernels) may import or call this file; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg do, and only
as the checker.

Every function cites the reference ``file:line`` (relative to /root/reference)
whose arithmetic it restates.  The restatement is deliberately *reference
faithful*: Cholesky-of-the-product predict with QR fallback, LU inverses,
``(I-KH) S`` update, P / P^-1 round trips, the rejected-step side effects.  The
time loop, PI controller, initdt and error norm live in the un-vendored
dependencies OrdinaryDiffEq 5.x / DiffEqBase 6.x (reference Project.toml:23,27,
no Manifest => unpinned); they are restated from their published algorithm
 and anchored on the reference's own golden vector
(test/specific_problems.jl:141-148), which this oracle reproduces
(tests/test_oracle_golden.py).

Pinning status: pinned against test/priors.jl:25-59 (literal matrices),
test/preconditioning.jl:29-38, test/filtering.jl (algebraic identities),
test/state_init.jl:21-45 (analytic derivatives), test/specific_problems.jl:148
(p-gradient through an adaptive EK1(order=3) solve).  Unpinned: the u0-gradient
(test/specific_problems.jl:155), log-likelihood, LV/VdP problem definitions,
the reject branch of the controller.

The code is dtype generic: float64 arrays use LAPACK (as the reference does
through Julia's LinearAlgebra); ``object`` arrays (forward-mode ``Dual`` numbers
or mpmath ``mpf``) use the plain-loop factorizations at the bottom.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------
# scalar helpers (work for float, Dual, mpf)
# ----------------------------------------------------------------------------


class Dual:
    """Forward-mode dual number with a vector of partials (ForwardDiff.Dual).

    Only what the golden-gradient test needs (test/specific_problems.jl:141-148).
    Comparisons act on the value, like ForwardDiff.
    """

    __slots__ = ("v", "p")
    __array_priority__ = 1000

    def __init__(self, v, p):
        self.v = float(v)
        self.p = np.asarray(p, dtype=float)

    @staticmethod
    def lift(x, n):
        return x if isinstance(x, Dual) else Dual(x, np.zeros(n))

    def _o(self, o):
        return o if isinstance(o, Dual) else Dual(o, np.zeros_like(self.p))

    def __add__(self, o):
        if not isinstance(o, _DUAL_SCALARS):
            return NotImplemented
        o = self._o(o)
        return Dual(self.v + o.v, self.p + o.p)

    __radd__ = __add__

    def __sub__(self, o):
        if not isinstance(o, _DUAL_SCALARS):
            return NotImplemented
        o = self._o(o)
        return Dual(self.v - o.v, self.p - o.p)

    def __rsub__(self, o):
        if not isinstance(o, _DUAL_SCALARS):
            return NotImplemented
        o = self._o(o)
        return Dual(o.v - self.v, o.p - self.p)

    def __mul__(self, o):
        if not isinstance(o, _DUAL_SCALARS):
            return NotImplemented
        o = self._o(o)
        return Dual(self.v * o.v, self.p * o.v + self.v * o.p)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if not isinstance(o, _DUAL_SCALARS):
            return NotImplemented
        o = self._o(o)
        return Dual(self.v / o.v, (self.p * o.v - self.v * o.p) / (o.v * o.v))

    def __rtruediv__(self, o):
        if not isinstance(o, _DUAL_SCALARS):
            return NotImplemented
        return self._o(o).__truediv__(self)

    def __neg__(self):
        return Dual(-self.v, -self.p)

    def __pos__(self):
        return self

    def __abs__(self):
        return self if self.v >= 0 else -self

    def __pow__(self, e):
        if isinstance(e, Dual):
            raise NotImplementedError
        return Dual(self.v ** e, e * self.v ** (e - 1) * self.p)

    def sqrt(self):
        s = math.sqrt(self.v)
        return Dual(s, self.p / (2.0 * s) if s != 0.0 else np.zeros_like(self.p))

    def log(self):
        return Dual(math.log(self.v), self.p / self.v)

    def __float__(self):
        return self.v

    def _cmp(self, o):
        return o.v if isinstance(o, Dual) else o

    def __lt__(self, o):
        return self.v < self._cmp(o)

    def __le__(self, o):
        return self.v <= self._cmp(o)

    def __gt__(self, o):
        return self.v > self._cmp(o)

    def __ge__(self, o):
        return self.v >= self._cmp(o)

    def __eq__(self, o):
        return self.v == self._cmp(o)

    def __ne__(self, o):
        return self.v != self._cmp(o)

    def __hash__(self):
        return hash(self.v)

    def __repr__(self):
        return f"Dual({self.v!r}, {self.p!r})"


_DUAL_SCALARS = (Dual, int, float, np.floating, np.integer)


def _value(x):
    return x.v if isinstance(x, Dual) else x


def _ssqrt(x):
    """sqrt of a scalar of any supported type."""
    if isinstance(x, (float, int, np.floating)):
        return math.sqrt(x)
    return x.sqrt()


def _slog(x):
    if isinstance(x, (float, int, np.floating)):
        return math.log(x)
    if isinstance(x, Dual):
        return x.log()
    import mpmath

    return mpmath.log(x)


def _is_float(a: np.ndarray) -> bool:
    return a.dtype == np.float64


def _sse(x):
    """DiffEqBase sse(): value^2 (+ sum of partials^2 for a Dual) -- App. B.2."""
    if isinstance(x, Dual):
        return x.v * x.v + float(np.dot(x.p, x.p))
    return x * x


def _totallength(u) -> int:
    n = 0
    for x in np.ravel(u):
        n += 1 + (len(x.p) if isinstance(x, Dual) else 0)
    return n


def internalnorm(u):
    """DiffEqBase.ODE_DEFAULT_NORM (App. B.2): RMS for arrays, abs for scalars.

    Returns a plain real number even for Duals (the partials enter the sum),
    which is why dt stays Float64 under ForwardDiff.
    """
    if np.ndim(u) == 0:
        x = u.item() if isinstance(u, np.ndarray) else u
        s = _sse(x)
        return _ssqrt(s) if not isinstance(s, float) else math.sqrt(s)
    flat = np.ravel(u)
    s = 0.0
    for x in flat:
        s = s + _sse(x)
    return _ssqrt(s / _totallength(u))


# ----------------------------------------------------------------------------
# dense factorizations, dtype generic
# ----------------------------------------------------------------------------


def chol_lower(a: np.ndarray):
    """cholesky!(Symmetric(A), check=false) -> (L, issuccess).

    float64: LAPACK potrf like Julia (src/filtering.jl:35-36).  object dtype: the
    textbook column loop.
    """
    if _is_float(a):
        if not np.all(np.isfinite(a)):
            return None, False
        try:
            return np.linalg.cholesky(a), True
        except np.linalg.LinAlgError:
            return None, False
    n = a.shape[0]
    L = np.zeros_like(a)
    L[...] = 0 * a[0, 0]
    for j in range(n):
        s = a[j, j]
        for k in range(j):
            s = s - L[j, k] * L[j, k]
        if not (_value(s) > 0):
            return None, False
        ljj = _ssqrt(s)
        L[j, j] = ljj
        for i in range(j + 1, n):
            t = a[i, j]
            for k in range(j):
                t = t - L[i, k] * L[j, k]
            L[i, j] = t / ljj
    return L, True


def inv(a: np.ndarray) -> np.ndarray:
    """inv(A) (LU, getrf+getri in the reference: src/filtering.jl:85,123,140)."""
    if _is_float(a):
        return np.linalg.inv(a)
    n = a.shape[0]
    m = np.empty((n, 2 * n), dtype=object)
    zero = 0 * a[0, 0]
    for i in range(n):
        for j in range(n):
            m[i, j] = a[i, j]
            m[i, n + j] = zero + (1.0 if i == j else 0.0)
    for c in range(n):
        piv = max(range(c, n), key=lambda r: abs(_value(m[r, c])))
        if piv != c:
            m[[c, piv]] = m[[piv, c]]
        pv = m[c, c]
        for j in range(2 * n):
            m[c, j] = m[c, j] / pv
        for r in range(n):
            if r != c:
                fct = m[r, c]
                for j in range(2 * n):
                    m[r, j] = m[r, j] - fct * m[c, j]
    return m[:, n:].copy()


def solve(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """A \\ b (src/diffusions.jl:78)."""
    if _is_float(a) and _is_float(np.asarray(b)):
        return np.linalg.solve(a, b)
    return inv(a) @ b


def qr_r(a: np.ndarray) -> np.ndarray:
    """R factor of qr(A) (src/filtering.jl:43,149; src/smoothing.jl:56)."""
    if _is_float(a):
        return np.linalg.qr(a, mode="r")
    m, n = a.shape
    r = a.copy()
    for k in range(min(m, n)):
        nrm2 = 0 * r[0, 0]
        for i in range(k, m):
            nrm2 = nrm2 + r[i, k] * r[i, k]
        if _value(nrm2) == 0:
            continue
        nrm = _ssqrt(nrm2)
        alpha = -nrm if _value(r[k, k]) >= 0 else nrm
        v = [r[i, k] for i in range(k, m)]
        v[0] = v[0] - alpha
        vtv = 0 * nrm2
        for x in v:
            vtv = vtv + x * x
        for j in range(k, n):
            dot = 0 * nrm2
            for i in range(k, m):
                dot = dot + v[i - k] * r[i, j]
            s = 2 * dot / vtv
            for i in range(k, m):
                r[i, j] = r[i, j] - s * v[i - k]
    return np.triu(r[:n, :]) if _is_float(r) else _triu_obj(r[:n, :])


def _triu_obj(r):
    out = r.copy()
    for i in range(out.shape[0]):
        for j in range(min(i, out.shape[1])):
            out[i, j] = 0 * out[i, j]
    return out


# ----------------------------------------------------------------------------
# SRMatrix / Gaussian (src/squarerootmatrix.jl:9-42, src/ProbNumDiffEq.jl:37-66)
# ----------------------------------------------------------------------------


class SRMatrix:
    """PSD matrix as factor + eager dense product (src/squarerootmatrix.jl:10-16)."""

    __slots__ = ("squareroot", "mat")

    def __init__(self, S, mat=None):
        self.squareroot = S
        self.mat = S @ S.T if mat is None else mat

    def copy(self):
        return SRMatrix(self.squareroot.copy(), self.mat.copy())


def X_A_Xt(M, X):
    """X*M*X' ; for an SRMatrix: SRMatrix(X*S) (src/squarerootmatrix.jl:38-39,
    src/ProbNumDiffEq.jl:37)."""
    if isinstance(M, SRMatrix):
        return SRMatrix(X @ M.squareroot)
    return X @ M @ X.T


def apply_diffusion(Q: SRMatrix, diffusion):
    """src/ProbNumDiffEq.jl:38-39: scalar -> SRMatrix(sqrt(s)*Q_L);
    Diagonal (given as a vector of its diagonal) -> X_A_Xt(Q, sqrt.(diffusion))."""
    if np.ndim(diffusion) == 0:
        return SRMatrix(_ssqrt(diffusion) * Q.squareroot)
    sd = np.array([_ssqrt(x) for x in diffusion], dtype=Q.squareroot.dtype)
    return SRMatrix(sd[:, None] * Q.squareroot)


@dataclass
class Gaussian:
    mu: np.ndarray
    Sigma: object  # SRMatrix or dense ndarray

    def copy(self):
        return Gaussian(self.mu.copy(), self.Sigma.copy())


def affine(M, g: Gaussian) -> Gaussian:
    """M * g::SRGaussian (src/ProbNumDiffEq.jl:58)."""
    return Gaussian(M @ g.mu, X_A_Xt(g.Sigma, M))


def _dense(Sigma):
    return Sigma.mat if isinstance(Sigma, SRMatrix) else Sigma


# ----------------------------------------------------------------------------
# Prior, preconditioner (src/priors.jl:7-99, src/preconditioning.jl:1-17)
# ----------------------------------------------------------------------------


def ibm(d: int, q: int):
    """Preconditioned IWP: A = Atilde (x) I_d, Q = SRMatrix(chol(Qtilde (x) I_d).L)
    (src/priors.jl:7-59)."""
    D = d * (q + 1)
    A = np.eye(D)
    val = 1.0
    for i in range(1, q + 1):
        val = val / i
        for j in range(d * (q + 1 - i)):
            A[j, j + d * i] = val
    Qb = np.zeros((D, D))
    for col in range(q + 1):
        for row in range(col, q + 1):
            idx = 2 * q + 1 - row - col
            v = 1.0 / (idx * math.factorial(q - row) * math.factorial(q - col))
            for i in range(d):
                Qb[col * d + i, row * d + i] = v
                Qb[row * d + i, col * d + i] = v
    QL = np.linalg.cholesky(Qb)
    return A, SRMatrix(QL)


def vanilla_ibm(d: int, q: int, h: float, sigma2: float = 1.0):
    """Un-preconditioned A(h), Q(h) (src/priors.jl:63-99; only used by tests)."""
    D = d * (q + 1)
    A = np.eye(D)
    val = 1.0
    for i in range(1, q + 1):
        val = val * h / i
        for j in range(d * (q + 1 - i)):
            A[j, j + d * i] = val
    Q = np.zeros((D, D))
    for col in range(q + 1):
        for row in range(col, q + 1):
            idx = 2 * q + 1 - row - col
            v = h ** idx / (idx * math.factorial(q - row) * math.factorial(q - col)) * sigma2
            for i in range(d):
                Q[col * d + i, row * d + i] = v
                Q[row * d + i, col * d + i] = v
    return A, Q


def preconditioner_diag(d: int, q: int, h):
    """diag of P(h) = diag(h^(j-q-1/2)) (x) I_d, built by repeated val *= h
    (src/preconditioning.jl:4-13)."""
    val = h ** (-q - 0.5)
    out = []
    for _ in range(q + 1):
        out.extend([val] * d)
        val = val * h
    if isinstance(val, (float, np.floating)):
        return np.array(out, dtype=float)
    return np.array(out, dtype=object)


def proj(d: int, q: int, deriv: int) -> np.ndarray:
    """Proj(deriv) = e_{deriv+1}' (x) I_d (src/caches.jl:63-64)."""
    E = np.zeros((d, d * (q + 1)))
    for i in range(d):
        E[i, deriv * d + i] = 1.0
    return E


# ----------------------------------------------------------------------------
# Kalman algebra (src/filtering.jl)
# ----------------------------------------------------------------------------


def predict_mean(x: Gaussian, A):
    """src/filtering.jl:22-25."""
    return A @ x.mu


def predict_cov(x: Gaussian, A, Q, stats=None):
    """src/filtering.jl:26-48: dense version, or SR version (Cholesky of the
    product first, QR of [A S, Q_L]' on failure)."""
    if not isinstance(x.Sigma, SRMatrix):
        return X_A_Xt(x.Sigma, A) + _dense(Q)
    L = np.concatenate([A @ x.Sigma.squareroot, Q.squareroot], axis=1)
    out_cov = L @ L.T
    # Symmetric(...) reads the upper triangle
    out_cov = np.triu(out_cov) + np.triu(out_cov, 1).T if _is_float(out_cov) else out_cov
    PpL, ok = chol_lower(out_cov)
    if ok:
        return SRMatrix(PpL)
    if stats is not None:
        stats["chol_fail"] = stats.get("chol_fail", 0) + 1
    R = qr_r(L.T)
    return SRMatrix(R.T.copy())


def predict(x: Gaussian, A, Q, stats=None) -> Gaussian:
    """src/filtering.jl:17-21,56-60."""
    return Gaussian(predict_mean(x, A), predict_cov(x, A, Q, stats))


def update(x_pred: Gaussian, meas: Gaussian, H) -> Gaussian:
    """src/filtering.jl:79-91 (R == 0): K = P H' inv(S); mu = m - K z;
    Sigma = X_A_Xt(P, I-KH)."""
    z, S = meas.mu, _dense(meas.Sigma)
    S_inv = inv(S)
    K = _dense(x_pred.Sigma) @ H.T @ S_inv
    mu = x_pred.mu + K @ (0 - z)
    I = np.eye(len(mu))
    Sigma = X_A_Xt(x_pred.Sigma, I - K @ H)
    return Gaussian(mu, Sigma)


def smooth(x_curr: Gaussian, x_next_s: Gaussian, A, Q, stats=None):
    """src/filtering.jl:119-154 (dense and SR versions)."""
    x_pred = predict(x_curr, A, Q, stats)
    P_p_inv = inv(_dense(x_pred.Sigma))
    G = _dense(x_curr.Sigma) @ A.T @ P_p_inv
    mean = x_curr.mu + G @ (x_next_s.mu - x_pred.mu)
    I = np.eye(len(mean))
    if isinstance(x_curr.Sigma, SRMatrix):
        R_ = np.concatenate(
            [
                x_curr.Sigma.squareroot.T @ (I - G @ A).T,
                Q.squareroot.T @ G.T,
                x_next_s.Sigma.squareroot.T @ G.T,
            ],
            axis=0,
        )
        P_s_R = qr_r(R_)
        cov = SRMatrix(P_s_R.T.copy())
    else:
        cov = X_A_Xt(x_curr.Sigma, I - G @ A) + X_A_Xt(_dense(Q), G) + X_A_Xt(_dense(x_next_s.Sigma), G)
    return Gaussian(mean, cov), G


# ----------------------------------------------------------------------------
# Vector-field catalogue (generic scalar arithmetic: float / Dual / Jet / mpf)
# ----------------------------------------------------------------------------


@dataclass
class VectorField:
    name: str
    d: int
    n_params: int
    f: Callable
    jac: Callable
    vf_kind: int = -1  # enum shared with include/pnde.h


def _fhn_readme_f(u, p, t):
    # README.md:36-40
    a, b, c = p
    return [c * (u[0] - u[0] * u[0] * u[0] / 3 + u[1]), -(1 / c) * (u[0] - a - b * u[1])]


def _fhn_readme_j(u, p, t):
    a, b, c = p
    return [[c * (1 - u[0] * u[0]), c + 0 * u[0]], [-(1 / c) + 0 * u[0], b / c + 0 * u[0]]]


def _fhn_lib_f(u, p, t):
    # DiffEqProblemLibrary prob_ode_fitzhughnagumo (App. B.4)
    a, b, tauinv, l = p
    v, w = u
    return [v - v * v * v / 3 - w + l, tauinv * (v + a - b * w)]


def _fhn_lib_j(u, p, t):
    a, b, tauinv, l = p
    v = u[0]
    return [[1 - v * v, -1 + 0 * v], [tauinv + 0 * v, -tauinv * b + 0 * v]]


def _lv_f(u, p, t):
    a, b, c, dd = p
    x, y = u
    return [a * x - b * x * y, -c * y + dd * x * y]


def _lv_j(u, p, t):
    a, b, c, dd = p
    x, y = u
    return [[a - b * y, -b * x], [dd * y, -c + dd * x]]


def _vdp_f(u, p, t):
    # library ordering u=(y,x): dy = mu((1-x^2) y - x), dx = y  (App. B.4)
    (mu,) = p
    y, x = u
    return [mu * ((1 - x * x) * y - x), y + 0 * x]


def _vdp_j(u, p, t):
    (mu,) = p
    y, x = u
    return [[mu * (1 - x * x), mu * (-2 * x * y - 1)], [1 + 0 * x, 0 * x]]


def _linear_f(u, p, t):
    # du_i = p_i u_i (test/state_init.jl:15 with p=(a,b); test/convergence.jl:10)
    return [p[i] * u[i] for i in range(len(u))]


def _linear_j(u, p, t):
    n = len(u)
    return [[(p[i] + 0 * u[i]) if i == j else 0 * u[i] for j in range(n)] for i in range(n)]


def _logistic_f(u, p, t):
    # test/specific_problems.jl:62
    return [p[0] * u[0] * (1 - u[0])]


def _logistic_j(u, p, t):
    return [[p[0] * (1 - 2 * u[0])]]


def _lorenz96_f(u, p, t):
    F = p[0]
    n = len(u)
    return [(u[(i + 1) % n] - u[(i - 2) % n]) * u[(i - 1) % n] - u[i] + F for i in range(n)]


def _lorenz96_j(u, p, t):
    n = len(u)
    J = [[0 * u[0] for _ in range(n)] for _ in range(n)]
    for i in range(n):
        J[i][(i + 1) % n] = J[i][(i + 1) % n] + u[(i - 1) % n]
        J[i][(i - 2) % n] = J[i][(i - 2) % n] - u[(i - 1) % n]
        J[i][(i - 1) % n] = J[i][(i - 1) % n] + (u[(i + 1) % n] - u[(i - 2) % n])
        J[i][i] = J[i][i] - 1
    return J


VF_FHN_README, VF_FHN_LIB, VF_LOTKA_VOLTERRA, VF_VANDERPOL, VF_LINEAR2, VF_LOGISTIC, VF_LORENZ96, VF_LINEAR1 = range(8)

CATALOGUE = {
    "fhn_readme": VectorField("fhn_readme", 2, 3, _fhn_readme_f, _fhn_readme_j, VF_FHN_README),
    "fhn_lib": VectorField("fhn_lib", 2, 4, _fhn_lib_f, _fhn_lib_j, VF_FHN_LIB),
    "lotka_volterra": VectorField("lotka_volterra", 2, 4, _lv_f, _lv_j, VF_LOTKA_VOLTERRA),
    "vanderpol": VectorField("vanderpol", 2, 1, _vdp_f, _vdp_j, VF_VANDERPOL),
    "linear2": VectorField("linear2", 2, 2, _linear_f, _linear_j, VF_LINEAR2),
    "logistic": VectorField("logistic", 1, 1, _logistic_f, _logistic_j, VF_LOGISTIC),
    "linear1": VectorField("linear1", 1, 1, _linear_f, _linear_j, VF_LINEAR1),
}


def lorenz96(d: int) -> VectorField:
    return VectorField(f"lorenz96_{d}", d, 1, _lorenz96_f, _lorenz96_j, VF_LORENZ96)


# ----------------------------------------------------------------------------
# Taylor-mode initial derivatives (src/state_initialization.jl:15-42)
# ----------------------------------------------------------------------------


class Jet:
    """Truncated power series in time, c[0] + c[1] tau + ... (order len(c)-1)."""

    __slots__ = ("c",)

    def __init__(self, c):
        self.c = list(c)

    def _o(self, o):
        if isinstance(o, Jet):
            return o
        return Jet([o] + [0 * o] * (len(self.c) - 1))

    def __add__(self, o):
        o = self._o(o)
        return Jet([a + b for a, b in zip(self.c, o.c)])

    __radd__ = __add__

    def __sub__(self, o):
        o = self._o(o)
        return Jet([a - b for a, b in zip(self.c, o.c)])

    def __rsub__(self, o):
        o = self._o(o)
        return Jet([b - a for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return Jet([-a for a in self.c])

    def __mul__(self, o):
        if not isinstance(o, Jet):
            return Jet([a * o for a in self.c])
        n = len(self.c)
        out = []
        for k in range(n):
            s = self.c[0] * o.c[k]
            for i in range(1, k + 1):
                s = s + self.c[i] * o.c[k - i]
            out.append(s)
        return Jet(out)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Jet):
            raise NotImplementedError("jet / jet")
        return Jet([a / o for a in self.c])


def _jet_sincos(a: "Jet"):
    n = len(a.c)
    s, c = [math.sin(_value(a.c[0])) + 0 * a.c[0]], [math.cos(_value(a.c[0])) + 0 * a.c[0]]
    for k in range(1, n):
        ss = sum(i * a.c[i] * c[k - i] for i in range(1, k + 1))
        cc = sum(i * a.c[i] * s[k - i] for i in range(1, k + 1))
        s.append(ss / k)
        c.append(-cc / k)
    return Jet(s), Jet(c)


def gsin(x):
    """sin for floats and jets (vector fields of the run-time compiled path use elementary functions)."""
    return _jet_sincos(x)[0] if isinstance(x, Jet) else math.sin(x)


def gcos(x):
    return _jet_sincos(x)[1] if isinstance(x, Jet) else math.cos(x)


def gexp(x):
    if not isinstance(x, Jet):
        return math.exp(x)
    r = [math.exp(x.c[0])]
    for k in range(1, len(x.c)):
        r.append(sum(i * x.c[i] * r[k - i] for i in range(1, k + 1)) / k)
    return Jet(r)


def get_derivatives(u0, vf: VectorField, p, t0, q: int):
    """[u'(t0), ..., u^(q)(t0)] by Taylor-mode AD (src/state_initialization.jl:15-42).

    The reference builds TaylorN polynomials in u and iterates
    df <- (d df/du) * f; this is the same quantity computed as the time-Taylor
    coefficients of the solution: c_{k+1} = [f(c(tau))]_k / (k+1), u^(k) = k! c_k.
    """
    d = len(u0)
    coeffs = [list(u0)]  # c_0
    for k in range(q):
        n = k + 1  # jets truncated after tau^k
        zero = 0 * u0[0]
        jets = [Jet([coeffs[j][i] if j < len(coeffs) else zero for j in range(n)]) for i in range(d)]
        fu = vf.f(jets, p, t0)
        ck = []
        for i in range(d):
            fi = fu[i]
            fk = fi.c[k] if isinstance(fi, Jet) else (fi if k == 0 else zero)
            ck.append(fk / (k + 1))
        coeffs.append(ck)
    out = []
    fact = 1.0
    for k in range(1, q + 1):
        fact *= k
        out.append([fact * c for c in coeffs[k]])
    return out


# ----------------------------------------------------------------------------
# Algorithm / problem descriptions (src/algorithms.jl:23-51)
# ----------------------------------------------------------------------------

DIFFUSIONS = ("dynamic", "fixed", "fixedMAP", "dynamicMV", "fixedMV")


@dataclass
class Alg:
    kind: str = "EK1"  # "EK0" | "EK1"
    order: int = 3
    diffusionmodel: str = "dynamic"
    smooth: bool = True
    linearize_at: object = None  # IEKS only: the previous iteration's Solution (src/ieks.jl:2-8)


def EK0(order=3, diffusionmodel="dynamic", smooth=True):
    return Alg("EK0", order, diffusionmodel, smooth)


def EK1(order=3, diffusionmodel="dynamic", smooth=True):
    return Alg("EK1", order, diffusionmodel, smooth)


def IEKS(order=1, diffusionmodel="dynamic", linearize_at=None):
    """src/ieks.jl:32-41: an EK1 whose Jacobian is evaluated at the previous iterate's dense output; smooth is forced on."""
    if linearize_at is not None:
        assert linearize_at.q == order and linearize_at.smoothed
    return Alg("EK1", order, diffusionmodel, True, linearize_at)


def solve_ieks(prob, alg: "Alg", iterations=10, **kwargs):
    """src/ieks.jl:53-61: fixed number of re-solves, each linearised at the previous solution; no stopping rule."""
    sol = None
    for _ in range(iterations):
        alg.linearize_at = sol
        sol = solve_ivp(prob, alg, **kwargs)
    return sol


@dataclass
class Problem:
    vf: VectorField
    u0: Sequence
    tspan: Sequence[float]
    p: Sequence


@dataclass
class Solution:
    t: List[float] = field(default_factory=list)
    u: List[np.ndarray] = field(default_factory=list)
    pu: List[Gaussian] = field(default_factory=list)
    x_filt: List[Gaussian] = field(default_factory=list)
    x_smooth: Optional[List[Gaussian]] = None
    diffusions: list = field(default_factory=list)
    log_likelihood: float = 0.0
    naccept: int = 0
    nreject: int = 0
    nf: int = 0
    njacs: int = 0
    retcode: str = "Default"
    stats: dict = field(default_factory=dict)
    # context for dense output / sampling
    d: int = 0
    q: int = 0
    A: np.ndarray = None
    Q: SRMatrix = None
    smoothed: bool = False


# ----------------------------------------------------------------------------
# The filter step (src/perform_step.jl) and its helpers
# ----------------------------------------------------------------------------


class _Cache:
    """GaussianODEFilterCache (src/caches.jl:5-114), only what the math needs."""

    def __init__(self, prob: Problem, alg: Alg, dtype):
        self.d = d = len(prob.u0)
        self.q = q = alg.order
        self.D = D = d * (q + 1)
        self.A, self.Q = ibm(d, q)
        if dtype is object:
            self.A = self.A.astype(object)
            self.Q = SRMatrix(self.Q.squareroot.astype(object))
        self.E0 = proj(d, q, 0)
        self.E1 = proj(d, q, 1)
        self.SolProj = self.E0
        self.dtype = dtype
        zero = np.zeros(D) if dtype is float else np.array([0.0] * D, dtype=object)
        eye = np.eye(D) if dtype is float else np.eye(D).astype(object)
        self.x = Gaussian(zero, SRMatrix(eye))
        self.x_pred = None
        self.x_filt = None
        self.H = None
        self.measurement = None
        self.u_pred = None
        self.u_filt = None
        self.local_diffusion = 1.0
        mv = alg.diffusionmodel in ("dynamicMV", "fixedMV")
        # initial_diffusion (src/diffusions.jl:8,84,116)
        self.global_diffusion = np.ones(D) if mv else 1.0
        self.log_likelihood = 0.0


def condition_on(x: Gaussian, H, data):
    """src/state_initialization.jl:45-53."""
    z = H @ x.mu
    S = X_A_Xt(x.Sigma, H)
    K = _dense(x.Sigma) @ H.T @ inv(S.mat)
    mu = x.mu + K @ (data - z)
    I = np.eye(len(mu))
    return Gaussian(mu, X_A_Xt(x.Sigma, I - K @ H))


def initial_update(cache: _Cache, prob: Problem, alg: Alg, t0):
    """src/state_initialization.jl:2-14: condition N(0,I) on u0 and on the q
    Taylor-mode derivatives."""
    d, q = cache.d, cache.q
    asarr = (lambda v: np.array(v, dtype=float)) if cache.dtype is float else (lambda v: np.array(list(v), dtype=object))
    x = condition_on(cache.x, proj(d, q, 0), asarr(prob.u0))
    derivs = get_derivatives(list(prob.u0), prob.vf, prob.p, t0, q)
    for o, df in zip(range(1, q + 1), derivs):
        x = condition_on(x, proj(d, q, o), asarr(df))
    cache.x = x


def measure(cache: _Cache, prob: Problem, alg: Alg, x_pred: Gaussian, PI, t, sol: Solution):
    """src/perform_step.jl:95-132."""
    d = cache.d
    du = prob.vf.f(list(cache.u_pred), prob.p, t)
    du = np.array(du, dtype=cache.dtype if cache.dtype is object else float)
    sol.nf += 1
    z = cache.E1 @ (PI * x_pred.mu) - du
    if alg.kind == "EK1":
        # IEKS: J at the previous solution's sol(t).mu, f still at u_pred (src/perform_step.jl:111-113, 116)
        lin = cache.u_pred if alg.linearize_at is None else dense_eval(alg.linearize_at, t).mu
        ddu = np.array(prob.vf.jac(list(lin), prob.p, t), dtype=cache.dtype if cache.dtype is object else float)
        sol.njacs += 1
        H = (cache.E1 - ddu @ cache.E0) * PI[None, :]
    else:
        H = cache.E1 * PI[None, :]
    cache.H = H
    S = X_A_Xt(x_pred.Sigma, H).mat
    cache.measurement = Gaussian(z, S)
    return cache.measurement


def estimate_diffusion(cache: _Cache, alg: Alg, sol: Solution, success_iter: int, PI):
    """src/diffusions.jl:11-153.  Returns (local, global)."""
    d, q = cache.d, cache.q
    model = alg.diffusionmodel
    meas = cache.measurement
    if model == "dynamic":
        z = meas.mu
        HQH = X_A_Xt(cache.Q, cache.H).mat
        s2 = z @ solve(HQH, z) / d  # :77-79
        return s2, s2
    if model == "fixed":
        v, S = meas.mu, meas.Sigma
        if all(_value(x) == 0 for x in v):
            # reference bug (src/diffusions.jl:18-20): returns a scalar, the caller's
            # destructuring then fails; we surface it instead of guessing.
            raise RuntimeError("FixedDiffusion with v == 0: reference throws (src/diffusions.jl:18-20)")
        diffusion_t = v @ inv(S) @ v / d
        if success_iter == 0:
            return diffusion_t, diffusion_t
        prev = sol.diffusions[-1]
        return diffusion_t, prev + (diffusion_t - prev) / success_iter  # :33
    if model == "fixedMAP":
        N = success_iter + 1
        v, S = meas.mu, meas.Sigma
        res_t = v @ inv(S) @ v / d
        alpha, beta = 0.5, 0.5
        if success_iter == 0:
            return res_t, (beta + 0.5 * res_t) / (alpha + N * d / 2 + 1)
        prev = sol.diffusions[-1]
        res_prev = (prev * (alpha + (N - 1) * d / 2 + 1) - beta) * 2
        res_sum_t = res_prev + res_t
        return res_t, (beta + 0.5 * res_sum_t) / (alpha + N * d / 2 + 1)
    if model == "dynamicMV":
        assert alg.kind == "EK0", "MV diffusions are EK0-only (src/diffusions.jl:97)"
        z = meas.mu
        HQH = cache.H @ cache.Q.mat @ cache.H.T
        Q0_11 = HQH[0, 0]
        Sii = z ** 2 / Q0_11
        Sii = np.maximum(Sii, np.finfo(float).eps)
        out = np.tile(Sii, q + 1)  # kron(I_{q+1}, Sigma) diagonal
        return out, out
    if model == "fixedMV":
        assert alg.kind == "EK0"
        v, S = meas.mu, meas.Sigma
        S_11 = S[0, 0]
        Sii = v ** 2 / S_11
        out = np.tile(Sii, q + 1)
        if success_iter == 0:
            return out, out
        prev = sol.diffusions[-1]
        return out, prev + (out - prev) / success_iter
    raise ValueError(model)


def estimate_errors(cache: _Cache):
    """src/perform_step.jl:148-158."""
    ld = cache.local_diffusion
    if np.ndim(ld) == 0 and not isinstance(ld, Dual) and math.isinf(float(ld)):
        return math.inf
    M = X_A_Xt(apply_diffusion(cache.Q, ld), cache.H).mat
    return np.array([_ssqrt(M[i, i]) for i in range(M.shape[0])], dtype=M.dtype)


def gaussian_logpdf_at_zero(mu, S):
    """GaussianDistributions.logpdf(N(mu, S), 0) (src/perform_step.jl:66)."""
    d = len(mu)
    L, ok = chol_lower(S)
    if not ok:
        return float("nan")
    y = solve(L, mu) if _is_float(L) else inv(L) @ mu
    logdet = 0.0
    for i in range(d):
        logdet = logdet + 2 * _slog(L[i, i])
    return -0.5 * (y @ y + logdet + d * math.log(2 * math.pi))


def perform_step(cache: _Cache, prob: Problem, alg: Alg, sol: Solution, t, dt, adaptive, abstol, reltol,
                 u_prev, success_iter):
    """One attempted step (src/perform_step.jl:27-93).  Returns (EEst, u_filt)."""
    d, q = cache.d, cache.q
    A, Q = cache.A, cache.Q
    tnew = t + dt
    P = preconditioner_diag(d, q, dt)
    if cache.dtype is object:
        P = P.astype(object)
    PI = 1 / P
    stats = sol.stats

    def pmul(diag, g: Gaussian) -> Gaussian:  # Diagonal * SRGaussian  (src/ProbNumDiffEq.jl:58)
        return Gaussian(diag * g.mu, SRMatrix(diag[:, None] * g.Sigma.squareroot))

    x = pmul(P, cache.x)  # :38
    dynamic = alg.diffusionmodel in ("dynamic", "dynamicMV")
    if dynamic:
        x_pred = Gaussian(predict_mean(x, A), x.Sigma)  # :43 (covariance not predicted yet)
        cache.u_pred = cache.SolProj @ (PI * x_pred.mu)  # :44
        # measure! computes S from the stale covariance (:129); it is overwritten at :54
        measure(cache, prob, alg, x_pred, PI, tnew, sol)  # :47
        cache.local_diffusion, cache.global_diffusion = estimate_diffusion(cache, alg, sol, success_iter, PI)  # :50
        x_pred = Gaussian(x_pred.mu, predict_cov(x, A, apply_diffusion(Q, cache.global_diffusion), stats))  # :53
        cache.measurement = Gaussian(cache.measurement.mu, X_A_Xt(x_pred.Sigma, cache.H).mat)  # :54
    else:
        x_pred = predict(x, A, Q, stats)  # :58
        cache.u_pred = cache.SolProj @ (PI * x_pred.mu)
        measure(cache, prob, alg, x_pred, PI, tnew, sol)
        cache.local_diffusion, cache.global_diffusion = estimate_diffusion(cache, alg, sol, success_iter, PI)
    # :66
    if cache.dtype is float:
        cache.log_likelihood = gaussian_logpdf_at_zero(cache.measurement.mu, cache.measurement.Sigma)
    else:
        cache.log_likelihood = 0.0
    x_filt = update(x_pred, cache.measurement, cache.H)  # :69
    u_filt = cache.SolProj @ (PI * x_filt.mu)  # :70
    # undo preconditioning (:73-75)
    cache.x = pmul(PI, x)
    cache.x_pred = pmul(PI, x_pred)
    cache.x_filt = pmul(PI, x_filt)
    EEst = None
    if adaptive:
        err = estimate_errors(cache)  # :79
        if np.ndim(err) == 0 and math.isinf(err):
            EEst = math.inf
        else:
            ut = dt * err
            res = np.empty(d, dtype=object)
            for i in range(d):  # calculate_residuals! (App. B.2)
                res[i] = ut[i] / (abstol + max(internalnorm(u_prev[i]), internalnorm(u_filt[i])) * reltol)
            EEst = internalnorm(res)  # :83
    # :86 (always) and :89-92
    accept_inside = (not adaptive) or (EEst < 1.0)
    if accept_inside:
        cache.x = cache.x_filt.copy()
        sol.log_likelihood = sol.log_likelihood + cache.log_likelihood
    return EEst, u_filt


# ----------------------------------------------------------------------------
# The external loop (OrdinaryDiffEq 5.x)
# ----------------------------------------------------------------------------


@dataclass
class Controller:
    qmin: float = 1 / 5
    qmax: float = 10.0
    gamma: float = 9 / 10
    qsteady_min: float = 1.0
    qsteady_max: float = 1.0
    qoldinit: float = 1e-4
    dtmin: float = 0.0
    maxiters: int = 100000


def initdt(prob: Problem, alg: Alg, t0, dtmax, abstol, reltol, sol: Solution):
    """ode_determine_initdt (Hairer; App. B.3).  Costs two f evaluations."""
    u0 = list(prob.u0)
    d = len(u0)
    f0 = prob.vf.f(u0, prob.p, t0)
    sk = [abstol + internalnorm(u0[i]) * reltol for i in range(d)]
    d0 = internalnorm(np.array([u0[i] / sk[i] for i in range(d)], dtype=object))
    d1 = internalnorm(np.array([f0[i] / sk[i] for i in range(d)], dtype=object))
    if d0 < 1e-5 or d1 < 1e-5:
        dt0 = 1e-6
    else:
        dt0 = (d0 / d1) / 100
    dt0 = min(dt0, dtmax)
    if dt0 < 10 * np.finfo(float).eps:
        return 1e-6
    u1 = [u0[i] + dt0 * f0[i] for i in range(d)]
    f1 = prob.vf.f(u1, prob.p, t0 + dt0)
    sol.nf += 2
    d2 = internalnorm(np.array([(f1[i] - f0[i]) / sk[i] for i in range(d)], dtype=object)) / dt0
    mx = max(d1, d2)
    if mx <= 1e-15:
        dt1 = max(1e-6, dt0 * 1e-3)
    else:
        dt1 = 10.0 ** (-(2 + math.log10(mx)) / (alg.order + 1))
    return min(100 * dt0, dt1, dtmax)


def _eps(x: float) -> float:
    return float(np.spacing(abs(x))) if x != 0 else float(np.finfo(float).tiny)


def solve_ivp(prob: Problem, alg: Alg, *, adaptive=True, dt=None, abstol=1e-6, reltol=1e-3,
              ctrl: Optional[Controller] = None, dtype=float, dtmax=None, record_attempts=None) -> Solution:
    """solve(prob, alg; abstol, reltol, adaptive, dt): __init + solve! + postamble!
    (src/perform_step.jl:2-12; src/integrator_utils.jl:2-48)."""
    if not adaptive and dt is None:
        raise ValueError("Fixed timestep methods require a choice of dt")  # test/errors.jl:16-20
    ctrl = ctrl or Controller()
    q = alg.order
    beta2 = 2 / (5 * (q + 1))  # src/alg_utils.jl:23
    beta1 = 7 / (10 * (q + 1))  # src/alg_utils.jl:24
    t0, t1 = float(prob.tspan[0]), float(prob.tspan[1])
    dtmax = (t1 - t0) if dtmax is None else dtmax
    if dtype is object:
        nparts = max([len(x.p) for x in list(prob.p) + list(prob.u0) if isinstance(x, Dual)] + [0])
        if nparts:  # DiffEqBase promote_u0: u0 becomes Dual when p is Dual
            prob = Problem(prob.vf, [Dual.lift(x, nparts) for x in prob.u0], prob.tspan, prob.p)
    cache = _Cache(prob, alg, dtype)
    sol = Solution(d=cache.d, q=q, A=cache.A, Q=cache.Q)
    initial_update(cache, prob, alg, t0)  # src/perform_step.jl:7
    sol.x_filt.append(cache.x.copy())  # :10
    sol.pu.append(affine(cache.SolProj, cache.x))  # :11
    sol.t.append(t0)
    asarr = (lambda v: np.array(v, dtype=float)) if dtype is float else (lambda v: np.array(list(v), dtype=object))
    u = asarr(prob.u0)
    sol.u.append(u.copy())

    if adaptive:
        cur_dt = initdt(prob, alg, t0, dtmax, abstol, reltol, sol) if dt is None else float(dt)
    else:
        cur_dt = float(dt)
    dt_user = cur_dt
    t = t0
    qold = ctrl.qoldinit
    q11 = 1.0
    success_iter = 0
    it = 0
    accepted_prev = None
    dtpropose = cur_dt
    while t < t1:
        # loopheader!
        if it > 0:
            if accepted_prev:
                success_iter += 1
                cur_dt = dtpropose
            else:
                cur_dt = cur_dt / min(1 / ctrl.qmin, q11 / ctrl.gamma)
        it += 1
        if it > ctrl.maxiters:
            sol.retcode = "MaxIters"
            break
        if adaptive:
            cur_dt = min(cur_dt, dtmax)
            cur_dt = max(cur_dt, ctrl.dtmin)
            cur_dt = min(cur_dt, t1 - t)
        else:
            cur_dt = min(dt_user, t1 - t)
        if not (cur_dt == cur_dt):
            sol.retcode = "DtNaN"
            break
        # check_error! (OrdinaryDiffEq): unstable_check(dt, u, p, t) = any(isnan, u) -> :Unstable
        try:
            if np.isnan(np.asarray(u, dtype=float)).any():
                sol.retcode = "Unstable"
                break
        except (TypeError, ValueError):
            pass  # generic scalar types (mpmath, duals): no such check
        EEst, u_filt = perform_step(cache, prob, alg, sol, t, cur_dt, adaptive, abstol, reltol, u, success_iter)
        if EEst is not None and not isinstance(EEst, float):
            EEst = float(EEst)  # mpmath arbiter runs: the controller itself stays in Float64 like dt and t
        u = u_filt  # src/perform_step.jl:86
        if record_attempts is not None:
            record_attempts.append((t, cur_dt, EEst))
        # loopfooter!
        ttmp = t + cur_dt
        if adaptive:
            if EEst == 0:
                qc = 1 / ctrl.qmax
            else:
                q11 = EEst ** beta1
                qc = q11 / (qold ** beta2)
                qc = max(1 / ctrl.qmax, min(1 / ctrl.qmin, qc / ctrl.gamma))
            accept = EEst <= 1.0
            if accept:
                sol.naccept += 1
                if ctrl.qsteady_min <= qc <= ctrl.qsteady_max:
                    qc = 1.0
                qold = max(EEst, ctrl.qoldinit)
                dtnew = cur_dt / qc
                t = t1 if abs(ttmp - t1) < 10 * _eps(max(t, t1)) else ttmp
                dtpropose = max(ctrl.dtmin, min(dtmax, dtnew))
            else:
                sol.nreject += 1
        else:
            accept = True
            sol.naccept += 1
            t = t1 if abs(ttmp - t1) < 10 * _eps(max(t, t1)) else ttmp
            dtpropose = cur_dt
        accepted_prev = accept
        if accept:
            # savevalues! (src/integrator_utils.jl:33-48)
            sol.t.append(t)
            sol.u.append(u.copy())
            sol.x_filt.append(cache.x.copy())
            gd = cache.global_diffusion
            sol.diffusions.append(gd.copy() if isinstance(gd, np.ndarray) else gd)
            sol.pu.append(affine(cache.SolProj, cache.x))
    else:
        sol.retcode = "Success"
    postamble(cache, alg, sol)
    return sol


def postamble(cache: _Cache, alg: Alg, sol: Solution):
    """src/integrator_utils.jl:2-30."""
    static = alg.diffusionmodel in ("fixed", "fixedMAP", "fixedMV")
    if static and len(sol.diffusions) > 0:
        sol.log_likelihood = float("nan")
        final = sol.diffusions[-1]
        sol.x_filt = [Gaussian(s.mu, apply_diffusion(s.Sigma, final)) for s in sol.x_filt]
        sol.diffusions = [final.copy() if isinstance(final, np.ndarray) else final for _ in sol.diffusions]
    if alg.smooth:
        smooth_all(cache, sol)
        sol.pu = [affine(cache.SolProj, x) for x in sol.x_smooth]
        sol.u = [g.mu.copy() for g in sol.pu]
        sol.smoothed = True


def smooth_all(cache: _Cache, sol: Solution):
    """src/smoothing.jl:4-28 (+ smooth! :31-63, which is filtering.jl's SR smooth)."""
    d, q = cache.d, cache.q
    A, Q = cache.A, cache.Q
    x = [g.copy() for g in sol.x_filt]
    t = sol.t
    n = len(x)
    for i in range(n - 2, 0, -1):  # Julia i = N:-1:2 (1-based) -> 0-based n-2 .. 1
        dt = t[i + 1] - t[i]
        if dt == 0:
            x[i] = x[i + 1].copy()
            continue
        P = preconditioner_diag(d, q, dt)
        PI = 1 / P
        Qh = apply_diffusion(Q, sol.diffusions[i])
        xi = Gaussian(P * x[i].mu, SRMatrix(P[:, None] * x[i].Sigma.squareroot))
        xn = Gaussian(P * x[i + 1].mu, SRMatrix(P[:, None] * x[i + 1].Sigma.squareroot))
        xs, _ = smooth(xi, xn, A, Qh, sol.stats)
        x[i] = Gaussian(PI * xs.mu, SRMatrix(PI[:, None] * xs.Sigma.squareroot))
    sol.x_smooth = x


# ----------------------------------------------------------------------------
# Dense output and sampling (src/solution.jl:165-215, src/solution_sampling.jl)
# ----------------------------------------------------------------------------


def posterior_at(sol: Solution, tval: float, smoothed: Optional[bool] = None) -> Gaussian:
    """GaussianODEFilterPosterior call (src/solution.jl:165-210): full state."""
    smoothed = sol.smoothed if smoothed is None else smoothed
    t = np.asarray(sol.t)
    d, q, A, Q = sol.d, sol.q, sol.A, sol.Q
    if tval < t[0]:
        raise ValueError("Invalid t<t0")
    idx = int(np.sum(t <= tval))  # 1-based count
    if np.any(t == tval):
        return (sol.x_smooth if smoothed else sol.x_filt)[idx - 1]
    prev_t = t[idx - 1]
    prev_rv = sol.x_filt[idx - 1]
    diffusion = sol.diffusions[min(idx, len(sol.diffusions)) - 1]
    h1 = tval - prev_t
    P = preconditioner_diag(d, q, h1)
    PI = 1 / P
    Qh = apply_diffusion(Q, diffusion)
    g = predict(Gaussian(P * prev_rv.mu, SRMatrix(P[:, None] * prev_rv.Sigma.squareroot)), A, Qh)
    goal_pred = Gaussian(PI * g.mu, SRMatrix(PI[:, None] * g.Sigma.squareroot))
    if (not smoothed) or tval >= t[-1]:
        return goal_pred
    next_t = t[idx]
    next_s = sol.x_smooth[idx]
    h2 = next_t - tval
    P = preconditioner_diag(d, q, h2)
    PI = 1 / P
    gp = Gaussian(P * goal_pred.mu, SRMatrix(P[:, None] * goal_pred.Sigma.squareroot))
    ns = Gaussian(P * next_s.mu, SRMatrix(P[:, None] * next_s.Sigma.squareroot))
    gs, _ = smooth(gp, ns, A, Qh)
    return Gaussian(PI * gs.mu, SRMatrix(PI[:, None] * gs.Sigma.squareroot))


def dense_eval(sol: Solution, tval: float) -> Gaussian:
    """sol(t) = SolProj * posterior(t) (src/solution.jl:211-214)."""
    E0 = proj(sol.d, sol.q, 0)
    return affine(E0, posterior_at(sol, tval))


def dense_sample_law(sol: Solution, times) -> List[Gaussian]:
    """Marginal law N(mean, cov) of dense_sample_states' draws at every time of `times`
    (src/solution_sampling.jl:63-74 -> :24-62): the reference draws backwards through the filtering
    posterior extrapolated to the grid (interp(...; smoothed=false), :66), x_k = m_k + G_k (x_{k+1} - m_k^-) + noise,
    with the diffusion of the solver interval that contains t_k (:41-42).  Replacing each draw by its mean and
    covariance is the RTS recursion over those extrapolated states; it ignores the measurements strictly inside a
    grid interval, so it equals sol(t) only when the grid contains the solver's."""
    assert sol.smoothed
    d, q = sol.d, sol.q
    xs = [posterior_at(sol, float(t), smoothed=False) for t in times]
    out = [None] * len(xs)
    out[-1] = cur = xs[-1]
    for i in range(len(xs) - 2, -1, -1):
        dt = times[i + 1] - times[i]
        diffusion = sol.diffusions[int(np.sum(np.asarray(sol.t) <= times[i])) - 1]
        P = preconditioner_diag(d, q, dt)
        sm, _ = smooth(Gaussian(P * xs[i].mu, SRMatrix(P[:, None] * xs[i].Sigma.squareroot)),
                       Gaussian(P * cur.mu, SRMatrix(P[:, None] * cur.Sigma.squareroot)), sol.A,
                       apply_diffusion(sol.Q, diffusion))
        out[i] = cur = Gaussian(sm.mu / P, SRMatrix(sm.Sigma.squareroot / P[:, None]))
    return out


def sample_states(sol: Solution, n: int, normals: np.ndarray) -> np.ndarray:
    """src/solution_sampling.jl:24-62 with the standard-normal draws supplied by
    the caller: normals[i, :, j] is the D-vector used at time index i for path j
    (the reference calls randn; the draws are an input here so that a device
    implementation with its own generator can be checked bit-for-bit)."""
    assert sol.smoothed
    d, q, A, Q = sol.d, sol.q, sol.A, sol.Q
    D = d * (q + 1)
    xs, ts = sol.x_filt, sol.t
    N = len(xs)
    path = np.zeros((N, D, n))
    x = xs[-1]
    path[-1] = x.mu[:, None] + x.Sigma.squareroot @ normals[-1]
    for i in range(N - 2, -1, -1):
        dt = ts[i + 1] - ts[i]
        i_diff = int(np.sum(np.asarray(ts) <= ts[i]))
        diffusion = sol.diffusions[i_diff - 1]
        Qh = apply_diffusion(Q, diffusion)
        P = preconditioner_diag(d, q, dt)
        PI = 1 / P
        for j in range(n):
            sample_p = P * path[i + 1, :, j]
            x_prev_p = Gaussian(P * xs[i].mu, SRMatrix(P[:, None] * xs[i].Sigma.squareroot))
            prev, _ = smooth(x_prev_p, Gaussian(sample_p, SRMatrix(np.zeros((D, D)))), A, Qh)
            draw = prev.mu + prev.Sigma.squareroot @ normals[i, :, j]
            path[i, :, j] = PI * draw
    return path
