"""numpy model of the arithmetic the CUDA filter kernel performs (TEST INFRASTRUCTURE).

This is NOT the reference algorithm (that is ``pnde_oracle.py``); it is a
line-for-line model of ``csrc/ek_dense.cuh``: reduced-rank square-root factor,
one structured Householder QR per step in measurement-aligned coordinates, the
update read straight off the triangular factor.  It exists so that the design
can be checked against the oracle on a CPU and so that a failing GPU parity
test can be bisected.  See DESIGN.md section 3 for the derivation.
"""
from __future__ import annotations

import math

import numpy as np

from pnde_oracle import ibm


def iwp_small(q):
    """Atilde (q+1 x q+1), Ltilde = chol(Qtilde), Qtilde  (src/priors.jl:7-59 with d=1)."""
    A, Q = ibm(1, q)
    return A, Q.squareroot, Q.mat


class DenseFilterModel:
    """State: m (D), S = [W | Lz] (D x (D-d)), both in P(h)-preconditioned coordinates."""

    def __init__(self, d, q, f, jac, ek1=True):
        self.d, self.q = d, q
        self.D = d * (q + 1)
        self.f, self.jac, self.ek1 = f, jac, ek1
        self.At, self.Lt, self.Qt = iwp_small(q)

    def A_apply(self, x):
        d, q = self.d, self.q
        out = np.zeros_like(x)
        for k in range(q + 1):
            for j in range(k, q + 1):
                out[k * d:(k + 1) * d] += self.At[k, j] * x[j * d:(j + 1) * d]
        return out

    def step(self, m, S, h, p, t, diffusion="dynamic"):
        """One attempted step from the preconditioned state (m, S).  Returns dict."""
        d, q, D = self.d, self.q, self.D
        r = D - d
        Lt, Qt = self.Lt, self.Qt
        pi1 = h ** (q - 0.5)
        pi0 = pi1 * h
        mm = self.A_apply(m)
        uhat = pi0 * mm[0:d]
        fu = np.array(self.f(list(uhat), p, t + h), dtype=float)
        J = np.array(self.jac(list(uhat), p, t + h), dtype=float) if self.ek1 else np.zeros((d, d))
        z = pi1 * mm[d:2 * d] - fu
        Jp = pi0 * J
        B = Qt[0, 0] * Jp @ Jp.T - pi1 * Qt[0, 1] * (Jp + Jp.T) + pi1 * pi1 * Qt[1, 1] * np.eye(d)
        if diffusion == "dynamic":
            Lb = np.linalg.cholesky(B)
            yb = np.linalg.solve(Lb, z)
            sigma2 = float(yb @ yb) / d
            sig = math.sqrt(sigma2)
        else:
            sigma2 = 1.0
            sig = 1.0
        # stack in primed coordinates [y, x0, x2, ..., xq]
        def col_of(k, a):  # primed column index of unprimed coordinate (k, a), k != 1
            return (d + a) if k == 0 else (k * d + a)

        top = np.zeros((D, D))
        for a in range(d):  # rows (1,a)
            top[a, a] = sig * pi1 * Lt[1, 1]
            for k in range(2, q + 1):
                top[a, col_of(k, a)] = sig * Lt[k, 1]
        for a in range(d):  # rows (0,a)
            rr = d + a
            for b in range(d):
                top[rr, b] = sig * ((pi1 * Lt[1, 0] if a == b else 0.0) - Lt[0, 0] * Jp[b, a])
            top[rr, col_of(0, a)] = sig * Lt[0, 0]
            for k in range(2, q + 1):
                top[rr, col_of(k, a)] = sig * Lt[k, 0]
        for k in range(2, q + 1):
            for a in range(d):
                rr = k * d + a
                for j in range(k, q + 1):
                    top[rr, col_of(j, a)] = sig * Lt[j, k]
        bot = np.zeros((r, D))
        for c in range(r):
            w = self.A_apply(S[:, c])
            bot[c, 0:d] = pi1 * w[d:2 * d] - Jp @ w[0:d]
            bot[c, d:2 * d] = w[0:d]
            bot[c, 2 * d:] = w[2 * d:]
        M = np.vstack([top, bot])
        R = np.linalg.qr(M, mode="r")
        # update
        G2 = R[0:d, 0:d].T
        y = np.linalg.solve(G2, z)
        mp = np.concatenate([np.zeros(d), mm[0:d], mm[2 * d:]])  # primed mean, y-part unused
        mp_new = mp.copy()
        mp_new[d:] = mp[d:] - R[0:d, d:].T @ y
        m_new = np.zeros(D)
        m_new[0:d] = mp_new[d:2 * d]
        m_new[2 * d:] = mp_new[2 * d:]
        m_new[d:2 * d] = (fu + Jp @ (m_new[0:d] - mm[0:d])) / pi1
        Sp = R[d:, d:].T  # (D-d) x (D-d) lower triangular, primed rows [x0, x2..]
        S_new = np.zeros((D, r))
        S_new[0:d, :] = Sp[0:d, :]
        S_new[2 * d:, :] = Sp[d:, :]
        S_new[d:2 * d, :] = (Jp @ S_new[0:d, :]) / pi1
        if diffusion != "dynamic":
            local = float(y @ y) / d
        else:
            local = sigma2
        err = np.sqrt(local * np.diag(B))
        loglik = -0.5 * (float(y @ y) + 2 * np.sum(np.log(np.abs(np.diag(R[0:d, 0:d])))) + d * math.log(2 * math.pi))
        return dict(m=m_new, S=S_new, sigma2=sigma2, local=local, err=err, u=pi0 * m_new[0:d], loglik=loglik,
                    m_pred=mm, z=z)
