"""ORACLE (test infrastructure).  Restatement of cellregmap/_math.py (structured covariance
algebra, score statistic, modified-Liu parameters, qmin).  PINNED by the known answers of
cellregmap/test/test_math.py:38-91 (see tests/test_oracle_goldens.py).
"""
import numpy as np
import scipy.linalg as sla
import scipy.stats as st

from .chiscore_port import liu_sf


def lstsq_solve(a, b):
    """_math.py:33-37  (lstsq with rcond=None)."""
    return np.linalg.lstsq(a, b, rcond=None)[0]


def qscov_dot(Q0, S0, a, b, v):
    """(a K + b I) v with K = Q0 S0 Q0'.  _math.py:53-56."""
    t = Q0.T @ v
    t = (S0 * t.T).T
    return (a * Q0) @ t + b * v          # same association as the reference: a * Q0 @ (...) binds as (a * Q0) @ (...)


def qscov_solve(Q0, S0, a, b, v):
    """(a K + b I)^-1 v.  _math.py:58-73."""
    R0 = 1.0 / (1.0 + (a / b) * S0)
    t = Q0.T @ v
    return (Q0 @ (R0 * t.T).T + v - Q0 @ t) / b


class Projection:
    """P = K^-1 - K^-1 W (W' K^-1 W)^-1 W' K^-1 for K = a Q0 S0 Q0' + b I.  _math.py:79-93."""

    def __init__(self, Q0, S0, a, b, W):
        self.Q0, self.S0, self.a, self.b, self.W = Q0, S0, a, b, W
        self.KiW = qscov_solve(Q0, S0, a, b, W)

    def dot(self, v):
        Kiv = qscov_solve(self.Q0, self.S0, self.a, self.b, v)
        return Kiv - self.KiW @ lstsq_solve(self.W.T @ self.KiW, self.KiW.T @ v)


def score_statistic_structured(P, sqrt_dK, y):
    """Q = 1/2 y' P dK P y with dK = sqrt_dK sqrt_dK'.  _math.py:114-117."""
    Py = P.dot(y)
    return float(Py.T @ sqrt_dK @ sqrt_dK.T @ Py / 2.0)


def weight_matrix_structured(P, sqrt_dK):
    """1/2 sqrt_dK' P sqrt_dK.  _math.py:119-124."""
    return sqrt_dK.T @ P.dot(sqrt_dK) / 2.0


def P_matrix(W, K):
    """Dense P.  _math.py:96-99."""
    KiW = np.linalg.solve(K, W)
    return np.linalg.inv(K) - KiW @ np.linalg.solve(W.T @ KiW, KiW.T)


def score_statistic(y, W, K, dK):
    """Dense Q.  _math.py:131-138."""
    P = P_matrix(W, K)
    return float(y.T @ P @ dK @ P @ y / 2.0)


def score_statistic_distr_weights(W, K, dK):
    """Non-zero eigenvalues of 1/2 sqrt(P) dK sqrt(P).  _math.py:150-160."""
    P = P_matrix(W, K)
    rP = sla.sqrtm(P)
    w = np.linalg.eigvalsh(rP @ dK @ rP) / 2.0
    return w[w > 1e-16]


def score_statistic_liu_params(q, weights):
    """_math.py:163-180."""
    n = len(weights)
    pv, dof_x, _, info = liu_sf(q, weights, [1] * n, [0] * n, True)
    return {"pv": float(pv), "mu_q": info["mu_q"], "sigma_q": info["sigma_q"], "dof_x": dof_x}


def qmin(liu_params):
    """Quantile matching across a rho grid.  _math.py:183-201."""
    T = min(p["pv"] for p in liu_params)
    out = np.zeros(len(liu_params))
    for i, p in enumerate(liu_params):
        q = st.chi2.ppf(1.0 - T, p["dof_x"])
        out[i] = (q - p["dof_x"]) / (2.0 * p["dof_x"]) ** 0.5 * p["sigma_q"] + p["mu_q"]
    return out
