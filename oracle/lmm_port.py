"""ORACLE (test infrastructure).  Restatement of glimix_core.lmm.LMM (3.1.x) and its FastScanner.

glimix-core>=3.1.12 (setup.cfg:28) is not vendored / not installed -> PARITY UNPINNED; the model is
the FaST-LMM two-component likelihood (Lippert et al. 2011), cross-checked against the dense
textbook ML/REML log-likelihoods in tests/test_oracle_lmm.py.

Reference call sites: cellregmap/_cellregmap.py:3,175-176,223-224,254-255,274-276,292-293,308-309,
351-352 (constructor + fit), and the attributes read afterwards: lml() :178,257,276,354;
v0, v1 :190-191,264-266,367-369,382-383; beta :186; mean() :188; get_fast_scanner() :308.

Model:  y ~ N(X beta, s ((1-delta) K + delta I)),  K = Q0 S0 Q0',  v0 = s (1-delta), v1 = s delta.
Only Q0 is supplied by the reference (economic_qs_linear(..., return_q1=False)), so the
complement-space terms are obtained by subtraction (y'y - |Q0'y|^2, ...).
"""
import math

import numpy as np

from . import brent_port
from .sugar_port import EPS_SMALL, EPS_TINY, economic_svd, rsolve

LOGMAX = math.log(np.finfo(float).max)  # bounds of the logistic variable
LOG2PI = math.log(2.0 * math.pi)


def logistic_delta(x):
    """delta(x): stable logistic, clipped to [tiny, 1 - tiny]."""
    if x > 0.0:
        v = 1.0 / (1.0 + math.exp(-x))
    else:
        v = math.exp(x)
        v = v / (v + 1.0)
    return min(max(v, EPS_TINY), 1.0 - EPS_TINY)


class LMM:
    def __init__(self, y, X, QS, restricted=False):
        y = np.asarray(y, float).ravel()
        if not np.all(np.isfinite(y)):
            raise ValueError("There are non-finite values in the outcome.")
        if y.size == 0:
            raise ValueError("The outcome array is empty.")
        X = np.atleast_2d(np.asarray(X, float).T).T
        if not np.all(np.isfinite(X)):
            raise ValueError("There are non-finite values in the covariates matrix.")
        Q0 = QS[0][0]
        S0 = np.asarray(QS[1], float)
        if Q0.shape[0] != y.shape[0]:
            raise ValueError("Sample size differs between outcome and covariance decomposition.")
        if X.shape[0] != y.shape[0]:
            raise ValueError("Sample size differs between outcome and covariates.")
        self._y = y
        self._Q0 = Q0
        self._S0 = S0
        self._restricted = bool(restricted)
        U, sv, Vt = economic_svd(X)          # rank-revealing reparametrisation of the design
        self._tX = U * sv                    # n x rank
        self._Vt = Vt
        self._X = X
        self._x = 0.0                        # logistic variable; delta = 0.5
        self._tbeta = np.zeros(sv.shape[0])
        self._scale = 1.0
        # rotations (the reference repeats these for every (SNP, rho1) pair)
        self._yr = Q0.T @ y                  # r
        self._Xr = Q0.T @ self._tX           # r x rank
        self._yy = float(y @ y)
        self._Xy = self._tX.T @ y
        self._XX = self._tX.T @ self._tX
        self._yy_res = self._yy - float(self._yr @ self._yr)
        self._Xy_res = self._Xy - self._Xr.T @ self._yr
        self._XX_res = self._XX - self._Xr.T @ self._Xr
        self.nfev = 0

    # -- sizes ---------------------------------------------------------------------------------
    @property
    def nsamples(self):
        return self._y.shape[0]

    @property
    def _df(self):
        return self.nsamples - self._tX.shape[1] if self._restricted else self.nsamples

    # -- parameters ----------------------------------------------------------------------------
    @property
    def delta(self):
        return logistic_delta(self._x)

    @property
    def scale(self):
        return self._scale

    @property
    def v0(self):
        return self._scale * (1.0 - self.delta)

    @property
    def v1(self):
        return self._scale * self.delta

    @property
    def beta(self):
        return rsolve(self._Vt, rsolve(self._tX, self.mean()))

    def mean(self):
        return self._tX @ self._tbeta

    # -- likelihood ----------------------------------------------------------------------------
    def _terms(self, delta):
        D0 = self._S0 * (1.0 - delta) + delta
        w = 1.0 / D0
        yKy = float((self._yr * self._yr) @ w) + self._yy_res / delta
        XKy = self._Xr.T @ (self._yr * w) + self._Xy_res / delta
        XKX = (self._Xr.T * w) @ self._Xr + self._XX_res / delta
        logdetK = float(np.sum(np.log(D0))) + (self.nsamples - D0.shape[0]) * math.log(delta)
        return yKy, XKy, XKX, logdetK

    def _evaluate(self, x):
        """lml at logistic value x with beta and scale at their conditional optima."""
        delta = logistic_delta(x)
        yKy, XKy, XKX, logdetK = self._terms(delta)
        tbeta = rsolve(XKX, XKy)
        scale = max((yKy - float(XKy @ tbeta)) / self._df, EPS_SMALL)
        n = self.nsamples
        lml = -self._df * LOG2PI - self._df - n * math.log(scale) - logdetK
        lml /= 2.0
        if self._restricted:
            sgn0, ld0 = np.linalg.slogdet(self._XX)
            if sgn0 != 1.0:
                raise ValueError("The determinant of X'X should be positive.")
            sgn1, ld1 = np.linalg.slogdet(XKX / scale)
            if sgn1 != 1.0:
                raise ValueError("The determinant of H should be positive.")
            lml += (ld0 - ld1) / 2.0
        return lml, tbeta, scale

    def lml(self):
        return self._evaluate(self._x)[0]

    def fit(self, verbose=True):
        """Maximise the lml over logit(delta) with bracket+Brent at rtol = atol = 1e-6
        (optimix scalar path of glimix_core), then refresh beta and scale."""
        def neg(x):
            self.nfev += 1
            return -self._evaluate(x)[0]

        x, _, _ = brent_port.minimize(neg, a=-LOGMAX, b=+LOGMAX, rtol=1e-6, atol=1e-6)
        self._x = x
        _, self._tbeta, self._scale = self._evaluate(x)

    # -- fast scanner (scan_association_fast, _cellregmap.py:307-309) ----------------------------
    def get_fast_scanner(self):
        return FastScanner(self._y, self._X, self._Q0, self.v0 * self._S0, self.v1)


class FastScanner:
    """ML scan of single candidate columns with the covariance *shape* v0 K + v1 I frozen and an
    overall scale re-estimated per candidate (glimix_core.lmm.FastScanner.fast_scan)."""

    def __init__(self, y, X, Q0, S, v):
        self._y = y
        self._X = X
        self._Q0 = Q0
        self._D = S + v
        self._v = v
        n = y.shape[0]
        self._n = n
        self._logdetK = float(np.sum(np.log(self._D))) + (n - S.shape[0]) * math.log(v)

    def _quad(self, A, B):
        Ar = self._Q0.T @ A
        Br = self._Q0.T @ B
        return (Ar.T / self._D) @ Br + (A.T @ B - Ar.T @ Br) / self._v

    def fast_scan(self, M, verbose=False):
        M = np.asarray(M, float)
        y = self._y[:, None]
        n = self._n
        lmls = np.empty(M.shape[1])
        eff1 = np.empty(M.shape[1])
        scales = np.empty(M.shape[1])
        for i in range(M.shape[1]):
            Z = np.concatenate([self._X, M[:, [i]]], axis=1)
            ZKZ = self._quad(Z, Z)
            ZKy = self._quad(Z, y)[:, 0]
            yKy = float(self._quad(y, y)[0, 0])
            beta = rsolve(ZKZ, ZKy)
            scale = max((yKy - float(ZKy @ beta)) / n, EPS_SMALL)
            lmls[i] = -0.5 * (n * LOG2PI + n + n * math.log(scale) + self._logdetK)
            eff1[i] = beta[-1]
            scales[i] = scale
        return {"lml": lmls, "effsizes1": eff1, "scale": scales}
