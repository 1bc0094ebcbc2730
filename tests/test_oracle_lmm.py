"""Oracle LMM / Brent restatement against independent formulations."""
import numpy as np
from scipy import optimize
from scipy.stats import multivariate_normal

from oracle import brent_port
from oracle.lmm_port import LMM, logistic_delta
from oracle.sugar_port import economic_qs_linear


def _problem(seed=0, n=40, r=6, p=2):
    rng = np.random.default_rng(seed)
    Gh = rng.standard_normal((n, r))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    y = X @ rng.standard_normal(p) + Gh @ rng.standard_normal(r) * 0.7 + rng.standard_normal(n)
    return y, X, Gh


def _dense_lml(y, X, K, delta, restricted):
    n, p = X.shape
    V = (1 - delta) * K + delta * np.eye(n)
    Vi = np.linalg.inv(V)
    beta = np.linalg.solve(X.T @ Vi @ X, X.T @ Vi @ y)
    r = y - X @ beta
    df = n - p if restricted else n
    s = (r @ Vi @ r) / df
    if not restricted:
        return multivariate_normal(X @ beta, s * V).logpdf(y)
    _, ldV = np.linalg.slogdet(s * V)
    _, ldH = np.linalg.slogdet(X.T @ Vi @ X / s)
    _, ldX = np.linalg.slogdet(X.T @ X)
    return -0.5 * ((n - p) * np.log(2 * np.pi) + ldV + ldH - ldX + (r @ Vi @ r) / s)


def test_lml_matches_dense_ml_and_reml():
    y, X, Gh = _problem()
    QS = economic_qs_linear(Gh, return_q1=False)
    K = Gh @ Gh.T
    for restricted in (False, True):
        lmm = LMM(y, X, QS, restricted=restricted)
        for x in (-2.0, 0.0, 1.3):
            got = lmm._evaluate(x)[0]
            want = _dense_lml(y, X, K, logistic_delta(x), restricted)
            assert abs(got - want) < 1e-9 * max(1.0, abs(want))


def test_fit_finds_the_maximum():
    y, X, Gh = _problem(seed=1)
    QS = economic_qs_linear(Gh, return_q1=False)
    lmm = LMM(y, X, QS, restricted=True)
    lmm.fit(verbose=False)
    res = optimize.minimize_scalar(lambda x: -lmm._evaluate(x)[0], bounds=(-30, 30), method="bounded", options={"xatol": 1e-10})
    assert lmm.lml() >= -res.fun - 1e-8
    assert abs(lmm.v0 + lmm.v1 - lmm.scale) < 1e-12
    assert 10 < lmm.nfev < 80


def test_brent_on_analytic_functions():
    x, fx, nfev = brent_port.minimize(lambda x: (x - 3.2) ** 2 + 1.0, a=-700, b=700, rtol=1e-6, atol=1e-6)
    assert abs(x - 3.2) < 1e-5 and abs(fx - 1.0) < 1e-9
    x, fx, _ = brent_port.minimize(lambda x: np.cosh(0.3 * (x + 7.0)), a=-700, b=700, rtol=1e-6, atol=1e-6)
    assert abs(x + 7.0) < 1e-4
    # monotone function: the search runs into the bound
    x, fx, _ = brent_port.minimize(lambda x: np.exp(-0.01 * x), a=-50.0, b=50.0, rtol=1e-6, atol=1e-6)
    assert x > 49.9


def test_rank_deficient_design_is_reduced():
    y, X, Gh = _problem(seed=2)
    X2 = np.column_stack([X, X[:, 0] * 2.0])
    QS = economic_qs_linear(Gh, return_q1=False)
    a = LMM(y, X, QS, restricted=True)
    b = LMM(y, X2, QS, restricted=True)
    a.fit(verbose=False)
    b.fit(verbose=False)
    assert abs(a.lml() - b.lml() - 0.0) < 5.0  # different X'X normalisation enters REML only through logdet terms
    assert b._tX.shape[1] == 2
