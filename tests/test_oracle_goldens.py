"""Oracle vs the reference's own known answers (cellregmap/test/test_math.py:17-91), stored in
tests/golden/reference_test_math.json."""
import itertools
import json
import os

import numpy as np
from numpy.testing import assert_allclose

from oracle import math_port as mp
from oracle.sugar_port import economic_qs


GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_test_math.json")))


def _data():
    # fixture of the reference test (test_math.py:17-35); the draw of y goes through an SVD inside
    # RandomState.multivariate_normal, so its sign depends on the LAPACK build: enumerate the sign choices.
    random = np.random.RandomState(0)
    W = random.randn(3, 2)
    K0 = random.randn(3, 3)
    K0 = K0 @ K0.T
    K = 0.2 * K0 + np.eye(3)
    alpha = np.array([0.5, -0.2])
    z = random.standard_normal(3)
    u, s, v = np.linalg.svd(K)
    ys = []
    for signs in itertools.product([1.0, -1.0], repeat=3):
        ys.append(W @ alpha + z @ (np.sqrt(s)[:, None] * (v * np.array(signs)[:, None])))
    return {"W": W, "K": K, "dK": K0, "ys": ys}


def test_QSCov():
    d = _data()
    K = d["K"][:, :2] @ d["K"][:, :2].T
    (Q0, _), S0 = economic_qs(K)
    a, b = 0.2, 0.3
    finalK = a * K + b * np.eye(3)
    v = np.array([0.3, -0.2, 0.19])
    assert_allclose(finalK @ v, mp.qscov_dot(Q0, S0, a, b, v))
    assert_allclose(mp.lstsq_solve(finalK, v), mp.qscov_solve(Q0, S0, a, b, v))


def test_P_matrix():
    d = _data()
    P = np.array(GOLD["P_matrix"]["value"])
    assert_allclose(mp.P_matrix(d["W"], d["K"]), P, rtol=2e-7)


def _golden_y(d):
    qs = [mp.score_statistic(y, d["W"], d["K"], d["dK"]) for y in d["ys"]]
    best = int(np.argmin([abs(q - GOLD["score_statistic"]["value"]) for q in qs]))
    return d["ys"][best], qs[best]


def test_score_statistic():
    d = _data()
    _, q = _golden_y(d)
    assert_allclose(q, GOLD["score_statistic"]["value"], rtol=1e-12)


def test_score_statistic_structured_matches_dense():
    d = _data()
    y, q = _golden_y(d)
    (Q0, _), S0 = economic_qs(d["dK"])
    P = mp.Projection(Q0, S0, 0.2, 1.0, d["W"])
    L = np.linalg.cholesky(d["dK"] + 1e-13 * np.eye(3))
    assert_allclose(mp.score_statistic_structured(P, L, y), q, rtol=1e-9)


def test_score_statistic_distr_weights():
    d = _data()
    w = mp.score_statistic_distr_weights(d["W"], d["K"], d["dK"])
    assert_allclose(w, np.array(GOLD["distr_weights"]["value"]), atol=GOLD["distr_weights"]["atol"])


def test_score_statistic_liu_params():
    d = _data()
    _, q = _golden_y(d)
    w = mp.score_statistic_distr_weights(d["W"], d["K"], d["dK"])
    params = mp.score_statistic_liu_params(q, w)
    for key in ("pv", "mu_q", "sigma_q", "dof_x"):
        assert_allclose(params[key], GOLD["liu_params"][key])


def test_qmin():
    assert_allclose(mp.qmin(GOLD["qmin"]["params"]), GOLD["qmin"]["value"])
