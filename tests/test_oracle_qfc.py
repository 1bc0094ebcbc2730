"""Oracle qfc (Davies AS 155 restatement) against independent answers: closed-form chi-square cases and
Imhof's formula integrated with scipy.  The acc=1e-6 of the reference path bounds the expected error."""
import numpy as np
import pytest
from scipy import integrate, stats

from oracle import chiscore_port as cp


def imhof_sf(lam, q):
    lam = np.asarray(lam, float)

    def integrand(u):
        theta = 0.5 * np.sum(np.arctan(lam * u)) - 0.5 * q * u
        rho = np.prod((1.0 + lam ** 2 * u ** 2) ** 0.25)
        return np.sin(theta) / (u * rho)

    val, _ = integrate.quad(integrand, 0, np.inf, limit=2000, epsabs=1e-12, epsrel=1e-12)
    return 0.5 + val / np.pi


def test_single_eigenvalue_uses_liu():
    # one chi2_1 term: the inversion integrand decays like u^-1/2, the routine needs more than lim=10000 terms
    # (ifault 1) -- the reason chiscore/SKAT return the (here exact) Liu value when only one eigenvalue is kept.
    for lam, q in [(1.0, 0.5), (2.5, 3.0), (0.3, 4.0)]:
        qf, ifault, _ = cp.qfc([lam], q)
        assert ifault in (0, 1)
        p, info = cp.pvalue_from_lambda(np.array([lam]), q)
        assert p == info["liu_pval"]
        assert abs(p - stats.chi2.sf(q / lam, 1)) < 1e-7 * p + 1e-12


def test_equal_eigenvalues_closed_form():
    qf, ifault, _ = cp.qfc([0.7] * 6, 5.0)
    assert ifault == 0
    assert abs(qf - stats.chi2.cdf(5.0 / 0.7, 6)) < 2e-6


@pytest.mark.parametrize("seed", range(8))
def test_against_imhof(seed):
    rng = np.random.default_rng(seed)
    lam = np.sort(rng.uniform(0.05, 2.0, size=rng.integers(2, 12)))[::-1]
    q = float(lam.sum() * rng.uniform(0.3, 3.0))
    qf, ifault, trace = cp.qfc(lam, q)
    assert ifault == 0
    assert abs((1.0 - qf) - imhof_sf(lam, q)) < 2e-6
    assert trace[6] < 200


def test_far_tail_falls_back_to_liu():
    lam = np.array([1.0, 0.5, 0.25, 0.1])
    p, info = cp.pvalue_from_lambda(lam, 200.0)
    assert info["Is_Converged"] == 0
    assert p == info["liu_pval"]
    assert 0 < p < 1e-20


def test_filter_lambda_rule():
    M = np.diag([1.0, 0.5, 1e-7, -1e-9, 0.0])
    lam = cp.filter_lambda(M)
    np.testing.assert_allclose(lam, [1.0, 0.5])
    with pytest.raises(RuntimeError):
        cp.filter_lambda(np.zeros((3, 3)))


def test_davies_pvalue_entry_point():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((6, 6))
    M = A @ A.T
    lam = np.linalg.eigvalsh(M)[::-1]
    q = float(lam.sum() * 1.7)
    p, info = cp.davies_pvalue(q, M, True)
    assert abs(p - imhof_sf(lam, q)) < 2e-6
    assert info["Is_Converged"] == 1
