"""ORACLE (test infrastructure).  Restatement of chiscore.liu_sf / chiscore.davies_pvalue.

chiscore>=0.2.3 (setup.cfg:27; wraps the C library chi2comb) is not vendored / not installed.
`liu_sf(..., kurtosis=True)` is PINNED by the reference's known answers at
cellregmap/test/test_math.py:76-83; `davies_pvalue` is PARITY UNPINNED (restates SKAT's
Get_Lambda / Get_PValue.Lambda rules around Davies' AS 155, see oracle/qfc_oracle.c).

Reference call sites: cellregmap/_cellregmap.py:333,435 (davies_pvalue(Q, M, True));
cellregmap/_math.py:169,179 (liu_sf).
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy.stats import ncx2

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libqfc_oracle.so")
_lib = None


def build_c_oracle(force=False):
    """Compile oracle/qfc_oracle.c with gcc (the checker is built, not shipped)."""
    src = os.path.join(_HERE, "qfc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _qfc_lib():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = ctypes.CDLL(_SO)
        _lib.crm_oracle_qfc.restype = ctypes.c_double
        _lib.crm_oracle_qfc.argtypes = [
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int),
            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double,
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    return _lib


def qfc(lambdas, q, noncentrality=None, dofs=None, sigma=0.0, lim=10000, acc=1e-6):
    """P(sum lambda_j chi2(dof_j, nc_j) + sigma N(0,1) < q) by Davies' method.
    Returns (qfval, ifault, trace[7])."""
    lb = np.ascontiguousarray(lambdas, dtype=np.float64)
    r = lb.shape[0]
    nc = np.zeros(r) if noncentrality is None else np.ascontiguousarray(noncentrality, dtype=np.float64)
    n = np.ones(r, dtype=np.int32) if dofs is None else np.ascontiguousarray(dofs, dtype=np.int32)
    trace = np.zeros(7)
    ifault = ctypes.c_int(0)
    val = _qfc_lib().crm_oracle_qfc(
        lb.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nc.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        n.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), r, float(sigma), float(q), int(lim), float(acc),
        trace.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(ifault))
    return float(val), int(ifault.value), trace


def liu_sf(t, lambs, dofs, deltas, kurtosis=False):
    """Liu-Tang-Zhang (2009) survival function of sum lambda_i chi2(h_i, delta_i) at t; with
    `kurtosis=True` the SKAT-O modification (match kurtosis instead of skewness).
    Returns (p, dof_x, delta_x, {mu_q, sigma_q})."""
    t = np.asarray(t, float)
    lambs = np.asarray(lambs, float)
    dofs = np.asarray(dofs, float)
    deltas = np.asarray(deltas, float)
    lp = [lambs ** i for i in range(1, 5)]
    c = [float(np.sum(lp[i] * dofs) + (i + 1) * np.sum(lp[i] * deltas)) for i in range(4)]
    s1 = c[2] / np.sqrt(c[1]) ** 3
    s2 = c[3] / c[1] ** 2
    s12 = s1 ** 2
    if s12 > s2:
        a = 1.0 / (s1 - np.sqrt(s12 - s2))
        delta_x = s1 * a ** 3 - a ** 2
        dof_x = a ** 2 - 2.0 * delta_x
    else:
        delta_x = 0.0
        if kurtosis:
            a = 1.0 / np.sqrt(s2)
            dof_x = 1.0 / s2
        else:
            a = 1.0 / s1
            dof_x = 1.0 / s12
    mu_q = c[0]
    sigma_q = np.sqrt(2.0 * c[1])
    mu_x = dof_x + delta_x
    sigma_x = np.sqrt(2.0 * (dof_x + 2.0 * delta_x))
    t_star = (t - mu_q) / sigma_q
    tfinal = t_star * sigma_x + mu_x
    p = ncx2.sf(tfinal, dof_x, np.maximum(delta_x, 1e-9))
    return p, dof_x, delta_x, {"mu_q": mu_q, "sigma_q": sigma_q}


def filter_lambda(M):
    """Eigenvalues of the symmetric matrix M, descending, keeping lambda > mean(lambda >= 0)/1e5."""
    lam = np.linalg.eigvalsh(np.asarray(M, float))[::-1]
    nonneg = lam[lam >= 0]
    if nonneg.size == 0:
        raise RuntimeError("No eigenvalue is bigger than 0!!")
    lam = lam[lam > nonneg.mean() / 100000.0]
    if lam.size == 0:
        raise RuntimeError("No eigenvalue is bigger than 0!!")
    return lam


def pvalue_from_lambda(lam, q, lim=10000, acc=1e-6):
    """Davies p-value with the modified-Liu fallbacks; returns (p, info)."""
    lam = np.asarray(lam, float)
    p_liu = float(liu_sf(q, lam, np.ones(lam.size), np.zeros(lam.size), True)[0])
    qfval, ifault, trace = qfc(lam, q, lim=lim, acc=acc)
    p = 1.0 - qfval
    converged = 1
    if lam.size == 1:
        p = p_liu
    elif ifault != 0:
        converged = 0
    if p > 1.0 or p <= 0.0:
        converged = 0
        p = p_liu
    return p, {"liu_pval": p_liu, "Is_Converged": converged, "ifault": ifault, "qfval": qfval, "trace": trace}


def davies_pvalue(q, w, return_info=False):
    """chiscore.davies_pvalue(q, w, return_info): w is the k x k matrix whose eigenvalues weight
    the chi-squares."""
    lam = filter_lambda(np.atleast_2d(np.asarray(w, float)))
    p, info = pvalue_from_lambda(lam, float(q))
    if return_info:
        return p, {"liu_pval": info["liu_pval"], "Is_Converged": info["Is_Converged"]}
    return p
