"""ORACLE (test infrastructure).  Restatement of the numpy-sugar routines the reference calls.

numpy-sugar>=1.5.1 (setup.cfg:29) is not vendored/installed -> PARITY UNPINNED, except that the
reference carries its own copy of economic_qs / economic_qs_linear in cellregmap/_math.py:204-256,
which is the semantics followed here.

Call sites in the reference: cellregmap/_cellregmap.py:16-17,106,114,129,172,192,238,415,540,544;
cellregmap/_math.py:29,54,73.
"""
import numpy as np
import scipy.linalg as sla

# numpy_sugar.epsilon  (used by _cellregmap.py:464-469 through lrt_pvalues, and inside glimix_core)
EPS_TINY = float(np.finfo(float).eps)          # epsilon.tiny
EPS_SMALL = float(np.sqrt(np.finfo(float).eps))  # epsilon.small
EPS_SUPER_TINY = float(np.finfo(float).tiny)   # epsilon.super_tiny


def ddot(L, R, left=None):
    """Product with a diagonal matrix given as a vector (numpy_sugar.ddot).

    `left=True` (or 1-d L when left is None): diag(L) @ R.  Otherwise L @ diag(R).
    Used at _math.py:54,73 and _cellregmap.py:415,544."""
    L = np.asarray(L, float)
    R = np.asarray(R, float)
    if left is None:
        left = L.ndim == 1
    if left:
        return (L * R.T).T if R.ndim > 1 else L * R
    return L * R


def economic_qs(K, epsilon=EPS_SMALL):
    """Eigendecomposition K = Q0 S0 Q0' keeping S >= epsilon (absolute).  _math.py:204-235.

    The reference retries with scipy's eigh when the first row of Q looks degenerate
    (_math.py:223-228); the retry changes the eigenvector representative only."""
    S, Q = np.linalg.eigh(K)
    first = Q[0]
    bad = abs(max(first.min(), first.max(), key=abs)) < epsilon
    bad = bad and abs(max(K.min(), K.max(), key=abs)) >= epsilon
    if bad:
        S, Q = sla.eigh(K)
    keep = S >= epsilon
    return (Q[:, keep], Q[:, ~keep]), S[keep]


def economic_qs_linear(G, return_q1=True):
    """Economic eigendecomposition of G G'.  _math.py:238-256 (numpy_sugar signature with return_q1).

    Tall G (n > m): thin SVD, S0 = sigma**2, *no* filtering of zero singular values.
    Otherwise: economic_qs(G G')."""
    G = np.asarray(G, float)
    if G.shape[0] > G.shape[1]:
        Q, sv, _ = np.linalg.svd(G, full_matrices=return_q1)
        S0 = sv ** 2
        if return_q1:
            keep = np.zeros(Q.shape[1], bool)
            keep[: S0.shape[0]] = True
            return (Q[:, keep], Q[:, ~keep]), S0
        return (Q,), S0
    (Q0, Q1), S0 = economic_qs(G @ G.T)
    if return_q1:
        return (Q0, Q1), S0
    return (Q0,), S0


def economic_svd(G, epsilon=EPS_SMALL):
    """Thin SVD keeping singular values >= epsilon (absolute).  Used at _cellregmap.py:540 and
    inside glimix_core.lmm.LMM for the fixed-effect design."""
    G = np.asarray(G, float)
    U, S, Vt = sla.svd(G, full_matrices=False, check_finite=False)
    keep = S >= epsilon
    return U[:, keep], S[keep], Vt[keep, :]


def rsolve(A, b, epsilon=EPS_SMALL):
    """numpy_sugar.linalg.rsolve: least-squares solve with rcond=epsilon; zeros when A is null."""
    A = np.asarray(A, float)
    b = np.asarray(b, float)
    if A.shape[0] == 0:
        return np.zeros((A.shape[1],))
    if A.shape[1] == 0:
        return np.zeros((0,))
    try:
        x = np.linalg.lstsq(A, b, rcond=epsilon)
        r = int(np.sum(x[3] > epsilon))
        if r == 0:
            return np.zeros(A.shape[1])
        return x[0]
    except (ValueError, np.linalg.LinAlgError):
        return np.linalg.solve(A, b)
