"""ORACLE (test infrastructure).  Restatement of brent-search's scalar minimiser.

brent-search is pulled in by glimix-core (setup.cfg:28) -> optimix -> brent_search.minimize and
is what `LMM.fit` runs on logit(delta) (call site: cellregmap/_cellregmap.py:351-352).
Not vendored / not installed -> PARITY UNPINNED.  `brent` is Brent's (1973) `localmin` with the
package's bookkeeping; `bracket` is the package's downhill bracketing: start at
x0 = clip(0, a, b), second point one tolerance-sized step away (`gfactor*(rtol*|x0|+atol)`),
then geometric expansion by `gfactor` until f rises or a bound is hit.
"""
import math

GOLD = 0.381966011250105097
_EPS = 1.4902e-08


def _initial_pair(x0, x1, a, b, gfactor, rtol, atol):
    xs = sorted(v for v in (x0, x1) if v is not None)
    if len(xs) == 0:
        x0 = min(max(0.0, a), b)
        x1 = None
    elif len(xs) == 1:
        x0, x1 = xs[0], None
    else:
        x0, x1 = xs
    if x1 is None:
        step = gfactor * (rtol * abs(x0) + atol)
        if x0 - a > b - x0:
            x1 = max(x0 - step, a)
        else:
            x1 = min(x0 + step, b)
    return x0, x1


def bracket(f, x0=None, x1=None, a=-math.inf, b=math.inf, gfactor=2.0, rtol=_EPS, atol=_EPS, maxiter=500):
    """Returns (xl, xm, xr, fl, fm, fr, nfev) with xl < xm < xr (or a degenerate triple at a bound)
    such that fm <= fl and fm <= fr whenever a proper bracket exists."""
    x0, x1 = _initial_pair(x0, x1, a, b, gfactor, rtol, atol)
    f0 = f(x0)
    f1 = f(x1)
    nfev = 2
    if f0 < f1:               # make x0 -> x1 the downhill direction
        x0, x1 = x1, x0
        f0, f1 = f1, f0
    x2, f2 = x1, f1
    it = 0
    while it < maxiter:
        it += 1
        step = (x1 - x0) * gfactor
        x2 = x1 + step
        x2 = min(max(x2, a), b)
        if x2 == x1:          # pinned at a bound: degenerate bracket
            f2 = f1
            break
        f2 = f(x2)
        nfev += 1
        if f2 > f1:
            break
        x0, f0 = x1, f1
        x1, f1 = x2, f2
    if x0 > x2:
        x0, x2 = x2, x0
        f0, f2 = f2, f0
    return x0, x1, x2, f0, f1, f2, nfev


def brent(f, a, b, x0, f0, rtol=_EPS, atol=_EPS, maxiter=500):
    """Brent's localmin on [a, b] started from the interior point (x0, f0).
    Returns (x, fx, nfev)."""
    x1 = x2 = x0
    f1 = f2 = f0
    d = e = 0.0
    nfev = 0
    for _ in range(maxiter):
        m = 0.5 * (a + b)
        tol = rtol * abs(x0) + atol
        tol2 = 2.0 * tol
        if abs(x0 - m) <= tol2 - 0.5 * (b - a):
            break
        p = q = r = 0.0
        if tol < abs(e):
            r = (x0 - x1) * (f0 - f2)
            q = (x0 - x2) * (f0 - f1)
            p = (x0 - x2) * q - (x0 - x1) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = d
        if abs(p) < abs(0.5 * q * r) and q * (a - x0) < p and p < q * (b - x0):
            d = p / q
            u = x0 + d
            if (u - a) < tol2 or (b - u) < tol2:
                d = tol if x0 < m else -tol
        else:
            e = (b if x0 < m else a) - x0
            d = GOLD * e
        if abs(d) >= tol:
            u = x0 + d
        elif d > 0.0:
            u = x0 + tol
        else:
            u = x0 - tol
        fu = f(u)
        nfev += 1
        if fu <= f0:
            if u < x0:
                b = x0
            else:
                a = x0
            x2, f2 = x1, f1
            x1, f1 = x0, f0
            x0, f0 = u, fu
        else:
            if u < x0:
                a = u
            else:
                b = u
            if fu <= f1 or x1 == x0:
                x2, f2 = x1, f1
                x1, f1 = u, fu
            elif fu <= f2 or x2 == x0 or x2 == x1:
                x2, f2 = u, fu
    return x0, f0, nfev


def minimize(f, x0=None, x1=None, a=-math.inf, b=math.inf, gfactor=2.0, rtol=_EPS, atol=_EPS, maxiter=500):
    """bracket + brent; returns (x, fx, nfev)."""
    xl, xm, xr, fl, fm, fr, n0 = bracket(f, x0, x1, a, b, gfactor, rtol, atol, maxiter)
    x, fx, n1 = brent(f, xl, xr, xm, fm, rtol, atol, maxiter)
    return x, fx, n0 + n1
