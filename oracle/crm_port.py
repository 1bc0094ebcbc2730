"""ORACLE (test infrastructure).  CPU restatement of cellregmap/_cellregmap.py: the model
object, the interaction scan (the hot path), the association scans, the effect-size estimator
and the functional API, on top of the dependency restatements in this package.

Keeps the reference's computational structure (explicit Q0 per rho1; a fresh LMM -- and a fresh
rotation of y and X -- for every (SNP, rho1) pair; serial SNP loop), because it doubles as the
"port" CPU baseline of bench.py.  End-to-end outputs: PARITY UNPINNED (see oracle/__init__.py).
"""
import numpy as np
from scipy.stats import chi2

from . import math_port as mp
from .chiscore_port import filter_lambda, pvalue_from_lambda
from .lmm_port import LMM
from .sugar_port import EPS_SUPER_TINY, EPS_TINY, economic_qs_linear, economic_svd


def rho_grid(has_background):
    """_cellregmap.py:101-108,119."""
    return np.linspace(0, 1, 11) if has_background else np.asarray([1.0])


def get_L_values(hK, E):
    """L_i = diag(U_i S_i) hK so that K o EE' = sum_i L_i L_i'.  _cellregmap.py:533-545."""
    U, S, _ = economic_svd(E)
    us = U * S
    hK = np.asarray(hK, float)
    return [us[:, i][:, None] * hK for i in range(us.shape[1])]


def lrt_pvalues(null_lml, alt_lmls, dof=1):
    """_cellregmap.py:443-469."""
    lrs = np.clip(-2.0 * null_lml + 2.0 * np.asarray(alt_lmls, float), EPS_SUPER_TINY, np.inf)
    return np.clip(chi2(df=dof).sf(lrs), EPS_SUPER_TINY, 1.0 - EPS_TINY)


def compute_maf(X):
    """numpy branch of _cellregmap.py:589-638."""
    X = np.asarray(X, float)
    s0 = np.nansum(X, axis=0) / (2.0 * np.logical_not(np.isnan(X)).sum(axis=0))
    return np.minimum(s0, 1.0 - s0)


def _qs_via_gram(hS):
    ev, V = np.linalg.eigh(hS.T @ hS)
    keep = ev > 1e-12 * ev[-1]
    ev, V = ev[keep][::-1], V[:, keep][:, ::-1]
    return ((hS @ V) / np.sqrt(ev),), ev


class CellRegMapOracle:
    """_cellregmap.py:23-440."""

    def __init__(self, y, E, W=None, Ls=None, E1=None, hK=None, qs_method="svd"):
        """`qs_method="gram"` obtains the same (Q0, S0) from the eigendecomposition of hS'hS instead of a thin SVD
        of hS; it exists only to bound the set-up time of the CPU baseline at n = 1e5 (bench.py) and is not the
        reference's route (numpy_sugar.economic_qs_linear)."""
        self.y = np.asarray(y, float).flatten()
        self.E0 = np.asarray(E, float)
        n = self.y.shape[0]
        self.W = np.ones((n, 1)) if W is None else np.asarray(W, float)
        self.E1 = self.E0 if E1 is None else np.asarray(E1, float)
        self.Ls = [np.asarray(L, float) for L in ([] if Ls is None else Ls)]
        assert self.W.ndim == 2 and self.E0.ndim == 2 and self.E1.ndim == 2
        assert n == self.W.shape[0] == self.E0.shape[0] == self.E1.shape[0]
        for L in self.Ls:
            assert L.ndim == 2 and L.shape[0] == n
        if len(self.Ls) == 0 and hK is None:      # :103-106
            blocks = None
        elif len(self.Ls) == 0:                   # :107-116
            blocks = [np.asarray(hK, float)]
        else:                                     # :117-131  (hK ignored when Ls is given)
            blocks = self.Ls
        self.rho1 = rho_grid(blocks is not None)
        self.QS = {}
        for rho in self.rho1:
            if blocks is None:
                hS = self.E1
            else:
                hS = np.concatenate([np.sqrt(rho) * self.E1] + [np.sqrt(1 - rho) * B for B in blocks], axis=1)
            self.QS[rho] = economic_qs_linear(hS, return_q1=False) if qs_method == "svd" else _qs_via_gram(hS)

    @property
    def n_samples(self):
        return self.y.shape[0]

    # ----------------------------------------------------------------------------------------
    def _best_fit(self, X, restricted, trace=None):
        """rho1 loop with strict '>' (first maximum wins).  _cellregmap.py:343-357,250-260."""
        best_lml, best_rho, best = -np.inf, 0, None
        for rho in self.rho1:
            lmm = LMM(self.y, X, self.QS[rho], restricted=restricted)
            lmm.fit(verbose=False)
            l = lmm.lml()
            if trace is not None:
                trace.append((float(rho), l, lmm.delta, lmm.scale, lmm._x, lmm.nfev))
            if l > best_lml:
                best_lml, best_rho, best = l, rho, lmm
        return best_lml, best_rho, best

    def scan_interaction(self, G, idx_E=None, idx_G=None, stages=None):
        """_cellregmap.py:317-440.  `stages`, when a dict, receives per-SNP intermediates
        (lml grid, Q, weight matrix, filtered eigenvalues, Davies/Liu details)."""
        G = np.asarray(G, float)
        pvalues, info = [], {"rho1": [], "e2": [], "g2": [], "eps2": []}
        if stages is not None:
            for key in ("fits", "lml", "v0", "v1", "Q", "M", "lambdas", "pinfo"):
                stages[key] = []
        for i in range(G.shape[1]):
            g = G[:, [i]]
            X = np.concatenate((self.W, g), axis=1)
            fits = [] if stages is not None else None
            lml, rho, lmm = self._best_fit(X, True, fits)
            info["rho1"].append(rho)
            info["e2"].append(lmm.v0 * rho)
            info["g2"].append(lmm.v0 * (1 - rho))
            info["eps2"].append(lmm.v1)
            Q0, S0 = self.QS[rho][0][0], self.QS[rho][1]
            P = mp.Projection(Q0, S0, lmm.v0, lmm.v1, X)
            E0 = self.E0 if idx_E is None else self.E0[idx_E, :]
            gtest = g.ravel() if idx_G is None else g.ravel()[idx_G]
            gE = gtest[:, None] * E0
            Q = mp.score_statistic_structured(P, gE, self.y)
            M = mp.weight_matrix_structured(P, gE)
            lam = filter_lambda(M)
            p, pinfo = pvalue_from_lambda(lam, Q)
            pvalues.append(p)
            if stages is not None:
                stages["fits"].append(fits); stages["lml"].append(lml)
                stages["v0"].append(lmm.v0); stages["v1"].append(lmm.v1)
                stages["Q"].append(Q); stages["M"].append(M); stages["lambdas"].append(lam)
                stages["pinfo"].append(pinfo)
        return np.asarray(pvalues, float), {k: np.asarray(v, float) for k, v in info.items()}

    def _null_association(self):
        lml, rho, lmm = self._best_fit(self.W, False)
        info = {"rho1": [rho], "e2": [lmm.v0 * rho], "g2": [lmm.v0 * (1 - rho)], "eps2": [lmm.v1]}
        return lml, rho, lmm, {k: np.asarray(v, float) for k, v in info.items()}

    def scan_association(self, G):
        """_cellregmap.py:246-281."""
        null_lml, rho, _, info = self._null_association()
        alt = []
        for i in range(G.shape[1]):
            X = np.concatenate((self.W, G[:, [i]]), axis=1)
            lmm = LMM(self.y, X, self.QS[rho], restricted=False)
            lmm.fit(verbose=False)
            alt.append(lmm.lml())
        return np.asarray(lrt_pvalues(null_lml, alt, dof=1), float), info

    def scan_association_fast(self, G):
        """_cellregmap.py:284-314."""
        null_lml, _, lmm, info = self._null_association()
        alt = lmm.get_fast_scanner().fast_scan(G, verbose=False)["lml"]
        return np.asarray(lrt_pvalues(null_lml, alt, dof=1), float), info

    def predict_interaction(self, G, MAF):
        """_cellregmap.py:137-205.  Returns (beta_g (p,), beta_gxe (1, n, p))."""
        G = np.asarray(G, float)
        maf = np.asarray(np.atleast_1d(MAF), float)
        norm = 1.0 / np.sqrt(2.0 * maf * (1.0 - maf))
        beta_g_s, beta_gxe_s = [], []
        for i in range(G.shape[1]):
            g = G[:, [i]]
            M = np.concatenate((self.W, g, self.E0), axis=1)
            gE = g * self.E0
            best_lml, best_rho, best, halves = -np.inf, 0, None, {}
            for rho in self.rho1:
                halves[rho] = np.concatenate([np.sqrt(rho) * gE] + [np.sqrt(1 - rho) * L for L in self.Ls], axis=1)
                lmm = LMM(self.y, M, economic_qs_linear(halves[rho], return_q1=False), restricted=True)
                lmm.fit(verbose=False)
                if lmm.lml() > best_lml:
                    best_lml, best_rho, best = lmm.lml(), rho, lmm
            beta_g = best.beta[self.W.shape[1]]
            yadj = (self.y - best.mean()).reshape(self.y.shape[0], 1)
            qs = economic_qs_linear(halves[best_rho], return_q1=False)
            v = mp.qscov_solve(qs[0][0], qs[1], best.v0, best.v1, yadj)
            beta_gxe = (best.v0 * best_rho) * self.E0 @ (gE.T @ v) * norm[i]
            beta_g_s.append(beta_g)
            beta_gxe_s.append(beta_gxe)
        return np.asarray(beta_g_s), np.stack(beta_gxe_s).T


def run_interaction(y, E, G, W=None, E1=None, E2=None, hK=None, idx_G=None, stages=None, qs_method="svd"):
    """_cellregmap.py:547-587  (idx_G lands on scan_interaction's idx_E, :586)."""
    E1 = E if E1 is None else E1
    E2 = E if E2 is None else E2
    Ls = None if hK is None else get_L_values(hK, E2)
    crm = CellRegMapOracle(y=y, E=E, W=W, E1=E1, Ls=Ls, qs_method=qs_method)
    return crm.scan_interaction(G, idx_G, stages=stages)


def run_association(y, W, E, G, hK=None):
    """_cellregmap.py:471-500  (positional call: W becomes the context matrix, E the covariates)."""
    return CellRegMapOracle(y, W, E, hK=hK).scan_association(G)


def run_association_fast(y, W, E, G, hK=None):
    """_cellregmap.py:502-531."""
    return CellRegMapOracle(y, W, E, hK=hK).scan_association_fast(G)


def estimate_betas(y, W, E, G, maf=None, E1=None, E2=None, hK=None):
    """_cellregmap.py:640-682."""
    E1 = E if E1 is None else E1
    E2 = E if E2 is None else E2
    Ls = None if hK is None else get_L_values(hK, E2)
    crm = CellRegMapOracle(y=y, E=E, W=W, E1=E1, Ls=Ls)
    if maf is None:
        maf = compute_maf(G)
    return crm.predict_interaction(G, maf)
