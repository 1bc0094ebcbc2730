"""Seeded synthetic genotype / expression / context data of the BASELINE.json shapes.

Scalable restatement of the *recipe* used by the reference's simulator
(cellregmap/_simulate.py:315-397 `sample_phenotype_gxe`, moment normalisation `:470-474`;
causal-SNP choices from cellregmap/test/test_struct_lmm2.py:15-24), without its n x n covariance
matrices, so that it works at n = 1e5: cells are assigned to donors, genotypes are Hardy-Weinberg
dosages expanded donor -> cell, the donor structure enters through a low-rank factor `hK`.
numpy only; used by tests/ and bench.py.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class SynthData:
    y: np.ndarray      # (n,)
    W: np.ndarray      # (n, c) covariates (intercept)
    E: np.ndarray      # (n, k) cellular contexts
    G: np.ndarray      # (n, p) expanded genotype dosages, float64, C-contiguous
    hK: np.ndarray     # (n, q) expanded low-rank factor of the donor kinship
    donor: np.ndarray  # (n,) donor index of each cell
    maf: np.ndarray    # (p,) population allele frequencies used for sampling


def _moments(v):
    v = v - v.mean()
    s = v.std()
    return v / s if s > 0 else v


def column_normalize(X):
    X = X - X.mean(0)
    s = X.std(0)
    s[s == 0] = 1.0
    return X / s


def make_data(n=500, donors=50, k=10, p=100, q=None, seed=0, n_covariates=1,
              v_env=0.25, v_bg=0.25, v_noise=0.4, v_persistent=0.05, v_gxc=0.05,
              causal_persistent=(5, 6), causal_gxc=(10, 11), normalize_G=False):
    rng = np.random.default_rng(seed)
    q = min(donors, 10) if q is None else q
    donor = np.sort(rng.integers(0, donors, n))
    maf = rng.uniform(0.05, 0.45, p)
    Gd = rng.binomial(2, maf, size=(donors, p)).astype(np.float64)
    for j in np.where(Gd.std(0) == 0)[0]:           # no monomorphic SNPs in the default data
        Gd[rng.integers(0, donors), j] += 1.0
    G = np.ascontiguousarray(Gd[donor])
    if normalize_G:
        G = np.ascontiguousarray(column_normalize(G))
    E = column_normalize(rng.standard_normal((n, k))) / np.sqrt(k)
    A = rng.standard_normal((donors, q)) / np.sqrt(q)
    hK = np.ascontiguousarray(A[donor])
    W = np.ones((n, n_covariates))
    if n_covariates > 1:
        W[:, 1:] = rng.standard_normal((n, n_covariates - 1))
    y = np.full(n, 0.3)
    y += np.sqrt(v_env) * _moments(E @ rng.standard_normal(k))
    bg = np.zeros(n)
    for i in range(k):
        bg += E[:, i] * (hK @ rng.standard_normal(q))
    y += np.sqrt(v_bg) * _moments(bg)
    y += np.sqrt(v_noise) * rng.standard_normal(n)
    cp = [j for j in causal_persistent if j < p]
    if cp:
        y += np.sqrt(v_persistent) * _moments(column_normalize(G[:, cp]) @ rng.standard_normal(len(cp)))
    cg = [j for j in causal_gxc if j < p]
    if cg:
        gx = np.zeros(n)
        for j in cg:
            gx += column_normalize(G[:, [j]])[:, 0] * (E @ rng.standard_normal(k))
        y += np.sqrt(v_gxc) * _moments(gx)
    if n_covariates > 1:
        y += W[:, 1:] @ (0.1 * rng.standard_normal(n_covariates - 1))
    return SynthData(y=y, W=W, E=E, G=G, hK=hK, donor=donor, maf=maf)
