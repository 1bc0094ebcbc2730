"""Device-backed versions of the distribution helpers of the reference's `cellregmap/_math.py`:
`score_statistic_liu_params` (:163-180) and `qmin` (:183-201), plus batched forms.  (The structured
covariance classes QSCov / PMat / ScoreStatistic of that module are fused into the score kernel of the scan.)"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._cellregmap import _device, _ptr, _stream, _to_dev


def liu_params_batch(q, weights):
    """Modified-Liu parameters for a batch: q (B,), weights (B, r) or a list of 1-d arrays.
    Returns an array (B, 4) = pv, mu_q, sigma_q, dof_x."""
    dev = _device()
    q = np.atleast_1d(np.asarray(q, float))
    rows = [np.asarray(w, float).ravel() for w in (weights if not isinstance(weights, np.ndarray) or weights.ndim != 2 else list(weights))]
    assert len(rows) == q.shape[0]
    ld = max(1, max(len(r) for r in rows))
    lam = np.zeros((len(rows), ld))
    nlam = np.zeros(len(rows), np.int32)
    for i, r in enumerate(rows):
        lam[i, : len(r)] = r
        nlam[i] = len(r)
    qd, ld_, nd = _to_dev(q, dev), _to_dev(lam, dev), torch.from_numpy(nlam).to(dev)
    out = torch.empty((len(rows), 4), dtype=torch.float64, device=dev)
    _lib.call("crm_liu_params", _ptr(qd), _ptr(ld_), _ptr(nd), ld, len(rows), _ptr(out), _stream())
    return out.cpu().numpy()


def score_statistic_liu_params(q, weights):
    """Pr(Q > q) for Q ~ sum_i w_i chi2(1) by the modified Liu approximation, with its parameters (reference :163-180)."""
    pv, mu_q, sigma_q, dof_x = liu_params_batch([q], [weights])[0]
    return {"pv": pv, "mu_q": mu_q, "sigma_q": sigma_q, "dof_x": dof_x}


def qmin_batch(params):
    """params (B, nrho, 4) = pv, mu_q, sigma_q, dof_x -> (B, nrho) quantile-matched statistics."""
    dev = _device()
    params = np.ascontiguousarray(np.asarray(params, float))
    assert params.ndim == 3 and params.shape[2] == 4
    pd = _to_dev(params, dev)
    out = torch.empty(params.shape[:2], dtype=torch.float64, device=dev)
    _lib.call("crm_qmin", _ptr(pd), params.shape[1], params.shape[0], _ptr(out), _stream())
    return out.cpu().numpy()


def qmin(liu_params):
    """Reference :183-201: list of {"pv", "mu_q", "sigma_q", "dof_x"} over a rho grid -> array of q_min(rho)."""
    arr = np.array([[p["pv"], p["mu_q"], p["sigma_q"], p["dof_x"]] for p in liu_params], float)[None]
    return qmin_batch(arr)[0]
