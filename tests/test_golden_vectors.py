"""Committed vectors (tests/golden/oracle_small.npz, made by tests/golden/make_oracle_vectors.py from the ORACLE):
the oracle still reproduces them (CPU), and the CUDA path matches them (GPU)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz")


def test_oracle_reproduces_committed_vectors():
    from oracle import crm_port
    g = np.load(GOLD)
    pv, info = crm_port.run_interaction(g["y"], g["E"], g["G"], W=g["W"], hK=g["hK"])
    np.testing.assert_array_equal(info["rho1"], g["rho1"])
    np.testing.assert_allclose(pv, g["pv"], rtol=1e-6)
    np.testing.assert_allclose(info["eps2"], g["eps2"], rtol=1e-6)
    pa, _ = crm_port.run_association(g["y"], g["W"], g["E"], g["G"], hK=g["hK"])
    np.testing.assert_allclose(pa, g["assoc_pv"], rtol=1e-6)


@pytest.mark.gpu
def test_cuda_path_matches_committed_vectors(cuda_device):
    from cellregmap_b200 import estimate_betas, run_association, run_association_fast, run_interaction
    from cellregmap_b200._cellregmap import _make_interaction_model
    g = np.load(GOLD)
    model = _make_interaction_model(g["y"], g["E"], g["W"], None, None, g["hK"])
    out = model._scan_interaction_device(g["G"], diagnostics=True)
    np.testing.assert_array_equal(out["rho1"].cpu().numpy(), g["rho1"])
    np.testing.assert_allclose(out["lml"].cpu().numpy(), g["lml"], rtol=1e-6)
    for key in ("e2", "g2", "eps2"):
        np.testing.assert_allclose(out[key].cpu().numpy(), g[key], rtol=1e-6, atol=1e-12)
    assert np.max(np.abs(np.log10(out["pv"].cpu().numpy()) - np.log10(g["pv"]))) <= 1e-4
    pa, ia = run_association(g["y"], g["W"], g["E"], g["G"], hK=g["hK"])
    assert np.max(np.abs(np.log10(pa) - np.log10(g["assoc_pv"]))) <= 1e-4
    np.testing.assert_array_equal(ia["rho1"], g["assoc_rho1"])
    pf, _ = run_association_fast(g["y"], g["W"], g["E"], g["G"], hK=g["hK"])
    assert np.max(np.abs(np.log10(pf) - np.log10(g["assoc_fast_pv"]))) <= 1e-4
    bg, bgxe = estimate_betas(g["y"], g["W"], g["E"], g["G"][:, :4], hK=g["hK"])
    np.testing.assert_allclose(bg, g["beta_g"], rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(bgxe, g["beta_gxe"], rtol=0, atol=2e-5 * np.abs(g["beta_gxe"]).max())
