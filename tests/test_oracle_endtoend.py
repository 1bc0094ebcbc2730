"""End-to-end behaviour of the oracle on small simulated data, in the spirit of the reference's statistical tests
(cellregmap/test/test_struct_lmm2.py:118-119,210-211: causal SNPs are found, null SNPs are calibrated), plus the host-side
helpers of the product that need no GPU."""
import numpy as np
import pytest

from cellregmap_b200.synth import make_data
from oracle import crm_port


@pytest.fixture(scope="module")
def data():
    return make_data(n=600, donors=60, k=6, p=40, q=6, seed=5, v_gxc=0.12, v_persistent=0.1, v_noise=0.3)


def test_interaction_finds_gxc_snps_and_is_calibrated(data):
    d = data
    pv, info = crm_port.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
    assert pv.shape == (40,) and set(info) == {"rho1", "e2", "g2", "eps2"}
    assert pv[10] < 1e-6 and pv[11] < 1e-6
    assert set(np.argsort(pv)[:2]) == {10, 11}
    assert np.all(np.isin(info["rho1"], np.linspace(0, 1, 11)))
    assert np.all(info["eps2"] > 0) and np.all(info["e2"] >= 0) and np.all(info["g2"] >= 0)


def test_interaction_is_calibrated_without_gxc():
    """No simulated GxC effect: p-values look uniform (reference thresholds: median > 0.3, min > 0.04 on fewer SNPs)."""
    d = make_data(n=600, donors=60, k=6, p=40, q=6, seed=8, causal_gxc=(), v_gxc=0.0)
    pv, _ = crm_port.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
    assert np.median(pv) > 0.3 and pv.min() > 1e-3


def test_association_finds_persistent_snps(data):
    d = data
    pv, info = crm_port.run_association(d.y, d.W, d.E, d.G, hK=d.hK)
    assert pv[5] < 1e-3 or pv[6] < 1e-3
    assert info["rho1"].shape == (1,)
    pf, _ = crm_port.run_association_fast(d.y, d.W, d.E, d.G, hK=d.hK)
    assert np.corrcoef(np.log10(pv), np.log10(pf))[0, 1] > 0.99


def test_permuted_contexts_destroy_the_interaction_signal(data):
    d = data
    idx = np.random.default_rng(0).permutation(d.y.shape[0])
    pv, _ = crm_port.run_interaction(d.y, d.E, d.G[:, [10, 11]], W=d.W, hK=d.hK)
    pv_perm, _ = crm_port.run_interaction(d.y, d.E, d.G[:, [10, 11]], W=d.W, hK=d.hK, idx_G=idx)
    assert np.all(pv_perm > pv)


def test_estimate_betas_shapes_and_sign(data):
    d = data
    bg, bgxe = crm_port.estimate_betas(d.y, d.W, d.E, d.G[:, [5, 10]], hK=d.hK)
    assert bg.shape == (2,) and bgxe.shape == (1, 600, 2)
    assert np.std(bgxe[0, :, 1]) > np.std(bgxe[0, :, 0])      # the GxC SNP has the larger per-cell effects


def test_host_helpers_match_the_oracle(data):
    from cellregmap_b200 import Term, compute_maf, get_L_values
    d = data
    np.testing.assert_allclose(compute_maf(d.G), crm_port.compute_maf(d.G))
    Xn = d.G.copy(); Xn[0, 0] = np.nan
    np.testing.assert_allclose(compute_maf(Xn), crm_port.compute_maf(Xn))
    Ls, Lr = get_L_values(d.hK, d.E), crm_port.get_L_values(d.hK, d.E)
    assert len(Ls) == len(Lr) == 6
    K1 = sum(L @ L.T for L in Ls); K2 = sum(L @ L.T for L in Lr)
    np.testing.assert_allclose(K1, K2, atol=1e-10)
    np.testing.assert_allclose(K1, (d.hK @ d.hK.T) * (d.E @ d.E.T), atol=1e-10)   # proof.md: K o EE' = sum_i L_i L_i'
    assert Term.FIXED.value == 1 and Term.RANDOM.value == 2


def test_synthetic_data_is_reproducible():
    a = make_data(n=200, donors=20, k=3, p=10, q=2, seed=9)
    b = make_data(n=200, donors=20, k=3, p=10, q=2, seed=9)
    np.testing.assert_array_equal(a.G, b.G)
    np.testing.assert_array_equal(a.y, b.y)
    assert a.G.flags["C_CONTIGUOUS"] and a.G.dtype == np.float64
    assert np.all(a.G.std(0) > 0)
