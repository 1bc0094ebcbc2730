"""Parity at the BASELINE.json sizes: oracle comparison on a sample of SNPs where the oracle finishes in seconds,
and size-independent properties of the scan at full size (shard invariance, scale invariance of the score test,
cell-order invariance, donor-level == expanded)."""
import numpy as np
import pytest

from _parity import assert_variance_components
from cellregmap_b200.synth import make_data

pytestmark = pytest.mark.gpu
DLOG10_P = 1e-4


def _dlog10(a, b):
    """max |log10 a - log10 b| where either is positive; entries that underflow to exactly 0 in both count as equal."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_zero = (a == 0) & (b == 0)
    with np.errstate(divide="ignore"):
        d = np.abs(np.log10(a[~both_zero]) - np.log10(b[~both_zero]))
    return float(d.max()) if d.size else 0.0


@pytest.fixture(scope="module")
def cfg2():
    """BASELINE configs[1]: n = 10k cells, 200 donors, k = 20, low-rank hK (q = 10 -> m = 220), 2k SNPs."""
    return make_data(n=10000, donors=200, k=20, p=2000, q=10, seed=42)


def test_config2_oracle_sample(cuda_device, cfg2):
    from cellregmap_b200 import run_interaction
    from oracle import crm_port
    d = cfg2
    pv, info = run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
    assert pv.shape == (2000,) and np.all((pv > 0) & (pv <= 1))
    sample = np.r_[0:12, 1990:2000]
    # the reference's own decomposition route: thin SVD of the half-covariance per rho1 (economic_qs_linear); the product decomposes
    # the Gram of the shared half-basis instead
    ref_pv, ref_info = crm_port.run_interaction(d.y, d.E, d.G[:, sample], W=d.W, hK=d.hK, qs_method="svd")
    np.testing.assert_array_equal(info["rho1"][sample], ref_info["rho1"])
    assert np.max(np.abs(np.log10(pv[sample]) - np.log10(ref_pv))) <= DLOG10_P
    assert_variance_components({k: info[k][sample] for k in ("e2", "g2", "eps2")}, ref_info)
    # the simulated GxC SNPs are the strongest hits
    assert set(np.argsort(pv)[:2]) == {10, 11}


def test_config2_properties(cuda_device, cfg2):
    import torch
    from cellregmap_b200._cellregmap import _make_interaction_model
    d = cfg2
    model = _make_interaction_model(d.y, d.E, d.W, None, None, d.hK)
    Gt = torch.from_numpy(d.G).cuda()
    pv, info = model.scan_interaction(Gt)
    # 1. SNP shards (what each GPU of a multi-GPU run sees) reproduce the full scan bit for bit
    parts = [model.scan_interaction(Gt[:, lo:hi].contiguous())[0] for lo, hi in ((0, 700), (700, 701), (701, 2000))]
    np.testing.assert_array_equal(np.concatenate(parts), pv)
    # 2. the score test is invariant to the scale of the genotype column (Q and all eigenvalues scale together)
    pv2, info2 = model.scan_interaction(Gt * 2.0)
    np.testing.assert_array_equal(info2["rho1"], info["rho1"])
    assert np.max(np.abs(np.log10(pv2) - np.log10(pv))) <= 1e-6
    # 3. p-values are ordered like the reference would order them: finite, in (0, 1], no NaN
    assert np.isfinite(pv).all() and pv.min() > 0 and pv.max() <= 1


def test_cell_order_invariance(cuda_device, cfg2):
    """Permuting the cells (rows of every input) leaves every output unchanged up to summation order."""
    from cellregmap_b200 import run_interaction
    d = cfg2
    sub = slice(0, 300)
    pv, info = run_interaction(d.y, d.E, d.G[:, sub], W=d.W, hK=d.hK)
    perm = np.random.default_rng(0).permutation(d.y.shape[0])
    pv_p, info_p = run_interaction(d.y[perm], d.E[perm], np.ascontiguousarray(d.G[perm][:, sub]), W=d.W[perm], hK=d.hK[perm])
    np.testing.assert_array_equal(info_p["rho1"], info["rho1"])
    assert np.max(np.abs(np.log10(pv_p) - np.log10(pv))) <= 1e-6


def test_config3_shapes_sample(cuda_device):
    """BASELINE configs[2] shapes (n = 100k cells, 1,000 donors, k = 20, m = 1,020) on 256 SNPs: donor-level ingress equals
    the expanded scan, shards are bit-exact, and a handful of SNPs are checked against the oracle."""
    import torch
    from cellregmap_b200._cellregmap import _make_interaction_model
    from oracle import crm_port
    d = make_data(n=100000, donors=1000, k=20, p=256, q=50, seed=7, v_gxc=0.002, v_persistent=0.002)
    model = _make_interaction_model(d.y, d.E, d.W, None, None, d.hK)
    Gt = torch.from_numpy(d.G).cuda()
    pv, info = model.scan_interaction(Gt)
    pv_a, _ = model.scan_interaction(Gt[:, :100].contiguous())
    np.testing.assert_array_equal(pv_a, pv[:100])
    Gd = np.zeros((1000, 256))
    Gd[d.donor] = d.G
    pv_d, info_d = model.scan_interaction(Gd, donor_index=d.donor)
    np.testing.assert_array_equal(info_d["rho1"], info["rho1"])
    assert _dlog10(pv_d, pv) <= 1e-6
    sample = [0, 10, 11]
    ref_pv, ref_info = crm_port.run_interaction(d.y, d.E, d.G[:, sample], W=d.W, hK=d.hK, qs_method="gram")
    np.testing.assert_array_equal(info["rho1"][sample], ref_info["rho1"])
    assert _dlog10(pv[sample], ref_pv) <= DLOG10_P
    assert pv[10] < 1e-6 and pv[11] < 1e-6


def test_config4_association_sample(cuda_device):
    """BASELINE configs[3] family: run_association (LRT) at n = 50k on a sample of SNPs against the oracle."""
    from cellregmap_b200 import run_association
    from oracle import crm_port
    d = make_data(n=50000, donors=500, k=20, p=64, q=10, seed=11)
    pv, info = run_association(d.y, d.W, d.E, d.G, hK=d.hK)
    sample = [0, 5, 6, 63]
    # the reference's positional quirk makes W the context matrix: the background is rho W W' + (1-rho) hK hK'
    ref_pv, ref_info = crm_port.run_association(d.y, d.W, d.E, d.G[:, sample], hK=d.hK)
    np.testing.assert_array_equal(info["rho1"], ref_info["rho1"])
    assert np.max(np.abs(np.log10(pv[sample]) - np.log10(ref_pv))) <= DLOG10_P


def test_config5_betas_sample(cuda_device):
    """BASELINE configs[4] family: estimate_betas at n = 50k cells, k = 20 on two SNPs against the oracle (thin SVD of the
    50k x 120 half-covariance per SNP and rho1, as the reference does at :166-171).  beta_G is a GLS coefficient at the fitted
    delta: its sensitivity to where Brent stops (|d logit delta| <= 4e-6) is far below 1e-6; the GxC betas carry the factor
    v0 * rho1 and follow the variance-component rule (rtol 1e-6 of the largest entry)."""
    from cellregmap_b200 import estimate_betas
    from oracle import crm_port
    d = make_data(n=50000, donors=500, k=20, p=2, q=5, seed=23, causal_gxc=(0,), causal_persistent=(1,), v_gxc=0.01)
    ref_bg, ref_bgxe = crm_port.estimate_betas(d.y, d.W, d.E, d.G, hK=d.hK)
    bg, bgxe = estimate_betas(d.y, d.W, d.E, d.G, hK=d.hK)
    assert bg.shape == (2,) and bgxe.shape == (1, 50000, 2)
    assert np.abs(ref_bgxe[0, :, 0]).max() > 0                   # SNP 0 carries the simulated GxC effect: rho1 > 0 selected
    np.testing.assert_allclose(bg, ref_bg, rtol=1e-6)
    np.testing.assert_allclose(bgxe, ref_bgxe, rtol=0, atol=4e-6 * np.abs(ref_bgxe).max())


def test_wide_background_basis(cuda_device):
    """m = k + k q = 2 404 columns in the half-covariance: the per-rho vectors no longer fit the fit kernel's shared memory
    (global-memory path) and the spectrum is long; checked against the oracle on a few SNPs."""
    from cellregmap_b200 import run_interaction
    from oracle import crm_port
    d = make_data(n=3000, donors=700, k=4, p=12, q=600, seed=15)
    ref_pv, ref_info = crm_port.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK, qs_method="gram")
    pv, info = run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
    np.testing.assert_array_equal(info["rho1"], ref_info["rho1"])
    assert _dlog10(pv, ref_pv) <= DLOG10_P
    assert_variance_components(info, ref_info, max_fraction_off_path=0.25)


@pytest.mark.parametrize("case", ["sigma ratio 1e-7", "duplicated columns", "small scale", "tiny scale wide"])
def test_ill_conditioned_background(cuda_device, case):
    """Rank decisions of the set-up against the reference's economic_qs_linear (thin SVD of the half-covariance in the tall branch,
    no filtering -- cellregmap/_math.py:250-253; eigh with the absolute cut S >= sqrt(eps) in the wide branch, :221-235).
    The product decomposes the m x m Gram D^1/2 H'H D^1/2 and drops directions below 1e-12 lambda_max (the Gram's own noise floor;
    a direction with S ~ 0 contributes like the complement space, D0 = delta, whether it is kept or not)."""
    from cellregmap_b200 import run_interaction
    from oracle import crm_port
    d = make_data(n=900, donors=60, k=5, p=16, q=6, seed=51)
    E, hK, y = d.E.copy(), d.hK.copy(), d.y
    if case == "sigma ratio 1e-7":          # one background direction 1e-7 of the others: S ratio 1e-14, below the Gram route's floor
        hK[:, 5] = 1e-7 * hK[:, 5]
    elif case == "duplicated columns":      # exactly and nearly collinear background columns
        hK[:, 4] = hK[:, 3]
        hK[:, 5] = hK[:, 2] * (1 + 1e-9) + 1e-12 * hK[:, 1]
    elif case == "small scale":             # S0 of order 1e-5: nothing is cut in either route (a background far below sqrt(eps) in absolute
        E, hK = 1e-2 * E, 3e-2 * hK         # terms explains no variance, rho1 is then unidentifiable and no parity statement can be made)
    else:                                   # wide branch (m > n) at tiny scale: the absolute cut S >= sqrt(eps) empties the background
        dd = make_data(n=60, donors=30, k=4, p=8, q=20, seed=52)
        pv, info = run_interaction(dd.y, 1e-3 * dd.E, dd.G, W=dd.W, hK=1e-3 * dd.hK)
        ref_pv, ref_info = crm_port.run_interaction(dd.y, 1e-3 * dd.E, dd.G, W=dd.W, hK=1e-3 * dd.hK)
        np.testing.assert_array_equal(info["rho1"], ref_info["rho1"])
        assert _dlog10(pv, ref_pv) <= DLOG10_P
        return
    pv, info = run_interaction(y, E, d.G, W=d.W, hK=hK)
    ref_pv, ref_info = crm_port.run_interaction(y, E, d.G, W=d.W, hK=hK, qs_method="svd")
    np.testing.assert_array_equal(info["rho1"], ref_info["rho1"])
    assert _dlog10(pv, ref_pv) <= DLOG10_P
    assert_variance_components(info, ref_info, max_fraction_off_path=0.25)
