"""Pins the oracle restatement `oracle/crm_port.py` (+ math_port) to the REFERENCE'S OWN SOURCE.

The reference's in-tree code (cellregmap/_cellregmap.py, _math.py, _simulate.py) runs unmodified over stand-ins for its
absent third-party dependencies (oracle/ref_shims.py): from /root/reference in this container, from the byte-compiled
oracle/_ref on the GPU box.  Every public entry point of the path must agree with the restatement bit for bit -- both sit on
the same dependency stand-ins, so any difference is a difference in the restated logic (grid, strict-'>' selection,
permutation hooks, projection algebra, positional quirks, output shapes).  The committed fixtures
tests/golden/reference_source_*.npz (made by tests/golden/make_reference_vectors.py) carry the same outputs to machines
without the reference."""
import glob
import os
import sys

import numpy as np
import pytest

from cellregmap_b200.synth import make_data
from oracle import crm_port, ref_shims

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    mod = ref_shims.load_reference()
    if mod is None:
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    return mod


def _same(a, b):
    pa, ia = a
    pb, ib = b
    np.testing.assert_array_equal(np.asarray(pa), np.asarray(pb))
    assert set(ia) == set(ib)
    for key in ia:
        np.testing.assert_array_equal(np.asarray(ia[key]), np.asarray(ib[key]))


@pytest.mark.parametrize("cfg", [dict(n=240, donors=24, k=4, p=10, q=3, seed=3), dict(n=180, donors=15, k=3, p=8, q=2, seed=41, n_covariates=2),
                                 dict(n=200, donors=20, k=4, p=8, q=3, seed=9, normalize_G=True)])
def test_port_equals_reference_source_interaction(ref, cfg):
    d = make_data(**cfg)
    _same(crm_port.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK), ref.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK))
    perm = np.random.default_rng(1).permutation(d.y.shape[0])
    _same(crm_port.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK, idx_G=perm), ref.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK, idx_G=perm))
    # E1 / E2 different from E; no covariates given
    E1, E2 = d.E[:, :2], d.E[:, 1:]
    _same(crm_port.run_interaction(d.y, d.E, d.G, E1=E1, E2=E2, hK=d.hK), ref.run_interaction(d.y, d.E, d.G, E1=E1, E2=E2, hK=d.hK))
    # model object: hK background, both permutation hooks; no background at all
    _same(crm_port.CellRegMapOracle(d.y, d.E, W=d.W, hK=d.hK).scan_interaction(d.G, idx_E=perm[::-1], idx_G=perm),
          ref.CellRegMap(d.y, d.E, W=d.W, hK=d.hK).scan_interaction(d.G, idx_E=perm[::-1], idx_G=perm))
    _same(crm_port.CellRegMapOracle(d.y, d.E, W=d.W).scan_interaction(d.G), ref.CellRegMap(d.y, d.E, W=d.W).scan_interaction(d.G))


def test_port_equals_reference_source_association_and_betas(ref):
    d = make_data(n=220, donors=22, k=4, p=9, q=3, seed=17)
    _same(crm_port.run_association(d.y, d.W, d.E, d.G, hK=d.hK), ref.run_association(d.y, d.W, d.E, d.G, hK=d.hK))
    _same(crm_port.run_association_fast(d.y, d.W, d.E, d.G, hK=d.hK), ref.run_association_fast(d.y, d.W, d.E, d.G, hK=d.hK))
    _same(crm_port.run_association(d.y, d.W, d.E, d.G), ref.run_association(d.y, d.W, d.E, d.G))
    bg0, bx0 = crm_port.estimate_betas(d.y, d.W, d.E, d.G[:, :3], hK=d.hK)
    bg1, bx1 = ref.estimate_betas(d.y, d.W, d.E, d.G[:, :3], hK=d.hK)
    assert bx1.shape == (1, d.y.shape[0], 3) and bx0.shape == bx1.shape
    np.testing.assert_array_equal(bg0, bg1)
    np.testing.assert_array_equal(bx0, bx1)
    refmod = sys.modules["cellregmap._cellregmap"]
    np.testing.assert_array_equal(crm_port.compute_maf(d.G), refmod.compute_maf(d.G))
    np.testing.assert_array_equal(crm_port.lrt_pvalues(-10.0, [-9.0, -3.0, -10.0, -11.0], dof=1), refmod.lrt_pvalues(-10.0, [-9.0, -3.0, -10.0, -11.0], dof=1))
    np.testing.assert_array_equal(crm_port.lrt_pvalues(-10.0, [-9.0, -3.0], dof=3), refmod.lrt_pvalues(-10.0, [-9.0, -3.0], dof=3))
    for a, b in zip(crm_port.get_L_values(d.hK, d.E), refmod.get_L_values(d.hK, d.E)):
        np.testing.assert_array_equal(a, b)


def test_port_equals_reference_source_on_the_reference_generator(ref):
    """configs[0] in miniature on the reference's own simulator (degenerate spectra, wide branch, column-normalised G)."""
    sim = sys.modules["cellregmap._simulate"]
    s = sim.sample_phenotype_gxe(offset=0.3, n_individuals=12, n_snps=14, n_cells=6, n_env_groups=4, maf_min=0.05, maf_max=0.45,
                                 g_causals=[5, 6], gxe_causals=[10, 11], variances=sim.create_variances(r0=0.5, v0=0.5),
                                 random=np.random.default_rng(20))
    _same(crm_port.run_interaction(y=s.y, E=s.E, G=s.G, W=s.M, hK=s.Lk), ref.run_interaction(y=s.y, E=s.E, G=s.G, W=s.M, hK=s.Lk))


def test_math_port_equals_reference_source(ref):
    """QSCov / PMat / ScoreStatistic of the reference (_math.py:40-128) against oracle/math_port.py on a random structured case."""
    from oracle import math_port as mp
    rmath = sys.modules["cellregmap._math"]
    rng = np.random.default_rng(0)
    n, r, k = 40, 7, 3
    Q0, _ = np.linalg.qr(rng.standard_normal((n, r)))
    S0 = rng.uniform(0.1, 3.0, r)
    W = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = rng.standard_normal(n)
    gE = rng.standard_normal((n, k))
    qscov = rmath.QSCov(Q0, S0, 0.7, 0.4)
    P = rmath.PMat(qscov, W)
    ss = rmath.ScoreStatistic(P, qscov, gE)
    Pm = mp.Projection(Q0, S0, 0.7, 0.4, W)
    assert ss.statistic(y) == mp.score_statistic_structured(Pm, gE, y)
    np.testing.assert_array_equal(ss.matrix_for_dist_weights(), mp.weight_matrix_structured(Pm, gE))
    np.testing.assert_array_equal(qscov.solve(gE), mp.qscov_solve(Q0, S0, 0.7, 0.4, gE))
    np.testing.assert_array_equal(qscov.dot(gE), mp.qscov_dot(Q0, S0, 0.7, 0.4, gE))


# ------------------------------------------------------------------------------------------------------------------
# committed fixtures (outputs of the reference source in the builder's container) still hold for the restatement
# ------------------------------------------------------------------------------------------------------------------
def _fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN, "reference_source_*.npz")))


def test_fixtures_are_committed():
    names = {os.path.basename(f) for f in _fixtures()}
    assert {"reference_source_cfg1.npz", "reference_source_synth_a.npz", "reference_source_synth_b.npz", "reference_source_synth_std.npz"} <= names


@pytest.mark.parametrize("path", _fixtures(), ids=lambda f: os.path.basename(f)[17:-4])
def test_port_reproduces_reference_fixture(path):
    """Same arithmetic on a possibly different BLAS thread count: selection exact, values to round-off of the optimiser path."""
    g = np.load(path)
    pv, info = crm_port.run_interaction(g["y"], g["E"], g["G"], W=g["W"], hK=g["hK"])
    np.testing.assert_array_equal(info["rho1"], g["info_rho1"])
    np.testing.assert_allclose(pv, g["pv"], rtol=1e-7)
    for key in ("e2", "g2", "eps2"):
        np.testing.assert_allclose(info[key], g["info_" + key], rtol=1e-7, atol=1e-12)
    assert np.array_equal(np.argsort(pv, kind="stable"), np.argsort(g["pv"], kind="stable"))
    pa, ia = crm_port.run_association(g["y"], g["W"], g["E"], g["G"], hK=g["hK"])
    np.testing.assert_array_equal(ia["rho1"], g["assoc_rho1"])
    np.testing.assert_allclose(pa, g["assoc_pv"], rtol=1e-6)
