"""CUDA path (through the C ABI) against outputs of the REFERENCE'S OWN SOURCE.

Fixtures tests/golden/reference_source_*.npz were produced by running /root/reference/cellregmap unmodified over the
dependency stand-ins (tests/golden/make_reference_vectors.py); `cfg1` is BASELINE configs[0] on the reference's own generator
(`sample_phenotype_gxe(..., default_rng(20))`: 10 distinct context rows, background rank 328 < n = 500, degenerate spectra,
column-normalised genotypes).  When the byte-compiled reference travelled to this box (oracle/_ref, built by
oracle/build_ref.py) the same comparison is repeated live on fresh inputs.

Tolerances are BASELINE.json's: selected rho1 and SNP ranking exact; variance components rtol 1e-6; |dlog10 p| <= 1e-4 for p >= 1e-12."""
import glob
import os

import numpy as np
import pytest

from _parity import DLOG10_P, assert_pvalues, assert_variance_components

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN, "reference_source_*.npz")))


def _check(pv, info, ref_pv, ref_info, ranking=True):
    np.testing.assert_array_equal(info["rho1"], ref_info["rho1"])
    assert_variance_components(info, ref_info)
    assert_pvalues(pv, ref_pv, ranking)


def _ref_info(g, prefix):
    return {k: g[f"{prefix}_{k}"] for k in ("rho1", "e2", "g2", "eps2")}


@pytest.mark.parametrize("path", _fixtures(), ids=lambda f: os.path.basename(f)[17:-4])
def test_run_interaction_matches_reference_source(cuda_device, path):
    from cellregmap_b200 import run_interaction
    g = np.load(path)
    pv, info = run_interaction(g["y"], g["E"], g["G"], W=g["W"], hK=g["hK"])
    _check(pv, info, g["pv"], _ref_info(g, "info"))


@pytest.mark.parametrize("rotation", ["dmma", "auto"])
def test_config1_reference_generator_both_routes(cuda_device, rotation, monkeypatch):
    """configs[0] on the reference generator through the fp64 tensor-core route and the default route selection."""
    from cellregmap_b200 import run_interaction
    if rotation != "auto":
        monkeypatch.setenv("CRM_ROTATION", rotation)
    g = np.load(os.path.join(GOLDEN, "reference_source_cfg1.npz"))
    assert g["G"].shape == (500, 100) and g["hK"].shape == (500, 50) and g["E"].shape == (500, 10)
    pv, info = run_interaction(g["y"], g["E"], g["G"], W=g["W"], hK=g["hK"])
    _check(pv, info, g["pv"], _ref_info(g, "info"))
    assert pv[10] < 1e-4 and pv[11] < 1e-4          # the two simulated GxC SNPs (reference test recipe, test_struct_lmm2.py:23-24)


@pytest.mark.parametrize("path", _fixtures(), ids=lambda f: os.path.basename(f)[17:-4])
def test_association_scans_match_reference_source(cuda_device, path):
    from cellregmap_b200 import run_association, run_association_fast
    g = np.load(path)
    for fn, key in ((run_association, "assoc"), (run_association_fast, "assoc_fast")):
        pv, info = fn(g["y"], g["W"], g["E"], g["G"], hK=g["hK"])
        ref_pv = g[key + "_pv"]
        np.testing.assert_array_equal(info["rho1"], g[key + "_rho1"])
        assert info["rho1"].shape == (1,)
        assert_variance_components(info, {k: g[f"{key}_{k}"] for k in ("e2", "g2", "eps2")})
        big = ref_pv >= 1e-12
        assert np.max(np.abs(np.log10(pv[big]) - np.log10(ref_pv[big]))) <= DLOG10_P


@pytest.mark.parametrize("name", ["synth_a", "synth_b", "synth_std"])
def test_model_object_variants_match_reference_source(cuda_device, name):
    """idx_G of run_interaction (lands on idx_E), CellRegMap(..., hK=) with permuted tested genotypes, no background."""
    from cellregmap_b200 import CellRegMap, run_interaction
    g = np.load(os.path.join(GOLDEN, f"reference_source_{name}.npz"))
    perm = g["perm"]
    pv, info = run_interaction(g["y"], g["E"], g["G"], W=g["W"], hK=g["hK"], idx_G=perm)
    _check(pv, info, g["perm_pv"], _ref_info(g, "perm_info"))
    pv, info = CellRegMap(g["y"], g["E"], W=g["W"], hK=g["hK"]).scan_interaction(g["G"], idx_G=perm)
    _check(pv, info, g["ctor_hk_idxg_pv"], _ref_info(g, "ctor_hk_idxg"))
    pv, info = CellRegMap(g["y"], g["E"], W=g["W"]).scan_interaction(g["G"])
    _check(pv, info, g["nobg_pv"], _ref_info(g, "nobg"))
    assert np.all(info["rho1"] == 1.0)


@pytest.mark.parametrize("name", ["synth_a", "synth_b", "synth_std"])
def test_estimate_betas_matches_reference_source(cuda_device, name):
    from cellregmap_b200 import estimate_betas
    g = np.load(os.path.join(GOLDEN, f"reference_source_{name}.npz"))
    nb = g["beta_g"].shape[0]
    maf = g["beta_maf"] if name == "synth_std" else None
    bg, bgxe = estimate_betas(g["y"], g["W"], g["E"], g["G"][:, :nb], maf=maf, hK=g["hK"])
    assert bgxe.shape == g["beta_gxe"].shape == (1, g["y"].shape[0], nb)
    np.testing.assert_allclose(bg, g["beta_g"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(bgxe, g["beta_gxe"], rtol=0, atol=1e-5 * np.abs(g["beta_gxe"]).max())


def test_live_reference_source_on_fresh_inputs(cuda_device):
    """The byte-compiled reference (oracle/_ref) run here, on inputs no fixture holds."""
    from cellregmap_b200 import run_association, run_interaction
    from cellregmap_b200.synth import make_data
    from oracle import ref_shims
    ref = ref_shims.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref did not travel to this box")
    for cfg in (dict(n=450, donors=45, k=5, p=30, q=4, seed=1001), dict(n=380, donors=30, k=7, p=26, q=3, seed=1002, normalize_G=True)):
        d = make_data(**cfg)
        ref_pv, ref_info = ref.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
        pv, info = run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
        _check(pv, info, ref_pv, ref_info)
        ref_pa, ref_ia = ref.run_association(d.y, d.W, d.E, d.G, hK=d.hK)
        pa, ia = run_association(d.y, d.W, d.E, d.G, hK=d.hK)
        np.testing.assert_array_equal(ia["rho1"], ref_ia["rho1"])
        big = ref_pa >= 1e-12
        assert np.max(np.abs(np.log10(pa[big]) - np.log10(ref_pa[big]))) <= DLOG10_P
