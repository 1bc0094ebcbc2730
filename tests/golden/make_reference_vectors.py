"""Generates tests/golden/reference_source_*.npz by RUNNING THE REFERENCE'S OWN SOURCE in this container.

`/root/reference/cellregmap/{_cellregmap,_math,_simulate}.py` are imported unmodified (oracle/ref_shims.py) over stand-ins
for the third-party packages that are absent from the image (glimix_core, numpy_sugar, chiscore: oracle/lmm_port.py,
sugar_port.py, chiscore_port.py).  Inputs and outputs of every public entry point of the path are stored, so that

  * the oracle restatement `oracle/crm_port.py` can be checked against the reference's own logic anywhere
    (tests/test_reference_source.py, CPU), and
  * the CUDA path can be checked against the reference's outputs on the GPU box, where /root/reference does not exist
    (tests/test_gpu_reference_goldens.py, -m gpu).

Case `cfg1` is BASELINE configs[0] on the reference's own generator: `_simulate.sample_phenotype_gxe(...,
random=default_rng(20))` (reference :315-397; call recipe from cellregmap/test/test_struct_lmm2.py:15-24,217) --
500 cells, 50 donors, 10 one-hot context groups, 100 column-normalised SNPs; E has 10 distinct rows, the background has
rank 328 < n (wide branch of economic_qs_linear) with massively degenerate spectra.

    python tests/golden/make_reference_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from cellregmap_b200.synth import make_data  # noqa: E402
from oracle import ref_shims  # noqa: E402


def _info(prefix, info):
    return {f"{prefix}_{k}": np.asarray(v, float) for k, v in info.items()}


def case_cfg1(ref):
    sim = sys.modules["cellregmap._simulate"]
    s = sim.sample_phenotype_gxe(offset=0.3, n_individuals=50, n_snps=100, n_cells=10, n_env_groups=10, maf_min=0.05, maf_max=0.45,
                                 g_causals=[5, 6], gxe_causals=[10, 11], variances=sim.create_variances(r0=0.5, v0=0.5),
                                 random=np.random.default_rng(20))
    pv, info = ref.run_interaction(y=s.y, E=s.E, G=s.G, W=s.M, hK=s.Lk)
    out = dict(y=s.y, E=s.E, G=s.G, W=s.M, hK=s.Lk, pv=pv, **_info("info", info))
    # the association scans and a few effect sizes on the same data (configs[3], configs[4] in miniature)
    pa, ia = ref.run_association(s.y, s.M, s.E, s.G, hK=s.Lk)
    pf, jf = ref.run_association_fast(s.y, s.M, s.E, s.G, hK=s.Lk)
    out.update(assoc_pv=pa, assoc_fast_pv=pf, **_info("assoc", ia), **_info("assoc_fast", jf))
    return out


def case_synth(ref, **cfg):
    d = make_data(**cfg)
    out = dict(y=d.y, E=d.E, G=d.G, W=d.W, hK=d.hK)
    pv, info = ref.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
    out.update(pv=pv, **_info("info", info))
    rng = np.random.default_rng(5)
    perm = rng.permutation(d.y.shape[0])
    pvp, infop = ref.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK, idx_G=perm)        # lands on idx_E (reference :586)
    out.update(perm=perm, perm_pv=pvp, **_info("perm_info", infop))
    crm = ref.CellRegMap(d.y, d.E, W=d.W, hK=d.hK)                                      # direct hK background (:107-116)
    pvh, infoh = crm.scan_interaction(d.G, idx_G=perm)                                  # permuted tested genotypes (:410-413)
    out.update(ctor_hk_idxg_pv=pvh, **_info("ctor_hk_idxg", infoh))
    pv0, info0 = ref.CellRegMap(d.y, d.E, W=d.W).scan_interaction(d.G)                  # no background: rho1 = [1.0] (:103-106)
    out.update(nobg_pv=pv0, **_info("nobg", info0))
    pa, ia = ref.run_association(d.y, d.W, d.E, d.G, hK=d.hK)
    pf, jf = ref.run_association_fast(d.y, d.W, d.E, d.G, hK=d.hK)
    out.update(assoc_pv=pa, assoc_fast_pv=pf, **_info("assoc", ia), **_info("assoc_fast", jf))
    nb = min(5, d.G.shape[1])
    # maf from compute_maf (:678-679) for dosages; standardised columns have no allele frequency: the sampling one is passed
    maf = d.maf[:nb] if cfg.get("normalize_G") else None
    bg, bgxe = ref.estimate_betas(d.y, d.W, d.E, d.G[:, :nb], maf=maf, hK=d.hK)
    out.update(beta_g=bg, beta_gxe=bgxe, beta_maf=np.asarray(sys.modules["cellregmap._cellregmap"].compute_maf(d.G[:, :nb]) if maf is None else maf, float))
    return out


CASES = {
    "cfg1": lambda ref: case_cfg1(ref),
    "synth_a": lambda ref: case_synth(ref, n=400, donors=40, k=5, p=24, q=4, seed=7),
    "synth_b": lambda ref: case_synth(ref, n=350, donors=25, k=6, p=20, q=3, seed=123, n_covariates=3),
    "synth_std": lambda ref: case_synth(ref, n=300, donors=30, k=4, p=16, q=3, seed=11, normalize_G=True),
}


def main():
    ref = ref_shims.load_reference()
    assert ref is not None and ref.__oracle_root__ == ref_shims.REF_SOURCE, "run where /root/reference exists"
    for name, fn in CASES.items():
        out = fn(ref)
        path = os.path.join(HERE, f"reference_source_{name}.npz")
        np.savez_compressed(path, shimmed=np.asarray(ref_shims.is_shimmed()), **out)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
