"""Generates tests/golden/oracle_small.npz: stage-wise outputs of the ORACLE (not of the reference, which cannot run in
this image) on a small seeded problem.  They freeze the oracle's behaviour, so that a change to the oracle or to the CUDA
path shows up as a diff against a committed file.    python tests/golden/make_oracle_vectors.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cellregmap_b200.synth import make_data  # noqa: E402
from oracle import crm_port  # noqa: E402

CONFIG = dict(n=300, donors=30, k=4, p=16, q=3, seed=77)


def main():
    d = make_data(**CONFIG)
    stages = {}
    pv, info = crm_port.run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK, stages=stages)
    pa, ia = crm_port.run_association(d.y, d.W, d.E, d.G, hK=d.hK)
    pf, _ = crm_port.run_association_fast(d.y, d.W, d.E, d.G, hK=d.hK)
    bg, bgxe = crm_port.estimate_betas(d.y, d.W, d.E, d.G[:, :4], hK=d.hK)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz")
    np.savez_compressed(
        out, y=d.y, W=d.W, E=d.E, G=d.G, hK=d.hK, pv=pv, rho1=info["rho1"], e2=info["e2"], g2=info["g2"], eps2=info["eps2"],
        lml=np.array([[f[1] for f in fits] for fits in stages["fits"]]), Q=np.array(stages["Q"]), M=np.array(stages["M"]),
        assoc_pv=pa, assoc_rho1=ia["rho1"], assoc_fast_pv=pf, beta_g=bg, beta_gxe=bgxe)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
