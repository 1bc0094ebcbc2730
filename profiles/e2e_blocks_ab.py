"""A/B of the feeder block width inside one process (box-to-box variance exceeds the effect): run_interaction from pageable numpy
genotypes at bench size, alternating CRM_FEEDER_BLOCK settings.   python profiles/e2e_blocks_ab.py 1792 2560 3584 0"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cellregmap_b200 as crm  # noqa: E402

widths = [int(x) for x in sys.argv[1:]] or [1792, 3584, 0]
sys.argv = [sys.argv[0]]
a = bench.parse_args()
gene = bench.make_gene(a)
Gd = bench.donor_genotypes(a, 0, a.snps)
G = np.ascontiguousarray(Gd[gene["donor"]])
times = {w: [] for w in widths}
for rnd in range(5):
    for w in widths:
        if w:
            os.environ["CRM_FEEDER_BLOCK"] = str(w)
        else:
            os.environ.pop("CRM_FEEDER_BLOCK", None)
        torch.cuda.synchronize(); t0 = time.time()
        crm.run_interaction(gene["y"], gene["E"], G, W=gene["W"], hK=gene["hK"])
        if rnd:
            times[w].append(1e3 * (time.time() - t0))
for w in widths:
    print(f"block {w or 'default':>7}: median {np.median(times[w]):7.1f} ms   all {np.round(times[w], 1)}")
