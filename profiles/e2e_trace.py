"""CRM_TRACE=1 python profiles/e2e_trace.py: timeline of run_interaction from pageable numpy genotypes at bench size (two calls)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cellregmap_b200 as crm  # noqa: E402

sys.argv = [sys.argv[0]]
a = bench.parse_args()
gene = bench.make_gene(a)
Gd = bench.donor_genotypes(a, 0, a.snps)
G = np.ascontiguousarray(Gd[gene["donor"]])
for rep in range(int(os.environ.get("REPS", "3"))):
    torch.cuda.synchronize(); t0 = time.time()
    pv, info = crm.run_interaction(gene["y"], gene["E"], G, W=gene["W"], hK=gene["hK"])
    print(f"[call {rep}] {1e3 * (time.time() - t0):.1f} ms", file=sys.stderr, flush=True)
