"""CRM_TRACE=1 python profiles/step_trace.py [--snps N]: per-phase device times of one device-resident run_interaction job at bench size."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellregmap_b200 import _cellregmap as api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--snps", type=int, default=10000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--keep", action="store_true", help="keep the model, set_phenotype per step")
ns = ap.parse_args()
sys.argv = [sys.argv[0], "--snps", str(ns.snps)]
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a)
Gd = bench.donor_genotypes(a, 0, a.snps)
y_d, W_d, E_d, hK_d = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
donor_d = torch.from_numpy(gene["donor"]).to(dev)
G_d = torch.from_numpy(Gd).to(dev)[donor_d].contiguous()
torch.cuda.synchronize()
keep = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev) if ns.keep else None
for rep in range(ns.reps):
    torch.cuda.synchronize(); t0 = time.time()
    if ns.keep:
        keep.set_phenotype(y_d + 0.01 * rep)
        model = keep
    else:
        model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
    torch.cuda.synchronize(); t1 = time.time()
    out = model._scan_interaction_device(G_d)
    torch.cuda.synchronize(); t2 = time.time()
    print(f"[step {rep}] model {1e3 * (t1 - t0):.1f} ms, scan {1e3 * (t2 - t1):.1f} ms", file=sys.stderr, flush=True)
