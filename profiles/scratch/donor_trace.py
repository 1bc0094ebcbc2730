"""CRM_TRACE=1 python profiles/scratch/donor_trace.py: phases of the donor-level arm at bench size."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from cellregmap_b200 import _cellregmap as api
sys.argv = [sys.argv[0]]
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a)
Gd = bench.donor_genotypes(a, 0, a.snps)
y_d, W_d, E_d, hK_d = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
donor_d = torch.from_numpy(gene["donor"]).to(dev)
Gdon_d = torch.from_numpy(Gd).to(dev)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
    torch.cuda.synchronize(); t1 = time.time()
    out = model._scan_interaction_device(Gdon_d, donor_index=donor_d)
    torch.cuda.synchronize(); t2 = time.time()
    print(f"[step {rep}] model {1e3 * (t1 - t0):.1f} ms, scan {1e3 * (t2 - t1):.1f} ms", file=sys.stderr, flush=True)
