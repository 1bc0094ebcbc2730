"""Wall/event timing of the phases of one run_interaction job at the bench workload (set-up vs scan)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cellregmap_b200 import _cellregmap as api
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0)
y, W, E, hK = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
G = torch.from_numpy(Gd).to(dev)[torch.from_numpy(gene["donor"]).to(dev)].contiguous()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.time(); r = fn(); torch.cuda.synchronize(); ts.append(time.time() - t0)
    return min(ts) * 1e3, r
ms_L, L = t(lambda: api._L_concat(hK, E))
ms_setup, model = t(lambda: api.CellRegMap(y=y, E=E, W=W, E1=E, Ls=L, device=dev))
ms_scan, _ = t(lambda: model._scan_interaction_device(G))
ms_eigh, _ = t(lambda: torch.linalg.eigh(torch.randn(11, 1020, 1020, dtype=torch.float64, device=dev).pow(2).cumsum(1) @ torch.eye(1020, dtype=torch.float64, device=dev)))
A = torch.randn(1020, 1020, dtype=torch.float64, device=dev); A = A @ A.T
ms_eigh1, _ = t(lambda: torch.linalg.eigh(A))
print(json.dumps({"L_concat_ms": ms_L, "setup_ms": ms_setup, "scan_ms": ms_scan, "torch_eigh_1020_ms": ms_eigh1, "snps": a.snps}))
