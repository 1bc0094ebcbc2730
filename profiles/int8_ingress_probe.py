import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import cellregmap_b200 as crm
from cellregmap_b200 import _cellregmap as api
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0)
G_d = torch.from_numpy(Gd).to(dev)[torch.from_numpy(gene["donor"]).to(dev)].contiguous()
G_h8 = torch.empty((a.cells, a.snps), dtype=torch.int8, pin_memory=True); G_h8.copy_(G_d); del G_d
y_h, W_h, E_h, hK_h = (torch.from_numpy(gene[k]).pin_memory() for k in ("y", "W", "E", "hK"))
out = {"pinned": G_h8.is_pinned(), "genotypes_ms": [], "run_ms": []}
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    g = api._Genotypes(G_h8, dev, a.cells); torch.cuda.synchronize()
    out["genotypes_ms"].append((time.time() - t0) * 1e3); del g
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    crm.run_interaction(y_h, E_h, G_h8, W=W_h, hK=hK_h); torch.cuda.synchronize()
    out["run_ms"].append((time.time() - t0) * 1e3)
print(json.dumps(out))
