import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import cellregmap_b200 as crm
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0)
G_d = torch.from_numpy(Gd).to(dev)[torch.from_numpy(gene["donor"]).to(dev)].contiguous()
G_h = torch.empty((a.cells, a.snps), dtype=torch.float64, pin_memory=True); G_h.copy_(G_d); del G_d
y_h, W_h, E_h, hK_h = (torch.from_numpy(gene[k]).pin_memory() for k in ("y", "W", "E", "hK"))
for i in range(3):
    print("=== call", i, file=sys.stderr, flush=True)
    crm.run_interaction(y_h, E_h, G_h, W=W_h, hK=hK_h); torch.cuda.synchronize()
