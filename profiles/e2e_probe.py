"""e2e of run_interaction from pinned host buffers: staged-ahead transfer vs block streaming (CRM_NO_STAGE=1), wall ms per call."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import cellregmap_b200 as crm
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0)
G_d = torch.from_numpy(Gd).to(dev)[torch.from_numpy(gene["donor"]).to(dev)].contiguous()
G_h = torch.empty((a.cells, a.snps), dtype=torch.float64, pin_memory=True); G_h.copy_(G_d)
y_h, W_h, E_h, hK_h = (torch.from_numpy(gene[k]).pin_memory() for k in ("y", "W", "E", "hK"))
y_d, W_d, E_d, hK_d = (t.to(dev) for t in (y_h, W_h, E_h, hK_h))
def run(G, *args):
    torch.cuda.synchronize(); t0 = time.time()
    pv, info = crm.run_interaction(args[0], args[2], G, W=args[1], hK=args[3])
    torch.cuda.synchronize(); return (time.time() - t0) * 1e3
out = {"device_resident": [run(G_d, y_d, W_d, E_d, hK_d) for _ in range(4)]}
out["host_" + ("streamed" if os.environ.get("CRM_NO_STAGE") == "1" else "staged")] = [run(G_h, y_h, W_h, E_h, hK_h) for _ in range(4)]
print(json.dumps(out))
