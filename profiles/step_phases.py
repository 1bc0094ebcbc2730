"""Per-iteration wall times of a bench step: model construction (set-up), scan, release; then the un-synchronised loop."""
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
rows = []
for it in range(5):
    torch.cuda.synchronize(); t0 = time.time()
    model = api._make_interaction_model(y, E, W, None, None, hK, device=dev)
    t1h = time.time(); torch.cuda.synchronize(); t1 = time.time()
    out = model._scan_interaction_device(G)
    t2h = time.time(); torch.cuda.synchronize(); t2 = time.time()
    del model, out
    torch.cuda.synchronize(); t3 = time.time()
    rows.append({"setup_ms": (t1 - t0) * 1e3, "setup_host_ms": (t1h - t0) * 1e3, "scan_ms": (t2 - t1) * 1e3, "scan_host_ms": (t2h - t1) * 1e3, "free_ms": (t3 - t2) * 1e3})
print(json.dumps(rows))
def step():
    model = api._make_interaction_model(y, E, W, None, None, hK, device=dev)
    out = model._scan_interaction_device(G)
    return torch.stack([out["pv"], out["rho1"]])
step(); torch.cuda.synchronize(); t0 = time.time()
for _ in range(5): r = step()
torch.cuda.synchronize(); print("loop ms/step", (time.time() - t0) * 1e3 / 5)
