import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cellregmap_b200 import _cellregmap as api
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0); gene = bench.add_causal_effects(gene, Gd, a)
y, W, E, hK = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
donor = torch.from_numpy(gene["donor"]).to(dev)
G = torch.from_numpy(Gd).to(dev)[donor].contiguous()
model = api._make_interaction_model(y, E, W, None, None, hK, device=dev)
o1 = model._scan_interaction_device(G, diagnostics=True)
o2 = model._scan_interaction_device(torch.from_numpy(Gd).to(dev), donor_index=donor, diagnostics=True)
for k in ("pv", "rho1", "e2", "eps2", "Q"):
    x, z = o1[k], o2[k]
    print(k, "nan dense", int(torch.isnan(x).sum()), "nan donor", int(torch.isnan(z).sum()), "maxrel", float(((x - z).abs() / x.abs().clamp_min(1e-300)).nan_to_num(0).max()))
bad = torch.isnan(o2["pv"]) | torch.isnan(o1["pv"])
print("bad idx", bad.nonzero().flatten()[:10].tolist(), "flags dense", o1["flags"][bad][:10].tolist(), "flags donor", o2["flags"][bad][:10].tolist())
i = bad.nonzero().flatten()
if len(i):
    j = int(i[0]); print("lam dense", o1["lam"][j].tolist(), o1["nlam"][j].item(), "lam donor", o2["lam"][j].tolist(), o2["nlam"][j].item(), "Q", o1["Q"][j].item(), o2["Q"][j].item(), "lml", o1["lml"][j].tolist(), o2["lml"][j].tolist())
print("pv min", float(o1["pv"].min()), float(o2["pv"].nan_to_num(1).min()))
