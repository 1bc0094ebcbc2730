"""CUPTI timeline of one run_interaction call from pinned host buffers (staged genotype transfer): copies vs kernels."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
import cellregmap_b200 as crm
from torch.profiler import profile, ProfilerActivity
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0)
G_d = torch.from_numpy(Gd).to(dev)[torch.from_numpy(gene["donor"]).to(dev)].contiguous()
G_h = torch.empty((a.cells, a.snps), dtype=torch.float64, pin_memory=True); G_h.copy_(G_d); del G_d
y_h, W_h, E_h, hK_h = (torch.from_numpy(gene[k]).pin_memory() for k in ("y", "W", "E", "hK"))
for _ in range(2): crm.run_interaction(y_h, E_h, G_h, W=W_h, hK=hK_h)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    crm.run_interaction(y_h, E_h, G_h, W=W_h, hK=hK_h); torch.cuda.synchronize()
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA))
t0 = ev[0][0]
last_copy_end = 0
rows = []
for s, e, nme in ev:
    big = e - s > 1000
    if "Memcpy HtoD" in nme and e - s > 500:
        last_copy_end = e; ncopy = globals().get("ncopy", 0) + 1; globals()["ncopy"] = ncopy
        if ncopy % 4 != 1: continue
    if big: rows.append("%9.2f -> %9.2f ms  %s" % ((s - t0) / 1e3, (e - t0) / 1e3, nme[:70]))
print("last big HtoD copy ends at %.2f ms; total span %.2f ms" % ((last_copy_end - t0) / 1e3, (ev[-1][1] - t0) / 1e3))
print("\n".join(rows[:80]))
