"""GPU timeline of one bench step from CUPTI (torch.profiler): busy time, idle gaps and what surrounds them."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cellregmap_b200 import _cellregmap as api
from torch.profiler import profile, ProfilerActivity
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a); Gd = bench.donor_genotypes(a, 0)
y, W, E, hK = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
G = torch.from_numpy(Gd).to(dev)[torch.from_numpy(gene["donor"]).to(dev)].contiguous()
def step():
    model = api._make_interaction_model(y, E, W, None, None, hK, device=dev)
    out = model._scan_interaction_device(G)
    return torch.stack([out["pv"], out["rho1"]])
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
t0, t1 = iv[0][0], max(x[1] for x in iv)
busy = 0.0; cur_end = t0; gaps = []; prev = iv[0][2]
for s, e, nme in iv:
    if s > cur_end:
        gaps.append((s - cur_end, prev, nme)); busy += e - s; cur_end = e
    else:
        if e > cur_end: busy += e - cur_end; cur_end = e
    if e >= cur_end: prev = nme
agg = {}
for s, e, nme in iv: agg[nme[:60]] = agg.get(nme[:60], 0.0) + (e - s)
print(json.dumps({"span_ms": (t1 - t0) / 1e3, "busy_ms": busy / 1e3, "idle_ms": (t1 - t0 - busy) / 1e3, "n_kernels": len(iv),
                  "n_gaps_over_20us": sum(1 for g in gaps if g[0] > 20), "idle_in_gaps_over_20us_ms": sum(g[0] for g in gaps if g[0] > 20) / 1e3}))
for g in sorted(gaps, reverse=True)[:25]: print("gap %.3f ms after [%s] before [%s]" % (g[0] / 1e3, g[1][:50], g[2][:50]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:25]: print("%9.3f ms  %s" % (v / 1e3, k))
