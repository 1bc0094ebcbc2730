#!/bin/bash
# First-call cost of cusolverDnXsyevBatched (CRM_EIG_BATCHED=1) at the bench shapes, in two consecutive processes.
for i in 1 2; do
  echo "process $i"; CRM_EIG_BATCHED=1 python - <<'PY'
import time, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from cellregmap_b200._cellregmap import _make_interaction_model
a = bench.parse_args()
gene = bench.make_gene(a)
dev = torch.device("cuda", 0)
y, W, E, hK = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time(); m = _make_interaction_model(y, E, W, None, None, hK, device=dev); torch.cuda.synchronize(); print("setup %d: %.1f ms" % (it, 1e3 * (time.time() - t0)), flush=True)
PY
done
