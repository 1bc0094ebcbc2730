#!/bin/bash
# Does the one-time initialisation of cusolverDnXsyevBatched persist across processes on one box (JIT cache)?
for i in 1 2; do
  echo "process $i"; CRM_EIG_BATCHED=1 python - <<'PY'
import time, sys, numpy as np, torch
sys.path.insert(0, ".")
from cellregmap_b200.synth import make_data
from cellregmap_b200._cellregmap import _make_interaction_model
d = make_data(n=3000, donors=100, k=20, p=8, q=50, seed=1)
for it in range(3):
    t0 = time.time(); m = _make_interaction_model(d.y, d.E, d.W, None, None, d.hK); torch.cuda.synchronize(); print("setup %d: %.1f ms" % (it, 1e3 * (time.time() - t0)))
PY
done
ls -la ~/.nv/ComputeCache 2>/dev/null | head -3
