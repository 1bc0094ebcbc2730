"""Is the box itself stalling?  A loop of small, identical, synchronised torch kernels (no code of this repo): prints the iterations
that took far longer than the median, with their time stamps.   python profiles/stall_probe.py [seconds]"""
import sys
import time

import numpy as np
import torch

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
a = torch.randn(2048, 2048, device="cuda", dtype=torch.float32)
torch.cuda.synchronize()
t_start = time.time()
stamps, durs = [], []
while time.time() - t_start < secs:
    t0 = time.time()
    b = a @ a                      # ~0.3 ms
    c = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")     # allocator traffic like a model object's
    del c
    torch.cuda.synchronize()
    durs.append(time.time() - t0); stamps.append(t0 - t_start)
durs = np.array(durs); stamps = np.array(stamps)
med = np.median(durs)
out = np.where(durs > max(10 * med, 5e-3))[0]
print(f"{len(durs)} iterations, median {1e3 * med:.3f} ms, p99 {1e3 * np.percentile(durs, 99):.3f} ms, max {1e3 * durs.max():.1f} ms; {len(out)} stalls > {1e3 * max(10 * med, 5e-3):.1f} ms")
for i in out[:40]:
    print(f"   t = {stamps[i]:7.3f} s   {1e3 * durs[i]:8.1f} ms")
