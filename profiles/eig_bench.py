"""Times the batched set-up eigensolver (crm_eigh_batched) on 9 + 2 matrices of the bench shape against torch.linalg.eigh."""
import ctypes, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cellregmap_b200 import _lib
n, batch = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (1020, 11)))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
X = torch.randn((batch, 4 * n, n), dtype=torch.float64, device=dev, generator=g)
A = torch.matmul(X.transpose(1, 2), X) / n
A = (A + A.transpose(1, 2)) / 2
W = torch.empty((batch, n), dtype=torch.float64, device=dev); V = torch.empty((batch, n, n), dtype=torch.float64, device=dev)
out = []
for it in range(3):
    q = (ctypes.c_double * batch)(); ms = ctypes.c_float(0.0)
    _lib.call("crm_eigh_batched", ctypes.c_void_p(A.data_ptr()), n, batch, ctypes.c_void_p(W.data_ptr()), ctypes.c_void_p(V.data_ptr()), q, ctypes.byref(ms), ctypes.c_void_p(0))
    out.append(ms.value)
Vt = V.transpose(1, 2)
res = (torch.matmul(A, Vt) - Vt * W[:, None, :]).abs().max().item() / A.abs().max().item()
orth = (torch.matmul(V, Vt) - torch.eye(n, device=dev, dtype=torch.float64)).abs().max().item()
torch.cuda.synchronize(); t0 = time.time()
for b in range(batch): torch.linalg.eigh(A[b])
torch.cuda.synchronize(); t1 = time.time()
print(json.dumps({"n": n, "batch": batch, "native_ms": out, "quality": max(q), "residual": res, "orthogonality": orth, "torch_eigh_sequential_ms": (t1 - t0) * 1e3}))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    q = (ctypes.c_double * batch)(); ms = ctypes.c_float(0.0)
    _lib.call("crm_eigh_batched", ctypes.c_void_p(A.data_ptr()), n, batch, ctypes.c_void_p(W.data_ptr()), ctypes.c_void_p(V.data_ptr()), q, ctypes.byref(ms), ctypes.c_void_p(0))
    torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:70]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print("%5d %9.3f ms  %s" % (c, t / 1e3, k))
