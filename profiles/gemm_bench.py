"""Times K1 <EXPAND> alone (CUDA events, best of 5) at the bench shapes: K = n cells, M = m + 2, `snps` SNPs x (1 + k) columns."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from cellregmap_b200 import _lib
n, m, k, snps = 100000, 1020, 20, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
Mx = m + 2
A = torch.randn(n, Mx, dtype=torch.float64, device=dev, generator=g)
G = torch.randint(0, 3, (n, snps), device=dev, generator=g).double()
kexp = k + 1
pitch = (kexp + 2) & ~1
while pitch % 16 not in (4, 12): pitch += 2
E = torch.zeros(n, pitch, dtype=torch.float64, device=dev); E[:, 0] = 1.0; E[:, 1:kexp] = torch.randn(n, k, dtype=torch.float64, device=dev, generator=g)
out = torch.empty(snps * kexp, Mx + (Mx & 1), dtype=torch.float64, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
def run():
    _lib.call("crm_gemm", 2, p(A), Mx, Mx, p(G), snps, snps, p(E), pitch, pitch, n, 0, Mx, 0, snps * kexp, p(out), out.shape[1], kexp, ctypes.c_void_p(0))
run(); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
flop_alg = 2.0 * n * m * kexp * snps
# spot check against torch
s = 7; ref = torch.cat([G[:, s:s+1], G[:, s:s+1] * E[:, 1:kexp]], 1).T @ A
err = float((out[s*kexp:(s+1)*kexp, :Mx] - ref).abs().max() / ref.abs().max())
print(json.dumps({"mt": os.environ.get("CRM_GEMM_MT", "default"), "snps": snps, "ms": best, "tflops_alg": flop_alg / best * 1e-9, "frac_of_37.1": flop_alg / best * 1e-9 / 37.1, "rel_err": err}))
