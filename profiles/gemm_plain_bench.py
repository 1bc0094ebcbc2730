"""Times K1 <PLAIN> alone (no Hadamard in the loop): C[N][M] = B^T A, K = 100000."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, ".")
from cellregmap_b200 import _lib
K, M, N = 100000, int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
A = torch.randn(K, M, dtype=torch.float64, device=dev); B = torch.randn(K, N, dtype=torch.float64, device=dev)
out = torch.empty(N, M, dtype=torch.float64, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
def run():
    _lib.call("crm_gemm", 0, p(A), M, M, p(B), N, N, ctypes.c_void_p(0), 0, 0, K, 0, M, 0, N, p(out), M, 1, ctypes.c_void_p(0))
run(); torch.cuda.synchronize()
best = 1e9
for _ in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
ref = B[:, :64].T @ A[:, :64]
err = float((out[:64, :64] - ref).abs().max() / ref.abs().max())
ctas = ((M + 127) // 128) * ((N + 127) // 128)
print(json.dumps({"mode": "plain", "mt": os.environ.get("CRM_GEMM_MT", "default"), "M": M, "N": N, "ctas": ctas, "waves": ctas / 148, "ms": best, "tflops": 2.0 * K * M * N / best * 1e-9, "rel_err": err}))
