"""Measured FP64 GEMM peak on this GPU (cuBLAS via torch.matmul) -- the denominator for the FP64
tensor roofline, measured the same way MEASURED_PEAKS.json measures bf16 (best of 10, CUDA events)."""
import json, sys, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(6):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# sustained: back to back for ~3 s
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = max(3, int(3000 / best)); e0.record()
for _ in range(reps): c = a @ b
e1.record(); torch.cuda.synchronize(); sus = e0.elapsed_time(e1) / reps
print(json.dumps({"n": n, "fp64_tflops_burst": 2 * n**3 / best * 1e-9, "fp64_tflops_sustained": 2 * n**3 / sus * 1e-9, "gpu": torch.cuda.get_device_name(0)}))
