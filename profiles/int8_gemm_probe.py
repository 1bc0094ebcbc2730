"""Probe: what an int8 tensor-core GEMM (cuBLASLt through torch._int_mm, tcgen05 on sm_100) delivers at the rotation's
shape -- the upper bound for an Ozaki-style exact int8 split of the fp64 rotation."""
import json, torch
dev = torch.device("cuda", 0)
res = []
for (M, K, N) in [(16384, 100096, 10240), (8192, 100096, 4096), (21504, 100096, 10240)]:
    a = torch.randint(-64, 64, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(0, 3, (N, K), dtype=torch.int8, device=dev)   # K-major B, passed as its transpose view
    bt = b.t()
    for _ in range(2): c = torch._int_mm(a, bt)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = torch._int_mm(a, bt); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    ref = (a[:4, :].double() @ b[:8, :].double().t())
    ok = bool((c[:4, :8].double() == ref).all())
    res.append({"M": M, "K": K, "N": N, "ms": best, "tops": 2.0 * M * K * N / best * 1e-9, "exact": ok})
    del a, b, c
print(json.dumps(res))
