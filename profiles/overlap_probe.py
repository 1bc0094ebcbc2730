"""Do the cuSOLVER eigendecompositions overlap with a large int8 tensor-core GEMM on another stream?"""
import json, time, torch
dev = torch.device("cuda", 0)
A = torch.randn(1020, 1020, dtype=torch.float64, device=dev); A = A @ A.T
a8 = torch.randint(-64, 64, (8 * 21504, 100096), dtype=torch.int8, device=dev)
b8 = torch.randint(0, 3, (10240, 100096), dtype=torch.int8, device=dev)
def eigs():
    for _ in range(11): torch.linalg.eigh(A)
def gemm():
    return torch._int_mm(a8, b8.t())
for _ in range(2): eigs(); gemm()
torch.cuda.synchronize()
def timeit(fn):
    torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); return 1e3 * (time.time() - t0)
t_e = timeit(eigs); t_g = timeit(gemm)
s1 = torch.cuda.Stream(priority=-1); s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s2): gemm()
    with torch.cuda.stream(s1): eigs()
t_b = timeit(both)
def both2():
    with torch.cuda.stream(s1): eigs()
    with torch.cuda.stream(s2): gemm()
t_b2 = timeit(both2)
print(json.dumps({"eig11_ms": t_e, "int8_gemm_ms": t_g, "concurrent_gemm_first_ms": t_b, "concurrent_eig_first_ms": t_b2}))
