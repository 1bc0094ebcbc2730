"""Times the int8 contraction of the exact int8 split at bench size through crm_int8_split_gemm: route 0 = fused tcgen05
kernel (oz_mma.cuh), route 1 = cuBLASLt int8 GEMM + recombination kernel."""
import ctypes, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from cellregmap_b200 import _lib
n, cols, B = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (100000, 21504, 10000)))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
X = torch.randn((n, cols), dtype=torch.float64, device=dev, generator=g)
G = torch.randint(0, 3, (n, B), device=dev, generator=g).to(torch.float64)
out = {}
res = {}
routes = (0, 0, 0) if len(sys.argv) > 4 else (0, 1, 0, 1)
for route in routes:
    C = torch.empty((B, cols), dtype=torch.float64, device=dev)
    flags = (ctypes.c_int32 * 2)(); ms = ctypes.c_float(0.0)
    _lib.call("crm_int8_split_gemm", ctypes.c_void_p(X.data_ptr()), cols, cols, ctypes.c_void_p(G.data_ptr()), B, B, n, route,
              ctypes.c_void_p(C.data_ptr()), cols, flags, ctypes.byref(ms), ctypes.c_void_p(0))
    ops = 2.0 * 8 * cols * float(n) * B
    out.setdefault(route, []).append({"ms": ms.value, "TOPs": ops / ms.value / 1e9})
    res[route] = C
    del C
print(json.dumps({"n": n, "cols": cols, "B": B, "fused_tcgen05": out[0], "cublaslt_plus_combine": out.get(1),
                  "bit_identical": bool(torch.equal(res[0], res[1])) if 1 in res else None}))
