"""Host-to-device rate of strided column-block copies (what staging a row-major n x p matrix needs) vs one contiguous copy."""
import json, sys, time
import torch
n, p = 100000, 10000
G = torch.empty((n, p), dtype=torch.float64, pin_memory=True); G.fill_(1.0)
D = torch.empty((n, p), dtype=torch.float64, device="cuda")
out = {}
def timed(fn):
    torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); return (time.time() - t0) * 1e3
out["contiguous_ms"] = timed(lambda: D.copy_(G, non_blocking=True))
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
def copy2d(width_cols, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    rq = (n + streams - 1) // streams
    def run():
        for c0 in range(0, p, width_cols):
            w = min(width_cols, p - c0)
            for q, s in enumerate(ss):
                r0, r1 = min(n, rq * q), min(n, rq * (q + 1))
                if r1 > r0:
                    rt.cudaMemcpy2DAsync(ctypes.c_void_p(D.data_ptr() + (r0 * p + c0) * 8), ctypes.c_size_t(p * 8), ctypes.c_void_p(G.data_ptr() + (r0 * p + c0) * 8),
                                         ctypes.c_size_t(p * 8), ctypes.c_size_t(w * 8), ctypes.c_size_t(r1 - r0), 1, ctypes.c_void_p(s.cuda_stream))
    return timed(run)
for w in (512, 1024, 2048, 3456):
    for st in (1, 2, 4, 8):
        out["cols%d_streams%d_ms" % (w, st)] = copy2d(w, st)
print(json.dumps(out))
