"""Host feeder alone: float64 -> int8 conversion rate of crm_host_narrow on a (cells x SNPs) block, by thread count.
    python profiles/host_narrow_bench.py            (CRM_NARROW_SCALAR=1: the scalar loop)"""
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from cellregmap_b200 import _lib
    lib = _lib.load()
    n, p, b = 100000, 6000, 1792
    G = np.random.default_rng(0).integers(0, 3, (n, p)).astype(np.float64)
    out = np.empty((n, b), dtype=np.int8)
    bad, gmax = ctypes.c_int32(), ctypes.c_int32()
    best = 1e9
    for rep in range(5):
        c0 = 1000 * rep
        t0 = time.time()
        _lib.call("crm_host_narrow", ctypes.c_void_p(G[:, c0:].ctypes.data), 0, p, n, b, ctypes.c_void_p(out.ctypes.data), b, ctypes.byref(bad), ctypes.byref(gmax))
        best = min(best, time.time() - t0)
    print(f"threads {lib.crm_host_threads():3d} scalar={os.environ.get('CRM_NARROW_SCALAR', '0')}: {1e3 * best:7.1f} ms per {n} x {b} block, {n * b * 8 / best * 1e-9:6.1f} GB/s read")
else:
    print("cpus", os.cpu_count(), flush=True)
    for scalar in ("0", "1"):
        for threads in (1, 4, 8, 16):
            env = dict(os.environ, CRM_HOST_THREADS=str(threads), CRM_NARROW_SCALAR=scalar)
            subprocess.run([sys.executable, __file__, "child"], env=env, check=True)
