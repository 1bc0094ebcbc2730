"""Where does an occasional slow model construction lose its time?  Times the pieces of _make_interaction_model (host mirror of
get_L_values, then CellRegMap.__init__ = finiteness read-back + crm_create + crm_setup) over many repetitions at bench size."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellregmap_b200 import _cellregmap as api  # noqa: E402

sys.argv = [sys.argv[0]]
a = bench.parse_args()
dev = torch.device("cuda", 0)
gene = bench.make_gene(a)
y_d, W_d, E_d, hK_d = (torch.from_numpy(gene[k]).to(dev) for k in ("y", "W", "E", "hK"))
sync = torch.cuda.synchronize
rows = []
for rep in range(60):
    sync(); t0 = time.time()
    R = torch.linalg.qr(E_d, mode="r").R
    sync(); t1 = time.time()
    Rh = R.cpu().numpy()
    _, S, Vh = np.linalg.svd(Rh)
    t2 = time.time()
    V = torch.from_numpy(np.ascontiguousarray(Vh.T)).to(dev)
    us = E_d @ V
    sync(); t3 = time.time()
    Ls = (us[:, :, None] * hK_d[:, None, :]).reshape(E_d.shape[0], -1).contiguous()
    sync(); t4 = time.time()
    model = api.CellRegMap(y=y_d, E=E_d, W=W_d, E1=E_d, Ls=Ls, device=dev)
    sync(); t5 = time.time()
    del model
    sync(); t6 = time.time()
    rows.append([1e3 * (b - a_) for a_, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6))])
rows = np.array(rows[3:])
names = ["qr", "R.cpu + svd", "V upload + E@V", "L blocks", "CellRegMap()", "del"]
print("median ms:", dict(zip(names, np.round(np.median(rows, 0), 2))))
print("max ms:   ", dict(zip(names, np.round(rows.max(0), 2))))
tot = rows.sum(1)
for i in np.where(tot > 1.5 * np.median(tot))[0]:
    print("slow rep", i + 3, dict(zip(names, np.round(rows[i], 1))))
