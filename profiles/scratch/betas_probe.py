import numpy as np
from cellregmap_b200.synth import make_data
from cellregmap_b200 import estimate_betas
from oracle import crm_port
for onehot in (False, True):
    d = make_data(n=400, donors=40, k=5, p=12, q=4, seed=21)
    E = d.E
    if onehot:
        lab = np.random.default_rng(3).integers(0, 5, 400)
        E = np.eye(5)[lab]; E = (E - E.mean(0)) / E.std(0) / np.sqrt(5)
    rb, rx = crm_port.estimate_betas(d.y, d.W, E, d.G, hK=d.hK)
    b, x = estimate_betas(d.y, d.W, E, d.G, hK=d.hK)
    print(onehot, "bg rel", np.abs(b/rb-1), "bgxe", np.abs(x-rx).max(axis=(0,1))/np.abs(rx).max(axis=(0,1)))
