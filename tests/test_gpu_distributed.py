"""Sharded scan with the shared set-up on a real device: two processes (ranks) on cuda:0 over gloo -- NCCL needs one device per rank,
the bench and the driver's SCALE runs cover that -- against the single-process call."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from cellregmap_b200.distributed import run_interaction_sharded
        from cellregmap_b200.synth import make_data
        d = make_data(n=700, donors=50, k=6, p=101, q=5, seed=21)
        res = {}
        for name, G in (("pageable", d.G), ("int8", d.G.astype(np.int8)), ("device", torch.from_numpy(d.G).cuda())):
            pv, info = run_interaction_sharded(d.y, d.E, G, W=d.W, hK=d.hK)
            res[name] = np.concatenate([pv] + [info[k] for k in ("rho1", "e2", "g2", "eps2")])
        pv, info = run_interaction_sharded(d.y, d.E, d.G, W=d.W, hK=d.hK, share_setup=False)
        res["replicated"] = np.concatenate([pv] + [info[k] for k in ("rho1", "e2", "g2", "eps2")])
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_with_shared_setup(cuda_device, world):
    from cellregmap_b200 import run_interaction
    from cellregmap_b200.synth import make_data
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    d = make_data(n=700, donors=50, k=6, p=101, q=5, seed=21)
    pv, info = run_interaction(d.y, d.E, d.G, W=d.W, hK=d.hK)
    want = np.concatenate([pv] + [info[k] for k in ("rho1", "e2", "g2", "eps2")])
    for rank in range(world):
        for name, got in results[rank].items():
            # every rank returns the full result; SNP sharding changes no bit of a SNP's arithmetic, and neither does sharing the set-up
            # (a grid point's decomposition does not depend on which other grid points are decomposed with it: the batched solver
            # keeps the CTA-group size of the whole grid)
            np.testing.assert_array_equal(got, want, err_msg=f"rank {rank} {name}")
