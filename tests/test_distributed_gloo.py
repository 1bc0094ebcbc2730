"""N > 1 host logic (SNP sharding + the all-gather of per-SNP results) on CPU with gloo, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cellregmap_b200.distributed import FIELDS, gather_results, scan_interaction_sharded, shard_range


def test_shard_range_partitions():
    for p in (0, 1, 7, 10, 10000, 10001):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(p, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == p
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _fake_scan(block):
    # deterministic function of the genotype column only (what SNP sharding must preserve)
    g = torch.as_tensor(np.asarray(block), dtype=torch.float64)
    s = g.sum(0)
    return torch.stack([torch.sigmoid(s * 1e-2), s * 0 + 0.3, s * 1e-3, s * 2e-3, (g * g).sum(0)])


def _worker(rank, world, port, p, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = np.random.default_rng(5).integers(0, 3, (50, p)).astype(float)
        pv, info = scan_interaction_sharded(None, G, scan=_fake_scan)
        want = _fake_scan(G).numpy()
        ok = np.array_equal(pv, want[0]) and all(np.array_equal(info[k], want[i]) for i, k in enumerate(FIELDS) if k != "pv")
        lo, hi = shard_range(p, rank, world)
        local = _fake_scan(G[:, lo:hi])
        ok = ok and np.array_equal(gather_results(local, p).numpy(), want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p", [(2, 101), (2, 8), (3, 10)])
def test_sharded_scan_equals_unsharded(world, p):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]
