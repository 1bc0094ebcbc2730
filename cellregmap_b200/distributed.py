"""SNP-sharded scans over the GPUs of one box (one process per GPU, torch.distributed).

SNP columns are independent units of work (reference: serial loop cellregmap/_cellregmap.py:340), so the path
shards with no data-path collective: every rank builds the same per-gene state, scans its own contiguous block
of SNP columns, and one all-gather of the 5 per-SNP outputs (40 B/SNP) over NCCL assembles the result on every
rank.  Per-SNP arithmetic does not depend on the sharding, so the result equals the 1-GPU result bit for bit.
"""
import numpy as np
import torch
import torch.distributed as dist

FIELDS = ("pv", "rho1", "e2", "g2", "eps2")


def shard_range(p, rank, world):
    """Contiguous block [lo, hi) of SNP columns owned by `rank`; the first p % world ranks get one more."""
    base, extra = divmod(int(p), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local, p, group=None):
    """All-gather of per-rank (5, p_local) result blocks into (5, p) on every rank (blocks may differ by one column)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    width = -(-int(p) // world)
    lo, hi = shard_range(p, rank, world)
    assert local.shape == (len(FIELDS), hi - lo), (tuple(local.shape), hi - lo)
    padded = torch.zeros((len(FIELDS), width), dtype=local.dtype, device=local.device)
    padded[:, : hi - lo] = local
    flat = torch.empty((world * len(FIELDS), width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, padded.contiguous(), group=group)
    out = flat.view(world, len(FIELDS), width)
    parts = []
    for r in range(world):
        a, b = shard_range(p, r, world)
        parts.append(out[r, :, : b - a])
    return torch.cat(parts, dim=1)


def scan_interaction_sharded(model, G, group=None, scan=None):
    """`model.scan_interaction` over this rank's block of the columns of G (the same G on every rank, or any
    object whose `[:, lo:hi]` slice yields the rank's columns), all-gathered.  Returns (pvalues, info) like the
    reference.  `scan` (tests) replaces the per-rank device scan: scan(G_block) -> tensor (5, p_local)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    p = int(G.shape[1])
    lo, hi = shard_range(p, rank, world)
    block = G[:, lo:hi]
    if scan is None:
        if isinstance(block, np.ndarray):
            block = np.ascontiguousarray(block)
        out = model._scan_interaction_device(block)
        flags = out["flags"]
        local = torch.stack([out[k] for k in FIELDS])
        bad = torch.tensor([int((flags & 1).any()), int((flags & 4).any())], device=local.device)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
        if int(bad[0]):
            raise RuntimeError("No eigenvalue is bigger than 0!!")
        if int(bad[1]):
            raise ValueError("The determinant of H should be positive.")
    else:
        local = scan(block)
    full = gather_results(local, p, group).cpu().numpy()
    return full[0], {k: full[i] for i, k in enumerate(FIELDS) if k != "pv"}
