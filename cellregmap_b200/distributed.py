"""SNP-sharded scans over the GPUs of one box (one process per GPU, torch.distributed over NCCL / NVLink).

SNP columns are independent units of work (reference: serial loop cellregmap/_cellregmap.py:340), so the path shards with no
data-path collective: every rank holds the same per-gene state, scans its own contiguous block of SNP columns, and one
all-gather of the 5 per-SNP outputs (40 B/SNP) assembles the result on every rank.  Per-SNP arithmetic does not depend on
the sharding, so the result equals the 1-GPU result computed with the same basis bit for bit.

The per-gene set-up is shared as well (`run_interaction_sharded`, `CellRegMap(..., _group=...)`): rank r decomposes the grid
points r, r + world, ... of the rho1 grid and one all-gather of the packed grid points (S0 and T_rho) gives every rank the
whole basis -- the R eigendecompositions are the largest replicated cost of a strong-scaled scan.
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


def column_block(G, rank, world):
    """This rank's block of SNP columns of G as a view (no copy: the library walks strided host and device matrices)."""
    lo, hi = shard_range(int(G.shape[1]), rank, world)
    return G[:, lo:hi]


def scan_interaction_sharded(model, G, group=None, scan=None, block=None):
    """`model.scan_interaction` over this rank's block of the columns of G (the same G on every rank, or any
    object whose `[:, lo:hi]` slice yields the rank's columns), all-gathered.  Returns (pvalues, info) like the
    reference.  `scan` (tests) replaces the per-rank device scan: scan(G_block) -> tensor (5, p_local).
    `block`: the rank's column block when the caller has sliced it already (e.g. to start its transfer early)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    p = int(G.shape[1])
    if block is None:
        block = column_block(G, rank, world)
    if scan is None:
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


def run_interaction_sharded(y, E, G, W=None, E1=None, E2=None, hK=None, group=None, device=None, share_setup=True):
    """`run_interaction` (reference :547-587) over the ranks of `group`: called with the same arguments on every rank, returns the
    full (pvalues, info) on every rank.  The set-up is shared between the ranks (see the module docstring; `share_setup=False`
    replicates it), each rank moves and scans only its own block of SNP columns (host matrices: a strided view, no copy)."""
    from ._cellregmap import _make_interaction_model
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    block = column_block(G, rank, world)
    model = _make_interaction_model(y, E, W, E1, E2, hK, device=device, prefetch=block,
                                    group=(True if group is None else group) if (share_setup and world > 1) else None)
    return scan_interaction_sharded(model, G, group=group, block=block)
