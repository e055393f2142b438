"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards:
  * the O(n^2) ordered-pair thal matrix (run_ntthal, od-msspe/src/delta_g.rs:61-153): pairs are independent, so
    the matrix is tiled in contiguous row blocks, one per rank; every rank computes its block with
    msspe_cross_dimer and the compacted results (pairs below the dG limit, structure-less pairs) are exchanged
    with one all_gather -- the only collective on this path.  Results are identical to a single-GPU run
    because the lists are keyed by the global pair index and merged in rank (= row) order.
  * independent genome sets (one design job per GPU): no collective at all (bench.py weak scaling).
The exact genome-sharded greedy loop (k-mer owner re-shard + per-iteration 256-byte all_gather) is described in
DESIGN.md as the next step; it is not implemented in this round.
"""
from __future__ import annotations

import numpy as np


def row_block(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous rows [begin, end) of rank `rank`; blocks differ by at most one row."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _all_gather_varlen(arr: np.ndarray, dist, device):
    """all_gather of a 1-D byte-viewable array whose length differs per rank."""
    import torch
    world = dist.get_world_size()
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    n = torch.tensor([raw.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw.copy()).to(device)
    outs = [torch.zeros(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [o[:s].cpu().numpy() for o, s in zip(outs, sizes)]


def cross_dimer_sharded(compute_rows, n: int, edge_dtype, dist=None, device="cpu"):
    """Row-tiled all-pairs dimer evaluation.

    compute_rows(row_begin, row_end) -> (edges[edge_dtype sorted by pair], nostruct[uint64 sorted]) for the
    rank's rows (on a GPU rank: Engine.cross_dimer).  Returns the merged lists, identical on every rank and
    identical to compute_rows(0, n).
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return compute_rows(0, n)
    rank, world = dist.get_rank(), dist.get_world_size()
    rb, re_ = row_block(n, rank, world)
    edges, nos = compute_rows(rb, re_)
    parts_e = _all_gather_varlen(np.ascontiguousarray(edges), dist, device)
    parts_n = _all_gather_varlen(np.ascontiguousarray(nos, dtype=np.uint64), dist, device)
    all_e = np.concatenate([p.view(edge_dtype) for p in parts_e]) if parts_e else np.zeros(0, dtype=edge_dtype)
    all_n = np.concatenate([p.view(np.uint64) for p in parts_n]) if parts_n else np.zeros(0, dtype=np.uint64)
    # row blocks are ordered by rank and each list is sorted inside its block -> already globally sorted
    return all_e, all_n
