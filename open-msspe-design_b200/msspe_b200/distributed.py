"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards:
  * the O(n^2) ordered-pair thal matrix (run_ntthal, od-msspe/src/delta_g.rs:61-153): pairs are independent, so
    the matrix is tiled in contiguous row blocks, one per rank; every rank computes its block with
    msspe_cross_dimer and the compacted results (pairs below the dG limit, structure-less pairs) are exchanged
    with one all_gather -- the only collective on this path.  Results are identical to a single-GPU run
    because the lists are keyed by the global pair index and merged in rank (= row) order.
  * independent genome sets (one design job per GPU): no collective at all (bench.py weak scaling).
  * the greedy selection of ONE job whose genomes are block-partitioned over the ranks (select_sharded): every rank
    recounts its own segments, the per-k-mer counts are merged with an all-reduce(sum) over NVLink, every rank then
    takes the same arg-max; the order-sensitive f32 tie score is rebuilt from the per-rank positions of the first
    live posting of each partition (rank order = global segment order).  Bit-identical to the single-GPU loop.
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


def cross_dimer_sharded_tensors(compute_rows_t, n: int, edge_dtype, dist=None, device="cuda"):
    """Row-tiled all-pairs dimer evaluation with the per-rank lists kept on the device (delta_g.rs:83-153).

    compute_rows_t(row_begin, row_end) -> (edges int64 [m, 2] = (pair, dG bit pattern), nostruct int64 [q]) as tensors on
    `device`, unsorted (on a GPU rank: Engine.cross_dimer_device, which leaves them in ctx-owned device buffers).
    Collectives: ONE all_gather_into_tensor of the two counts (the only host read before the end) and ONE
    all_gather_into_tensor of the packed lists; the merged lists are sorted by pair index on the device and come back
    with one device-to-host copy each.  Result identical on every rank and to a single-rank run."""
    import torch
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    rb, re_ = row_block(n, rank, world)
    e, q = compute_rows_t(rb, re_)
    if world > 1:
        sz = torch.tensor([e.shape[0], q.shape[0]], dtype=torch.int64, device=device)
        all_sz = torch.empty(2 * world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(all_sz, sz)
        all_sz = all_sz.cpu().view(world, 2)
        cap_e, cap_q = max(1, int(all_sz[:, 0].max())), max(1, int(all_sz[:, 1].max()))
        buf = torch.empty(2 * cap_e + cap_q, dtype=torch.int64, device=device)
        buf[:2 * e.shape[0]] = e.reshape(-1)
        buf[2 * cap_e:2 * cap_e + q.shape[0]] = q
        out = torch.empty(world * buf.numel(), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(out, buf)
        out = out.view(world, -1)
        e = torch.cat([out[r, :2 * int(all_sz[r, 0])].view(-1, 2) for r in range(world)])
        q = torch.cat([out[r, 2 * cap_e:2 * cap_e + int(all_sz[r, 1])] for r in range(world)])
    if e.shape[0]:
        e = e[torch.argsort(e[:, 0])]
    q = torch.sort(q).values
    edges = e.contiguous().cpu().numpy().view(edge_dtype).reshape(-1)
    return edges, q.cpu().numpy().view(np.uint64)


def partition_range(n_part: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous partitions [begin, end) of rank `rank` (the column shard of the multi-GPU loop)."""
    return row_block(n_part, rank, world)


def column_shard(genomes: np.ndarray, window: int, step: int, rank: int, world: int) -> np.ndarray:
    """The columns of this rank's partitions for EVERY genome of an equal-length alignment (n x L uint8): partition p =
    columns [p * step, p * step + window) (main.rs:173-181), so a contiguous partition range is a contiguous column
    range and the rank's windows are exactly the global windows of its partitions."""
    n, L = genomes.shape
    n_part = (L - window) // step + 1 if L >= window else 0
    p0, p1 = partition_range(n_part, rank, world)
    if p1 <= p0:
        return np.zeros((n, 0), dtype=np.uint8)
    return np.ascontiguousarray(genomes[:, p0 * step:(p1 - 1) * step + window])


NO_LOCAL_ID = 0xFFFFFFFF


def _all_reduce(t, dist, op=None):
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM)
    return t


def _all_gather_same(t, dist):
    """all_gather of equally shaped tensors -> stacked [world, ...]."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return t.unsqueeze(0)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t.contiguous())
    return torch.stack(outs)


def select_sharded(se, direction: int, max_iterations: int, max_mismatch_segments: int, dist=None, device="cuda"):
    """find_candidates_kmers (od-msspe/src/main.rs:331-406) for one direction over genome shards.

    `se` is this rank's shard engine (msspe_b200.Engine after load_genomes/build_index on the rank's contiguous
    block of records, or any object with the same shard_* methods).  Returns (candidates, evals, iterations):
    candidates as a CANDIDATE_DTYPE array identical on every rank and identical to Engine.select on the whole input.
    """
    import torch
    from . import CANDIDATE_DTYPE
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    se.shard_begin(direction)
    codes = se.shard_codes(direction).to(device)
    n_local = int(codes.numel())
    # ---- global k-mer dictionary: union of the per-rank sorted code lists ----
    if world > 1:
        sizes = _all_gather_same(torch.tensor([n_local], dtype=torch.int64, device=device), dist).flatten().tolist()
        cap = max(max(sizes), 1)
        padded = torch.full((cap,), -1, dtype=torch.int64, device=device)
        padded[:n_local] = codes
        allc = _all_gather_same(padded, dist)
        union = torch.cat([allc[r, :sizes[r]] for r in range(world)])
        ug = torch.unique(union, sorted=True)
    else:
        ug = codes.clone()
    n_global = int(ug.numel())
    l2g = torch.searchsorted(ug, codes) if n_local else torch.zeros(0, dtype=torch.int64, device=device)
    g2l = torch.full((max(n_global, 1),), -1, dtype=torch.int64, device=device)
    if n_local:
        g2l[l2g] = torch.arange(n_local, dtype=torch.int64, device=device)
    n_part = int(_all_reduce(torch.tensor([se.shard_n_part()], dtype=torch.int64, device=device), dist,
                             dist.ReduceOp.MAX if world > 1 else None).item())
    cov = np.zeros(max(n_part, 1), dtype=np.uint32)      # partition_coverage, replicated
    out, evals, iterations = [], 0, 0
    one = np.float32(1.0)
    for _ in range(max_iterations):
        freq, _live = se.shard_count(direction)
        g = torch.zeros(max(n_global, 1), dtype=torch.int32, device=device)
        if n_local:
            g[l2g] = freq.to(device)
        _all_reduce(g, dist)                               # "per-window counts merged by NCCL reduce over NVLink"
        iterations += 1
        evals += int(g.sum(dtype=torch.int64).item())
        gmax = int(g.max().item()) if n_global else 0
        if gmax <= 1:                                      # None / freq == 1: stop before the push (main.rs:353-366)
            break
        tied = torch.nonzero(g == gmax).flatten()          # ascending global id == ascending word
        lids = g2l[tied]
        lids_np = np.where(lids.cpu().numpy() < 0, NO_LOCAL_ID, lids.cpu().numpy()).astype(np.uint32)
        fp = se.shard_firstpos(direction, lids_np, n_part)                              # [n_tied, n_part] u32
        fpw = _all_gather_same(torch.from_numpy(fp.view(np.int32).copy()).to(device), dist).cpu().numpy().view(np.uint32)
        has = fpw != np.uint32(0xFFFFFFFF)                                               # [world, n_tied, n_part]
        anyp = has.any(axis=0)
        first_rank = has.argmax(axis=0)
        pos = np.take_along_axis(fpw, first_rank[None], axis=0)[0]
        key = np.where(anyp, (first_rank.astype(np.uint64) << np.uint64(32)) | pos.astype(np.uint64), np.uint64(0xFFFFFFFFFFFFFFFF))
        order = np.argsort(key, axis=1, kind="stable")
        term = (one / (cov[:n_part].astype(np.float32) + one)).astype(np.float32)       # 1.0 / (already_covered as f32 + 1.0)
        terms = np.where(anyp, term[None, :], np.float32(0.0)).astype(np.float32)
        ts = np.take_along_axis(terms, order, axis=1)
        score = np.zeros(len(tied), dtype=np.float32)
        for t in range(int(anyp.sum(axis=1).max()) if len(tied) else 0):                 # sequential f32 adds, first-seen order
            score = (score + ts[:, t]).astype(np.float32)
        w = int(np.argmax(score))                                                        # first maximum = smallest word
        gid = int(tied[w].item())
        out.append((int(ug[gid].item()), gmax, len(tied), float(score[w]), 0))
        lid = int(g2l[gid].item())
        flags = se.shard_apply(direction, lid if lid >= 0 else NO_LOCAL_ID, n_part)
        ft = _all_reduce(torch.from_numpy(flags.astype(np.int32)).to(device), dist, dist.ReduceOp.MAX if world > 1 else None)
        cov[:n_part] += ft.cpu().numpy().astype(np.uint32)
        if gmax < max_mismatch_segments or len(out) >= max_iterations:                    # main.rs:387-390, :344
            break
    res = np.zeros(len(out), dtype=CANDIDATE_DTYPE)
    for i, (c, f, nt, sc, _) in enumerate(out):
        res[i] = (c, f, nt, sc, 0)
    return res, evals, iterations
