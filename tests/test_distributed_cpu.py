"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row tiling of the ordered-pair thal matrix and the
all_gather merge.  The per-rank compute is stood in for by the oracle (tests may use it); on GPU ranks it is
Engine.cross_dimer -- the merge logic under test is the same."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_compute(words, limit):
    from oracle import oracle as O
    import msspe_b200 as m
    cond = O.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    n = len(words)

    def compute(rb, re_):
        e, nos = [], []
        for i in range(rb, re_):
            for j in range(n):
                o = O.thal(words[i], words[j], 1, cond)
                if o.no_structure:
                    nos.append(i * n + j)
                elif o.dg < limit:
                    e.append((i * n + j, o.dg))
        return np.array(e, dtype=m.EDGE_DTYPE), np.array(nos, dtype=np.uint64)
    return compute


def _worker(rank, world, port, words, limit, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
    import torch.distributed as dist
    import msspe_b200 as m
    from msspe_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e, nos = D.cross_dimer_sharded(_make_compute(words, limit), len(words), m.EDGE_DTYPE, dist, "cpu")
    # the device-list path (Engine.cross_dimer_device on a GPU rank): unsorted (pair, dG bits) tensors, one all_gather of the
    # counts, one of the packed lists, sorted after the merge
    import torch

    def compute_t(rb, re_):
        ee, nn = _make_compute(words, limit)(rb, re_)
        perm = np.random.default_rng(rank).permutation(len(ee))
        t = torch.from_numpy(np.ascontiguousarray(ee[perm]).view(np.int64).reshape(-1, 2).copy())
        return t, torch.from_numpy(nn[::-1].copy().view(np.int64))
    e2, nos2 = D.cross_dimer_sharded_tensors(compute_t, len(words), m.EDGE_DTYPE, dist, "cpu")
    assert e2.tobytes() == e.tobytes() and nos2.tobytes() == nos.tobytes()
    q.put((rank, e.tobytes(), nos.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_row_blocks_partition_the_matrix():
    sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
    from msspe_b200 import distributed as D
    for n in (0, 1, 7, 64, 20000):
        for world in (1, 2, 3, 8):
            blocks = [D.row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_sharded_cross_dimer_equals_single_process(oracle_lib):
    sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
    import msspe_b200 as m
    from msspe_b200 import synth
    k, n = 13, 21   # odd row count: ragged blocks
    words = [m.decode_word(c, k) for c in synth.random_primers(n - 2, k, 4)] + ["AACCACACACCAA", "CACACAACCACAC"]
    limit = -1500.0
    want_e, want_n = _make_compute(words, limit)(0, n)
    assert len(want_e) > 5 and len(want_n) >= 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, words, limit, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, eb, nb in got:
        assert eb == want_e.tobytes() and nb == want_n.tobytes(), "rank %d merged result differs" % rank


# ---------------------------------------------------------------------------------------------------------------
# genome-sharded greedy selection: two gloo ranks, each with half of the records, must reproduce the single-process
# reference loop (winners, frequencies, tie counts, f32 scores bit for bit).
def _sharded_worker(rank, world, port, fasta, params, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from msspe_b200 import distributed as D
    from oracle import kmer_oracle as ko
    from _cpu_shard import CpuShard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, S, w, k, max_iter, mms = params
    recs = ko.to_records(fasta)
    lo, hi = D.row_block(len(recs), rank, world)
    se = CpuShard(recs[lo:hi], W, S, w, k)
    res = []
    for d in (0, 1):
        cand, evals, iters = D.select_sharded(se, d, max_iter, mms, dist, "cpu")
        res.append((cand.tobytes(), evals, iters))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _reference_select(fasta, params):
    from oracle import kmer_oracle as ko
    import msspe_b200 as m
    W, S, w, k, max_iter, mms = params
    recs = ko.to_records(fasta)
    segs = ko.get_segment_manager(recs, W, S, w, k)
    out = []
    for d in (0, 1):
        tr = []
        ko.find_candidates_kmers(segs, d, max_iter, mms, tr)
        a = np.zeros(len(tr), dtype=m.CANDIDATE_DTYPE)
        for i, (wd, f, nt, sc) in enumerate(tr):
            a[i] = (ko.encode(wd), f, nt, sc, 0)
        out.append(a.tobytes())
    return out


@pytest.mark.timeout(600)
@pytest.mark.parametrize("case", ["zika_first30", "random_ties", "one_rank_empty"])
def test_sharded_selection_equals_reference(case, zika_fasta):
    sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
    from oracle import kmer_oracle as ko
    if case == "zika_first30":
        recs = ko.to_records(zika_fasta)[:30]
        fasta = "".join(">%s\n%s\n" % (r.name, r.sequence[:4000]) for r in recs).encode()
        params = (500, 250, 50, 13, 40, 1)
    elif case == "random_ties":
        rng = np.random.default_rng(5)
        anc = rng.integers(0, 4, 900)
        lines = []
        for i in range(14):
            s = anc.copy()
            mut = rng.random(900) < 0.02
            s[mut] = rng.integers(0, 4, int(mut.sum()))
            lines.append(">r%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
        fasta = "".join(lines).encode()
        params = (100, 50, 20, 7, 30, 1)
    else:  # the second rank's records are too short for a single window: it owns no segment at all
        rng = np.random.default_rng(6)
        anc = rng.integers(0, 4, 600)
        lines = [">a%d\n%s\n" % (i, "".join("ACGT"[x] for x in anc)) for i in range(3)] + [">s%d\nACGTACGT\n" % i for i in range(3)]
        fasta = "".join(lines).encode()
        params = (100, 50, 20, 7, 10, 1)
    want = _reference_select(fasta, params)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, fasta, params, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in got:
        for d in (0, 1):
            assert res[d][0] == want[d], "rank %d direction %d differs from the reference loop" % (rank, d)
    assert got[0][1][0][1] == got[1][1][0][1]  # evals agree across ranks
