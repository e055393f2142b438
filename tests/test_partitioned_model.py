"""The per-partition decomposition of the greedy loop (MSSPE_SELECT_PARTITIONED, csrc/select_part.cu) proved on the CPU:
tests/_partitioned_model.py runs the algorithm of the CUDA kernels step for step (unit sequences -> merge -> check of the
multi-partition lists -> external winner + roll-back) and must reproduce the oracle's loop (od-msspe/src/main.rs:331-406)
bit for bit -- winners, frequencies, tie counts, f32 tie scores, reference-equivalent evals -- on the Zika fixture, the
reference's unit vector, random alignments (incl. k = 5, where EVERY list spans several partitions and every iteration
is an external winner), a tie storm, ragged record lengths, every stop rule and several look-ahead chunk sizes."""
import numpy as np
import pytest

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _partitioned_model as pm  # noqa: E402


def _random_alignment(seed, n, L, p=0.03, gaps=True):
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, L)
    out = []
    for i in range(n):
        s = anc.copy()
        mut = rng.random(L) < p
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        ch = np.array(list("ACGT"))[s]
        if gaps:
            ch[rng.random(L) < 0.004] = "-"
            ch[rng.random(L) < 0.002] = "N"
        out.append(">g%d x y\n%s\n" % (i, "".join(ch)))
    return "".join(out).encode()


def _check(O, fa, W, S, w, k, max_iter, mms, **kw):
    stats = []
    for d in (0, 1):
        slots, part = O.segment_slots(fa, W, S, w, k, d)
        want = O.select(fa, W, S, w, k, d, max_iter, mms)
        got = pm.select(slots, part, max_iter, mms, **kw)
        assert got["codes"] == want["codes"].tolist(), "direction %d" % d
        assert got["freqs"] == want["freqs"].tolist()
        assert got["n_tied"] == want["n_tied"].tolist()
        assert got["score_bits"] == want["scores"].view(np.uint32).tolist()
        assert got["evals"] == want["evals"]
        stats.append((got["rounds"], got["rollbacks"]))
    return stats


def test_reference_unit_vector(oracle_lib):
    fa = b">seq1\nAACCTTGGAACCTTGG\n>seq2\nAACCTTGGAACCTTG-\n>seq3\n-ACCTTGGAACCTT-G\n"
    _check(oracle_lib, fa, 10, 5, 5, 3, 10, 1)


@pytest.mark.parametrize("chunk0,chunk", [(None, 4), (1, 1), (3, 2), (1000, 1000)])
def test_zika_fixture(oracle_lib, zika_fasta, chunk0, chunk):
    st = _check(oracle_lib, zika_fasta, 500, 250, 50, 13, 1000, 2, chunk0=chunk0, chunk=chunk)
    assert all(r >= 1 for r, _ in st)


def test_zika_stop_rules(oracle_lib, zika_fasta):
    _check(oracle_lib, zika_fasta, 500, 250, 50, 13, 7, 2)       # max_iterations binds
    _check(oracle_lib, zika_fasta, 500, 250, 50, 13, 1000, 94)   # frequency threshold: pushed, then break
    _check(oracle_lib, zika_fasta, 500, 250, 50, 13, 0, 2)
    _check(oracle_lib, zika_fasta, 500, 250, 50, 13, 1, 2)


@pytest.mark.parametrize("seed,n,L,W,S,w,k,mms", [
    (0, 12, 700, 100, 50, 20, 7, 1),
    (1, 40, 2000, 200, 100, 60, 9, 1),
    (2, 30, 1501, 500, 250, 50, 13, 1),
    (3, 64, 1000, 128, 64, 64, 5, 2),       # tiny k: every list spans several partitions
    (4, 20, 1200, 300, 150, 100, 31, 1),
    (5, 25, 900, 90, 45, 45, 17, 1),
    (6, 30, 1200, 100, 50, 30, 6, 3),       # a mix of single- and multi-partition lists with many external winners
])
def test_random_alignments(oracle_lib, seed, n, L, W, S, w, k, mms):
    fa = _random_alignment(seed, n, L)
    _check(oracle_lib, fa, W, S, w, k, 60, mms)
    _check(oracle_lib, fa, W, S, w, k, 60, mms, chunk0=1, chunk=1)


def test_tie_storm_and_ragged(oracle_lib):
    _check(oracle_lib, _random_alignment(7, 50, 3000, p=0.0, gaps=False), 500, 250, 50, 13, 40, 1)
    rng = np.random.default_rng(21)
    anc = rng.integers(0, 4, 2600)
    lines = []
    for i, L in enumerate([2600, 1800, 2600, 999, 1200, 2599, 500, 499, 2600]):
        s = anc[:L].copy()
        mut = rng.random(L) < 0.02
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        lines.append(">x%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    _check(oracle_lib, "".join(lines).encode(), 500, 250, 50, 13, 60, 1)
    _check(oracle_lib, b">a\n" + b"-" * 1200 + b"\n>b\n" + b"ACGT" * 20 + b"\n", 500, 250, 50, 13, 10, 1)


def test_repeats_couple_partitions(oracle_lib):
    """A genome family with a repeated block: the same words sit in two partitions of every genome, so the strongest
    lists are multi-partition and win early -- external winners, roll-backs and partition_coverage of already covered
    postings (main.rs:371-378) all in play."""
    rng = np.random.default_rng(5)
    anc = rng.integers(0, 4, 3000)
    anc[2000:2300] = anc[500:800]
    out = []
    for i in range(40):
        s = anc.copy()
        mut = rng.random(3000) < 0.03
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        out.append(">r%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    fa = "".join(out).encode()
    st = _check(oracle_lib, fa, 500, 250, 50, 13, 200, 1)
    assert sum(rb for _, rb in st) > 0
    _check(oracle_lib, fa, 500, 250, 50, 13, 200, 3, chunk0=2, chunk=1)


# ---- the multi-GPU protocol (csrc/select_dist.cu) on two gloo ranks -------------------------------------------------------
def _dist_worker(rank, world, port, fasta, params, q):
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import torch.distributed as dist
    import _partitioned_model as pm2
    from oracle import oracle as O, kmer_oracle as ko
    _os.environ["MASTER_ADDR"] = "127.0.0.1"
    _os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Comm:
        def all_gather(self, obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

    W, S, w, k, max_iter, mms, kw = params
    recs = ko.to_records(fasta)
    L = min(len(r.sequence) for r in recs)
    n_part = (L - W) // S + 1
    base, rem = divmod(n_part, world)
    p0 = rank * base + min(rank, rem)
    p1 = p0 + base + (1 if rank < rem else 0)
    local = "".join(">%s\n%s\n" % (r.name, r.sequence[p0 * S:(p1 - 1) * S + W] if p1 > p0 else "") for r in recs).encode()
    res = []
    for d in (0, 1):
        slots, part = O.segment_slots(local, W, S, w, k, d)
        m = pm2.DistributedPartitionedSelect(slots, part, p1 - p0, Comm(), rank, world, max_iter, mms, **kw).run()
        res.append(dict(codes=[o[0] for o in m.out], freqs=[o[1] for o in m.out], n_tied=[o[2] for o in m.out],
                        score_bits=[pm2.f32_bits(o[3]) for o in m.out], evals=m.evals, rollbacks=m.rollbacks, asks=m.asks, rounds=m.rounds))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _run_dist(fasta, params):
    import multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, fasta, params, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


@pytest.mark.timeout(900)
def test_two_rank_column_sharded_loop_equals_the_oracle(oracle_lib, zika_fasta):
    """World-size-2 gloo run of the multi-GPU protocol (unit entries all-gathered and merged on every rank, cross-rank
    lists decided from exchanged cover-time histograms, the first-seen order of three or more partitions asked for when a
    tie needs it): every rank must return the oracle's loop on the WHOLE alignment, for the Zika fixture, for a family
    with blocks repeated across the rank boundary (cross-rank external winners) and for k = 6 (nearly every word on both
    ranks: ties over many partitions)."""
    rng = np.random.default_rng(5)
    anc = rng.integers(0, 4, 6000)
    anc[4000:4300] = anc[500:800]
    anc[5200:5500] = anc[1500:1800]
    rep = []
    for i in range(40):
        s = anc.copy()
        mut = rng.random(6000) < 0.03
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        rep.append(">r%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    cases = [("".join(rep).encode(), (500, 250, 50, 13, 150, 1, dict(kmax=8, wmax=16, chunk0=4, chunk=2))),
             (zika_fasta, (500, 250, 50, 13, 1000, 2, dict(kmax=4, wmax=8, chunk0=2, chunk=1))),
             (_random_alignment(6, 30, 1200, gaps=False), (100, 50, 30, 6, 60, 3, dict(kmax=8, wmax=16, chunk0=4, chunk=2)))]
    total_rb = total_asks = 0
    for fasta, params in cases:
        W, S, w, k, max_iter, mms, _ = params
        got = _run_dist(fasta, params)
        for d in (0, 1):
            want = oracle_lib.select(fasta, W, S, w, k, d, max_iter, mms)
            for rank in (0, 1):
                g = got[rank][d]
                assert g["codes"] == want["codes"].tolist(), "rank %d direction %d" % (rank, d)
                assert g["freqs"] == want["freqs"].tolist()
                assert g["n_tied"] == want["n_tied"].tolist()
                assert g["score_bits"] == want["scores"].view(np.uint32).tolist()
                assert g["evals"] == want["evals"]
            total_rb += got[0][d]["rollbacks"]
            total_asks += got[0][d]["asks"]
    assert total_rb > 0 and total_asks > 0
