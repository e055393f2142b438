"""GPU parity tests, through the C ABI, for the k-mer engine: K1 (segments), K2 (inverted index), K3 (greedy
selection).  Bit-exact against the oracle (integer/index work; the f32 tie score is compared by bit pattern)."""
import gzip
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import kmer_oracle as ko  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
# every loop variant: persistent recount, persistent incremental, AUTO, per-partition greedy sequences + merge, and the
# launch-per-phase forms of the first two
ALL_MODES = (0, 1, 2, 3, 0x100, 0x101)


def _fasta_to_arrays(fa: bytes):
    import msspe_b200 as m
    recs = ko.to_records(fa)
    return recs, m.pack_records([r.sequence.encode() for r in recs])


def _check_select(eng, O, fa, W, S, w, k, max_iter, mms, mode):
    for d in (0, 1):
        got = eng.select(d, max_iter, mms, mode)
        want = O.select(fa, W, S, w, k, d, max_iter, mms)
        assert got["code"].tolist() == want["codes"].tolist(), "direction %d winners differ" % d
        assert got["freq"].tolist() == want["freqs"].tolist()
        assert got["n_tied"].tolist() == want["n_tied"].tolist()
        assert got["tie_score"].astype(np.float32).tobytes() == want["scores"].tobytes()
        t = eng.timing()
        assert t.select_evals[d] == want["evals"], "coverage evals differ from the reference-equivalent count"


def _load_golden(name):
    """tests/golden/<name>_candidates.json: written offline by tools/gen_size_goldens.py from the CPU oracle
    (oracle_select = main.rs:285-406 with a full recount per iteration) on the same synthetic input."""
    with open(os.path.join(GOLDEN, "%s_candidates.json" % name)) as f:
        return json.load(f)


def _check_golden(eng, gold, mode, both=True):
    """Every loop variant is compared with the ORACLE's winners / frequencies / tie counts / f32 tie scores (bit
    patterns) / reference-equivalent evals at this size -- not with another variant of the CUDA path."""
    it, mms = gold["max_iterations"], gold["max_mismatch_segments"]
    if both:
        res = eng.select_both(it, mms, mode)
        evals = tuple(eng.timing().select_evals)
    else:
        res, evals = [], []
        for d in (0, 1):
            res.append(eng.select(d, it, mms, mode))
            evals.append(eng.timing().select_evals[d])
    for d in (0, 1):
        want, got = gold["dirs"][d], res[d]
        assert got["code"].tolist() == want["codes"], "mode %#x direction %d: winners differ from the oracle golden" % (mode, d)
        assert got["freq"].tolist() == want["freqs"]
        assert got["n_tied"].tolist() == want["n_tied"]
        assert got["tie_score"].astype(np.float32).view(np.uint32).tolist() == want["score_bits"]
        assert int(evals[d]) == want["evals"], "mode %#x direction %d: coverage evals differ from the oracle" % (mode, d)
    return res


@pytest.fixture(scope="module")
def zika_engine(zika_fasta, oracle_lib):
    import msspe_b200 as m
    recs, (bases, offs) = _fasta_to_arrays(zika_fasta)
    eng = m.Engine(13, 500, 250, 50)
    eng.load_genomes(bases, offs)
    eng.build_index()
    yield eng, recs
    eng.close()


def test_zika_segments_match_oracle(zika_engine, zika_fasta, oracle_lib):
    eng, recs = zika_engine
    g, maxp, s = eng.segment_info()
    assert (g, maxp, s) == (5088, 52, 38)
    for d in (0, 1):
        want, part = oracle_lib.segment_slots(zika_fasta, 500, 250, 50, 13, d)
        got = eng.segment_kmers(d)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_zika_inverted_index_matches_reference_mapping(zika_engine):
    eng, recs = zika_engine
    segs = ko.get_segment_manager(recs, 500, 250, 50, 13)
    mp = ko.make_kmer_segments_windows_mapping(segs)
    for d in (0, 1):
        codes, offs, post = eng.index(d)
        want = {ko.encode(w): v for (w, dd), v in mp.items() if dd == d}
        assert len(codes) == len(want) == (7471 if d == 0 else 7737)
        assert np.all(np.diff(codes.astype(np.int64)) > 0)
        assert int(offs[-1]) == len(post) == (129943 if d == 0 else 139451)
        for i in list(range(0, len(codes), 97)) + [len(codes) - 1]:
            assert post[int(offs[i]):int(offs[i + 1])].tolist() == want[int(codes[i])]


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 0x100, 0x101])
def test_zika_greedy_selection_bit_exact(zika_engine, zika_fasta, oracle_lib, mode):
    eng, _ = zika_engine
    _check_select(eng, oracle_lib, zika_fasta, 500, 250, 50, 13, 1000, 2, mode)
    with open(os.path.join(GOLDEN, "snapshot_zika96_candidates.json")) as f:
        gold = json.load(f)
    fwd, rev = eng.select_both(1000, 2, mode)
    assert [[ko.decode(int(c), 13), int(f)] for c, f in zip(fwd["code"], fwd["freq"])] == gold["fwd"]
    assert [[ko.decode(int(c), 13), int(f)] for c, f in zip(rev["code"], rev["freq"])] == gold["rev"]


def test_zika_stop_rules(zika_engine, zika_fasta, oracle_lib):
    """max_iterations bound (main.rs:344) and the frequency threshold break after the push (main.rs:387-390)."""
    eng, _ = zika_engine
    for mode in (0, 3):
        _check_select(eng, oracle_lib, zika_fasta, 500, 250, 50, 13, 7, 2, mode)
        _check_select(eng, oracle_lib, zika_fasta, 500, 250, 50, 13, 1000, 94, mode)
        _check_select(eng, oracle_lib, zika_fasta, 500, 250, 50, 13, 0, 2, mode)
        _check_select(eng, oracle_lib, zika_fasta, 500, 250, 50, 13, 1, 2, mode)


def test_reference_unit_vectors_through_the_engine(oracle_lib):
    """main.rs:897-947 (test_get_segments): 3 records, window 10 / step 5 / search 5 / k 3 -> 6 segments."""
    import msspe_b200 as m
    fa = b">seq1\nAACCTTGGAACCTTGG\n>seq2\nAACCTTGGAACCTTG-\n>seq3\n-ACCTTGGAACCTT-G\n"
    recs, (bases, offs) = _fasta_to_arrays(fa)
    eng = m.Engine(3, 10, 5, 5)
    eng.load_genomes(bases, offs)
    eng.build_index()
    g, maxp, s = eng.segment_info()
    assert (g, maxp, s) == (6, 1, 3)
    f = eng.segment_kmers(0)
    r = eng.segment_kmers(1)
    assert [ko.decode(int(c), 3) for c in f[0] if c != m.NO_KMER] == ["AAC", "ACC", "CCT"]
    assert len([c for c in r[1] if c != m.NO_KMER]) == 3
    _check_select(eng, oracle_lib, fa, 10, 5, 5, 3, 10, 1, 0)
    _check_select(eng, oracle_lib, fa, 10, 5, 5, 3, 10, 1, 3)
    eng.close()


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
            ch[rng.random(L) < 0.001] = "R"
            ch[rng.random(L) < 0.002] = "u"   # lower case / U are normalised by to_records
        out.append(">g%d x y\n%s\n" % (i, "".join(ch)))
    return "".join(out).encode()


@pytest.mark.parametrize("seed,n,L,W,S,w,k,mms", [
    (0, 12, 700, 100, 50, 20, 7, 1),
    (1, 40, 2000, 200, 100, 60, 9, 1),
    (2, 30, 1501, 500, 250, 50, 13, 1),     # ragged tail: only full windows count
    (3, 64, 1000, 128, 64, 64, 5, 2),       # tiny k: heavy within-window duplicates, massive ties
    (4, 20, 1200, 300, 150, 100, 31, 1),    # maximum k
    (5, 25, 900, 90, 45, 45, 17, 1),        # k > 16: 64-bit codes
    (6, 30, 1200, 100, 50, 30, 6, 3),       # single- and multi-partition lists mixed: many external winners in the partitioned loop
    (7, 20, 1500, 300, 150, 90, 11, 1),     # search window > 64 bases with k <= 16: the general (shared-memory) one-pass encoder
    (8, 16, 1300, 260, 130, 64, 16, 1),     # widest packed window with the longest packed word
])
def test_random_alignments_bit_exact(oracle_lib, seed, n, L, W, S, w, k, mms):
    import msspe_b200 as m
    fa = _random_alignment(seed, n, L)
    recs, (bases, offs) = _fasta_to_arrays(fa)
    eng = m.Engine(k, W, S, w)
    eng.load_genomes(bases, offs)
    eng.build_index()
    for d in (0, 1):
        want, part = oracle_lib.segment_slots(fa, W, S, w, k, d)
        assert np.array_equal(eng.segment_kmers(d), want)
    for mode in (0, 1, 2, 3, 0x100, 0x101):   # persistent recount / incremental / AUTO / partitioned / launch-per-phase recount / incremental
        _check_select(eng, oracle_lib, fa, W, S, w, k, 60, mms, mode)
    eng.close()


def test_low_complexity_windows_repeat_words(oracle_lib):
    """Homopolymer runs, di- and tri-nucleotide repeats: most words of a search window repeat an earlier one, which the
    packed encoder finds by comparing the window with itself at every distance (itertools unique(), main.rs:163-171)."""
    import msspe_b200 as m
    rng = np.random.default_rng(77)
    L, n = 1500, 14
    anc = rng.integers(0, 4, L)
    for start, unit in ((40, [0]), (260, [0, 3]), (520, [1, 2, 1]), (800, [2]), (1010, [0, 1, 2, 3]), (1290, [3, 3, 0])):
        rep = np.array((unit * 200)[:170])
        anc[start:start + len(rep)] = rep
    recs = []
    for i in range(n):
        s = anc.copy()
        mut = rng.random(L) < 0.01
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        ch = np.array(list("ACGT"))[s]
        ch[rng.random(L) < 0.003] = "N"
        recs.append(">r%d\n%s\n" % (i, "".join(ch)))
    fa = "".join(recs).encode()
    _, (bases, offs) = _fasta_to_arrays(fa)
    for (W, S, w, k) in ((250, 125, 50, 13), (250, 125, 64, 9), (200, 100, 40, 15)):
        eng = m.Engine(k, W, S, w)
        eng.reserve_pool(64 << 20)           # msspe_reserve_pool: background first touch, the build waits for it
        eng.load_genomes(bases, offs)
        eng.build_index()
        for d in (0, 1):
            want, _ = oracle_lib.segment_slots(fa, W, S, w, k, d)
            assert np.array_equal(eng.segment_kmers(d), want)
        for mode in (0, 2, 3):
            _check_select(eng, oracle_lib, fa, W, S, w, k, 40, 1, mode)
        eng.close()


def test_index_sort_with_narrow_digits(oracle_lib, monkeypatch):
    """MSSPE_SORT_BITS=8: four 8-bit passes with ballot ranking instead of three 10-11-bit passes (kmer_build_fast.cu); the
    index must not depend on the digit width."""
    import msspe_b200 as m
    fa = _random_alignment(11, 60, 3000)
    _, (bases, offs) = _fasta_to_arrays(fa)
    ref = None
    for bits in (None, "8", "5"):
        if bits:
            monkeypatch.setenv("MSSPE_SORT_BITS", bits)
        else:
            monkeypatch.delenv("MSSPE_SORT_BITS", raising=False)
        eng = m.Engine(15, 500, 250, 50)
        eng.load_genomes(bases, offs)
        eng.build_index()
        got = [tuple(np.asarray(x).tobytes() for x in eng.index(d)) for d in (0, 1)]
        if ref is None:
            ref = got
            for d in (0, 1):
                want, _ = oracle_lib.segment_slots(fa, 500, 250, 50, 15, d)
                assert np.array_equal(eng.segment_kmers(d), want)
        else:
            assert got == ref
        eng.close()


def test_identical_genomes_tie_storm(oracle_lib):
    """p = 0: every k-mer of a partition ties with every other; the tie-break decides everything."""
    import msspe_b200 as m
    fa = _random_alignment(7, 50, 3000, p=0.0, gaps=False)
    recs, (bases, offs) = _fasta_to_arrays(fa)
    eng = m.Engine(13, 500, 250, 50)
    eng.load_genomes(bases, offs)
    eng.build_index()
    for mode in (0, 1, 3, 0x100, 0x101):
        _check_select(eng, oracle_lib, fa, 500, 250, 50, 13, 40, 1, mode)
    eng.close()


def test_empty_and_degenerate_inputs(oracle_lib):
    import msspe_b200 as m
    # all-gap windows and sequences shorter than the window: zero k-mers, zero segments
    fa = b">a\n" + b"-" * 1200 + b"\n>b\n" + b"ACGT" * 20 + b"\n"
    recs, (bases, offs) = _fasta_to_arrays(fa)
    eng = m.Engine(13, 500, 250, 50)
    eng.load_genomes(bases, offs)
    eng.build_index()
    g, maxp, s = eng.segment_info()
    assert g == 3 and maxp == 2
    assert np.all(eng.segment_kmers(0) == m.NO_KMER)
    assert len(eng.select(0, 10, 1)) == 0 and len(eng.select(1, 10, 1, 1)) == 0 and len(eng.select(0, 10, 1, 3)) == 0
    fa2 = b">only\nACGTACGT\n"
    recs, (bases, offs) = _fasta_to_arrays(fa2)
    eng.load_genomes(bases, offs)
    eng.build_index()
    assert eng.segment_info()[0] == 0
    assert len(eng.select(0, 10, 1)) == 0 and len(eng.select(0, 10, 1, 3)) == 0 and len(eng.select(1, 10, 1, 2)) == 0
    with pytest.raises(m.MsspeError):
        eng.load_genomes(np.zeros(0, np.uint8), np.zeros(1, np.uint64))  # "No sequences found", main.rs:652-654
    eng.close()
    # Segment.partition_no is `j as u16` (main.rs:84,227): more than 65,536 windows per record would wrap in the reference;
    # the engine refuses instead of mixing wrapped (K1) and unwrapped (tie score) partition numbers
    e2 = m.Engine(1, 2, 1, 1)
    with pytest.raises(m.MsspeError) as ei:
        e2.load_genomes(np.full(70_000, ord("A"), np.uint8), np.array([0, 70_000], np.uint64))
    assert ei.value.code == m.ERR_INVALID
    e2.load_genomes(np.full(65_537, ord("A"), np.uint8), np.array([0, 65_537], np.uint64))   # exactly 65,536 windows: accepted
    e2.close()


def test_cfg1_shape_full_size(oracle_lib):
    """BASELINE configs[0]: 50 x 10 kb, k=13, 500/250/50 -- the reference's own CPU-runnable case, full size."""
    import msspe_b200 as m
    from msspe_b200 import synth
    g, k = synth.make_config("cfg1")
    fa = synth.to_fasta(g)
    eng = m.Engine(k, 500, 250, 50)
    eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
    eng.build_index()
    assert eng.segment_info()[0] == 50 * 39
    for mode in (0, 1, 2, 3, 0x100):
        _check_select(eng, oracle_lib, fa, 500, 250, 50, k, 1000, 1, mode)
    eng.close()


def test_cfg2_full_size_vs_oracle_golden(monkeypatch):
    """BASELINE configs[1] at full size (1000 x 30 kb, 2 x ~605 iterations, 1.1e9 reference-equivalent evals): every loop
    variant against tests/golden/cfg2_candidates.json (the CPU oracle run offline, tools/gen_size_goldens.py), incl. the
    variants that only exist at scale: default-threshold stream compaction, the 148 x 1024 block shape
    (MSSPE_PERSIST_1024), the AUTO switch, launch-per-phase."""
    import msspe_b200 as m
    from msspe_b200 import synth
    gold = _load_golden("cfg2")
    g, k = synth.make_config("cfg2")
    assert g.shape == (gold["genomes"], gold["length"]) and k == gold["kmer_size"]
    eng = m.Engine(k, 500, 250, 50)
    eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
    eng.build_index()
    G, maxp, s = eng.segment_info()
    assert G == 1000 * 119 and maxp == 118
    for mode in ALL_MODES:
        a0, b0 = _check_golden(eng, gold, mode)
    _check_golden(eng, gold, m.SELECT_RECOUNT, both=False)
    monkeypatch.setenv("MSSPE_PERSIST_1024", "1")
    _check_golden(eng, gold, m.SELECT_RECOUNT)
    monkeypatch.delenv("MSSPE_PERSIST_1024")
    for x in (a0, b0):
        assert np.all(np.diff(x["freq"].astype(np.int64)) <= 0) and len(set(x["code"].tolist())) == len(x)
    for d in (0, 1):
        codes, offs, post = eng.index(d)
        assert int(offs[-1]) == len(post)
        assert np.all(np.diff(codes.astype(np.int64)) > 0)
    cov, part, rec = eng.coverage(a0["code"], b0["code"])
    assert cov.sum() > 0.5 * G and part.max() == 118 and rec.max() == 999
    eng.close()


def test_sharded_selection_primitives_world1(zika_engine):
    """msspe_shard_* primitives driven by select_sharded without a process group == the fused device loop."""
    from msspe_b200 import distributed as D
    eng, _ = zika_engine
    for d in (0, 1):
        want = eng.select(d, 1000, 2, 0)
        want_evals = eng.timing().select_evals[d]
        got, evals, iters = D.select_sharded(eng, d, 1000, 2, None, "cuda")
        assert got.tobytes() == want.tobytes()
        assert evals == want_evals and iters == eng.timing().select_iterations[d]


def test_cfg3_full_size_vs_oracle_golden(monkeypatch):
    """BASELINE configs[2] at full size (10,000 x 11 kb, k=15, --max-mismatch-segments=2, 2 x 1000 iterations, 6.4e9
    reference-equivalent evals) -- the bench.py headline workload: every loop variant against
    tests/golden/cfg3_candidates.json (CPU oracle, offline)."""
    import msspe_b200 as m
    from msspe_b200 import synth
    gold = _load_golden("cfg3")
    g, k = synth.make_config("cfg3")
    assert g.shape == (gold["genomes"], gold["length"]) and k == gold["kmer_size"]
    eng = m.Engine(k, 500, 250, 50)
    eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
    eng.build_index()
    G, maxp, s = eng.segment_info()
    assert G == 10_000 * 43 and maxp == 42 and s == 36
    for mode in ALL_MODES:
        a0, b0 = _check_golden(eng, gold, mode)
    monkeypatch.setenv("MSSPE_PERSIST_1024", "1")
    _check_golden(eng, gold, m.SELECT_RECOUNT)
    monkeypatch.delenv("MSSPE_PERSIST_1024")
    assert len(a0) == 1000 and np.all(np.diff(a0["freq"].astype(np.int64)) <= 0)
    first = eng.select(0, 1, 2, 0)
    assert len(first) == 1 and eng.timing().select_evals[0] == eng.index(0)[1][-1]
    eng.close()


def test_ragged_record_lengths_use_the_search_path(oracle_lib):
    """Records of different lengths (different partition counts): the per-segment record lookup is a binary search."""
    import msspe_b200 as m
    rng = np.random.default_rng(21)
    anc = rng.integers(0, 4, 2600)
    lines = []
    for i, L in enumerate([2600, 1800, 2600, 999, 1200, 2599, 500, 499, 2600]):
        s = anc[:L].copy()
        mut = rng.random(L) < 0.02
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        lines.append(">x%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    fa = "".join(lines).encode()
    recs, (bases, offs) = _fasta_to_arrays(fa)
    eng = m.Engine(13, 500, 250, 50)
    eng.load_genomes(bases, offs)
    eng.build_index()
    for d in (0, 1):
        want, part = oracle_lib.segment_slots(fa, 500, 250, 50, 13, d)
        assert np.array_equal(eng.segment_kmers(d), want)
    _check_select(eng, oracle_lib, fa, 500, 250, 50, 13, 60, 1, 0)
    _check_select(eng, oracle_lib, fa, 500, 250, 50, 13, 60, 1, 3)
    cov, part2, rec = eng.coverage([], [])
    assert part2.tolist() == part.tolist() and rec.max() == 8
    eng.close()


@pytest.mark.parametrize("compact_min", ["1", "4096"])
def test_stream_compaction_on_small_inputs(zika_fasta, oracle_lib, compact_min, monkeypatch):
    """The persistent kernel leaves for a dead-posting compaction of its scoring stream when less than half of what it
    streams is live.  By default only streams of >= 2^20 postings do (cfg2, cfg3); MSSPE_COMPACT_MIN=1 forces many
    compactions on the small fixtures: the Zika alignment (incl. the partition_coverage of already covered postings,
    main.rs:371-378, which the compacted stream no longer holds), random alignments with massive ties, a tie storm."""
    import msspe_b200 as m
    monkeypatch.setenv("MSSPE_COMPACT_MIN", compact_min)
    cases = [(zika_fasta, 500, 250, 50, 13, 1000, 2), (_random_alignment(3, 64, 1000), 128, 64, 64, 5, 60, 2),
             (_random_alignment(1, 40, 2000), 200, 100, 60, 9, 200, 1), (_random_alignment(7, 50, 3000, p=0.0, gaps=False), 500, 250, 50, 13, 40, 1)]
    for fa, W, S, w, k, it, mms in cases:
        recs, (bases, offs) = _fasta_to_arrays(fa)
        eng = m.Engine(k, W, S, w)
        eng.load_genomes(bases, offs)
        eng.build_index()
        _check_select(eng, oracle_lib, fa, W, S, w, k, it, mms, 0)
        a, b = eng.select_both(it, mms, 0)          # both directions in one launch sequence; a second run on the same index
        assert a.tobytes() == eng.select(0, it, mms, 0).tobytes() and b.tobytes() == eng.select(1, it, mms, 0).tobytes()
        eng.close()


def test_global_bitmask_variant_many_segments():
    """More segments than a shared-memory bitmask can hold (2,000 genomes x 1,000 partitions = 2.0 M segments > 1.8 M
    bits): the persistent kernel reads the global bitmask, which block 0 updates inside the launch (grid barrier
    after the update).  Every loop variant against tests/golden/bitmask2m_candidates.json (CPU oracle, offline)."""
    import msspe_b200 as m
    from msspe_b200 import synth
    gold = _load_golden("bitmask2m")
    g = synth.synth_genomes(2000, 30_000, 11, clades=16, p_clade=0.08, p_leaf=0.01)
    eng = m.Engine(13, 30, 30, 20)
    eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
    eng.build_index()
    G, maxp, s = eng.segment_info()
    assert G == 2000 * 1000 and maxp == 999 and s == 8
    os.environ["MSSPE_FORCE_RECOUNT"] = "1"     # without it a recount request of this size is served by the incremental kernel
    try:
        _check_golden(eng, gold, m.SELECT_RECOUNT)
    finally:
        del os.environ["MSSPE_FORCE_RECOUNT"]
    for mode in ALL_MODES:                      # mode 0 now = the automatic fallback to the incremental kernel
        a0, b0 = _check_golden(eng, gold, mode)
    assert len(a0) == 60 and np.all(np.diff(a0["freq"].astype(np.int64)) <= 0)
    first = eng.select(0, 1, 10, 0)
    assert len(first) == 1 and eng.timing().select_evals[0] == eng.index(0)[1][-1]
    eng.close()


def test_cfg5_shard_vs_oracle_golden(monkeypatch):
    """One GPU's eighth of BASELINE configs[4] (12,500 x 30 kb, 1.49 M segments, 2 x 56.5 M postings, larger than L2):
    the first 100 iterations per direction of every loop variant against tests/golden/cfg5shard_candidates.json (CPU
    oracle, offline, 1.1e10 evals), incl. the u32 posting offsets at 5.6e7 postings and the 148 x 1024 block shape;
    then 300 iterations: the variants agree with each other beyond the golden's horizon and with its prefix."""
    import msspe_b200 as m
    from msspe_b200 import synth
    gold = _load_golden("cfg5shard")
    g = synth.synth_genomes(12_500, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
    assert g.shape == (gold["genomes"], gold["length"])
    eng = m.Engine(13, 500, 250, 50)
    eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
    eng.build_index()
    assert eng.segment_info() == (12_500 * 119, 118, 38)
    for mode in ALL_MODES:
        if mode & m.SELECT_BATCHED and (mode & 0xFF) == m.SELECT_RECOUNT:
            continue    # 2 x 100 stand-alone recounts of 226 MB: covered by cfg2/cfg3, skipped here for GPU time
        _check_golden(eng, gold, mode)
    monkeypatch.setenv("MSSPE_PERSIST_1024", "1")
    _check_golden(eng, gold, m.SELECT_RECOUNT)
    monkeypatch.delenv("MSSPE_PERSIST_1024")
    out = {}
    for mode in (m.SELECT_RECOUNT, m.SELECT_INCREMENTAL, m.SELECT_AUTO, m.SELECT_PARTITIONED):
        a, b = eng.select_both(300, 10, mode)
        out[mode] = (a.tobytes(), b.tobytes(), tuple(eng.timing().select_evals))
        assert len(a) == 300 and len(b) == 300
        assert a["code"][:100].tolist() == gold["dirs"][0]["codes"] and b["code"][:100].tolist() == gold["dirs"][1]["codes"]
    assert out[m.SELECT_RECOUNT] == out[m.SELECT_INCREMENTAL] == out[m.SELECT_AUTO] == out[m.SELECT_PARTITIONED]
    eng.close()


@pytest.mark.parametrize("chunks", [None, ("1", "1"), ("3", "2")])
def test_partitioned_loop_with_repeats_and_small_chunks(oracle_lib, zika_fasta, chunks, monkeypatch):
    """MSSPE_SELECT_PARTITIONED where its hard cases live: a genome family with a repeated block (the strongest lists span
    two partitions and win early: external winners, roll-backs, partition_coverage of already covered postings), with the
    look-ahead chunk sizes forced down so that extension rounds, horizon limits and roll-backs interleave."""
    import msspe_b200 as m
    if chunks:
        monkeypatch.setenv("MSSPE_PART_CHUNK0", chunks[0])
        monkeypatch.setenv("MSSPE_PART_CHUNK", chunks[1])
    rng = np.random.default_rng(5)
    anc = rng.integers(0, 4, 3000)
    anc[2000:2300] = anc[500:800]
    out = []
    for i in range(40):
        s = anc.copy()
        mut = rng.random(3000) < 0.03
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        out.append(">r%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    cases = [("".join(out).encode(), 500, 250, 50, 13, 200, 1), ("".join(out).encode(), 500, 250, 50, 13, 200, 3),
             (zika_fasta, 500, 250, 50, 13, 1000, 2), (_random_alignment(6, 30, 1200), 100, 50, 30, 6, 60, 3)]
    for fa, W, S, w, k, it, mms in cases:
        recs, (bases, offs) = _fasta_to_arrays(fa)
        eng = m.Engine(k, W, S, w)
        eng.load_genomes(bases, offs)
        eng.build_index()
        _check_select(eng, oracle_lib, fa, W, S, w, k, it, mms, m.SELECT_PARTITIONED)
        a, b = eng.select_both(it, mms, m.SELECT_PARTITIONED)
        assert a.tobytes() == eng.select(0, it, mms, 0).tobytes() and b.tobytes() == eng.select(1, it, mms, 0).tobytes()
        eng.close()
