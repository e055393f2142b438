"""conflict graph + greedy vertex cover (main.rs:754-798) and the aggregation of the coverage
report (main.rs:518-574).  CPU part: the oracle's restatement on hand-worked cases.  GPU part: the device versions
behind the C ABI against the oracle."""
import numpy as np
import pytest

from oracle import kmer_oracle as ko


# ---------------------------------------------------------------- oracle (CPU)
def test_vertex_cover_oracle_hand_cases():
    # star: the centre has 3 active conflicts, everybody else 1 -> only the centre goes
    assert ko.greedy_vertex_cover(list("ABCD"), [("A", "B"), ("A", "C"), ("A", "D")]) == {"A"}
    # a single edge is a 1:1 tie: the lexicographically GREATEST word is removed (main.rs:789)
    assert ko.greedy_vertex_cover(["AAC", "AAG"], [("AAC", "AAG")]) == {"AAG"}
    # self conflict: the primer conflicts with itself and is removed
    assert ko.greedy_vertex_cover(["ACG", "TTT"], [("ACG", "ACG")]) == {"ACG"}
    # path a-b-c-d: b and c tie at 2 -> c goes first, then a-b remains -> b (greatest of the tie a/b)
    assert ko.greedy_vertex_cover(list("abcd"), [("a", "b"), ("b", "c"), ("c", "d")]) == {"c", "b"}
    # both orientations of an edge collapse (HashSet)
    assert ko.greedy_vertex_cover(list("ab"), [("a", "b"), ("b", "a")]) == {"b"}
    assert ko.greedy_vertex_cover(list("ab"), []) == set()


def test_degree_restatement_used_at_cfg4_size_equals_the_oracle():
    """tests/test_gpu_thermo.py checks the 20,000-node vertex cover against an integer-degree restatement (the oracle's
    string version is quadratic in Python); here that restatement is pinned to the oracle on random graphs."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "open-msspe-design_b200"))
    import msspe_b200 as m
    from test_gpu_thermo import _vertex_cover_by_degrees
    for seed, (n, pe, ps) in enumerate([(1, 0.0, 1.0), (2, 1.0, 0.0), (40, 0.05, 0.1), (120, 0.02, 0.05), (33, 0.5, 0.5), (200, 0.004, 0.0)]):
        rng = np.random.default_rng(seed)
        codes, ea, eb = _random_graph(rng, n, 13, pe, ps)
        words = [m.decode_word(int(c), 13) for c in codes]
        want = ko.greedy_vertex_cover(words, [(words[x], words[y]) for x, y in zip(ea.tolist(), eb.tolist())])
        got = _vertex_cover_by_degrees(codes, ea, eb)
        assert {words[i] for i in np.nonzero(got)[0]} == want


# ---------------------------------------------------------------- device (GPU)
def _random_graph(rng, n, k, p_edge, p_self):
    from msspe_b200 import synth
    codes = synth.random_primers(n, k, int(rng.integers(1 << 30)))
    a, b = np.nonzero(rng.random((n, n)) < p_edge)
    selfs = np.nonzero(rng.random(n) < p_self)[0]
    ea = np.concatenate([a, selfs]).astype(np.uint32)
    eb = np.concatenate([b, selfs]).astype(np.uint32)
    return codes, ea, eb


@pytest.mark.gpu
@pytest.mark.parametrize("n,p_edge,p_self,seed", [(1, 0.0, 1.0, 0), (2, 1.0, 0.0, 1), (40, 0.05, 0.1, 2), (300, 0.01, 0.02, 3),
                                                  (300, 0.2, 0.0, 4), (1500, 0.002, 0.01, 5), (33, 0.5, 0.5, 6)])
def test_vertex_cover_matches_oracle(n, p_edge, p_self, seed):
    import msspe_b200 as m
    rng = np.random.default_rng(seed)
    k = 13
    codes, ea, eb = _random_graph(rng, n, k, p_edge, p_self)
    words = [m.decode_word(int(c), k) for c in codes]
    want = ko.greedy_vertex_cover(words, [(words[x], words[y]) for x, y in zip(ea.tolist(), eb.tolist())])
    eng = m.Engine(k, 500, 250, 50)
    got = eng.vertex_cover(codes, ea, eb)
    assert {words[i] for i in np.nonzero(got)[0]} == want
    eng.close()


@pytest.mark.gpu
def test_vertex_cover_argument_errors():
    import msspe_b200 as m
    eng = m.Engine(13, 500, 250, 50)
    assert eng.vertex_cover(np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32)).size == 0
    with pytest.raises(m.MsspeError):  # duplicate word
        eng.vertex_cover(np.array([5, 5], np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    with pytest.raises(m.MsspeError):  # edge names a node that does not exist
        eng.vertex_cover(np.array([5, 6], np.uint64), np.array([0], np.uint32), np.array([2], np.uint32))
    eng.close()


@pytest.mark.gpu
def test_vertex_cover_cfg4_shape():
    """20,000 primers, ~0.5 % of the ordered pairs in conflict (BASELINE configs[3]): properties that do not need the
    O(n^2) python oracle -- the result is a vertex cover, and no removed primer was isolated."""
    import msspe_b200 as m
    from msspe_b200 import synth
    n = 20_000
    rng = np.random.default_rng(44)
    codes = synth.random_primers(n, 13, 4)
    ne = 2_000_000
    ea = rng.integers(0, n, ne).astype(np.uint32); eb = rng.integers(0, n, ne).astype(np.uint32)
    eng = m.Engine(13, 500, 250, 50)
    got = eng.vertex_cover(codes, ea, eb).astype(bool)
    assert np.all(got[ea] | got[eb])          # every conflict edge lost an endpoint
    touched = np.zeros(n, bool); touched[ea] = True; touched[eb] = True
    assert not np.any(got & ~touched)         # nothing without conflicts was removed
    assert 0 < got.sum() < n
    eng.close()


@pytest.mark.gpu
def test_coverage_summary_matches_full_coverage():
    import gzip, os
    import msspe_b200 as m
    here = os.path.dirname(os.path.abspath(__file__))
    with gzip.open(os.path.join(here, "golden", "zika96_aligned.fa.gz"), "rb") as f:
        recs = ko.to_records(f.read())
    bases, offs = m.pack_records([r.sequence.encode() for r in recs])
    eng = m.Engine(13, 500, 250, 50)
    eng.load_genomes(bases, offs)
    eng.build_index()
    a, b = eng.select_both(25, 2, m.SELECT_RECOUNT)
    for fc, rc in ((a["code"], b["code"]), (a["code"][:3], b["code"][:0]), (a["code"][:0], b["code"][:0])):
        cov, part, rec = eng.coverage(fc, rc)
        rcv, rtot, pcv, ptot, ncov = eng.coverage_summary(fc, rc, len(recs))
        assert ncov == int(cov.sum())
        assert np.array_equal(rtot, np.bincount(rec, minlength=len(recs)))
        assert np.array_equal(rcv, np.bincount(rec, weights=cov, minlength=len(recs)).astype(np.uint32))
        assert np.array_equal(ptot, np.bincount(part, minlength=len(ptot)))
        assert np.array_equal(pcv, np.bincount(part, weights=cov, minlength=len(pcv)).astype(np.uint32))
    eng.close()
