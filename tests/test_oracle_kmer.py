"""Pins the k-mer oracle (oracle/kmer_oracle.py and its C++ port) to the reference's own unit tests
(od-msspe/src/main.rs:863-1236, delta_g.rs:162-194) and cross-checks the two restatements."""
import json
import os

import numpy as np
import pytest

from oracle import kmer_oracle as ko

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reverse_complement():  # main.rs:868-871
    assert ko.reverse_complement("ATCGAA") == "TTCGAT"


def test_get_search_windows():  # main.rs:874-879
    assert ko.get_sequence_on_search_windows("AACCTTGGAACCTTG-", 5) == ("AACCT", "CTTG-")


def test_find_kmers():  # main.rs:882-894
    k = ko.find_kmers("AACCTTGGAACCTTG-", 5)
    assert len(k) == 8
    assert set(k) == {"AACCT", "ACCTT", "CCTTG", "CTTGG", "TTGGA", "TGGAA", "GGAAC", "GAACC"}


def test_partitioning_sequence():  # main.rs:950-956
    assert ko.partitioning_sequence("AACCTTGGAACCTTGG", 10, 5) == ["AACCTTGGAA", "TGGAACCTTG"]


def test_get_segments():  # main.rs:897-947
    recs = [ko.SequenceRecord("seq1", "AACCTTGGAACCTTGG"), ko.SequenceRecord("seq2", "AACCTTGGAACCTTG-"),
            ko.SequenceRecord("seq3", "-ACCTTGGAACCTT-G")]
    segs = ko.get_segment_manager(recs, 10, 5, 5, 3)
    assert len(segs) == 6
    assert len(segs[0].kmers[0]) == 3
    assert len(segs[1].kmers[1]) == 3


def _two_segment_manager(second_fwd):
    s0 = ko.Segment(0, 0, 0)
    s0.kmers = [["ACT", "CTG", "TGA"], ["TAA", "AAT", "ATA"]]
    s1 = ko.Segment(1, 1, 1)
    s1.kmers = [second_fwd, ["TTC", "TCC", "CCA"]]
    return [s0, s1]


def test_make_kmer_segments_mapping():  # main.rs:959-1135
    m = ko.make_kmer_segments_windows_mapping(_two_segment_manager(["ACT", "CTG", "TGA"]))
    assert len(m) == 9
    for w in ("ACT", "CTG", "TGA"):
        assert len(m[(w, 0)]) == 2
    for w in ("TAA", "AAT", "ATA", "TTC", "TCC", "CCA"):
        assert len(m[(w, 1)]) == 1


def test_find_most_freq_kmer():  # main.rs:1138-1235
    segs = _two_segment_manager(["ACT", "CAG", "TGG"])
    m = ko.make_kmer_segments_windows_mapping(segs)
    word, freq, _, _ = ko.find_most_freq_kmer(segs, 0, set(), m, {})
    assert (word, freq) == ("ACT", 2)


def test_auto_max_mismatch_segments():  # main.rs:658-660
    assert [ko.auto_max_mismatch_segments(n) for n in (1, 50, 51, 96, 500, 501, 100000)] == [1, 1, 2, 2, 10, 10, 10]


def test_is_run():  # main.rs:478-490 -- only the trailing run counts
    assert ko.is_run("ACGTAAAAAA") and not ko.is_run("AAAAAAACGT") and not ko.is_run("ACGTAAAAA")


def test_fasta_loader(zika_fasta):
    recs = ko.to_records(zika_fasta)
    assert len(recs) == 96
    assert sorted({len(r.sequence) for r in recs}) == [13747, 13748]
    assert all("U" not in r.sequence and r.sequence == r.sequence.upper() for r in recs)
    assert " " not in recs[0].name


def test_zika_greedy_matches_survey_golden(zika_fasta):
    """SURVEY.md section 8a provisional golden, derived there by an independent restatement."""
    recs = ko.to_records(zika_fasta)
    segs = ko.get_segment_manager(recs, 500, 250, 50, 13)
    assert len(segs) == 5088
    assert (sum(len(s.kmers[0]) for s in segs), sum(len(s.kmers[1]) for s in segs)) == (129943, 139451)
    mms = ko.auto_max_mismatch_segments(len(recs))
    fwd = ko.find_candidates_kmers(segs, 0, 1000, mms)
    rev = ko.find_candidates_kmers(segs, 1, 1000, mms)
    assert len(fwd) == 107 and len(rev) == 106
    assert fwd[:5] == [("CTTGGAGTGCTTG", 96), ("ACACATGAGATGT", 95), ("AAGCAAGAATGCT", 93), ("ACCAACAACACCA", 93),
                       ("AGAGAGATCATAC", 93)]
    assert rev[:5] == [("CCACCGCCATCTG", 94), ("ACTGCTGTTGTCA", 93), ("AATGGCATCCCTT", 92), ("ATTGTGTCAATGT", 92),
                       ("CCACCTCCATACA", 92)]
    with open(os.path.join(GOLDEN, "snapshot_zika96_candidates.json")) as f:
        gold = json.load(f)
    assert [list(x) for x in fwd] == gold["fwd"] and [list(x) for x in rev] == gold["rev"]


def test_cpp_port_equals_python_oracle_on_zika(zika_fasta, oracle_lib):
    O = oracle_lib
    recs = ko.to_records(zika_fasta)
    segs = ko.get_segment_manager(recs, 500, 250, 50, 13)
    for d in (0, 1):
        tr = []
        py = ko.find_candidates_kmers(segs, d, 1000, 2, tr)
        cc = O.select(zika_fasta, 500, 250, 50, 13, d, 1000, 2)
        assert [ko.encode(w) for w, _ in py] == cc["codes"].tolist()
        assert [f for _, f in py] == cc["freqs"].tolist()
        assert [t[2] for t in tr] == cc["n_tied"].tolist()
        assert np.array([t[3] for t in tr], dtype=np.float32).tobytes() == cc["scores"].tobytes()
        codes, part = O.segment_slots(zika_fasta, 500, 250, 50, 13, d)
        assert codes.shape == (5088, 38)
        for g in (0, 1, 77, 5087):
            want = [ko.encode(w) for w in segs[g].kmers[d]]
            got = [int(c) for c in codes[g] if c != np.uint64(0xFFFFFFFFFFFFFFFF)]
            assert got == want and part[g] == segs[g].partition_no


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cpp_port_equals_python_oracle_random(seed, oracle_lib):
    """Small random alignments with gaps/N/IUPAC and heavy ties."""
    O = oracle_lib
    rng = np.random.default_rng(seed)
    L, n = 700, 12
    anc = rng.integers(0, 4, L)
    lines = []
    for i in range(n):
        s = anc.copy()
        mut = rng.random(L) < 0.03
        s[mut] = rng.integers(0, 4, mut.sum())
        chars = np.array(list("ACGT"))[s]
        chars[rng.random(L) < 0.004] = "-"
        chars[rng.random(L) < 0.002] = "N"
        chars[rng.random(L) < 0.001] = "R"
        lines.append(">g%d some description\n%s\n" % (i, "".join(chars)))
    fa = "".join(lines).encode()
    recs = ko.to_records(fa)
    segs = ko.get_segment_manager(recs, 100, 50, 20, 7)
    for d in (0, 1):
        py = ko.find_candidates_kmers(segs, d, 50, 1)
        cc = O.select(fa, 100, 50, 20, 7, d, 50, 1)
        assert [ko.encode(w) for w, _ in py] == cc["codes"].tolist()
        assert [f for _, f in py] == cc["freqs"].tolist()


def test_format_and_parse_ntthal_text_semantics(oracle_lib, zika_fasta):
    """Whole-pipeline regression on the Zika fixture (oracle-derived golden, NOT a reference output)."""
    O = oracle_lib
    r = O.run_pipeline(zika_fasta, O.default_config())
    with open(os.path.join(GOLDEN, "snapshot_zika96_default.csv")) as f:
        assert r.csv == f.read()
    with open(os.path.join(GOLDEN, "snapshot_zika96_default.report.txt")) as f:
        assert r.report == f.read()
    r.close()
