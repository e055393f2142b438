"""the multi-threaded FASTA ingest behind the C ABI against the oracle's to_records (main.rs:108-122).
The parser is host code, so its parity tests run without a GPU; the load path (chunked H2D) is the GPU test."""
import gzip
import os

import numpy as np
import pytest

from oracle import kmer_oracle as ko

CASES = {
    "plain": b">a desc here\nACGT\nACGT\n>b\nTTTT\n",
    "crlf": b">a x\r\nACGT\r\nAC\r\n>b\r\nGG\r\n",
    "lower_u": b">r1\nacgun\nNNuU-\n",
    "no_trailing_newline": b">a\nAC\n>b\nGT",
    "empty_lines": b"\n\n>a\nAC\n\nGT\n\n>b\n\n",
    "empty_record": b">a\n>b\nAC\n>c\n",
    "gt_inside_line": b">a\nAC>GT\nTT\n>b\nA\n",
    "header_only_spaces": b">a  two  spaces\nAC\n> leading\nGG\n",
    "cr_inside": b">a\nA\rC\nG\n",
    "empty_file": b"",
    "only_newlines": b"\n\n",
}


def _check(tmp_path, name, data, threads):
    import msspe_b200 as m
    p = tmp_path / (name + ".fa")
    p.write_bytes(data)
    want = ko.to_records(data)
    f = m.fasta_open(str(p), threads)
    assert f.names == [r.name for r in want]
    assert f.sequences() == [r.sequence for r in want]
    f.close()


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads", [1, 4])
def test_fasta_parser_matches_to_records(tmp_path, name, threads):
    _check(tmp_path, name, CASES[name], threads)


def test_fasta_invalid_start_and_missing_file(tmp_path):
    import msspe_b200 as m
    p = tmp_path / "bad.fa"
    p.write_bytes(b"ACGT\n>a\nAC\n")
    with pytest.raises(m.MsspeError) as e:
        m.fasta_open(str(p))
    assert e.value.code == m.ERR_INVALID and "InvalidStart" in str(e.value)
    with pytest.raises(m.MsspeError) as e:
        m.fasta_open(str(tmp_path / "missing.fa"))
    assert e.value.code == m.ERR_IO


def test_fasta_zika_fixture_and_many_chunks(tmp_path, zika_fasta):
    """The reference's own fixture, and a file large enough (80 MB of sequence) to cross several 32 MB chunks and to
    give every thread several record ranges."""
    _check(tmp_path, "zika", zika_fasta, 0)
    rng = np.random.default_rng(7)
    recs = []
    for i in range(400):
        seq = np.frombuffer(b"ACGTacgtuUN-", np.uint8)[rng.integers(0, 12, 200_000)].tobytes()
        lines = b"\n".join(seq[j:j + 70] for j in range(0, len(seq), 70))
        recs.append(b">g%d some description\n" % i + lines + (b"\r\n" if i % 3 == 0 else b"\n"))
    data = b"".join(recs)
    import msspe_b200 as m
    p = tmp_path / "big.fa"
    p.write_bytes(data)
    f = m.fasta_open(str(p), 8)
    assert f.names == ["g%d" % i for i in range(400)]
    assert int(f.offsets[-1]) == 400 * 200_000
    lut = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"acgtu", b"ACGTU"):
        lut[a] = b
    lut[ord("U")] = ord("T"); lut[ord("u")] = ord("T")
    raw = np.frombuffer(data, np.uint8)
    # independent numpy restatement: drop header lines and line ends, map through the LUT
    want = ko.to_records(data[: len(recs[0]) + len(recs[1])])
    assert f.sequences()[:2] == [r.sequence for r in want]
    body = np.concatenate([np.frombuffer(r[r.index(b"\n") + 1:], np.uint8) for r in recs])
    body = body[(body != 10) & (body != 13)]
    assert np.array_equal(f.bases, lut[body])
    f.close()


@pytest.mark.gpu
def test_load_fasta_equals_load_genomes(tmp_path, zika_fasta):
    import msspe_b200 as m
    p = tmp_path / "zika.fa"
    p.write_bytes(zika_fasta)
    recs = ko.to_records(zika_fasta)
    bases, offs = m.pack_records([r.sequence.encode() for r in recs])
    e1 = m.Engine(13, 500, 250, 50); e1.load_genomes(bases, offs); e1.build_index()
    e2 = m.Engine(13, 500, 250, 50); f = e2.load_fasta(str(p), 4); e2.build_index()
    assert f.names == [r.name for r in recs] and np.array_equal(f.offsets, offs)
    assert e1.segment_info() == e2.segment_info()
    for d in (0, 1):
        assert np.array_equal(e1.segment_kmers(d), e2.segment_kmers(d))
        assert e1.select(d, 50, 2).tobytes() == e2.select(d, 50, 2).tobytes()
    with pytest.raises(m.MsspeError):
        e2.load_fasta(str(tmp_path / "nope.fa"))
    e1.close(); e2.close(); f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", ["0", "1"])
def test_load_fasta_pageable_and_pinned_paths(tmp_path, pinned, monkeypatch):
    """Inputs above 256 MB are parsed into a PAGEABLE buffer whose 32 MB chunks go to the device while the next chunk is
    normalised (fasta.cu host_alloc); MSSPE_FASTA_PINNED forces either path at any size.  A 100 MB multi-chunk file with
    lower case, U, CR/LF and ragged line lengths through both paths: names, offsets, device-side k-mers and the greedy
    winners must equal the oracle's to_records + msspe_load_genomes."""
    import msspe_b200 as m
    monkeypatch.setenv("MSSPE_FASTA_PINNED", pinned)
    rng = np.random.default_rng(8)
    anc = rng.integers(0, 4, 50_000)
    lines = []
    for i in range(2000):
        s = anc.copy()
        mut = rng.random(len(s)) < 0.01
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        seq = "".join(np.array(list("ACGU" if i % 3 == 0 else "acgt" if i % 3 == 1 else "ACGT"))[s])
        wrap = 60 + (i % 7) * 13
        body = ("\r\n" if i % 5 == 0 else "\n").join(seq[j:j + wrap] for j in range(0, len(seq), wrap))
        lines.append(">rec%d some description\n%s\n" % (i, body))
    data = "".join(lines).encode()
    assert len(data) > 3 * (32 << 20)      # more than three 32 MB chunks
    p = tmp_path / "big.fa"
    p.write_bytes(data)
    recs = ko.to_records(data)
    bases, offs = m.pack_records([r.sequence.encode() for r in recs])
    e1 = m.Engine(13, 500, 250, 50); e1.load_genomes(bases, offs); e1.build_index()
    e2 = m.Engine(13, 500, 250, 50); f = e2.load_fasta(str(p), 0); e2.build_index()
    assert f.names == [r.name for r in recs] and np.array_equal(f.offsets, offs) and np.array_equal(f.bases, bases)
    assert e1.segment_info() == e2.segment_info()
    for d in (0, 1):
        c1, o1, p1 = e1.index(d)
        c2, o2, p2 = e2.index(d)
        assert np.array_equal(c1, c2) and np.array_equal(o1, o2) and np.array_equal(p1, p2)
        assert e1.select(d, 30, 10, m.SELECT_AUTO).tobytes() == e2.select(d, 30, 10, m.SELECT_AUTO).tobytes()
    e1.close(); e2.close(); f.close()
