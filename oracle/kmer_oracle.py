"""kmer_oracle.py -- TEST INFRASTRUCTURE ONLY (oracle).

Pure-Python restatement of the reference's k-mer engine and host glue, written to be obviously
equal to the Rust (same containers: str words, dict/set keyed by (word, direction), a full recount
of every live segment per greedy iteration).  Slow by construction; used on small cases and on the
Zika fixture, and to cross-check the C++ port in oracle/kmer_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.

Every function cites the reference lines it follows (od-msspe/src/...).  Pinning: the seven unit
tests of main.rs:863-1236 are replayed in tests/test_oracle_kmer.py.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

SEQ_DIR_FWD = 0  # constants.rs:22
SEQ_DIR_REV = 1  # constants.rs:23


def f32(x: float) -> float:
    """Round a Python float to the nearest IEEE binary32 (Rust `as f32` / f32 arithmetic result)."""
    return struct.unpack("f", struct.pack("f", x))[0]


@dataclass
class SequenceRecord:  # main.rs:21-24
    name: str
    sequence: str


@dataclass
class Segment:  # main.rs:82-87
    record: int  # index into records (the Rust holds a reference)
    partition_no: int  # u16
    index: int
    kmers: list = field(default_factory=lambda: [[], []])  # [fwd words, rev words] in first-seen order


def to_records(src: bytes) -> list[SequenceRecord]:
    """main.rs:108-122 with seq_io 0.3.2 semantics: id() = header up to the first space, full_seq() =
    all sequence lines joined; then to_uppercase().replace("U","T")."""
    records = []
    name, chunks = None, []
    for raw in src.split(b"\n"):
        line = raw.rstrip(b"\r")
        if line.startswith(b">"):
            if name is not None:
                records.append(SequenceRecord(name, b"".join(chunks).decode().upper().replace("U", "T")))
            head = line[1:].decode()
            name = head.split(" ")[0]
            chunks = []
        elif name is not None:
            chunks.append(line)
    if name is not None:
        records.append(SequenceRecord(name, b"".join(chunks).decode().upper().replace("U", "T")))
    return records


def reverse_complement(s: str) -> str:  # main.rs:148-161
    m = {"A": "T", "T": "A", "U": "A", "C": "G", "G": "C"}
    return "".join(m.get(c, c) for c in reversed(s))


def find_kmers(sequence: str, k: int) -> list[str]:
    """main.rs:163-171: all n-grams of length k (ngrams crate on an iterator: no padding, pinned by
    main.rs:882-894), keep those made only of "ATCGU", itertools unique() = first occurrence order."""
    seen, out = set(), []
    for i in range(0, len(sequence) - k + 1):
        w = sequence[i:i + k]
        if all(c in "ATCGU" for c in w) and w not in seen:
            seen.add(w)
            out.append(w)
    return out


def partitioning_sequence(sequence: str, size: int, step: int) -> list[str]:
    """main.rs:173-181: slice.windows(size).step_by(step) -- full windows only; step 0 panics."""
    if step == 0:
        raise ValueError("step_by(0) panics")
    if size == 0:
        raise ValueError("windows(0) panics")
    return [sequence[s:s + size] for s in range(0, len(sequence) - size + 1, step)] if len(sequence) >= size else []


def get_sequence_on_search_windows(sequence: str, w: int) -> tuple[str, str]:  # main.rs:183-187
    return sequence[:w], sequence[len(sequence) - w:]


def get_segment_manager(records, segment_size, overlap_size, window_size, kmer_size) -> list[Segment]:
    """main.rs:196-235.  Returns SegmentManager.segments."""
    if overlap_size < window_size:
        raise ValueError("Overlap windows size must be greater or equal than search windows size")
    segments: list[Segment] = []
    for r, record in enumerate(records):
        for j, part in enumerate(partitioning_sequence(record.sequence, segment_size, overlap_size)):
            start, end = get_sequence_on_search_windows(part, window_size)
            seg = Segment(r, j & 0xFFFF, len(segments))
            seg.kmers[0] = find_kmers(start, kmer_size)
            seg.kmers[1] = [reverse_complement(x) for x in find_kmers(end, kmer_size)]
            segments.append(seg)
    return segments


def make_kmer_segments_windows_mapping(segments) -> dict:
    """main.rs:237-255: (word, direction) -> ascending list of segment indices."""
    m: dict = {}
    for seg in segments:
        for d in (0, 1):
            for wd in seg.kmers[d]:
                m.setdefault((wd, d), []).append(seg.index)
    return m


def partition_tie_score(key, kmer_to_segments, segments, ignored, partition_coverage) -> float:
    """main.rs:261-283: sequential f32 accumulation in postings order over first-seen live partitions."""
    seen = set()
    score = f32(0.0)
    for idx in kmer_to_segments.get(key, []):
        if idx in ignored:
            continue
        p = segments[idx].partition_no
        if p not in seen:
            seen.add(p)
            already = partition_coverage.get(p, 0)
            term = f32(f32(1.0) / f32(f32(float(already)) + f32(1.0)))
            score = f32(score + term)
    return score


def find_most_freq_kmer(segments, direction, ignored, kmer_to_segments, partition_coverage):
    """main.rs:285-329.  Returns (word, freq, n_tied, score) or None."""
    freq: dict = {}
    for idx, seg in enumerate(segments):
        if idx in ignored:
            continue
        for wd in seg.kmers[direction]:
            freq[wd] = freq.get(wd, 0) + 1
    if not freq:
        return None
    max_freq = max(freq.values())
    best = None
    n_tied = 0
    for wd, f in freq.items():
        if f != max_freq:
            continue
        n_tied += 1
        s = partition_tie_score((wd, direction), kmer_to_segments, segments, ignored, partition_coverage)
        # max_by(s1.partial_cmp(s2).then(k2.word.cmp(k1.word))): higher score, then smaller word
        if best is None or s > best[1] or (s == best[1] and wd < best[0]):
            best = (wd, s)
    return best[0], max_freq, n_tied, best[1]


def find_candidates_kmers(segments, direction, max_iterations, max_mismatch_segments, trace=None):
    """main.rs:331-406.  Returns [(word, freq)] in selection order (empty list == the Rust's None)."""
    out = []
    kmer_to_segments = make_kmer_segments_windows_mapping(segments)
    ignored: set = set()
    partition_coverage: dict = {}
    for _ in range(max_iterations):
        r = find_most_freq_kmer(segments, direction, ignored, kmer_to_segments, partition_coverage)
        if r is None:
            break
        wd, fq, n_tied, score = r
        if fq == 1:
            break
        out.append((wd, fq))
        if trace is not None:
            trace.append((wd, fq, n_tied, score))
        newly = set()
        for idx in kmer_to_segments[(wd, direction)]:
            ignored.add(idx)
            newly.add(segments[idx].partition_no)
        for p in newly:
            partition_coverage[p] = partition_coverage.get(p, 0) + 1
        if fq < max_mismatch_segments:
            break
    return out


def auto_max_mismatch_segments(n_records: int) -> int:
    """main.rs:658-660: records.len().div_ceil(50).clamp(1, 10)."""
    return min(10, max(1, -(-n_records // 50)))


def is_run(kmer: str) -> bool:  # main.rs:478-490
    runs, last = 0, " "
    for c in kmer:
        runs = runs + 1 if c == last else 0
        last = c
    return runs >= 5


def encode(word: str) -> int:
    """2-bit big-endian code (A0 C1 G2 T3); order-preserving for equal-length words."""
    v = 0
    for c in word:
        v = (v << 2) | "ACGT".index(c)
    return v


def decode(code: int, k: int) -> str:
    return "".join("ACGT"[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))


def coverage_report(fwd_words, rev_words, segments, records) -> str:
    """main.rs:518-594; returns the text println! would write (HashMap order never reaches the output)."""
    sf, sr = set(fwd_words), set(rev_words)
    covered = set()
    for seg in segments:
        if any(w in sf for w in seg.kmers[0]) or any(w in sr for w in seg.kmers[1]):
            covered.add(seg.index)
    total = len(segments)
    seq_stats: dict = {}
    part_stats: dict = {}
    for seg in segments:
        nm = records[seg.record].name
        se = seq_stats.setdefault(nm, [0, 0])
        se[1] += 1
        pe = part_stats.setdefault(seg.partition_no, [0, 0])
        pe[1] += 1
        if seg.index in covered:
            se[0] += 1
            pe[0] += 1
    covs = [f32(f32(f32(float(c)) / f32(float(t))) * f32(100.0)) for c, t in seq_stats.values()]
    min_cov = min(covs) if covs else math.inf
    max_cov = max(covs) if covs else -math.inf
    well = sum(1 for c in covs if c >= 80.0)
    unc = sorted(p for p, (c, _) in part_stats.items() if c == 0)
    pct = f32(f32(100.0) * f32(float(len(covered)))) / f32(float(total)) if total else math.nan
    pct = f32(pct)
    lines = ["", "Coverage report:",
             "  Segments:  %d/%d covered (%s%%)" % (len(covered), total, rust_fmt(pct, 1)),
             "  Sequences: %d/%d at ≥80%% coverage (min %s%%, max %s%%)" % (
                 well, len(seq_stats), rust_fmt(min_cov, 1), rust_fmt(max_cov, 1))]
    if not unc:
        lines.append("  All partitions have primer coverage")
    else:
        lines.append("  Uncovered partitions: [%s]" % ", ".join(str(p) for p in unc))
    return "\n".join(lines) + "\n"


def rust_fmt(x: float, prec: int) -> str:
    """Rust `{:.N}` on an f32: exact decimal expansion of the f32, round-half-even on the exact value."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    from decimal import Decimal, ROUND_HALF_EVEN
    q = Decimal(1).scaleb(-prec)
    d = Decimal(x).quantize(q, rounding=ROUND_HALF_EVEN)
    s = format(d, "f")
    if d == 0 and math.copysign(1.0, x) < 0:
        s = "-" + s.lstrip("-")
    return s


def greedy_vertex_cover(words: list[str], edges: list[tuple[str, str]]) -> set[str]:
    """main.rs:754-798: conflicts[a] += b, conflicts[b] += a for every conflict edge (a == b is a self conflict);
    then repeatedly delete the not-yet-deleted primer with the most not-yet-deleted neighbours (count > 0), ties ->
    lexicographically greatest word (`c1.cmp(c2).then(p1.cmp(p2))` under max_by)."""
    conflicts: dict[str, set[str]] = {}
    for a, b in edges:
        conflicts.setdefault(a, set()).add(b)
        conflicts.setdefault(b, set()).add(a)
    deleted: set[str] = set()
    while True:
        worst = None
        for p, nbs in conflicts.items():
            if p in deleted:
                continue
            active = sum(1 for n in nbs if n not in deleted)
            if active > 0 and (worst is None or (active, p) > worst):
                worst = (active, p)
        if worst is None:
            return deleted
        deleted.add(worst[1])
