#!/usr/bin/env python3
"""Reference output of `primer3_core` for PRIMER_TASK=check_primers, produced by the REFERENCE'S OWN executable
(od-msspe/bin/primer3_core, Primer3 2.6.1, Mach-O arm64; primer.rs:125-140 spawns it without arguments and writes the Boulder-IO
records of primer.rs:113-127 to its stdin) executed under tools/a64emu.

Writes tests/golden/primer3_core_emulated.json: per primer the PRIMER_LEFT_0_* values the reference parses (primer.rs:67-111:
TM, GC_PERCENT, SELF_ANY_TH, SELF_END_TH, HAIRPIN_TH) plus PENALTY and END_STABILITY, as printed.  The script first checks that
the emulated executable reproduces the one record the reference keeps in its tests (primer.rs:238-250) and refuses to write
anything otherwise.

Run here (the GPU box has no /root/reference):  python tools/gen_primer3_core_emulated_golden.py
"""
import json
import multiprocessing as mp
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "a64emu"))
from emu import run_primer3_core  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
KEYS = ["TM", "GC_PERCENT", "SELF_ANY_TH", "SELF_END_TH", "HAIRPIN_TH", "PENALTY", "END_STABILITY"]


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def rnd(r, n):
    return "".join(r.choice("ACGT") for _ in range(n))


def record(p, min_tm=30.0, max_tm=60.0):
    """format_primer3_input, primer.rs:113-127."""
    return ("SEQUENCE_ID=%s\nSEQUENCE_PRIMER=%s\nPRIMER_TASK=check_primers\nPRIMER_MIN_SIZE=13\nPRIMER_MIN_TM=%.2f\n"
            "PRIMER_MAX_TM=%.2f\nPRIMER_OPT_TM=%.2f\nPRIMER_PICK_ANYWAY=1\n=\n" % (p, p, min_tm, max_tm, max_tm))


def primers():
    r = random.Random(20261020)
    out = ["AGCCCGTGTAAAC"]                                    # primer.rs:238-250
    out += [rnd(r, 13) for _ in range(120)]                    # od-msspe's default k
    out += [rnd(r, k) for k in (14, 15, 16, 18, 20, 22, 25) for _ in range(8)]
    for i in range(70):                                        # stem-loops: HAIRPIN_TH > 0
        stem = rnd(r, r.randrange(3, 6))
        loop = rnd(r, r.randrange(3, 8))
        s = rnd(r, r.randrange(0, 3)) + stem + loop + rc(stem) + rnd(r, r.randrange(0, 3))
        if 13 <= len(s) <= 25:
            out.append(s)
    for i in range(50):                                        # self-complementary stretches: SELF_ANY_TH / SELF_END_TH > 0
        h = rnd(r, r.randrange(3, 7))
        core = h + rc(h)
        pad = 13 - len(core) if len(core) < 13 else r.randrange(0, 4)
        s = (rnd(r, pad) + core) if i % 2 else (core + rnd(r, pad))
        if 13 <= len(s) <= 25:
            out.append(s)
    out += ["ACACACACACACA", "AAAAAAAAAAAAA", "GGGGGGGGGGGGG", "ATATATATATATA", "GCGCGCGCGCGCG"]   # low complexity
    # outside PRIMER_MIN_SIZE=13 .. PRIMER_MAX_SIZE=27: with PRIMER_PICK_ANYWAY=1 the executable warns ("Too short" / "Too long")
    # and still reports every number, so od-msspe --kmer-size 8 .. 12 and 28 .. 32 get values too
    out += [rnd(r, k) for k in (8, 9, 10, 11, 12, 26, 28, 30, 32) for _ in range(4)]
    for k in (28, 30, 32):
        for _ in range(3):
            stem = rnd(r, r.randrange(4, 8))
            loop = rnd(r, r.randrange(3, 9))
            core = stem + loop + rc(stem)
            out.append((rnd(r, k - len(core)) + core)[:k] if len(core) < k else core[:k])
    seen, uniq = set(), []
    for p in out:
        if p not in seen:
            seen.add(p)
            uniq.append(p)
    return uniq


def parse(stdout):
    """Boulder-IO records -> list of dicts."""
    recs, cur = [], {}
    for line in stdout.split("\n"):
        if line == "=":
            recs.append(cur)
            cur = {}
        elif "=" in line:
            k, v = line.split("=", 1)
            cur[k] = v
    return recs


def run(batch):
    o, e, code, n = run_primer3_core("".join(record(p) for p in batch).encode())
    if code != 0:
        raise SystemExit("primer3_core exit %s: %s" % (code, e))
    recs = parse(o)
    assert [x["SEQUENCE_ID"] for x in recs] == batch
    return [{"primer": p, **{k: x.get("PRIMER_LEFT_0_" + k) for k in KEYS}} for p, x in zip(batch, recs)], n


def main():
    kat, _ = run(["AGCCCGTGTAAAC"])
    want = {"TM": "43.727", "GC_PERCENT": "53.846", "SELF_ANY_TH": "0.00", "SELF_END_TH": "0.00", "HAIRPIN_TH": "0.00"}
    if {k: kat[0][k] for k in want} != want:
        raise SystemExit("emulated primer3_core does not reproduce primer.rs:238-250: %r" % kat)
    print("self check: primer.rs:238-250 reproduced")
    ps = primers()
    nb = 16
    batches = [ps[i::nb] for i in range(nb)]
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(run, batches)
    by = {x["primer"]: x for b, _ in res for x in b}
    doc = {"source": "od-msspe/bin/primer3_core (Primer3 2.6.1, Mach-O arm64) executed by tools/a64emu on the Boulder-IO records "
                     "of primer.rs:113-127 (PRIMER_MIN_TM 30.00, PRIMER_MAX_TM = PRIMER_OPT_TM 60.00); values as printed",
           "primers": [by[p] for p in ps]}
    path = os.path.join(GOLDEN, "primer3_core_emulated.json")
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
        f.write("\n")
    pos = {k: sum(1 for x in doc["primers"] if x[k] and float(x[k]) > 0) for k in ("SELF_ANY_TH", "SELF_END_TH", "HAIRPIN_TH")}
    print("wrote %s: %d primers, > 0: %r, %.1f M instructions" % (os.path.relpath(path), len(ps), pos, sum(n for _, n in res) / 1e6))


if __name__ == "__main__":
    main()
