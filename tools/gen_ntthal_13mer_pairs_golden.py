#!/usr/bin/env python3
"""Hot loop #2 at od-msspe's own shape: 2000 ordered pairs of random 13-mers through the REFERENCE'S OWN ntthal executable
(od-msspe/bin/ntthal under tools/a64emu) exactly as delta_g.rs:93-110 runs it - `-a ANY -mv 50.00 -dv 3.00 -n 0.00 -d 250.00
-t 25.00 -i`, pairs on stdin - keeping line 0 of every block (dS, dH, dG, t as printed; the parser reads dG, delta_g.rs:33-36)
or null for a pair that printed nothing.  Writes tests/golden/ntthal_emulated_13mer_pairs.json.

Run here (needs /root/reference; ~1 minute on 8 cores):  python tools/gen_ntthal_13mer_pairs_golden.py
"""
import json
import multiprocessing as mp
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "a64emu"))
from emu import run_ntthal  # noqa: E402

ARGS = ["-a", "ANY", "-mv", "50.00", "-dv", "3.00", "-n", "0.00", "-d", "250.00", "-t", "25.00", "-i"]
N, K, SESSION = 2000, 13, 50


def pairs():
    r = random.Random(20261021)
    rnd = lambda alphabet="ACGT": "".join(r.choice(alphabet) for _ in range(K))  # noqa: E731
    out = []
    for i in range(N):
        if i % 40 == 39:                       # {A,C}-only against {A,C}-only: nothing can pair
            out.append((rnd("AC"), rnd("AC")))
        elif i % 40 == 19:                     # a primer with itself
            a = rnd()
            out.append((a, a))
        else:
            out.append((rnd(), rnd()))
    return out


def session(chunk):
    """One -i session; structure-less pairs print nothing, so each pair is also run alone when the block count is short."""
    out, _, code, _ = run_ntthal(ARGS, stdin="".join("%s,%s\n" % p for p in chunk).encode())
    assert code == 0
    lines = out.split("\n")[:-1]
    blocks = [lines[i] for i in range(0, len(lines), 5)]
    if len(blocks) == len(chunk):
        return [b.split() for b in blocks]
    res = []
    for p in chunk:                            # find out which pairs were silent
        o, _, code, _ = run_ntthal(ARGS, stdin=("%s,%s\n" % p).encode())
        assert code == 0
        res.append(o.split("\n")[0].split() if o else None)
    assert [x for x in res if x is not None] == [b.split() for b in blocks]
    return res


def main():
    ps = pairs()
    chunks = [ps[i:i + SESSION] for i in range(0, N, SESSION)]
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        res = [x for r in pool.map(session, chunks) for x in r]
    rows = []
    for (a, b), tok in zip(ps, res):
        rows.append([a, b] + ([tok[7], tok[10], tok[13], tok[16]] if tok else [None] * 4))
    doc = {"source": "od-msspe/bin/ntthal (Primer3 2.6.1, Mach-O arm64) under tools/a64emu; argv " + " ".join(ARGS),
           "columns": ["a", "b", "dS", "dH", "dG", "t"], "pairs": rows}
    path = os.path.join(HERE, "..", "tests", "golden", "ntthal_emulated_13mer_pairs.json")
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
        f.write("\n")
    silent = sum(1 for r in rows if r[2] is None)
    print("wrote %s: %d pairs, %d silent, dG < -9000: %d, dG < -3000: %d" %
          (os.path.relpath(path), len(rows), silent, sum(1 for r in rows if r[4] and float(r[4]) < -9000),
           sum(1 for r in rows if r[4] and float(r[4]) < -3000)))


if __name__ == "__main__":
    main()
