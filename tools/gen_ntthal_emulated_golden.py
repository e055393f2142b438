#!/usr/bin/env python3
"""Reference output for Primer3's thal in all four alignment modes, produced by the REFERENCE'S OWN `ntthal` executable
(od-msspe/bin/ntthal, Primer3 2.6.1, Mach-O arm64; delta_g.rs:90-108 spawns it) executed under tools/a64emu.

Writes tests/golden/ntthal_emulated.json: a list of {"args": [...], "stdin": "...", "stdout": "..."}.  The emulator is
first made to reproduce the five blocks delta_g.rs:197-230 holds (three ways: compiled-in tables, -path parameter files,
-i stdin mode); the script refuses to write anything otherwise.

What the case list covers (seeded, 13-mers = od-msspe's default k, plus longer oligos):
  ANY / END1 / END2   random pairs, self pairs, perfect duplexes, partial overlaps; od-msspe's conditions, Primer3's
                      check_primers conditions, random salts / DNA concentrations / temperatures, -maxloop 0 .. 30
  HAIRPIN             random oligos (mostly structure-less), stem-loops with 3 .. 9 nt loops (every special triloop /
                      tetraloop family gets hits through random loops), bulges and interior loops in the stem
Run here (the GPU box has no /root/reference):  python tools/gen_ntthal_emulated_golden.py
"""
import json
import multiprocessing as mp
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "a64emu"))
from emu import run_ntthal  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")
PARAMS = "/root/reference/od-msspe/primer3_config"
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def rnd(r, n):
    return "".join(r.choice("ACGT") for _ in range(n))


def cond_args(mv, dv, n, d, t):
    return ["-mv", "%.2f" % mv, "-dv", "%.2f" % dv, "-n", "%.2f" % n, "-d", "%.2f" % d, "-t", "%.2f" % t]


def mutate(r, s, k):
    s = list(s)
    for _ in range(k):
        i = r.randrange(len(s))
        s[i] = r.choice([c for c in "ACGT" if c != s[i]])
    return "".join(s)


def cases():
    r = random.Random(20261019)
    msspe = cond_args(50, 3, 0, 250, 25)          # od-msspe defaults (config.rs), annealing temperature 25
    p3 = cond_args(50, 1.5, 0.6, 50, 37)          # Primer3 check_primers conditions (primer.rs:125-140)
    out = []

    def salts():
        return cond_args(r.choice([10, 50, 100, 200]), r.choice([0, 0.5, 1.5, 3, 5]), r.choice([0, 0.2, 0.6, 0.8, 1.5]),
                         r.choice([10, 50, 250, 500]), r.choice([20, 25, 37, 50, 60, 65]))
    # --- dimers
    for mode in ("ANY", "END1", "END2"):
        for i in range(60):                       # random 13-mer pairs
            a, b = rnd(r, 13), rnd(r, 13)
            out.append((["-a", mode] + (msspe if i % 2 else p3) + ["-s1", a, "-s2", b], ""))
        for i in range(20):                       # self pairs (the symmetric concentration term) and palindromes
            a = rnd(r, 13) if i % 2 else (lambda h: h + rc(h))(rnd(r, r.choice([5, 6, 7, 8])))
            out.append((["-a", mode] + (msspe if i % 4 < 2 else p3) + ["-s1", a, "-s2", a], ""))
        for i in range(20):                       # duplexes: perfect, mismatched, shifted
            a = rnd(r, r.choice([13, 15, 18, 22, 26, 30]))
            b = rc(a)
            if i % 3 == 1:
                b = mutate(r, b, r.choice([1, 2, 3]))
            elif i % 3 == 2:
                sh = r.randrange(1, 6)
                b = rnd(r, sh) + b[:-sh]
            out.append((["-a", mode] + salts() + ["-s1", a, "-s2", b], ""))
        for i in range(20):                       # mixed lengths, random salts, maxloop
            a, b = rnd(r, r.randrange(8, 33)), rnd(r, r.randrange(8, 33))
            extra = ["-maxloop", str(r.choice([0, 3, 8, 15, 30]))] if i % 2 else []
            out.append((["-a", mode] + salts() + extra + ["-s1", a, "-s2", b], ""))
    # --- hairpins
    for i in range(40):
        out.append((["-a", "HAIRPIN"] + (p3 if i % 2 else msspe) + ["-s1", rnd(r, r.choice([13, 13, 15, 20, 25]))], ""))
    for i in range(120):
        stem = rnd(r, r.randrange(3, 9))
        loop = rnd(r, 3 + i % 7)
        left, right = stem, rc(stem)
        kind = i % 5
        if kind == 1:                             # bulge
            p = r.randrange(1, len(left))
            left = left[:p] + rnd(r, r.choice([1, 2, 3])) + left[p:]
        elif kind == 2:                           # interior loop
            p = r.randrange(1, len(left))
            q = len(right) - p
            left = left[:p] + rnd(r, r.choice([1, 2])) + left[p:]
            right = right[:q] + rnd(r, r.choice([1, 2, 3])) + right[q:]
        elif kind == 3:                           # mismatch in the stem
            right = mutate(r, right, 1)
        s = rnd(r, r.randrange(0, 5)) + left + loop + right + rnd(r, r.randrange(0, 5))
        c = p3 if i % 3 == 0 else msspe if i % 3 == 1 else salts()
        extra = ["-maxloop", str(r.choice([3, 8, 30]))] if i % 10 == 9 else []
        out.append((["-a", "HAIRPIN"] + c + extra + ["-s1", s], ""))
    # --- the reference's own way of calling it: -path <dir> -i, pairs on stdin (delta_g.rs:93-110)
    for i in range(6):
        pairs = [(rnd(r, 13), rnd(r, 13)) for _ in range(8)]
        out.append((["-a", "ANY"] + msspe + ["-path", "primer3_config/", "-i"],
                    "".join("%s,%s\n" % p for p in pairs)))
    return out


def run(case):
    args, stdin = case
    o, e, code, n = run_ntthal(args, stdin=stdin.encode(), file_root=PARAMS if "-path" in args else None)
    return {"args": args, "stdin": stdin, "stdout": o, "stderr": e, "exit": code, "instructions": n}


def self_check():
    gold = json.load(open(os.path.join(GOLDEN, "ntthal_delta_g_rs.json")))

    def block(g):
        v = g["values"]
        head = "Calculated thermodynamical parameters for dimer:\tdS = %s\tdH = %s\tdG = %s\tt = %s\n" % (v["dS"], v["dH"], v["dG"], v["t"])
        w = max(len(b) for _, b in g["lines"])
        return head + "".join("%s\t%s\n" % (t, b.ljust(w)) for t, b in g["lines"])
    for g in gold:
        c = g["cond"]
        ca = cond_args(c["mv"], c["dv"], c["dntp"], c["dna"], c["t"])
        want = block(g)
        got1 = run_ntthal(["-a", "ANY"] + ca + ["-s1", g["a"], "-s2", g["b"]])[0]
        got2 = run_ntthal(["-a", "ANY"] + ca + ["-path", "primer3_config/", "-s1", g["a"], "-s2", g["b"]], file_root=PARAMS)[0]
        got3 = run_ntthal(["-a", "ANY"] + ca + ["-path", "primer3_config/", "-i"], stdin=("%s,%s\n" % (g["a"], g["b"])).encode(),
                          file_root=PARAMS)[0]
        # -s1/-s2 mode ends with one more "\n"; the trailing blanks of the drawing are the executable's own
        for got in (got1, got2, got3):
            if got.rstrip("\n") != want.rstrip("\n"):
                raise SystemExit("emulated ntthal does not reproduce delta_g.rs:197-230:\n%r\n%r" % (got, want))
    print("self check: the five delta_g.rs blocks reproduced (compiled-in tables, -path files, -i)")


def main():
    self_check()
    cs = cases()
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(run, cs, chunksize=4)
    bad = [x for x in res if x["exit"] != 0]
    if bad:
        raise SystemExit("non-zero exit: %r" % bad[:3])
    doc = {"source": "od-msspe/bin/ntthal (Primer3 2.6.1, Mach-O arm64) executed by tools/a64emu; stdout verbatim",
           "cases": [{"args": x["args"], "stdin": x["stdin"], "stdout": x["stdout"]} for x in res]}
    path = os.path.join(GOLDEN, "ntthal_emulated.json")
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
        f.write("\n")
    ns = sum("No secondary structure" in x["stdout"] for x in res)
    print("wrote %s: %d cases (%d structure-less), %d bytes, %.1f M instructions" %
          (os.path.relpath(path), len(res), ns, os.path.getsize(path), sum(x["instructions"] for x in res) / 1e6))


if __name__ == "__main__":
    main()
