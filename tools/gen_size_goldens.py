#!/usr/bin/env python3
"""gen_size_goldens.py -- oracle-pinned goldens of the greedy loop at BASELINE.json sizes.

Runs the CPU oracle (oracle/liboracle.so: `oracle_select`, the C++ restatement of od-msspe/src/main.rs:285-406 with
the reference's own containers and a FULL recount per iteration) offline on the synthetic inputs bench.py and the GPU
tests use, and writes winners / frequencies / tie counts / f32 tie scores (bit patterns) / reference-equivalent evals
to tests/golden/<name>_candidates.json.  The GPU tests compare EVERY loop variant of the CUDA engine with these files
(not with each other).  Minutes of CPU per file; run once, commit the output.

  python tools/gen_size_goldens.py cfg2 cfg3 cfg5shard        # all (directions run as parallel processes)
  python tools/gen_size_goldens.py --one cfg3 1               # one (workload, direction) -> partial file

Workloads (msspe_b200/synth.py; the seeds and mutation rates are frozen there):
  cfg2       1,000 x 30 kb, k=13, max_mismatch_segments auto (=10), 1000 iterations        BASELINE configs[1]
  cfg3       10,000 x 11 kb, k=15, --max-mismatch-segments=2, 1000 iterations              BASELINE configs[2]
  cfg5shard  12,500 x 30 kb (one GPU's eighth of configs[4]), k=13, mms 10, first 100 iterations per direction
  bitmask2m  2,000 x 30 kb cut into 1,000 windows of 30 (search 20): 2.0 M segments, more than a shared-memory bitmask holds
  cfg5part   100,000 genomes x columns [0, 1000) of the cfg5 alignment (3 partitions: one partition range of the
             column-sharded multi-GPU job), k=13, mms 10, first 60 iterations per direction
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))

import numpy as np  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    "cfg2": dict(synth="cfg2", max_iter=1000, mms=10),
    "cfg3": dict(synth="cfg3", max_iter=1000, mms=2),
    "cfg5shard": dict(synth="cfg5", n=12_500, max_iter=100, mms=10),
    "cfg5part": dict(synth="cfg5", cols=1000, max_iter=60, mms=10),
    "bitmask2m": dict(synth=dict(n=2000, length=30_000, seed=11, clades=16, p_clade=0.08, p_leaf=0.01, k=13), wsw=(30, 30, 20),
                      max_iter=60, mms=10),
}


def make_input(name: str):
    """The genomes of a workload, exactly as the GPU tests / bench.py generate them."""
    from msspe_b200 import synth
    w = WORKLOADS[name]
    c = dict(synth.CONFIGS[w["synth"]]) if isinstance(w["synth"], str) else dict(w["synth"])
    k = c.pop("k")
    if "n" in w:
        c["n"] = w["n"]
    g = synth.synth_genomes(**c)
    if "cols" in w:
        g = np.ascontiguousarray(g[:, :w["cols"]])
    return g, k, w["max_iter"], w["mms"], w.get("wsw", (500, 250, 50))


def run_one(name: str, direction: int) -> dict:
    from oracle import oracle as O
    from msspe_b200 import synth
    O.build()
    g, k, max_iter, mms, (W, S, w) = make_input(name)
    fa = synth.to_fasta(g)
    t0 = time.time()
    r = O.select(fa, W, S, w, k, direction, max_iter, mms)
    return {"codes": [int(x) for x in r["codes"]], "freqs": [int(x) for x in r["freqs"]],
            "n_tied": [int(x) for x in r["n_tied"]],
            "score_bits": [int(x) for x in r["scores"].view(np.uint32)],
            "evals": int(r["evals"]), "oracle_seconds": round(time.time() - t0, 1)}


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--one":
        name, d = sys.argv[2], int(sys.argv[3])
        out = run_one(name, d)
        with open(os.path.join(GOLDEN, ".%s_dir%d.partial.json" % (name, d)), "w") as f:
            json.dump(out, f)
        return
    names = sys.argv[1:] or list(WORKLOADS)
    procs = []
    for name in names:
        for d in (0, 1):
            procs.append((name, d, subprocess.Popen([sys.executable, os.path.abspath(__file__), "--one", name, str(d)])))
    for name, d, p in procs:
        if p.wait() != 0:
            raise SystemExit("oracle run failed: %s dir %d" % (name, d))
    for name in names:
        g, k, max_iter, mms, (W, S, w) = make_input(name)
        doc = {"generator": "tools/gen_size_goldens.py (oracle/kmer_oracle.cpp oracle_select = main.rs:285-406, full recount per iteration)",
               "workload": name, "synth": WORKLOADS[name], "genomes": int(g.shape[0]), "length": int(g.shape[1]), "kmer_size": k,
               "window": W, "step": S, "search": w, "max_iterations": max_iter, "max_mismatch_segments": mms, "dirs": []}
        for d in (0, 1):
            part = os.path.join(GOLDEN, ".%s_dir%d.partial.json" % (name, d))
            with open(part) as f:
                doc["dirs"].append(json.load(f))
            os.remove(part)
        with open(os.path.join(GOLDEN, "%s_candidates.json" % name), "w") as f:
            json.dump(doc, f, separators=(",", ":"))
            f.write("\n")
        print(name, "fwd", len(doc["dirs"][0]["codes"]), "rev", len(doc["dirs"][1]["codes"]),
              "evals", doc["dirs"][0]["evals"] + doc["dirs"][1]["evals"], flush=True)


if __name__ == "__main__":
    main()
