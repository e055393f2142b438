#!/usr/bin/env python3
"""Times the greedy loop variants on a synthetic config (default cfg2): python tools/prof_part.py [cfg2|cfg3|cfg5shard] [reps]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import numpy as np
import msspe_b200 as m
from msspe_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if name == "cfg5shard":
    g, k, it, mms = synth.synth_genomes(12_500, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01), 13, 1000, 10
elif name == "cfg5":
    g, k, it, mms = synth.make_config("cfg5")[0], 13, 1000, 10
else:
    g, k = synth.make_config(name)
    it, mms = 1000, (2 if name == "cfg3" else min(10, max(1, -(-g.shape[0] // 50))))
eng = m.Engine(k, 500, 250, 50)
eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
t0 = time.perf_counter(); eng.build_index(); t1 = time.perf_counter()
tm = eng.timing()
print("%s: build wall %.2f ms (encode %.2f, index %.2f)" % (name, 1e3 * (t1 - t0), tm.encode_ms, tm.index_ms))
ref = None
modes = [(m.SELECT_PARTITIONED, "partitioned"), (m.SELECT_INCREMENTAL, "incremental"), (m.SELECT_RECOUNT, "recount"), (m.SELECT_AUTO, "auto")]
if len(sys.argv) > 3:
    modes = [x for x in modes if x[1] in sys.argv[3].split(",")]
for mode, label in modes:
    for r in range(reps):
        t0 = time.perf_counter()
        a, b = eng.select_both(it, mms, mode)
        dt = time.perf_counter() - t0
        tm = eng.timing()
        print("  %-12s rep %d: wall %.3f ms, device %.3f ms, winners %d/%d, evals %d, launches %d" % (
            label, r, 1e3 * dt, tm.select_ms[0], len(a), len(b), tm.select_evals[0] + tm.select_evals[1], tm.kernel_launches), flush=True)
    key = (a.tobytes(), b.tobytes(), tuple(tm.select_evals))
    if ref is None:
        ref = key
    print("  %-12s identical to the first mode: %s" % (label, key == ref))
eng.close()
