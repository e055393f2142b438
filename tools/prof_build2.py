#!/usr/bin/env python3
"""Index build (K1 + K2) of a synthetic config, for ncu: python tools/prof_build2.py [cfg3|cfg2|cfg5shard] [reps]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import msspe_b200 as m
from msspe_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if name == "cfg5shard":
    g, k = synth.synth_genomes(12_500, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01), 13
else:
    g, k = synth.make_config(name)
eng = m.Engine(k, 500, 250, 50)
eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
for r in range(reps):
    t0 = time.perf_counter(); eng.build_index(); dt = time.perf_counter() - t0
    tm = eng.timing()
    print("%s build rep %d: wall %.3f ms (encode %.3f, index %.3f), records %d + %d" % (name, r, 1e3 * dt, tm.encode_ms, tm.index_ms, eng.index(0)[1][-1], eng.index(1)[1][-1]), flush=True)
eng.close()
