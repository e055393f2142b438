"""Race hunting without a sanitizer: the greedy loop repeated many times on the same index (and across compaction
thresholds and block shapes) must return byte-identical winners, tie counts, f32 scores and evals every time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for name, mms in (("cfg2", 10), ("cfg3", 2)):
    g, k = synth.make_config(name)
    eng = m.Engine(k, 500, 250, 50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
    ref = None
    for rep in range(reps):
        os.environ["MSSPE_COMPACT_MIN"] = ["1048576", "65536", "0", "300000"][rep % 4]
        if rep % 5 == 4: os.environ["MSSPE_PERSIST_1024"] = "1"
        else: os.environ.pop("MSSPE_PERSIST_1024", None)
        a, b = eng.select_both(1000, mms, 0)
        sig = (a.tobytes(), b.tobytes(), tuple(eng.timing().select_evals), tuple(eng.timing().select_iterations))
        if ref is None: ref = sig
        assert sig == ref, "run %d of %s differs" % (rep, name)
    print(name, "identical over", reps, "runs:", len(a), len(b), "winners, evals", ref[2])
    eng.close()
