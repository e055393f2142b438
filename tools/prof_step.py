#!/usr/bin/env python3
"""Wall time of every C-ABI call of one bench step (cfg3 by default): python tools/prof_step.py [cfg] [reps]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import numpy as np, torch
import msspe_b200 as m
from msspe_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g, k = synth.make_config(name)
mms = 2 if name == "cfg3" else min(10, max(1, -(-g.shape[0] // 50)))
offs = synth.offsets_for(g)
dev = torch.device("cuda", 0)
hp = torch.from_numpy(g.reshape(-1)).pin_memory()
db = hp.to(dev)
eng = m.Engine(k, 500, 250, 50)
eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
fcfg = m.default_filter_cfg()
for r in range(reps):
    t = [time.perf_counter()]
    eng.load_genomes_device(db.data_ptr(), offs, keepalive=db); t.append(time.perf_counter())
    eng.build_index(); t.append(time.perf_counter())
    a, b = eng.select_both(1000, mms, m.SELECT_AUTO); t.append(time.perf_counter())
    st = eng.kmer_stats_both(a["code"], b["code"], fcfg); t.append(time.perf_counter())
    tm = eng.timing()
    d = [1e3 * (t[i + 1] - t[i]) for i in range(4)]
    print("%s rep %d: load %.3f | build %.3f (dev %.3f) | select %.3f (dev %.3f) | stats %.3f (dev %.3f) | total %.3f ms" % (
        name, r, d[0], d[1], tm.encode_ms + tm.index_ms, d[2], tm.select_ms[0], d[3], tm.thermo_ms, sum(d)), flush=True)
eng.close()
