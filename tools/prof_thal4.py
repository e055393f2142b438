#!/usr/bin/env python3
"""A row block of the cfg4 pair matrix (20,000-primer pool) through msspe_cross_dimer_device, for ncu: python tools/prof_thal4.py [rows]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import msspe_b200 as m
from msspe_b200 import synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
pool = synth.random_primers(20_000, 13, 4)
eng = m.Engine(13, 500, 250, 50)
cond = m.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
for rep in range(2):
    t0 = time.perf_counter()
    e, s = eng.cross_dimer_device(pool, cond, -8999.0, 0, rows, edge_capacity=rows * 20_000 // 50, nostruct_capacity=1 << 16)
    dt = time.perf_counter() - t0
    print("rows %d: %d pairs in %.3f s (%.3e pairs/s; kernel %.3f ms), %d edges, %d structure-less" % (rows, rows * 20_000, dt, rows * 20_000 / dt, eng.timing().dimer_ms, e.shape[0], s.shape[0]))
eng.close()
