"""Scaling probe: cfg5/8 shard (12,500 x 30 kb) through build + a bounded number of greedy iterations."""
import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
t=time.time(); g = synth.synth_genomes(n, 30000, 5, clades=256, p_clade=0.10, p_leaf=0.01); print('synth s', time.time()-t, flush=True)
eng = m.Engine(13,500,250,50)
t=time.time(); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); print('load s', time.time()-t)
t=time.time(); eng.build_index(); print('build s', time.time()-t); tm = eng.timing(); print('encode ms', tm.encode_ms, 'index ms', tm.index_ms)
print(eng.segment_info())
for rep in range(2):
    t=time.time(); a,b = eng.select_both(iters, 10, 0); dt=time.time()-t
    tm = eng.timing()
    ev = tm.select_evals[0]+tm.select_evals[1]; pr = tm.select_postings_read[0]+tm.select_postings_read[1]
    ck = tm.count_kernel_ms[0]+tm.count_kernel_ms[1]
    print('select wall ms', dt*1e3, 'iters', len(a), len(b), 'evals', ev, 'postings read', pr, 'count ms', ck,
          'phys GB/s', 4*pr/ck/1e6 if ck else 0, 'alg GB/s', 4*ev/ck/1e6 if ck else 0)
