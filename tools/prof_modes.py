"""Recount (persistent) vs incremental (launch per phase) greedy loop on inputs with more segments than a shared-memory bitmask holds."""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
g = synth.synth_genomes(n, 30000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
eng = m.Engine(13,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
print('segments', eng.segment_info(), 'build ms', eng.timing().encode_ms, eng.timing().index_ms, flush=True)
res = {}
for mode, name in ((0,'recount'),(1,'incremental')):
    for rep in range(2):
        t=time.time(); a,b = eng.select_both(iters, 10, mode); dt=time.time()-t
    res[name]=(a.tobytes(), b.tobytes())
    print(name, 'select wall ms', dt*1e3, 'iters', len(a), len(b), flush=True)
assert res['recount'] == res['incremental']
print('identical')
