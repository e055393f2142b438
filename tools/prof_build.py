import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
g = synth.synth_genomes(n, 30000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
eng = m.Engine(13,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
best = (1e9, 1e9)
for rep in range(6):
    t=time.time(); eng.build_index(); dt=time.time()-t
    best = min(best, (eng.timing().index_ms, eng.timing().encode_ms))
print('best of 6 builds: index ms %.2f encode ms %.2f' % best, flush=True)
