import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
g = synth.synth_genomes(n, 30000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
eng = m.Engine(13,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g))
for rep in range(3):
    t=time.time(); eng.build_index(); dt=time.time()-t
    print('build wall ms %.1f' % (dt*1e3), 'encode', eng.timing().encode_ms, 'index', eng.timing().index_ms, flush=True)
