import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
g,k = synth.make_config(name)
eng = m.Engine(k,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
mms = 2 if name == 'cfg3' else 10
for rep in range(2):
    t=time.time(); a,b = eng.select_both(1000,mms,0); dt=time.time()-t
    tm=eng.timing()
    print(name, 'select wall ms', dt*1e3, 'iters', len(a), len(b), 'evals', sum(tm.select_evals), 'postings read', sum(tm.select_postings_read), 'count ms', sum(tm.count_kernel_ms))
