import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
g,k = synth.make_config('cfg2')
eng = m.Engine(k,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
for d in (0,1):
    codes, offs, post = eng.index(d); ln = np.diff(offs.astype(np.int64))
    print('dir',d,'D',len(codes),'R',len(post),'maxlist',ln.max(),'mean',ln.mean())
for bps in (1,2,3,4):
    os.environ['MSSPE_PERSIST_BLOCKS_PER_SM']=str(bps)
    for rep in range(2):
        t=time.time(); a,b = eng.select_both(1000,10,0); dt=time.time()-t
    print('bps',bps,'wall ms',dt*1e3,'ntied mean',a['n_tied'].mean(), 'max', a['n_tied'].max(), 'iters', len(a), len(b))
