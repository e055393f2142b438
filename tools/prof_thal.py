import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 13
pool = synth.random_primers(n, k, 4)
eng = m.Engine(k,500,250,50)
cond = m.ThalCond(50,3,0,250,25.0,30,0)
for rep in range(3):
    t=time.time(); e,nos = eng.cross_dimer(pool, cond, -8999.0, edge_capacity=1<<22, nostruct_capacity=1<<20); dt=time.time()-t
    print('pairs', n*n, 'wall ms', dt*1e3, 'kernel ms', eng.timing().dimer_ms, 'pairs/s (kernel)', n*n/eng.timing().dimer_ms*1e3, 'edges', len(e), 'nostruct', len(nos))
