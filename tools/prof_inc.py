import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import msspe_b200 as m
from msspe_b200 import synth
g,k = synth.make_config('cfg2')
eng = m.Engine(k,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
for iters in (1, 2, 50, 200, 600):
    for mode in (0, 1):
        for rep in range(3):
            t=time.time(); a,b = eng.select_both(iters, 10, mode); dt=time.time()-t
        print('iters', iters, 'mode', mode, 'wall ms %.3f' % (dt*1e3), len(a), len(b))
