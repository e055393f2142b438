"""Phase breakdown of the persistent greedy kernel (MSSPE_DEBUG_TIMERS) and the tie statistics of a configuration."""
import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
os.environ['MSSPE_DEBUG_TIMERS'] = '1'
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
g,k = synth.make_config(cfg)
eng = m.Engine(k,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
mi, mms = (1000, 2) if cfg == 'cfg3' else (1000, 10)
for mode in (0, 1):
    for rep in range(3):
        t=time.time(); a,b = eng.select_both(mi, mms, mode); dt=time.time()-t
    nt = np.concatenate([a['n_tied'], b['n_tied']])
    print('mode', mode, 'wall ms %.3f' % (dt*1e3), 'iters', len(a), len(b), 'n_tied mean %.2f median %d p90 %d max %d' % (nt.mean(), np.median(nt), np.percentile(nt, 90), nt.max()),
          'freq first/last', a['freq'][0], a['freq'][-1], flush=True)
