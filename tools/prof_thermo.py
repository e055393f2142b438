import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
eng = m.Engine(13,500,250,50)
for n in (600, 1200, 2000):
    codes = synth.random_primers(n, 13, 5)
    for rep in range(3):
        t=time.time(); st = eng.kmer_stats(codes); dt=time.time()-t
    print('n', n, 'kmer_stats wall ms %.3f' % (dt*1e3), 'device thermo ms %.3f' % eng.timing().thermo_ms, 'sum hairpin', float(st['hairpin_th'].sum()))
