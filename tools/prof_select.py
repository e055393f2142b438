import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
g,k = synth.make_config('cfg2')
eng = m.Engine(k,500,250,50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
for rep in range(2):
    a,b = eng.select_both(1000,10,0)
print(len(a), len(b))
