import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, torch, msspe_b200 as m
from msspe_b200 import synth
g,k = synth.make_config('cfg2'); offs = synth.offsets_for(g)
dev = torch.from_numpy(g.reshape(-1)).pin_memory().to('cuda')
eng = m.Engine(k,500,250,50); fcfg = m.default_filter_cfg()
acc = {}
def T(name, f):
    torch.cuda.synchronize(); t=time.perf_counter(); r=f(); torch.cuda.synchronize(); acc[name]=acc.get(name,0)+time.perf_counter()-t; return r
for rep in range(8):
    if rep == 3: acc.clear()
    T('load', lambda: eng.load_genomes_device(dev.data_ptr(), offs, keepalive=dev))
    T('build', lambda: eng.build_index())
    fwd, rev = T('select', lambda: eng.select_both(1000, 10, 0))
    T('stats_fwd', lambda: eng.kmer_stats(fwd['code'], fcfg))
    T('stats_rev', lambda: eng.kmer_stats(rev['code'], fcfg))
tm = eng.timing()
print({k: round(v/5*1e3,3) for k,v in acc.items()}, 'sum', round(sum(acc.values())/5*1e3,3))
print('engine timers: encode', tm.encode_ms, 'index', tm.index_ms, 'select', tm.select_ms[0], 'thermo', tm.thermo_ms)
