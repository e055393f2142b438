import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/open-msspe-design_b200')
import numpy as np, torch, msspe_b200 as m
from msspe_b200 import synth
g,k = synth.make_config('cfg2')
offs = synth.offsets_for(g)
hp = torch.from_numpy(g.reshape(-1)).pin_memory()
dev = hp.to('cuda')
eng = m.Engine(k,500,250,50)
def T(f, n=5):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
for rep in range(2):
    print('load host pinned ms', T(lambda: eng.load_genomes(hp.numpy(), offs)), 'h2d_ms', eng.timing().h2d_ms)
    print('load device ms', T(lambda: eng.load_genomes_device(dev.data_ptr(), offs, keepalive=dev)))
    eng.load_genomes(hp.numpy(), offs)
    print('build ms', T(lambda: eng.build_index()))
    print('select ms', T(lambda: eng.select_both(1000, 10, 0)))
    def full(host):
        if host: eng.load_genomes(hp.numpy(), offs)
        else: eng.load_genomes_device(dev.data_ptr(), offs, keepalive=dev)
        eng.build_index(); eng.select_both(1000,10,0)
    print('full host ms', T(lambda: full(True)), 'full dev ms', T(lambda: full(False)))
