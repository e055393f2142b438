"""count_kernel alone (launch-per-phase mode) on cfg2 or a cfg5/8 shard: for ncu instruction counts / timings."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/open-msspe-design_b200')
import numpy as np, msspe_b200 as m
from msspe_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if which == 'cfg2':
    g, k = synth.make_config('cfg2')
else:
    g, k = synth.synth_genomes(int(which), 30000, 5, clades=256, p_clade=0.10, p_leaf=0.01), 13
eng = m.Engine(k, 500, 250, 50); eng.load_genomes(g.reshape(-1), synth.offsets_for(g)); eng.build_index()
m.load_library().msspe_set_profiling(eng.h, 1)
a = eng.select(0, iters, 10, m.SELECT_RECOUNT | m.SELECT_BATCHED)
t = eng.timing()
print(which, 'iters', len(a), 'count ms', t.count_kernel_ms[0], 'postings', t.select_postings_read[0],
      'GB/s', 4 * t.select_postings_read[0] / max(t.count_kernel_ms[0], 1e-9) / 1e6)
