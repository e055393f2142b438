"""One GPU's shard of BASELINE configs[4] (12,500 x 30 kb) through the drop-in od-msspe executable, end to end:
FASTA file -> multi-threaded ingest -> index -> greedy selection (1000 iterations) -> thermo filters -> all-pairs
cross dimers -> vertex cover -> coverage report -> CSV.  Prints wall times; run on the GPU box."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import numpy as np
from msspe_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
t = time.time()
g = synth.synth_genomes(n, 30000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
path = "/tmp/shard.fa"
with open(path, "wb") as f:
    for i in range(n):
        f.write(b">g%d synthetic\n" % i); f.write(g[i].tobytes()); f.write(b"\n")
print("synth + write %.1f s, %.0f MB" % (time.time() - t, os.path.getsize(path) / 1e6), flush=True)
exe = os.path.join(ROOT, "open-msspe-design_b200", "bin", "od-msspe")
for rep in range(2):
    t = time.time()
    r = subprocess.run([exe, "-i", path, "-o", "/tmp/shard.csv", "--do-align=false"], capture_output=True, text=True, env=dict(os.environ, RUST_LOG="info"))
    dt = time.time() - t
    print("od-msspe exit %d, wall %.2f s" % (r.returncode, dt))
    print(r.stdout.strip()); print(r.stderr.strip())
print(open("/tmp/shard.csv").read().count("\n") - 1, "primers in the CSV")
