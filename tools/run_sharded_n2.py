"""torchrun --nproc-per-node N tools/run_sharded_n2.py [cfg] : genome-sharded exact greedy selection over N GPUs,
checked on rank 0 against the single-GPU fused loop on the whole input."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import numpy as np, torch, torch.distributed as dist
import msspe_b200 as m
from msspe_b200 import synth, distributed as D

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g, k = synth.make_config(cfg)
n = g.shape[0]
mms = min(10, max(1, -(-n // 50)))
lo, hi = D.row_block(n, rank, world)
eng = m.Engine(k, 500, 250, 50, device=lr)
eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
eng.load_genomes(g[lo:hi].reshape(-1), synth.offsets_for(g[lo:hi]))
eng.build_index()
res = []
for d in (0, 1):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    cand, evals, iters = D.select_sharded(eng, d, max_iter, mms, dist, dev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res.append((cand, evals, iters, dt))
if rank == 0:
    full = m.Engine(k, 500, 250, 50, device=lr)
    full.load_genomes(g.reshape(-1), synth.offsets_for(g))
    full.build_index()
    ok = True
    for d in (0, 1):
        want = full.select(d, max_iter, mms, 0)
        same = want.tobytes() == res[d][0].tobytes() and full.timing().select_evals[d] == res[d][1]
        ok &= same
        print(json.dumps({"cfg": cfg, "world": world, "direction": d, "identical_to_single_gpu": bool(same), "winners": len(want),
                          "iterations": res[d][2], "evals": res[d][1], "sharded_seconds": res[d][3],
                          "single_gpu_ms": float(full.timing().select_ms[d])}))
    assert ok
dist.barrier()
dist.destroy_process_group()
