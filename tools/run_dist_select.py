#!/usr/bin/env python3
"""One design job over N GPUs (torchrun): every rank loads the columns of its partition range, builds its index, and
msspe_select_both_dist must return exactly what one GPU returns for the whole alignment.
  torchrun --nproc-per-node 2 tools/run_dist_select.py [case ...]      cases: small, repeats, cfg2, cfg5shard, cfg3xN, cfg5xN, tiny_k"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
import numpy as np
import torch, torch.distributed as dist
import msspe_b200 as m
from msspe_b200 import synth, distributed as D

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def case(name):
    W, S, w = 500, 250, 50
    if name == "small":
        return synth.synth_genomes(200, 30_000, 2, clades=8, p_clade=0.08, p_leaf=0.01, gap_rate=1e-4), 13, 400, 4, (W, S, w)
    if name == "repeats":   # a block repeated in distant partitions: the strongest lists sit on two ranks and win early
        rng = np.random.default_rng(5)
        anc = rng.integers(0, 4, 12_000)
        anc[9000:9600] = anc[1000:1600]
        anc[6000:6300] = anc[500:800]
        g = np.empty((60, 12_000), np.uint8)
        for i in range(60):
            s = anc.copy(); mut = rng.random(12_000) < 0.02; s[mut] = rng.integers(0, 4, int(mut.sum()))
            g[i] = np.frombuffer(b"ACGT", np.uint8)[s]
        return g, 13, 300, 2, (W, S, w)
    if name == "tiny_k":    # k = 6: nearly every word occurs on every rank; expected to hit the documented limits cleanly
        rng = np.random.default_rng(6)
        anc = rng.integers(0, 4, 3000)
        g = np.empty((30, 3000), np.uint8)
        for i in range(30):
            s = anc.copy(); mut = rng.random(3000) < 0.03; s[mut] = rng.integers(0, 4, int(mut.sum()))
            g[i] = np.frombuffer(b"ACGT", np.uint8)[s]
        return g, 6, 60, 3, (100, 50, 30)
    if name == "cfg2":
        g, k = synth.make_config("cfg2")
        return g, k, 1000, 10, (W, S, w)
    if name == "cfg5shard":
        return synth.synth_genomes(12_500, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01), 13, 1000, 10, (W, S, w)
    if name == "cfg5xN":    # 12,500 x world genomes of 30 kb: at 8 ranks exactly BASELINE configs[4] (100,000 x 30 kb)
        return synth.synth_genomes(12_500 * world, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01), 13, 1000, 10, (W, S, w)
    if name == "cfg3xN":    # bench.py --gpus N: N x 10,000 genomes of the cfg3 shape
        cfgd = dict(synth.CONFIGS["cfg3"]); k = cfgd.pop("k"); cfgd["n"] *= world
        return synth.synth_genomes(**cfgd), k, 1000, 2, (W, S, w)
    raise SystemExit("unknown case " + name)


for name in (sys.argv[1:] or ["small", "repeats", "cfg2"]):
    g, k, it, mms, (W, S, w) = case(name)
    shard = D.column_shard(g, W, S, rank, world)
    eng = m.Engine(k, W, S, w, device=lr)
    eng.load_genomes(shard.reshape(-1), synth.offsets_for(shard))
    eng.build_index()
    eng.build_index()
    build_ms = eng.timing().index_ms + eng.timing().encode_ms
    eng.dist_init(dist, dev)
    res = None
    try:
        for rep in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            a, b = eng.select_both_dist(it, mms)
            dt = time.perf_counter() - t0
            tm = eng.timing()
        res = (a, b, int(tm.select_evals[0] + tm.select_evals[1]))
        msg = "%d genomes x %d columns, shard index build %.2f ms, %d + %d winners, evals %d, %.3f ms (device %.3f)" % (g.shape[0], g.shape[1], build_ms, len(a), len(b), res[2], 1e3 * dt, tm.select_ms[0])
    except m.MsspeError as e:
        msg = "error: %s" % e
    ok = None
    if rank == 0:
        e1 = m.Engine(k, W, S, w, device=lr)
        e1.load_genomes(g.reshape(-1), synth.offsets_for(g))
        e1.build_index()
        e1.build_index()
        build1_ms = e1.timing().index_ms + e1.timing().encode_ms
        a1, b1 = e1.select_both(it, mms, m.SELECT_PARTITIONED)
        t0 = time.perf_counter(); a1, b1 = e1.select_both(it, mms, m.SELECT_PARTITIONED); dt1 = time.perf_counter() - t0
        t1 = e1.timing()
        if res is not None:
            ok = a.tobytes() == a1.tobytes() and b.tobytes() == b1.tobytes() and res[2] == int(t1.select_evals[0] + t1.select_evals[1])
            if not ok:
                for x, y, lab in ((a, a1, "fwd"), (b, b1, "rev")):
                    n = min(len(x), len(y))
                    bad = [i for i in range(n) if x[i].tobytes() != y[i].tobytes()]
                    print("  %s: len %d vs %d, first differences %s" % (lab, len(x), len(y), [(i, x[i], y[i]) for i in bad[:3]]))
                print("  evals %d vs %d" % (res[2], int(t1.select_evals[0] + t1.select_evals[1])))
        print("[%s] world %d: %s | one GPU: index build %.2f ms, %d + %d winners, %.3f ms | identical: %s" % (name, world, msg, build1_ms, len(a1), len(b1), 1e3 * dt1, ok), flush=True)
        e1.close()
    eng.close()
    dist.barrier()
dist.destroy_process_group()
