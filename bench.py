#!/usr/bin/env python3
"""bench.py -- headline benchmark of the od-msspe hot path on B200 (contract: see the task statement).

A "step" is one pass of the hot path over one synthetic input batch.  Workload at --gpus 1: BASELINE.json configs[2],
the largest named single-GPU configuration (10,000 x 11 kb dengue-like pre-aligned genomes, k=15, window 500 / step 250 /
search 50, --max-mismatch-segments=2, --max-iterations 1000, hairpin + self-dimer checks, Tm filters):
load -> K1 encode -> K2 index -> K3 greedy selection (both directions) -> K4-K6 primer thermodynamics -> filters.
The winners of every run are compared with tests/golden/cfg3_candidates.json (CPU oracle, offline).

  value         reference-equivalent k-mer coverage evals / s with the genomes already resident in HBM (one eval = one
                execution of od-msspe/src/main.rs:302-307; the count equals the oracle's, tests/test_gpu_kmer.py)
  e2e           the same metric through the C ABI with HOST buffers (pinned host -> device copy of the genomes and the
                device -> host reads of the results inside the timed region)
  roofline      the kernel class with the largest share of device time in the timed steps (CUDA events around every launch,
                msspe_get_kernel_profile): algorithmic bytes per launch / average launch time / measured HBM peak
  thal          secondary metric of BASELINE.json: all-ordered-pairs thal dimer pairs / s on the 20,000-primer pool of
                configs[3] (4.0e8 pairs), row-tiled across ranks, with its own (SM-issue) roofline entry
  cpu_baseline  the CPU restatement of the reference (oracle/, single-threaded like the Rust binary, which cannot be
                built in this image) on a bounded sample of the same workload

`--impl reference` times that CPU restatement alone: the complete cfg3 alignment (every genome, every column), each step
= the first REF_ITERS greedy iterations of both directions (the full 2 x 1000 iterations take the port ~50 minutes).
With --gpus N > 1 (torchrun, one process per GPU, NCCL): ONE design job of N x 10,000 genomes of the same shape (weak
scaling: the work per GPU stays that of cfg3), partitions (alignment columns) sharded over the ranks, every rank building
the index of its own columns, the greedy loop of the whole job through msspe_select_both_dist (an all-gather and an
all-reduce per ROUND of the per-partition loop, nothing per iteration); rank 0 re-runs the whole job on its own GPU after
the timed region and refuses to report unless the winners are bit-identical.  See run_multi().
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))

import numpy as np  # noqa: E402

WORKLOAD = ("cfg3: synthetic 10,000 x 11 kb dengue-like pre-aligned genomes, k=15, window 500/step 250/search 50, "
            "--max-mismatch-segments=2, --max-iterations 1000, hairpin+self-dimer checks, Tm min/max + stddev filters")
CFG, K_WINDOW = "cfg3", (500, 250, 50)
MAX_ITER, MMS = 1000, 2
REF_ITERS = 1            # greedy iterations per direction and step of the CPU arm (complete alignment)
THAL_POOL = 20_000       # BASELINE configs[3]
THAL_COND = (50, 3, 0, 250, 25.0, 30, 0)
THAL_LIMIT = -9000.0 + 1.0
# algorithmic bytes of one launch of each kernel class are counted by the library (DESIGN.md section 4 states the
# per-unit figures); flop of one 13-mer ordered pair through thal ANY: SURVEY.md 8(d), candidates x 25 on a uniform pool
THAL_FLOP_PER_PAIR = 1.8e4


def clock_sampler(stop, out, gpu_index):
    """SM clock and throttle reasons DURING the timed region.  In-process NVML (a query costs microseconds); spawning
    nvidia-smi every few milliseconds takes a driver lock often enough to perturb the host-side calls being timed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not stop.is_set():
            r = get_reasons(h)
            flag = lambda bit: "Active" if r & bit else "Not Active"  # noqa: E731
            out.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx), flag(0x8), flag(0x40), flag(0x20), flag(0x4)])
            stop.wait(0.004)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.05)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max(int(s[1]) for s in samples if s[1].isdigit()) if any(s[1].isdigit() for s in samples) else None
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(samples)}


def load_golden():
    with open(os.path.join(ROOT, "tests", "golden", "%s_candidates.json" % CFG)) as f:
        return json.load(f)


def cpu_arm(genomes, k, steps, warmup):
    """The oracle's C++ port of find_candidates_kmers (main.rs:331-406: string k-mers, hash maps, a full recount per
    iteration), single-threaded like the reference, on the COMPLETE alignment; the segment manager is built once
    (untimed, as the GPU arm's genomes are resident); one step = REF_ITERS iterations of both directions."""
    from oracle import oracle as O
    from msspe_b200 import synth
    O.build()
    W, S, w = K_WINDOW
    t0 = time.perf_counter()
    mgr = O.Manager(synth.to_fasta(genomes), W, S, w, k)
    t_build = time.perf_counter() - t0
    gold = load_golden()
    times, evals = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ev = 0
        for d in (0, 1):
            r = mgr.select(d, REF_ITERS, MMS)
            ev += r["evals"]
            assert r["codes"].tolist() == gold["dirs"][d]["codes"][:REF_ITERS]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            evals += ev
    mgr.close()
    total = sum(times)
    sample = ("the complete cfg3 alignment (10,000 genomes x 11,000 columns, %d segments); one step = the first %d greedy iterations "
              "of both directions (2 x 1000 iterations take this port ~50 min); segment manager built once, untimed (%.0f s)" % (
                  mgr.n_segments, REF_ITERS, t_build))
    return {"value": evals / total, "unit": "evals/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": os.cpu_count(),
            "note": "restated reference baseline (oracle port, not the Rust binary: no cargo/rustc in this image); the reference is single-threaded",
            "ms_per_step": 1e3 * total / max(1, len(times)), "steps": len(times)}


def run_reference(args, rank):
    if rank != 0:
        return
    from msspe_b200 import synth
    g, k = synth.make_config(CFG)
    cpu = cpu_arm(g, k, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "kmer_coverage_evals_per_s", "value": cpu["value"], "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_multi(args, rank, world, local_rank, real_stdout):
    """One job over `world` GPUs (see the module docstring).  value = reference-equivalent evals of the WHOLE job / time."""
    import torch
    import torch.distributed as dist
    import msspe_b200 as m
    from msspe_b200 import synth, distributed as D
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    cfgd = dict(synth.CONFIGS[CFG])
    k = cfgd.pop("k")
    cfgd["n"] = cfgd["n"] * world                       # the same alignment on every rank (same seed); each keeps its columns
    genomes = synth.synth_genomes(**cfgd)
    W, S, w = K_WINDOW
    n_rec, L = genomes.shape
    n_part = (L - W) // S + 1
    p0, p1 = D.partition_range(n_part, rank, world)
    shard = D.column_shard(genomes, W, S, rank, world)
    offs = synth.offsets_for(shard)
    host_pinned = torch.from_numpy(shard.reshape(-1)).pin_memory()
    dev_bases = host_pinned.to(dev)
    fcfg = m.default_filter_cfg()
    eng = m.Engine(k, W, S, w, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    eng.dist_init(dist, dev)
    d2h_bytes = [0]

    def one_step(device_resident):
        if device_resident:
            eng.load_genomes_device(dev_bases.data_ptr(), offs, keepalive=dev_bases)
        else:
            eng.load_genomes(host_pinned.numpy(), offs)
        eng.build_index()
        fwd, rev = eng.select_both_dist(MAX_ITER, MMS)
        st = eng.kmer_stats_both(fwd["code"], rev["code"], fcfg)     # 2 x <= 1000 primers: every rank, no exchange
        d2h_bytes[0] = fwd.nbytes + rev.nbytes + st[0].nbytes + st[1].nbytes
        t = eng.timing()
        return int(t.select_evals[0] + t.select_evals[1]), fwd, rev, t

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        res = [fn() for _ in range(steps)]
        e1.record(stream)
        torch.cuda.synchronize(dev)
        wall = 1e3 * (time.perf_counter() - t0)
        barrier()
        return max(e0.elapsed_time(e1), 0.0), wall, res

    # the sampler starts before the warm-up (NVML initialisation takes longer than the timed region on some boxes); what
    # it saw before the timed region is dropped
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, samples, local_rank), daemon=True)
    th.start()
    for _ in range(max(1, args.warmup - 1)):
        one_step(True)
    one_step(False)
    del samples[:]
    eng.reset_timing()
    ms_dev, wall_dev, res_dev = timed(lambda: one_step(True), args.steps)
    launches = eng.timing().kernel_launches
    eng.set_profiling(True)
    eng.reset_timing()
    ms_prof, wall_prof, _ = timed(lambda: one_step(True), args.steps)
    kprof = eng.kernel_profile()
    eng.set_profiling(False)
    ms_e2e, wall_e2e, res_e2e = timed(lambda: one_step(False), args.steps)
    stop.set()
    th.join(timeout=2)
    evals_step = res_dev[0][0]
    assert all(r[0] == evals_step for r in res_dev + res_e2e)
    tt = torch.tensor([max(ms_dev, wall_dev) / 1e3, max(ms_e2e, wall_e2e) / 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)                     # max over ranks
    t_dev, t_e2e = float(tt[0]), float(tt[1])
    ln = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    thal = None
    if not args.no_thal:
        thal = thal_section(eng, m, synth, dist, world, rank, dev, barrier)
    # parity: the whole job once more on ONE GPU (rank 0), winners and evals must be identical
    verified = None
    if rank == 0:
        e1 = m.Engine(k, W, S, w, device=local_rank)
        e1.load_genomes(genomes.reshape(-1), synth.offsets_for(genomes))
        e1.build_index()
        t0 = time.perf_counter()
        a1, b1 = e1.select_both(MAX_ITER, MMS, m.SELECT_AUTO)
        one_gpu_loop_ms = 1e3 * (time.perf_counter() - t0)
        t1 = e1.timing()
        ok = (a1.tobytes() == res_dev[0][1].tobytes() and b1.tobytes() == res_dev[0][2].tobytes()
              and int(t1.select_evals[0] + t1.select_evals[1]) == evals_step)
        if world == 1 or n_rec == 10_000:
            gold = load_golden()
            ok = ok and all(x["code"].tolist() == gold["dirs"][d]["codes"] for d, x in enumerate((a1, b1)))
        e1.close()
        verified = {"identical_to_one_gpu_run_of_the_whole_job": bool(ok), "one_gpu_greedy_loop_ms_first_call": one_gpu_loop_ms,
                    "what": "winners, frequencies, tie counts, f32 tie scores (bytes of msspe_candidate) and reference-equivalent evals of both directions"}
        if not ok:
            raise SystemExit("bench.py: the %d-GPU job differs from the one-GPU run of the same input -- refusing to report a number" % world)
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        classes = [{"kernel": r["name"].decode(), "ms_per_step": float(r["ms"]) / args.steps, "launches_per_step": int(r["launches"]) // args.steps,
                    "alg_bytes_per_step": int(r["alg_bytes"]) // args.steps} for r in kprof]
        classes.sort(key=lambda c: -c["ms_per_step"])
        dom = classes[0] if classes else None
        roofline = None
        if dom:
            lp = max(1, dom["launches_per_step"])
            alg = dom["alg_bytes_per_step"] / lp
            avg_s = dom["ms_per_step"] / 1e3 / lp
            ach = alg / avg_s / 1e9 if avg_s > 0 else 0.0
            roofline = {"bound": "hbm" if alg > 0 else "latency/L2", "kernel": dom["kernel"] + " (rank 0)", "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                        "frac": ach / peak_gbs, "traffic": None, "algorithmic_bytes_per_launch": alg, "avg_launch_us": 1e6 * avg_s, "launches_per_step": lp,
                        "timing": "CUDA events around every launch of this class on rank 0's stream, a second pass of the same K steps"}
        tm = res_dev[0][3]
        total_evals = float(evals_step) * args.steps
        line = {
            "metric": "kmer_coverage_evals_per_s", "value": total_evals / t_dev, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "ONE job: %d x 11 kb genomes of the cfg3 shape (%d x cfg3's 10,000), k=15, window 500/step 250/search 50, --max-mismatch-segments=2, "
                                   "--max-iterations 1000; the %d partitions (alignment columns) sharded over %d GPUs, greedy loop of the whole job by msspe_select_both_dist" % (
                                       n_rec, world, n_part, world),
                       "genomes": n_rec, "genome_length": L, "partitions_of_rank0": [p0, p1], "max_iterations": MAX_ITER, "max_mismatch_segments": MMS,
                       "collectives": "per ROUND of the per-partition loop: one all_gather of the ranks' not-yet-final entries (%d B per rank and direction) and one all_reduce of the "
                                      "cross-rank lists' cover-time histograms (fixed buffer); set-up per call: all_gather of the ranks' distinct words. Nothing per greedy iteration." % 0,
                       "parity": verified, "evals_per_step": evals_step, "candidates": [int(len(res_dev[0][1])), int(len(res_dev[0][2]))],
                       "l2": "per-rank inputs (genome columns + 2 x 62 MB postings + forward index) exceed the 126 MB L2; every step rebuilds the index"},
            "e2e": {"value": total_evals / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": int(shard.size) * world + 8 * (n_rec + 1) * world,
                    "d2h_bytes_per_step": int(d2h_bytes[0]) * world, "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(ln[0]),
            "clocks": summarize_clocks(samples),
            "roofline": roofline,
            "kernel_classes": classes,
            "stage_ms_rank0": {"build": float(tm.encode_ms + tm.index_ms), "select": float(tm.select_ms[0]), "thermo": float(tm.thermo_ms)},
            "cpu_baseline": None,
            "thal": thal,
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def thal_section(eng, m, synth, dist, world, rank, dev, barrier):
    """All ordered pairs of the cfg4 pool through thal ANY (delta_g.rs:61-153), rows tiled across the ranks; the compacted
    lists stay on the device, one all_gather of the counts and one of the lists (msspe_b200/distributed.py)."""
    import torch
    from msspe_b200 import distributed as D
    pool = synth.random_primers(THAL_POOL, 13, 4)
    cond = m.ThalCond(*THAL_COND)
    rb, re_ = D.row_block(THAL_POOL, rank, world)
    ecap, ncap = max(1 << 16, (re_ - rb) * THAL_POOL // 50), max(1 << 16, (re_ - rb) * THAL_POOL // 2000)
    eng.cross_dimer_device(pool, cond, THAL_LIMIT, rb, min(re_, rb + 64), edge_capacity=ecap, nostruct_capacity=ncap)  # warm-up
    kernel_ms = []

    def compute_rows(b0, b1):
        r = eng.cross_dimer_device(pool, cond, THAL_LIMIT, b0, b1, edge_capacity=ecap, nostruct_capacity=ncap)
        kernel_ms.append(float(eng.timing().dimer_ms))
        return r

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    edges, nos = D.cross_dimer_sharded_tensors(compute_rows, THAL_POOL, m.EDGE_DTYPE, dist if world > 1 else None, dev)
    e1.record()
    torch.cuda.synchronize(dev)
    t_thal = time.perf_counter() - t0
    tt = torch.tensor([t_thal, sum(kernel_ms) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_thal, k_s = float(tt[0]), float(tt[1])
    pairs = THAL_POOL * THAL_POOL
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2s6_thal_thread_metrics.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass
    return {"metric": "thal_dimer_pairs_per_s", "value": pairs / t_thal, "unit": "pairs/s", "pairs": pairs,
            "pool": "%d uniform-random 13-mers (BASELINE configs[3], seed 4), mv 50 dv 3 dNTP 0 DNA 250 nM 25 C" % THAL_POOL,
            "kernel_only_pairs_per_s": pairs / k_s if k_s > 0 else None, "conflict_edges_below_-9000": int(len(edges)),
            "structureless_pairs": int(len(nos)), "seconds": t_thal,
            "scaling": "strong (rows tiled across ranks; one all_gather of counts + one of the device-resident lists)",
            "roofline": {"bound": "sm_issue", "kernel": "thal_dimer_thread_kernel (one thread per ordered pair)", "achieved": pairs / k_s if k_s > 0 else None, "unit": "pairs/s",
                         "flop_per_pair": THAL_FLOP_PER_PAIR, "achieved_fp64_gflops": pairs / k_s * THAL_FLOP_PER_PAIR / 1e9 if k_s > 0 else None,
                         "issue_active_pct": ncu.get("issue_active_pct"), "fp64_pipe_pct": ncu.get("fp64_pipe_pct"),
                         "lanes_active_per_inst": ncu.get("lanes_active_per_inst"), "ncu_source": ncu.get("source"),
                         "note": "16 B in / <= 48 B out per pair against ~1e4 instructions: HBM is irrelevant (ncu: < 0.01 % DRAM); the bound is SM issue slots"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "partitioned", "recount", "incremental"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-thal", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the cfg5/8-shard probe of the recount kernel (12,500 x 30 kb, inputs larger than L2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # libraries (NCCL) print banners on fd 1: keep fd 1 clean for the single JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import msspe_b200 as m
    from msspe_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    if world > 1:
        run_multi(args, rank, world, local_rank, real_stdout)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    mode = {"auto": m.SELECT_AUTO, "partitioned": m.SELECT_PARTITIONED, "recount": m.SELECT_RECOUNT, "incremental": m.SELECT_INCREMENTAL}[args.mode]
    # ---- synthetic input (N > 1: every rank its own genome set of the same shape, no data-path collective) ----
    cfgd = dict(synth.CONFIGS[CFG])
    k = cfgd.pop("k")
    cfgd["seed"] = cfgd["seed"] + 1000 * rank
    genomes = synth.synth_genomes(**cfgd)
    n_rec, L = genomes.shape
    offs = synth.offsets_for(genomes)
    host_pinned = torch.from_numpy(genomes.reshape(-1)).pin_memory()
    dev_bases = host_pinned.to(dev, non_blocking=False)
    fcfg = m.default_filter_cfg()
    gold = load_golden() if rank == 0 else None

    eng = m.Engine(k, *K_WINDOW, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    d2h_bytes = [0]

    def one_step(device_resident: bool):
        if device_resident:
            eng.load_genomes_device(dev_bases.data_ptr(), offs, keepalive=dev_bases)
        else:
            eng.load_genomes(host_pinned.numpy(), offs)
        eng.build_index()
        fwd, rev = eng.select_both(MAX_ITER, MMS, mode)
        # get_kmer_stats + filter_kmers for both directions (main.rs:723-732, 408-516) through the C ABI, one device batch
        st = eng.kmer_stats_both(fwd["code"], rev["code"], fcfg)
        kept = [s["code"][s["keep"] != 0] for s in st]
        d2h_bytes[0] = fwd.nbytes + rev.nbytes + st[0].nbytes + st[1].nbytes   # what came back to the host this step
        t = eng.timing()
        return int(t.select_evals[0] + t.select_evals[1]), fwd, rev, kept, t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps, device-timed with CUDA events on the launching stream; returns (ms_total, wall ms, results)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        res = [fn() for _ in range(steps)]
        e1.record(stream)
        torch.cuda.synchronize(dev)
        wall = 1e3 * (time.perf_counter() - t0)
        barrier()
        return max(e0.elapsed_time(e1), 0.0), wall, res

    # cold start: the very first step of the process (stream-ordered pool maps fresh memory, partition view built)
    t0 = time.perf_counter()
    _, _, _, _, tcold = one_step(True)
    cold_ms = 1e3 * (time.perf_counter() - t0)
    # the sampler starts before the warm-up (NVML initialisation takes longer than the timed region on some boxes); what
    # it saw before the timed region is dropped
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, samples, local_rank), daemon=True)
    th.start()
    for _ in range(max(0, args.warmup - 1)):
        one_step(True)
    one_step(False)
    del samples[:]
    eng.reset_timing()
    ms_dev, wall_dev, res_dev = timed(lambda: one_step(True), args.steps)
    launches = eng.timing().kernel_launches
    # the same K steps once more with CUDA events around every kernel launch of the library (two event records per launch
    # cost ~0.5 ms per step, so the headline value is the unprofiled pass; both are inside bench.py's timed regions)
    eng.set_profiling(True)
    eng.reset_timing()
    ms_prof, wall_prof, _ = timed(lambda: one_step(True), args.steps)
    kprof = eng.kernel_profile()
    eng.set_profiling(False)
    ms_e2e, wall_e2e, res_e2e = timed(lambda: one_step(False), args.steps)
    stop.set()
    th.join(timeout=2)

    evals_step = res_dev[0][0]
    assert all(r[0] == evals_step for r in res_dev + res_e2e)
    # the device timeline includes host gaps (the engine synchronises between stages); use the larger of event time and
    # wall time so that nothing is hidden
    t_dev = max(ms_dev, wall_dev) / 1e3
    t_e2e = max(ms_e2e, wall_e2e) / 1e3

    golden = None
    if rank == 0:
        fwd, rev = res_dev[0][1], res_dev[0][2]
        ok = all(x["code"].tolist() == gold["dirs"][d]["codes"] and x["freq"].tolist() == gold["dirs"][d]["freqs"] for d, x in enumerate((fwd, rev)))
        ok = ok and evals_step == gold["dirs"][0]["evals"] + gold["dirs"][1]["evals"]
        golden = {"file": "tests/golden/%s_candidates.json" % CFG, "matched": bool(ok),
                  "what": "winners, frequencies and reference-equivalent evals of both directions vs the CPU oracle (tools/gen_size_goldens.py)"}
        if not ok:
            raise SystemExit("bench.py: winners differ from the oracle golden -- refusing to report a number")

    # ---- the other exact loop implementations on the same index (reported separately) ----
    variants = {}
    if rank == 0:
        for label, md in (("partitioned", m.SELECT_PARTITIONED), ("incremental", m.SELECT_INCREMENTAL), ("recount", m.SELECT_RECOUNT)):
            eng.select_both(MAX_ITER, MMS, md)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            f2, r2 = eng.select_both(MAX_ITER, MMS, md)
            dt = time.perf_counter() - t0
            same = f2.tobytes() == res_dev[0][1].tobytes() and r2.tobytes() == res_dev[0][2].tobytes()
            variants[label] = {"greedy_loop_ms": 1e3 * dt, "identical_winners": bool(same),
                               "loop_evals_per_s": evals_step / dt}

    # ---- roofline of the dominant kernel class of the timed steps ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    classes = [{"kernel": r["name"].decode(), "ms_per_step": float(r["ms"]) / args.steps, "launches_per_step": int(r["launches"]) // args.steps,
                "alg_bytes_per_step": int(r["alg_bytes"]) // args.steps} for r in kprof]
    classes.sort(key=lambda c: -c["ms_per_step"])
    dom = classes[0] if classes else None
    roofline = None
    if dom:
        traffic = None
        bound_note = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2s5_dominant_kernel_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get("kernels", {}).get(dom["kernel"], {}).get("dram_bytes_per_launch")
            bound_note = tj.get("kernels", {}).get(dom["kernel"], {}).get("bound")
        except Exception:
            pass
        lp = max(1, dom["launches_per_step"])
        alg_per_launch = dom["alg_bytes_per_step"] / lp
        avg_launch_s = dom["ms_per_step"] / 1e3 / lp
        achieved = alg_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
        bound = "hbm"
        if traffic is not None and alg_per_launch > 0 and traffic < 0.25 * alg_per_launch:
            bound = "latency/L2"
        if bound_note:
            bound = bound_note          # what the ncu capture of this kernel shows (profiles/r2s5_dominant_kernel_traffic.json)
        roofline = {"bound": bound, "kernel": dom["kernel"], "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": achieved / peak_gbs if peak_gbs else None, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_per_launch, "avg_launch_us": 1e6 * avg_launch_s, "launches_per_step": lp,
                    "share_of_kernel_time": dom["ms_per_step"] / max(1e-9, sum(c["ms_per_step"] for c in classes)),
                    "timing": "CUDA events around every launch of this class on the launching stream, inside the timed steps (msspe_get_kernel_profile)",
                    "per_unit": "algorithmic bytes per launch as the library counts them; DESIGN.md section 4 states the per-unit figure of every class"}
    # the greedy loop as a whole: reference-equivalent evals x 4 B against HBM (the partitioned loop skips the recount)
    sel_ms = float(res_dev[0][4].select_ms[0])
    greedy = {"loop_ms": sel_ms, "reference_equivalent_gbs": 4.0 * evals_step / (sel_ms / 1e3) / 1e9 if sel_ms > 0 else None, "peak": peak_gbs,
              "note": "4 B x reference-equivalent evals / loop time: above the HBM peak because the per-partition loop examines each forward-index "
                      "record about twice in the whole run instead of once per iteration; an algorithmic speed-up, not a bandwidth figure"}

    # ---- the recount kernel where it is HBM-bound: one GPU's shard of BASELINE configs[4] (first 50 iterations) ----
    roofline_large = None
    if not args.no_large and rank == 0:
        big = synth.synth_genomes(12_500, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
        e2 = m.Engine(13, 500, 250, 50, device=local_rank)
        e2.load_genomes(big.reshape(-1), synth.offsets_for(big))
        e2.build_index()
        tb0 = e2.timing()
        e2.build_index()
        tb = e2.timing()
        iters = 50
        e2.select_both(iters, 10, m.SELECT_RECOUNT)      # warm-up
        e2.select_both(iters, 10, m.SELECT_RECOUNT)
        t2 = e2.timing()
        ev2 = int(t2.select_evals[0] + t2.select_evals[1])
        pr2 = int(t2.select_postings_read[0] + t2.select_postings_read[1])
        ck2 = float(t2.count_kernel_ms[0] + t2.count_kernel_ms[1])
        n2 = int(t2.count_kernel_launches[0] + t2.count_kernel_launches[1])
        G2 = e2.segment_info()[0]
        t0 = time.perf_counter()
        e2.select_both(1000, 10, m.SELECT_AUTO)
        auto_ms = 1e3 * (time.perf_counter() - t0)
        roofline_large = {"workload": "one GPU's shard of cfg5: 12,500 x 30 kb genomes (%d segments), first %d greedy iterations per direction, MSSPE_SELECT_RECOUNT" % (G2, iters),
                          "bound": "hbm", "kernel": "coverage-scoring phase of greedy_persistent_kernel", "achieved": 4.0 * ev2 / ck2 / 1e6 if ck2 > 0 else None,
                          "physical_gbs": 4.0 * pr2 / ck2 / 1e6 if ck2 > 0 else None, "peak": peak_gbs, "unit": "GB/s",
                          "frac": 4.0 * ev2 / ck2 / 1e6 / peak_gbs if ck2 > 0 else None,
                          "algorithmic_bytes_per_launch": 4.0 * ev2 / max(1, n2), "avg_launch_us": 1e3 * ck2 / max(1, n2), "launches": n2,
                          "l2": "226 MB of postings per direction per iteration: larger than the 126 MB L2",
                          "encode_ms_first_build": float(tb0.encode_ms), "index_ms_first_build": float(tb0.index_ms),
                          "encode_ms": float(tb.encode_ms), "index_ms": float(tb.index_ms),
                          "auto_loop_2x1000_iterations_ms": auto_ms,
                          "timing": "in-kernel %globaltimer of block 0 around the phase (includes the grid barrier that ends it)"}
        e2.close()
        del big

    thal = None
    if not args.no_thal:
        thal = thal_section(eng, m, synth, dist, world, rank, dev, barrier)

    # ---- aggregate over ranks ----
    agg = torch.tensor([float(evals_step * args.steps), t_dev, t_e2e, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        s = agg.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        total_evals, t_dev, t_e2e, launches = float(s[0]), float(mx[1]), float(mx[2]), int(s[3])
    else:
        total_evals = float(agg[0])

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_arm(genomes, k, 1, 0)
        fwd, rev, kept, tm = res_dev[0][1], res_dev[0][2], res_dev[0][3], res_dev[0][4]
        h2d = int(genomes.size) + 8 * (n_rec + 1)
        line = {
            "metric": "kmer_coverage_evals_per_s", "value": total_evals / t_dev, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "select_mode": args.mode, "genomes_per_gpu": n_rec, "genome_length": L,
                       "max_iterations": MAX_ITER, "max_mismatch_segments": MMS, "golden": golden,
                       "multi_gpu": None if world == 1 else "every rank runs its own cfg3-shaped job (weak scaling, no data-path collective)",
                       "l2": "inputs per step (110 MB of genomes, 2 x 62 MB of postings, 2 x 62 MB forward index) exceed the 126 MB L2; each step rebuilds the index from the genome bytes, nothing is cached across steps",
                       "iterations": [int(tm.select_iterations[0]), int(tm.select_iterations[1])],
                       "candidates": [int(len(fwd)), int(len(rev))], "kept_after_filters": [int(len(kept[0])), int(len(kept[1]))],
                       "evals_per_step": evals_step},
            "e2e": {"value": total_evals / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h_bytes[0]),
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": summarize_clocks(samples),
            "roofline": roofline,
            "kernel_classes": classes,
            "kernel_classes_ms_per_step_profiled_pass": 1e3 * max(ms_prof, wall_prof) / 1e3 / args.steps,
            "greedy_loop": greedy,
            "loop_variants": variants,
            "roofline_large_shard": roofline_large,
            "cpu_baseline": cpu,
            "thal": thal,
            "stage_ms": {"encode": float(tm.encode_ms), "index": float(tm.index_ms), "select": float(tm.select_ms[0]), "thermo": float(tm.thermo_ms)},
            "cold_start": {"first_step_ms": cold_ms, "encode_ms": float(tcold.encode_ms), "index_ms": float(tcold.index_ms),
                           "note": "first step of the process: the stream-ordered pool maps fresh device memory, the partition view is built"},
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
