#!/usr/bin/env python3
"""bench.py -- headline benchmark of the od-msspe hot path on B200 (contract: see the task statement).

A "step" is one pass of the hot path over one synthetic input batch: BASELINE.json configs[1]
(1,000 x 30 kb coronavirus-like pre-aligned genomes, k=13, window 500 / step 250 / search 50,
--check-hairpin --check-self-dimers, Tm sigma filter on, cross-dimers off): load -> K1 encode -> K2 index ->
K3 greedy selection (both directions, up to 1000 iterations) -> K4-K6 primer thermodynamics -> filters.

  value     reference-equivalent k-mer coverage evals / s with the genomes already resident in HBM
            (one eval = one execution of od-msspe/src/main.rs:302-307; counted on the device, equal to
            the oracle's count -- tests/test_gpu_kmer.py)
  e2e       the same metric through the C ABI with HOST buffers (pinned host -> device copy and the
            device -> host result reads inside the timed region)
  roofline  the count kernel (K3 coverage scoring): algorithmic bytes = 4 B per eval
  thal      secondary metric of BASELINE.json: all-ordered-pairs thal dimer pairs / s on a cfg4-style pool

`--impl reference` times the CPU restatement of the reference (oracle/, single-threaded like the Rust binary,
which cannot be built in this image) on a bounded sample of the same workload.
With --gpus N (torchrun) every rank runs an independent genome set of the same shape (weak scaling, no
data-path collective) and the thal pair matrix is row-tiled across the ranks.
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

WORKLOAD = "cfg2: synthetic 1000 x 30 kb pre-aligned genomes, k=13, window 500/step 250/search 50, hairpin+self-dimer checks, Tm stddev filter"
CPU_SAMPLE_COLS = 8000
MAX_ITER = 1000
THAL_POOL = 4096


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
            stop.wait(0.01)
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


def run_reference(args, rank, world):
    """CPU arm: the oracle's C++ port of the reference pipeline, single-threaded, on a bounded sample."""
    if rank != 0:
        return
    from oracle import oracle as O
    from msspe_b200 import synth
    O.build()
    g, k = synth.make_config("cfg2")
    fa = synth.to_fasta(g[:, :CPU_SAMPLE_COLS])
    cfg = O.default_config(check_cross_dimers=0)
    times, evals = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = O.run_pipeline(fa, cfg, 0)
        dt = time.perf_counter() - t0
        ev = sum(r.evals)
        r.close()
        if it >= args.warmup:
            times.append(dt)
            evals += ev
    total = sum(times)
    val = evals / total
    sample = "columns [0,%d) of the cfg2 alignment (1000 genomes x %d partitions), full greedy loop both directions + primer thermo + filters" % (
        CPU_SAMPLE_COLS, (CPU_SAMPLE_COLS - 500) // 250 + 1)
    line = {
        "impl": "reference", "metric": "kmer_coverage_evals_per_s", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "restated reference baseline (not the Rust binary: no cargo/rustc in this image); the reference is single-threaded, host has %d cores" % (os.cpu_count() or 0)},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="recount", choices=["recount", "incremental"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-thal", action="store_true")
    ap.add_argument("--batched", action="store_true", help="one launch per phase instead of the persistent kernel (for per-launch ncu numbers)")
    ap.add_argument("--no-large", action="store_true", help="skip the cfg5/8-shard roofline probe (12,500 x 30 kb, inputs larger than L2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
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
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    mode = m.SELECT_RECOUNT if args.mode == "recount" else m.SELECT_INCREMENTAL
    if args.batched:
        mode |= m.SELECT_BATCHED
    # ---- synthetic input: every rank gets its own genome set of the cfg2 shape (weak scaling) ----
    cfgd = dict(synth.CONFIGS["cfg2"])
    k = cfgd.pop("k")
    cfgd["seed"] = cfgd["seed"] + 1000 * rank
    genomes = synth.synth_genomes(**cfgd)
    n_rec, L = genomes.shape
    offs = synth.offsets_for(genomes)
    host_pinned = torch.from_numpy(genomes.reshape(-1)).pin_memory()
    dev_bases = host_pinned.to(dev, non_blocking=False)
    mms = min(10, max(1, -(-n_rec // 50)))  # main.rs:658-660
    fcfg = m.default_filter_cfg()

    eng = m.Engine(k, 500, 250, 50, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)

    def one_step(device_resident: bool):
        if device_resident:
            eng.load_genomes_device(dev_bases.data_ptr(), offs, keepalive=dev_bases)
        else:
            eng.load_genomes(host_pinned.numpy(), offs)
        eng.build_index()
        fwd, rev = eng.select_both(MAX_ITER, mms, mode)
        # get_kmer_stats + filter_kmers for both directions (main.rs:723-732, 408-516) through the C ABI, one device batch
        kept = [st["code"][st["keep"] != 0] for st in eng.kmer_stats_both(fwd["code"], rev["code"], fcfg)]
        t = eng.timing()
        return int(t.select_evals[0] + t.select_evals[1]), fwd, rev, kept, t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps, device-timed with CUDA events on the launching stream; returns (ms_total, results)."""
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

    for _ in range(args.warmup):
        one_step(True)
    one_step(False)

    samples, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, samples, local_rank), daemon=True)
    th.start()
    eng.reset_timing()
    ms_dev, wall_dev, res_dev = timed(lambda: one_step(True), args.steps)
    launches = eng.timing().kernel_launches
    ms_e2e, wall_e2e, res_e2e = timed(lambda: one_step(False), args.steps)
    stop.set()
    th.join(timeout=2)

    evals_step = res_dev[0][0]
    assert all(r[0] == evals_step for r in res_dev + res_e2e)
    # the device timeline includes host gaps (the engine synchronises between batches); use the larger of
    # event time and wall time so that nothing is hidden
    t_dev = max(ms_dev, wall_dev) / 1e3
    t_e2e = max(ms_e2e, wall_e2e) / 1e3

    # ---- the incremental loop (identical winners, no recount after the first): reported separately ----
    inc = None
    if args.mode == "recount":
        def inc_step():
            eng.load_genomes_device(dev_bases.data_ptr(), offs, keepalive=dev_bases)
            eng.build_index()
            f2, r2 = eng.select_both(MAX_ITER, mms, m.SELECT_INCREMENTAL)
            eng.kmer_stats_both(f2["code"], r2["code"], fcfg)
            return f2, r2
        inc_step()
        ms_i, wall_i, res_i = timed(inc_step, args.steps)
        same = res_i[0][0].tobytes() == res_dev[0][1].tobytes() and res_i[0][1].tobytes() == res_dev[0][2].tobytes()
        t_i = max(ms_i, wall_i) / 1e3
        inc = {"value": evals_step * args.steps / t_i, "unit": "reference-equivalent evals/s", "ms_per_step": 1e3 * t_i / args.steps,
               "identical_winners": bool(same),
               "note": "MSSPE_SELECT_INCREMENTAL: exact counts maintained through the forward index, one persistent kernel; no posting is re-streamed, so this is an algorithmic speed-up and not a bandwidth figure (on this rank)"}

    # ---- roofline of the dominant kernel (count_kernel), from one extra profiled step ----
    L_ = m.load_library()
    L_.msspe_set_profiling(eng.h, 1)
    _, _, _, _, tprof = one_step(True)
    L_.msspe_set_profiling(eng.h, 0)
    ck_ms = float(tprof.count_kernel_ms[0] + tprof.count_kernel_ms[1])
    ck_n = int(tprof.count_kernel_launches[0] + tprof.count_kernel_launches[1])
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg_bytes_per_launch = 4.0 * evals_step / max(1, ck_n)
    phys_bytes_per_launch = 4.0 * float(tprof.select_postings_read[0] + tprof.select_postings_read[1]) / max(1, ck_n)
    avg_launch_s = (ck_ms / 1e3) / max(1, ck_n)
    achieved = alg_bytes_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "count_kernel (K3 coverage scoring)", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs if peak_gbs else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "physical_bytes_per_launch": phys_bytes_per_launch,
                "physical_gbs": phys_bytes_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0,
                "avg_launch_us": 1e6 * avg_launch_s, "launches": ck_n,
                "timing": "in-kernel %globaltimer of block 0 around every coverage-scoring phase of the persistent kernel (one phase = one reference recount, both directions; includes the grid barrier that ends it); with --batched: CUDA events around each count_kernel launch",
                "note": "cfg2 postings (2 x 18 MB) are L2-resident, the loop is latency-bound at this size; roofline_large_shard shows the same kernel where it is HBM-bound; see DESIGN.md"}

    # ---- the same kernel where it is HBM-bound: one GPU's shard of BASELINE configs[4] (100,000 x 30 kb over 8 GPUs) ----
    roofline_large = None
    if not args.no_large and rank == 0:
        big = synth.synth_genomes(12_500, 30_000, 5, clades=256, p_clade=0.10, p_leaf=0.01)
        e2 = m.Engine(13, 500, 250, 50, device=local_rank)
        e2.load_genomes(big.reshape(-1), synth.offsets_for(big))
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
        roofline_large = {"workload": "one GPU's shard of cfg5: 12,500 x 30 kb genomes (%d segments), first %d greedy iterations per direction" % (G2, iters),
                          "bound": "hbm", "kernel": "coverage-scoring phase of greedy_persistent_kernel", "achieved": 4.0 * ev2 / ck2 / 1e6,
                          "physical_gbs": 4.0 * pr2 / ck2 / 1e6, "peak": peak_gbs, "unit": "GB/s", "frac": 4.0 * ev2 / ck2 / 1e6 / peak_gbs,
                          "algorithmic_bytes_per_launch": 4.0 * ev2 / max(1, n2), "avg_launch_us": 1e3 * ck2 / max(1, n2), "launches": n2,
                          "postings_bytes_per_direction": int(4 * pr2 / max(1, n2)), "l2": "226 MB per direction per iteration: larger than the 126 MB L2",
                          "encode_ms": float(tb.encode_ms), "index_ms": float(tb.index_ms),
                          "timing": "in-kernel %globaltimer of block 0 around the phase (includes the grid barrier that ends it)"}
        e2.close()
        del big

    # ---- secondary metric: thal dimer pairs / s, pair matrix row-tiled across ranks ----
    thal = None
    if not args.no_thal:
        from msspe_b200 import distributed as D
        pool = synth.random_primers(THAL_POOL, 13, 4)
        cond = m.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
        rb, re_ = D.row_block(THAL_POOL, rank, world)
        eng.cross_dimer(pool, cond, -9000.0 + 1.0, rb, min(re_, rb + 64), edge_capacity=1 << 20, nostruct_capacity=1 << 16)  # warm-up
        kernel_ms = []

        def compute_rows(b0, b1):
            r = eng.cross_dimer(pool, cond, -9000.0 + 1.0, b0, b1, edge_capacity=1 << 22, nostruct_capacity=1 << 20)
            kernel_ms.append(float(eng.timing().dimer_ms))
            return r

        barrier()
        t0 = time.perf_counter()
        # rows tiled across ranks, compacted edge lists merged with one NCCL all_gather (msspe_b200/distributed.py)
        edges, nos = D.cross_dimer_sharded(compute_rows, THAL_POOL, m.EDGE_DTYPE, dist if world > 1 else None, dev)
        torch.cuda.synchronize(dev)
        t_thal = time.perf_counter() - t0
        tt = torch.tensor([t_thal, sum(kernel_ms) / 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_thal, k_s, n_edges = float(tt[0]), float(tt[1]), int(len(edges))
        pairs = THAL_POOL * THAL_POOL
        thal = {"metric": "thal_dimer_pairs_per_s", "value": pairs / t_thal, "unit": "pairs/s", "pairs": pairs,
                "pool": "%d uniform-random 13-mers (cfg4 shape, seed 4), mv 50 dv 3 dNTP 0 DNA 250 nM 25 C" % THAL_POOL,
                "kernel_only_pairs_per_s": pairs / k_s if k_s > 0 else None, "conflict_edges_below_-9000": n_edges,
                "scaling": "strong (rows tiled across ranks)", "bound": "SM issue (FP64/ALU), not HBM"}

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
            from oracle import oracle as O
            O.build()
            fa = synth.to_fasta(genomes[:, :CPU_SAMPLE_COLS])
            t0 = time.perf_counter()
            r = O.run_pipeline(fa, O.default_config(check_cross_dimers=0), 0)
            dt = time.perf_counter() - t0
            cpu = {"value": sum(r.evals) / dt, "unit": "evals/s", "cores": 1, "kind": "port",
                   "sample": "columns [0,%d) of the cfg2 alignment (1000 genomes x %d partitions), full pipeline, %.1f s" % (
                       CPU_SAMPLE_COLS, (CPU_SAMPLE_COLS - 500) // 250 + 1, dt),
                   "host_cores": os.cpu_count(),
                   "note": "restated reference baseline (oracle port, not the Rust binary); the reference is single-threaded"}
            r.close()
        fwd, rev, kept = res_dev[0][1], res_dev[0][2], res_dev[0][3]
        h2d = int(genomes.size) + 8 * (n_rec + 1)
        d2h = int((len(fwd) + len(rev)) * (24 + 5 * 8 + 3 * 40)) + 2 * 128
        line = {
            "metric": "kmer_coverage_evals_per_s", "value": total_evals / t_dev, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "select_mode": args.mode + ("+batched-launches" if args.batched else "+persistent-kernel" if args.mode == "recount" else ""), "genomes_per_gpu": n_rec, "genome_length": L,
                       "max_iterations": MAX_ITER, "max_mismatch_segments": mms,
                       "l2": "inputs per step (30 MB genomes + 2 x 18 MB postings) are smaller than L2; each step rebuilds the index from the genome bytes, nothing is cached across steps",
                       "iterations": [int(res_dev[0][4].select_iterations[0]), int(res_dev[0][4].select_iterations[1])],
                       "candidates": [int(len(fwd)), int(len(rev))], "kept_after_filters": [int(len(kept[0])), int(len(kept[1]))],
                       "evals_per_step": evals_step},
            "e2e": {"value": total_evals / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": summarize_clocks(samples),
            "roofline": roofline,
            "roofline_large_shard": roofline_large,
            "incremental": inc,
            "cpu_baseline": cpu,
            "thal": thal,
            "stage_ms": {"encode": float(res_dev[0][4].encode_ms), "index": float(res_dev[0][4].index_ms),
                         "select_fwd": float(res_dev[0][4].select_ms[0]), "select_rev": float(res_dev[0][4].select_ms[1]),
                         "thermo_last_dir": float(res_dev[0][4].thermo_ms)},
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
