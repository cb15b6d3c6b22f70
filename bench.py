#!/usr/bin/env python3
"""Benchmark of the esloco hot path (read placement + target tally) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]

Metric (BASELINE.json): simulated reads placed per second (and iterations per second), hg38-size genome, 30x.
Workload at N=1: BASELINE.json configs[1] -- I mode, synthetic hg38-size reference (3.1 Gb, 24 contigs),
1 barcode, 30x, 100 random insertions of 5 kb, 100 iterations.  One STEP = one pass of the hot path over one
batch = the config's 100 iterations in one esl_run_batch call (every rank runs its own 100 iterations: weak
scaling over independent iterations; for N > 1 the per-target moment accumulators are combined by one NCCL
reduce inside the step).

Keys of the JSON line are described in DESIGN.md ("Measurement").  `--impl reference` times the CPU port of the
reference's algorithm (oracle/, plain C, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_READ = 96     # staged pipeline, u32 coordinates
ALG_BYTES_PER_TARGET = 32   # per (target x iteration)


def make_workload(name):
    from esloco_b200 import synthetic as S
    G, _ = S.chromosome_dir(S.HG38)
    w = dict(name=name, genome_size=G, coverage=30, mean_read_length=15000, sigma=1, min_read_length=1,
             consuming=False, min_overlap=1, scaling=1.0, barcode_weights=None, mask=None, seed=20261019)
    if name == "cfg2":
        w.update(n_barcodes=1, iterations=100, desc="I mode, hg38-size (3.1 Gb, 24 contigs), 1 barcode, 30x, 100 insertions of 5 kb, 100 iterations")
        w["targets"] = S.random_insertions(G, 100, 5000, 1, seed=2)
    elif name.startswith("sweep:"):  # a cell of BASELINE.json configs[4]: "sweep:<barcodes>:<coverage>" on the config-2 shape
        _, nb, cov = name.split(":")
        w.update(n_barcodes=int(nb), coverage=float(cov), iterations=100,
                 desc=f"config-2 shape (hg38-size, 100 insertions of 5 kb per barcode), {nb} barcode(s), {cov}x, 100 iterations")
        w["targets"] = S.random_insertions(G, 100, 5000, int(nb), seed=2)
    elif name == "cfg3":
        w.update(n_barcodes=24, iterations=125, barcode_weights={"0": 10, "1": 5},
                 desc="I mode, hg38-size, 24 barcodes (weights 0:10, 1:5), 30x, 10k insertions per barcode, 125 iterations per GPU")
        w["targets"] = S.random_insertions(G, 10000, 5000, 24, seed=3)
    elif name == "cfg4":
        w.update(n_barcodes=1, iterations=100, min_overlap=1000, mean_read_length=8000,
                 desc="ROI mode, hg38-size, 50k ROIs, blocked centromere/gap mask, min_overlap 1000, 100 iterations")
        w["targets"] = S.barcode_rois(S.random_rois(S.HG38, 50000, 1000, 100000, seed=4), 1)
        w["mask"] = S.gap_mask(S.HG38, 800, seed=5)
    elif name == "chr":
        w.update(n_barcodes=1, iterations=100,
                 desc="ROI mode, hg38-size, 24 whole-chromosome ROIs + 76 ROIs of 1-100 kb, 1 barcode, 30x, 100 iterations")
        total, cdir = S.chromosome_dir(S.HG38)
        rois = {f"C{i}": {"start": s, "end": e, "weight": 1, "barcode": []} for i, (s, e) in enumerate(cdir.values())}
        rois.update(S.random_rois(S.HG38, 76, 1000, 100000, seed=6))
        w["targets"] = S.barcode_rois(rois, 1)
    else:
        raise SystemExit(f"unknown workload {name}")
    w["pool"] = S.lognormal_lengths(w["mean_read_length"], w["sigma"], w["min_read_length"], seed=1)
    return w


# ---------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi needs a few hundred ms to start: make sure it is already sampling when the timed region begins."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06]
        if not rows and self.rows:  # region shorter than one sampling period: take the sample nearest to it
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))[1]]
        sm, smax, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU baseline

def cpu_port_sample(w, n_threads, iters_per_thread=1, seed0=1000):
    """Reference algorithm on host cores: oracle.process_combination (numpy-legacy stream, 1 M log-normal pool,
    10x resample, placement loop, faithful O(T*R) tally) for one iteration per thread, all threads at once."""
    from oracle import oracle as O
    from esloco_b200.utils import targets_to_arrays
    O.lib()
    keys, ts, te, tb = targets_to_arrays(w["targets"], w["n_barcodes"])
    mask = None
    if w["mask"]:
        from esloco_b200.utils import build_mask_tree, mask_tree_to_arrays
        mask = mask_tree_to_arrays(build_mask_tree(w["mask"]), w["n_barcodes"])
    reads = [0] * n_threads

    def work(i):
        for j in range(iters_per_thread):
            rng = O.Rng(seed0 + i * 7919 + j)
            out = O.process_combination(rng, w["mean_read_length"], w["coverage"], w["genome_size"], ts, te, tb,
                                        w["n_barcodes"], w["barcode_weights"], w["consuming"], mask, w["sigma"],
                                        w["min_read_length"], w["scaling"], w["min_overlap"])
            reads[i] += out["placement"]["n_placed"]

    threads = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    t0 = time.perf_counter()
    for t in threads: t.start()
    for t in threads: t.join()
    dt = time.perf_counter() - t0
    return sum(reads), n_threads * iters_per_thread, dt


def host_threads(w):
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_iter_gb = 1.5 + 40e-9 * w["coverage"] * w["genome_size"] / w["mean_read_length"] * 10  # 10x resample + read arrays
    try:
        avail_gb = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) / 1e6
        n = max(1, min(n, int(avail_gb * 0.6 / per_iter_gb)))
    except Exception:
        pass
    return n


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_thr = host_threads(w)
    cpu_port_sample(dict(w, coverage=0.05), min(2, n_thr))  # warm the library / page in
    tot_reads = tot_it = 0
    tot_t = 0.0
    done = 0
    for _ in range(max(1, args.steps)):  # bounded: stop after ~200 s of CPU sampling even if fewer than K steps ran
        r, it, dt = cpu_port_sample(w, n_thr)
        tot_reads += r; tot_it += it; tot_t += dt
        done += 1
        if tot_t > 200.0:
            break
    val = tot_reads / tot_t
    sample = f"{n_thr} iterations per step (one per thread) of the same workload, {done} of {args.steps} step(s) run (time-bounded)"
    line = {"impl": "reference", "metric": "simulated reads placed per second (hg38-size, 30x)", "value": val,
            "unit": "reads/s", "iterations_per_s": tot_it / tot_t, "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, done), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": w["desc"], "impl": "CPU port of the reference algorithm (oracle/esloco_oracle.c), numpy-legacy MT19937 stream"},
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": n_thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm

def run_ours(args, w):
    import torch
    import torch.distributed as dist
    from esloco_b200.combined_calculations import run_simulation_batch
    from esloco_b200.session import get_session
    from esloco_b200.utils import build_mask_tree

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sess = get_session(local)
    eng = sess.engine
    n_iter = w["iterations"]
    params = dict(n_barcodes=w["n_barcodes"], barcode_weights=w["barcode_weights"], seed=w["seed"], scaling=w["scaling"],
                  combinations=[(w["mean_read_length"], w["coverage"])], sequenced_data_path=None, sigma=w["sigma"],
                  min_read_length=w["min_read_length"], consuming=w["consuming"], min_overlap_for_detection=w["min_overlap"])
    masked_tree = build_mask_tree(w["mask"]) if w["mask"] else None
    # host inputs of the end-to-end leg live in pinned memory (the set-up calls copy straight from them)
    pool_pinned = torch.from_numpy(w["pool"].astype(np.uint32)).pin_memory()
    pool = pool_pinned.numpy()
    sess.configure(w["genome_size"], w["targets"], masked_tree, w["n_barcodes"], w["barcode_weights"])
    sess.set_pool(("bench", 0), lambda: pool)
    T = eng.n_targets
    out = eng.alloc_outputs(n_iter)
    acc = torch.zeros((T, 8), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(k):
        # rank r owns iterations [r*n_iter + k*world*n_iter ...): disjoint iteration ids across ranks and steps
        it0 = (k * world + rank) * n_iter
        if args.path == "staged":
            eng.run_batch_staged(w["seed"], 0, it0, n_iter, w["coverage"], w["consuming"], w["min_overlap"], out=out)
        else:
            eng.run_batch(w["seed"], 0, it0, n_iter, w["coverage"], w["consuming"], w["min_overlap"], w["scaling"], out=out)
        if world > 1:
            acc.zero_()
            eng.moments(out.tallies, acc)
            dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()  # before the warm-up, so the GPU does not idle between warm-up and timing
    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.set_profiling(True)
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    bulk_ms = cut_ms = 0.0
    reads = 0
    t_wall0 = time.time()
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush between timed steps, outside the step's events
        ev[k][0].record()
        step(args.warmup + k)
        ev[k][1].record()
        b, c = eng.last_timing_ms()  # waits for the step
        bulk_ms += b; cut_ms += c
        reads += int(out.n_reads.sum().item())
        assert not (out.status & 2).any().item()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = eng.launch_count() - launches0 + (args.steps if world > 1 else 0)
    eng.set_profiling(False)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    stats = torch.tensor([dev_ms, float(reads)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, reads_all = mx[0].item(), sm[1].item()
    else:
        reads_all = float(reads)
    value = reads_all / (dev_ms * 1e-3)
    iters_all = n_iter * args.steps * world

    # ---- end to end through the public host API with HOST inputs: every step re-uploads targets / mask / length
    # pool from host memory and reads all results back
    h2d = int(T * (8 + 8 + 4) + pool.size * 4 + (len(w["mask"]) * 28 if w["mask"] else 0) + w["n_barcodes"] * 8)
    d2h = int(n_iter * (T * 3 * 8 + (w["n_barcodes"] + 2) * 8 + 8 + 8 + 4))
    e2e_reads = 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sess.invalidate()
        its = range((10_000 + k * world + rank) * n_iter, (10_000 + k * world + rank) * n_iter + n_iter)
        res, _ = run_simulation_batch(its, params, w["genome_size"], w["targets"], masked_tree, device=local, to_host=True,
                                      length_pools={0: pool}, reuse_host_buffers=True)
        e2e_reads += int(res[0].n_reads.sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = torch.tensor([e2e_s, float(e2e_reads)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = e2e.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2e.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_s, e2e_reads = mx[0].item(), sm[1].item()
    e2e_val = e2e_reads / e2e_s

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # Algorithmic bytes of what the bulk kernel processed in one step.  Single bulk pass: exactly n_bulk draws per
        # iteration.  Masked + non-consuming runs add a second bulk pass whose per-iteration extent is decided on the
        # device; there all placed reads are counted against the bulk passes AND the cut-off kernel's time.
        two_pass = bool(w["mask"]) and not w["consuming"]
        if two_pass or args.path == "staged":
            reads_kernel, kernel_ms = reads_all / world / args.steps, (bulk_ms + cut_ms) / args.steps
        else:
            reads_kernel, kernel_ms = float(eng.bulk_reads(w["coverage"]) * n_iter), bulk_ms / args.steps
        alg_bytes = ALG_BYTES_PER_READ * reads_kernel + ALG_BYTES_PER_TARGET * T * n_iter
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, summary = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and args.path == "fused":
            tj = json.load(open(tp))
            traffic = tj.get(w["name"])  # bytes per launch, from the committed ncu capture
            summary = tj.get("summaries", {}).get(w["name"])
        # what actually binds the fused kernel, from the committed ncu capture of this workload (not measured live)
        binding = None
        sp = os.path.join(ROOT, "profiles", summary or "none")
        if os.path.exists(sp):
            ps = json.load(open(sp))
            binding = {"source": os.path.relpath(sp, ROOT), "warp_instructions_per_read": round(ps["warp_instructions_per_read"], 1),
                       "issue_slots_busy_pct": float(ps["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                       "fmaheavy_pipe_pct": float(ps["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
                       "dram_bytes_per_launch": sum(float(ps[k] or 0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(ps["units"].get(k), 1)
                                                    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))}
        n_thr = host_threads(w)
        if args.no_cpu_baseline or world > 1:  # the CPU baseline is reported at N = 1 only
            cpu = None
        else:
            r, it, dt = cpu_port_sample(w, n_thr)
            cpu = {"value": r / dt, "unit": "reads/s", "cores": n_thr, "kind": "port",
                   "iterations_per_s": it / dt,
                   "sample": f"{it} iterations (one per thread) of the same workload, C port of the reference algorithm"}
        line = {"metric": "simulated reads placed per second (hg38-size, 30x)", "value": value, "unit": "reads/s",
                "iterations_per_s": iters_all / (dev_ms * 1e-3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32" if w["genome_size"] <= 0xFFFFFFF0 else "u64", "data": "synthetic",
                "config": {"workload": w["desc"], "path": args.path, "iterations_per_step_per_gpu": n_iter, "targets": T,
                           "reads_per_iteration": reads_all / iters_all, "rng": "philox4x32-10",
                           "l2": "flushed between steps (256 MiB write) outside the per-step CUDA events",
                           "parallelism": f"iterations sharded over {world} GPU(s)" + (", one NCCL sum-reduce of per-target moments per step" if world > 1 else "")},
                "roofline": {"bound": "hbm", "kernel": "k_place_tally_bulk" if args.path == "fused" else "staged pipeline (20 kernels per iteration)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": kernel_ms,
                             "launches_per_step": 2 if two_pass else 1,
                             "kernel_share_of_step": bulk_ms / dev_ms if world == 1 else None,
                             "cutoff_kernel_ms_per_launch": cut_ms / args.steps,
                             "note": "achieved = ALGORITHMIC bytes of the staged design (96 B/read + 32 B/target/iteration, SURVEY 8d) / kernel time; "
                                     "the fused kernel keeps reads in registers, so its real DRAM traffic is `traffic` and frac may exceed 1",
                             "binding_resource": binding},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", default="fused", choices=["fused", "staged"],
                    help="fused = the product kernels; staged = the north star's sort + sweep pipeline kept as a measured baseline")
    args = ap.parse_args()
    w = make_workload(args.workload)
    if args.impl == "reference":
        run_reference_arm(args, w)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, w)


if __name__ == "__main__":
    main()
