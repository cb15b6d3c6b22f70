#!/usr/bin/env python3
"""BASELINE.json configs[4]: sweep coverage x barcodes x iterations on the config-2 genome / target shape and record
reads placed/s and iterations/s per cell.  One GPU per process; under torchrun the iterations of a cell are sharded
over the ranks (contiguous blocks, as the CLI does) and the time is the max over ranks.  `--cpu` adds, per (barcodes,
coverage), the C port of the reference algorithm on all host threads (one iteration per thread).  JSON lines:
    python profiles/sweep.py --cpu > profiles/r1j_sweep_n1.jsonl
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/sweep.py > profiles/r1j_sweep_n8.jsonl"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from esloco_b200 import synthetic as S
    from esloco_b200.distributed import shard_iterations
    from esloco_b200.session import get_session
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G, _ = S.chromosome_dir(S.HG38)
    pool = S.lognormal_lengths(15000, 1, 1, seed=1)
    sess = get_session(local)
    eng = sess.engine
    for n_barcodes in (1, 8, 24, 96):
        targets = S.random_insertions(G, 100, 5000, n_barcodes, seed=2)
        sess.invalidate()
        sess.configure(G, targets, None, n_barcodes, None)
        sess.set_pool(("sweep",), lambda: pool)
        cpu = {}
        for coverage, n_iter in itertools.product((5, 10, 30, 60), (10, 100, 1000, 10000)):
            mine = shard_iterations(n_iter, rank, world)  # this rank's iteration ids
            first, n_mine = (mine[0], len(mine)) if len(mine) else (0, 0)
            out = eng.alloc_outputs(max(n_mine, 1))
            if n_mine:
                eng.run_batch(1, 0, first, min(n_mine, 100), coverage, out=out)  # warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            reps = 3 if n_iter * coverage <= 30000 else 1
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for r in range(reps):
                if n_mine:
                    eng.run_batch(1, 0, (r + 1) * n_iter + first, n_mine, coverage, out=out)
            ev1.record()
            torch.cuda.synchronize()
            stats = torch.tensor([ev0.elapsed_time(ev1) / reps, float(out.n_reads[:n_mine].sum().item()) if n_mine else 0.0],
                                 dtype=torch.float64, device="cuda")
            if world > 1:
                mx, sm = stats.clone(), stats.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                ms, reads = mx[0].item(), sm[1].item()
            else:
                ms, reads = stats[0].item(), stats[1].item()
            row = {"n_gpus": world, "n_barcodes": n_barcodes, "targets": len(targets), "coverage": coverage, "iterations": n_iter,
                   "ms_per_batch": round(ms, 3), "reads_per_s": reads / (ms * 1e-3), "iterations_per_s": n_iter / (ms * 1e-3),
                   "reads_per_iteration": reads / n_iter}
            if "--cpu" in sys.argv and rank == 0:
                if coverage not in cpu:  # the CPU port's rate does not depend on the iteration count: one sample per coverage
                    # bench.py's reference arm on this cell (the CPU port is test / baseline infrastructure: bench.py runs it)
                    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                          "--workload", f"sweep:{n_barcodes}:{coverage}"], capture_output=True, text=True, check=True)
                    line = json.loads(ref.stdout.strip().splitlines()[-1])
                    cpu[coverage] = {"cpu_reads_per_s": line["value"], "cpu_iterations_per_s": line["iterations_per_s"],
                                     "cpu_threads": line["cpu_baseline"]["cores"]}
                row.update(cpu[coverage])
            if rank == 0:
                print(json.dumps(row), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
