#!/usr/bin/env python3
"""BASELINE.json configs[4]: sweep coverage x barcodes x iterations on the config-2 genome / target shape and record
reads placed/s and iterations/s per cell (one GPU per process; run under torchrun for more).  Writes JSON lines.
    python profiles/sweep.py > profiles/r1_sweep_1gpu.jsonl"""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from esloco_b200 import synthetic as S
    from esloco_b200.session import get_session
    G, _ = S.chromosome_dir(S.HG38)
    pool = S.lognormal_lengths(15000, 1, 1, seed=1)
    sess = get_session(0)
    eng = sess.engine
    for n_barcodes in (1, 8, 24, 96):
        targets = S.random_insertions(G, 100, 5000, n_barcodes, seed=2)
        sess.invalidate()
        sess.configure(G, targets, None, n_barcodes, None)
        sess.set_pool(("sweep",), lambda: pool)
        for coverage, n_iter in itertools.product((5, 10, 30, 60), (10, 100, 1000)):
            out = eng.alloc_outputs(n_iter)
            eng.run_batch(1, 0, 0, n_iter, coverage, out=out)  # warm-up
            torch.cuda.synchronize()
            reps = 3
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for r in range(reps):
                eng.run_batch(1, 0, (r + 1) * n_iter, n_iter, coverage, out=out)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / reps
            reads = int(out.n_reads.sum().item())
            print(json.dumps({"n_barcodes": n_barcodes, "targets": len(targets), "coverage": coverage, "iterations": n_iter,
                              "ms_per_batch": round(ms, 3), "reads_per_s": reads / (ms * 1e-3), "iterations_per_s": n_iter / (ms * 1e-3),
                              "reads_per_iteration": reads / n_iter}), flush=True)


if __name__ == "__main__":
    main()
