#!/usr/bin/env python3
"""Multi-iteration SAMPLES of the unmodified reference for the statistical check (north_star check 2).

Run in the build container only (needs /root/reference):

    python tests/golden/make_stat_golden.py         # rewrites tests/golden/stat/*.npz (a few KB each)

The native engine draws from a Philox stream, so it cannot reproduce the reference's MT19937 results read for
read; what it must reproduce is their DISTRIBUTION: "per-target mean coverage and detection rate must agree with
the reference within 3 standard errors over the same iteration count".  This script calls the reference's own
run_simulation_iteration (combined_calculations.py:97-126) N_ITER times per case, iteration i in-process after
np.random.seed(seed + i) (the reference's own seeding only works sequentially), and records
per iteration and target on_target_bases, full and partial matches, plus N_bases, reads placed and the label
histogram.  Nothing from the reference's source is copied; only its outputs are recorded.

Cases
  stat_cfg2_reduced_50mb   the inputs of the golden `cfg2_reduced_50mb` (config-2 shape at G = 5e7, 5x, 100
                           insertions of 5 kb, 1 barcode)
  stat_masked_weighted     3 barcodes with barcode_weights {"0": 4}, blocked regions (weight 1 for all barcodes,
                           weight 0.5 for barcode 1 only), not consuming, min_overlap 200, 40 ROIs x 3 barcodes
"""
import json
import logging
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "baseline", "shims"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, ".."))

import esloco.combined_calculations as cc  # noqa: E402
from esloco.utils import build_mask_tree  # noqa: E402
from golden_util import Golden  # noqa: E402

logging.disable(logging.CRITICAL)
N_ITER = 64
LABELS_BLOCKED = ("Blocked/Consuming", "Blocked/NotConsuming")

_job = {}


def one_iteration(i):
    c = _job
    np.random.seed(c["seed"] + i)
    results, dists = cc.run_simulation_iteration(i, c["params"], c["genome_size"], c["targets"], c["tree"], os.devnull)
    det, dist = results[0], dists[0]
    B = c["params"]["n_barcodes"]
    labels = [dist.get(str(k), 0) for k in range(B)] + [dist.get(k, 0) for k in LABELS_BLOCKED]
    return (np.array([int(d["on_target_bases"]) for d in det], np.int64), np.array([d["full_matches"] for d in det], np.int32),
            np.array([d["partial_matches"] for d in det], np.int32), int(dist["N_bases"]), np.array(labels, np.int64))


def run_case(name, genome_size, targets, masked, params, seed):
    _job.update(seed=seed, params=params, genome_size=genome_size, targets=targets,
                tree=build_mask_tree(masked) if masked else None)
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        rows = pool.map(one_iteration, range(N_ITER))
    keys = list(targets.keys())
    ts = np.array([v[0] if isinstance(v, list) else v["start"] for v in targets.values()], np.int64)
    te = np.array([v[1] if isinstance(v, list) else v["end"] for v in targets.values()], np.int64)
    ms = [(v["start"], v["end"], float(v["weight"]), v["barcode"]) for v in (masked or {}).values()]
    meta = dict(name=name, genome_size=int(genome_size), target_keys=keys, n_iter=N_ITER, seed=seed,
                mask=[[int(s), int(e), w, b] for s, e, w, b in ms],
                params={k: v for k, v in params.items() if k not in ("output_path_plots",)})
    os.makedirs(os.path.join(HERE, "stat"), exist_ok=True)
    path = os.path.join(HERE, "stat", name + ".npz")
    np.savez_compressed(path, target_start=ts, target_end=te,
                        otb=np.stack([r[0] for r in rows]), full=np.stack([r[1] for r in rows]),
                        partial=np.stack([r[2] for r in rows]), n_bases=np.array([r[3] for r in rows], np.int64),
                        label_counts=np.stack([r[4] for r in rows]),
                        meta_json=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    otb = np.stack([r[0] for r in rows]).astype(float)
    print(name, "bytes", os.path.getsize(path), "mean depth", (otb / (te - ts)).mean(), flush=True)


def base_params(**kw):
    p = dict(sequenced_data_path=None, sigma=1, min_read_length=1, n_barcodes=1, barcode_weights=None, consuming=False,
             output_path_plots="/tmp", scaling=1.0, min_overlap_for_detection=1, no_cov_plots=True)
    p.update(kw)
    return p


def main():
    g = Golden("cfg2_reduced_50mb")
    m = g.meta
    targets = {k: [int(s), int(e)] for k, s, e in zip(m["target_keys"], g["target_start"], g["target_end"])}
    run_case("stat_cfg2_reduced_50mb", m["genome_size"], targets, None,
             base_params(combinations=[tuple(m["combinations"][0])]), seed=5000)

    r = np.random.RandomState(77)
    G = 20_000_000
    rois = {}
    starts = np.sort(r.randint(0, G - 60_000, 40))
    for i, s in enumerate(starts):
        ln = int(r.randint(500, 40_000))
        for b in range(3):
            rois[f"R{i}_{b}"] = {"start": int(s), "end": int(s) + ln, "weight": 1, "barcode": []}
    # ROI barcoding is b-major in the reference ({id}_{b} for b in range(n) for id ...); order does not matter here
    masked = {"cen": {"start": 9_000_000, "end": 10_500_000, "weight": 1, "barcode": []},
              "gapA": {"start": 2_000_000, "end": 2_300_000, "weight": 1, "barcode": []},
              "half1": {"start": 14_000_000, "end": 17_000_000, "weight": "0.5", "barcode": ["1"]}}
    run_case("stat_masked_weighted", G, rois, masked,
             base_params(n_barcodes=3, barcode_weights={"0": 4}, min_overlap_for_detection=200, combinations=[(5000, 8)]),
             seed=6000)


if __name__ == "__main__":
    main()
