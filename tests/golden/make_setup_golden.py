#!/usr/bin/env python3
"""Record the UNMODIFIED reference's host set-up outputs (build container only) for BED variants the hot-path
fixtures do not cover: header-less BEDs, BED-guided insertion placement (length-weighted, explicitly weighted, and the
"one insertion per row" case), Poisson insertion counts.  -> tests/golden/setup_golden.json"""
import json
import logging
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "baseline", "shims"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, HERE)
from make_golden import write_bed, write_fasta  # noqa: E402
from esloco.genome_generation import create_barcoded_insertion_genome, create_barcoded_roi_genome  # noqa: E402

logging.disable(logging.CRITICAL)
CONTIGS = [("chr1", 500_000), ("chr2", 300_000), ("chrX", 200_000), ("chr2_random", 1000)]
HDR = ["chrom", "start", "end", "id", "barcode", "weight"]

CASES = [
    dict(name="guided_length_weighted", mode="I", n_barcodes=2, insertion_length=700, insertion_numbers=9.0, dist=None,
         guide=(None, [("chr1", 1000, 50_000), ("chr2", 10_000, 20_000), ("chrX", 100, 150_000)]), blocked=None, seed=11),
    dict(name="guided_explicit_weights", mode="I", n_barcodes=2, insertion_length=500, insertion_numbers=6.0, dist=None,
         guide=(HDR, [("chr1", 1000, 50_000, "a", "[]", 5), ("chr2", 10_000, 20_000, "b", "[]", 1), ("chrX", 100, 150_000, "c", "[]", 2)]),
         blocked=None, seed=12),
    dict(name="guided_one_per_row", mode="I", n_barcodes=1, insertion_length=300, insertion_numbers=3.0, dist=None,
         guide=(None, [("chr1", 1000, 50_000), ("chr2", 10_000, 20_000), ("chrX", 100, 150_000)]), blocked=None, seed=13),
    dict(name="poisson_count_headerless_blocked", mode="I", n_barcodes=3, insertion_length=1000, insertion_numbers=5.0, dist="poisson",
         guide=None, blocked=(None, [("chr1", 100, 2000), ("chrX", 0, 0), ("chr9", 5, 50)]), seed=14),
    dict(name="roi_headerless_is_unmatchable", mode="ROI", n_barcodes=2, roi=(None, [("chr1", 10, 2000), ("chr2", 0, 0)]),
         blocked=(HDR, [("chr2", 5, 500, "m", '["1"]', 0.3)]), seed=15),
]


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


if __name__ == "__main__":
    out = []
    for c in CASES:
        with tempfile.TemporaryDirectory() as tmp:
            fasta = os.path.join(tmp, "g.fa"); write_fasta(fasta, CONTIGS)
            def bed(spec, name):
                if spec is None:
                    return None
                p = os.path.join(tmp, name); write_bed(p, spec[0], spec[1]); return p
            blocked = bed(c.get("blocked"), "blocked.bed")
            np.random.seed(c["seed"])
            if c["mode"] == "I":
                G, targets, masked, cdir = create_barcoded_insertion_genome(
                    1, fasta, bed(c.get("guide"), "guide.bed"), blocked, None, c["insertion_length"], c["insertion_numbers"],
                    c["dist"], c["n_barcodes"], os.path.join(tmp, "log"))
            else:
                targets, _bed, masked, G, cdir = create_barcoded_roi_genome(fasta, None, bed(c["roi"], "roi.bed"), c["n_barcodes"], blocked)
            tail = float(np.random.rand())  # position of the stream after set-up
        out.append(dict(case=jsonable(c), contigs=CONTIGS, genome_size=int(G), targets=jsonable(targets), masked=jsonable(masked),
                        chromosome_dir=jsonable(cdir), stream_tail=tail))
        print(c["name"], len(targets), "targets", 0 if not masked else len(masked), "masked")
    json.dump(out, open(os.path.join(HERE, "setup_golden.json"), "w"), indent=1)
