#!/usr/bin/env python3
"""Run the UNMODIFIED reference CLI (`esloco --config`) on two tiny configs and record the FORMAT of its output
tables (column order, the RNG-independent name columns, row order) as tests/golden/cli_format.json.
Build container only (needs /root/reference); see make_golden.py for the stand-in modules."""
import json
import os
import subprocess
import sys
import tempfile

import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import write_bed, write_fasta  # noqa: E402

CONTIGS = [("chr1", 300_000), ("chr2", 200_000), ("chrX", 100_000), ("chr1_alt", 50_000), ("chrM", 16_000)]
COMMON = """[COMMON]
mode = {mode}
reference_genome_path = {fasta}
output_path = {out}/
experiment_name = {name}
n_barcodes = 2
iterations = 3
parallel_jobs = 2
coverages = [2, 4]
mean_read_lengths = [3000]
no_cov_plots = True
seed = 7
min_overlap_for_detection = 10
blocked_regions_bedpath = {blocked}
barcode_weights = {{"0": 3}}
"""


def run_case(mode, tmp):
    fasta = os.path.join(tmp, "ref.fa"); write_fasta(fasta, CONTIGS)
    blocked = os.path.join(tmp, "blocked.bed")
    write_bed(blocked, ["chrom", "start", "end", "id", "barcode", "weight"], [("chr2", 1000, 90_000, "b1", "[]", 1)])
    out = os.path.join(tmp, "out_" + mode); os.makedirs(out)
    ini = os.path.join(tmp, mode + ".ini")
    text = COMMON.format(mode=mode, fasta=fasta, out=out, name="exp" + mode, blocked=blocked)
    if mode == "ROI":
        roi = os.path.join(tmp, "roi.bed")
        write_bed(roi, ["chrom", "start", "end", "id"], [("chr1", 1000, 5000, "RA"), ("chrX", 0, 0, "RX"), ("chr2", 100_000, 120_000, "RB")])
        text += f"\n[ROI]\nroi_bedpath = {roi}\n"
    else:
        text += "\n[I]\ninsertion_length = 2000\ninsertion_numbers = 4\n"
    open(ini, "w").write(text)
    env = dict(os.environ, PYTHONPATH=os.path.join(HERE, "..", "..", "baseline", "shims") + ":/root/reference/src")
    code = f"import sys; sys.argv=['esloco','--config',{ini!r}]; from esloco.esloco import main; main()"
    subprocess.run([sys.executable, "-c", code], env=env, check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    res = {"ini": text.replace(tmp, "{tmp}")}
    for table in ("matches_table.csv", "barcode_distribution_table.csv", "summary.csv") + (("insertion_locations.bed",) if mode == "I" else ()):
        df = pd.read_csv(os.path.join(out, f"exp{mode}_{table}"), sep="\t", dtype=str)
        entry = {"columns": list(df.columns), "rows": len(df)}
        if table == "matches_table.csv":
            for c in ("target_region", "target", "barcode", "iteration", "mean_read_length", "coverage") + (("insertion",) if mode == "I" else ()):
                entry[c] = df[c].tolist()
        if table == "summary.csv":
            for c in ("target", "mean_read_length", "coverage"):
                entry[c] = df[c].tolist()
        if table == "barcode_distribution_table.csv":
            for c in ("coverage", "mean_read_length", "iteration"):
                entry[c] = df[c].tolist()
        if table == "insertion_locations.bed":
            entry["item"] = df["item"].tolist()
        res[table] = entry
    return res


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        out = {mode: run_case(mode, tmp) for mode in ("ROI", "I")}
    json.dump(out, open(os.path.join(HERE, "cli_format.json"), "w"), indent=1)
    for m, v in out.items():
        print(m, {k: (e["columns"], e["rows"]) for k, e in v.items() if k != "ini"})
