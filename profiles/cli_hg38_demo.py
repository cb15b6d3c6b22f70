#!/usr/bin/env python3
"""End-to-end `esloco --config` at BASELINE.json configs[1] scale on the GPU box: writes a synthetic hg38-sized FASTA
(24 contigs, 3.1 Gb), an INI for I mode / 1 barcode / 30x / 100 insertions of 5 kb / 100 iterations, runs the CLI
and reports wall-clock per phase plus sanity checks of the tables.  usage: python profiles/cli_hg38_demo.py [workdir]"""
import json
import os
import subprocess
import sys
import time

import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esloco_b200.synthetic import HG38  # noqa: E402


def main():
    work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/esl_hg38_demo"
    os.makedirs(work, exist_ok=True)
    fasta = os.path.join(work, "hg38_synthetic.fa")
    t0 = time.time()
    line = ("ACGT" * 15 + "\n").encode()  # 60 bases per line like UCSC hg38.fa
    with open(fasta, "wb") as fh:
        for name, n in HG38:
            fh.write(f">{name}\n".encode())
            full, rest = divmod(n, 60)
            block = line * 100000
            for _ in range(full // 100000):
                fh.write(block)
            fh.write(line * (full % 100000))
            if rest:
                fh.write(line[:rest] + b"\n")
    t_fasta = time.time() - t0
    ini = os.path.join(work, "cfg2.ini")
    open(ini, "w").write(f"""[COMMON]
mode = I
reference_genome_path = {fasta}
output_path = {work}/out/
experiment_name = cfg2
n_barcodes = 1
iterations = 100
coverages = [30]
mean_read_lengths = [15000]
no_cov_plots = False
seed = 20261019

[I]
insertion_length = 5000
insertion_numbers = 100
""")
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "esloco_b200", "--config", ini], cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True)
    t_cli = time.time() - t0
    out = os.path.join(work, "out")
    res = {"fasta_bytes": os.path.getsize(fasta), "write_fasta_s": round(t_fasta, 1), "cli_wall_s": round(t_cli, 1), "cli_rc": r.returncode}
    if r.returncode == 0:
        m = pd.read_csv(os.path.join(out, "cfg2_matches_table.csv"), sep="\t")
        d = pd.read_csv(os.path.join(out, "cfg2_barcode_distribution_table.csv"), sep="\t")
        s = pd.read_csv(os.path.join(out, "cfg2_summary.csv"), sep="\t")
        cov = pd.read_csv(os.path.join(out, "15000_30_coverage.tsv"), sep="\t")
        res.update(matches_rows=len(m), summary_rows=len(s), mean_depth_on_targets=float((m["on_target_bases"] / 5000).mean()),
                   reads_per_iteration=float(d["total_Reads"].mean()), n_bases_min=int(d["N_bases"].min()),
                   coverage_track_bins=len(cov), coverage_track_mean=float(cov["Combined"].mean()),
                   log_tail=[ln for ln in open(os.path.join(out, "cfg2_log.log")).read().splitlines() if "seconds" in ln][-3:])
    else:
        res["stderr"] = r.stderr[-1500:]
    # the same CLI at the config-3 shape: 24 weighted barcodes, 10 k insertions per barcode (240 k targets), 40 iterations
    # = 9.6 M rows of matches_table.csv through the pipelined native row writer
    ini3 = os.path.join(work, "cfg3.ini")
    open(ini3, "w").write(f"""[COMMON]
mode = I
reference_genome_path = {fasta}
output_path = {work}/out3/
experiment_name = cfg3
n_barcodes = 24
barcode_weights = {{"0": 10, "1": 5}}
iterations = 40
coverages = [30]
mean_read_lengths = [15000]
no_cov_plots = True
seed = 20261019
label_columns = reference

[I]
insertion_length = 5000
insertion_numbers = 10000
""")
    t0 = time.time()
    prof = os.path.join(work, "cfg3.prof")
    cmd = [sys.executable] + (["-m", "cProfile", "-o", prof] if "--profile" in sys.argv else []) + ["-m", "esloco_b200", "--config", ini3]
    r3 = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True)
    if "--profile" in sys.argv and os.path.exists(prof):
        import io, pstats
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(45)
        open(os.path.join(ROOT, "gpurun_out", "cli_cfg3_cprofile.txt"), "w").write(buf.getvalue())
    res["cfg3"] = {"cli_wall_s": round(time.time() - t0, 1), "cli_rc": r3.returncode}
    if r3.returncode == 0:
        out3 = os.path.join(work, "out3")
        mt = os.path.join(out3, "cfg3_matches_table.csv")
        d3 = pd.read_csv(os.path.join(out3, "cfg3_barcode_distribution_table.csv"), sep="\t")
        s3 = pd.read_csv(os.path.join(out3, "cfg3_summary.csv"), sep="\t")
        with open(mt, "rb") as fh:
            rows = sum(chunk.count(b"\n") for chunk in iter(lambda: fh.read(64 << 20), b"")) - 1
        res["cfg3"].update(matches_rows=rows, matches_bytes=os.path.getsize(mt), summary_rows=len(s3),
                           label_columns=list(d3.columns[:list(d3.columns).index("coverage")])[:6],
                           mean_depth_barcode0=float(s3[s3["target"].str.startswith("Barcode_0_")]["mean_bases_on_target"].mean() / 5000),
                           log_tail=[ln for ln in open(os.path.join(out3, "cfg3_log.log")).read().splitlines() if "seconds" in ln][-3:])
    else:
        res["cfg3"]["stderr"] = r3.stderr[-1500:]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
