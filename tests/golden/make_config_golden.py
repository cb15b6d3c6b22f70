#!/usr/bin/env python3
"""Record what the UNMODIFIED reference's parse_config returns for a few INI files (build container only):
tests/golden/config_golden.json = [{"ini": text, "params": dict}], paths kept relative to a {tmp} placeholder."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "baseline", "shims"))
sys.path.insert(0, "/root/reference/src")
from esloco.config_handler import parse_config  # noqa: E402

INIS = [
    "[COMMON]\nmode = ROI\nreference_genome_path = {tmp}/g.fa\noutput_path = {tmp}/o/\nseed = 1\n\n[ROI]\nroi_bedpath = {tmp}/r.bed\n",
    "[COMMON]\nmode = I\nreference_genome_path = {tmp}/g.fa\noutput_path = {tmp}/o2/\nexperiment_name = e\nsigma = 2\n"
    "min_overlap_for_detection = 250\nn_barcodes = 4\nbarcode_weights = {\"0\": 10, \"2\": 3}\niterations = 12\nscaling = 0.5\n"
    "min_read_length = 300\nno_cov_plots = True\nparallel_jobs = 3\nseed = 77\ncoverages = [0.5, 5, 30]\nmean_read_lengths = [1000, 15000]\n"
    "blocked_regions_bedpath = {tmp}/b.bed\nconsuming = True\nchr_restriction = unrestricted\noutput_path_plots = {tmp}/plots/\n\n"
    "[I]\ninsertion_length = 4321\ninsertion_number_distribution = poisson\nbedpath = {tmp}/i.bed\ninsertion_numbers = 12.5\n",
    "[COMMON]\nmode = I\nreference_genome_path = {tmp}/g.fa\noutput_path = {tmp}/o3/\nseed = 3\ncoverages = 7\nmean_read_lengths = 9000\n\n[I]\n",
    "[COMMON]\nmode = ROI\nreference_genome_path = {tmp}/g.fa\noutput_path = {tmp}/o4/\nseed = 4\nsequenced_data_path = {tmp}/reads.fastq\n"
    "min_read_length = 20\n\n[ROI]\nroi_bedpath = {tmp}/r.bed\n",
]

if __name__ == "__main__":
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "reads.fastq"), "w") as fh:
            for i, n in enumerate((10, 30, 55, 100, 21)):
                fh.write(f"@r{i}\n{'A' * n}\n+\n{'I' * n}\n")
        for k, text in enumerate(INIS):
            ini = os.path.join(tmp, f"c{k}.ini")
            open(ini, "w").write(text.replace("{tmp}", tmp))
            p = parse_config(ini)
            p = json.loads(json.dumps(p, default=str).replace(tmp, "{tmp}"))
            out.append({"ini": text, "params": p})
    json.dump(out, open(os.path.join(HERE, "config_golden.json"), "w"), indent=1)
    print(len(out), "configs recorded;", sorted(out[1]["params"]))
