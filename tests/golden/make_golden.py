#!/usr/bin/env python3
"""Generate golden vectors from the UNMODIFIED reference (aweich/esloco v1.0.6).

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference has no tests or fixtures of its own, so every
fixture here is produced by importing /root/reference/src through the stand-in
modules in baseline/shims/ (intervaltree, Bio.SeqIO, plotly, pyfiglet,
tqdm_joblib are not installed) and calling the reference's own hot-path
functions in-process after `np.random.seed(seed)`:

  * process_combination            combined_calculations.py:11-95
  * generate_reads_based_on_coverage   read_operations.py:86-123
  * count_matches / count_barcode_occurrences   counting.py:7-71
  * create_barcoded_insertion_genome / create_barcoded_roi_genome /
    build_mask_tree for the inputs (genome_generation.py:25-98, utils.py:89-101)

Each fixture stores the inputs (as plain arrays), the seed, and the reference's
outputs: every placed read (start, end, label) in draw order, covered_length,
the label histogram in first-appearance order, and per-target full / partial /
on_target_bases / overlap list.  Nothing from the reference's source is copied;
only its outputs are recorded.
"""
import hashlib
import json
import logging
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
sys.path.insert(0, os.path.join(HERE, "..", "..", "baseline", "shims"))
sys.path.insert(0, REF_SRC)

import esloco.combined_calculations as cc  # noqa: E402
import esloco.read_operations as ro  # noqa: E402
import esloco.counting as counting  # noqa: E402
from esloco.genome_generation import create_barcoded_insertion_genome, create_barcoded_roi_genome  # noqa: E402
from esloco.utils import build_mask_tree  # noqa: E402

logging.disable(logging.CRITICAL)

LABEL_BLOCKED_CONSUMING = -1
LABEL_BLOCKED_NOTCONSUMING = -2


def write_fasta(path, contigs):
    with open(path, "w") as fh:
        for name, n in contigs:
            fh.write(f">{name}\n")
            line = "N" * 100000
            full, rest = divmod(n, len(line))
            for _ in range(full):
                fh.write(line + "\n")
            if rest:
                fh.write("N" * rest + "\n")


def write_bed(path, header, rows):
    with open(path, "w") as fh:
        if header:
            fh.write("\t".join(header) + "\n")
        for r in rows:
            fh.write("\t".join(str(x) for x in r) + "\n")


def write_fastq(path, lengths):
    with open(path, "w") as fh:
        for i, n in enumerate(lengths):
            fh.write(f"@r{i}\n{'A' * n}\n+\n{'I' * n}\n")


def label_code(label):
    if label == "Blocked/Consuming":
        return LABEL_BLOCKED_CONSUMING
    if label == "Blocked/NotConsuming":
        return LABEL_BLOCKED_NOTCONSUMING
    return int(label)


def targets_to_arrays(target_regions):
    keys, s, e = [], [], []
    for k, v in target_regions.items():
        keys.append(k)
        if isinstance(v, list):
            s.append(int(v[0])); e.append(int(v[1]))
        else:
            s.append(int(v["start"])); e.append(int(v["end"]))
    return keys, np.array(s, np.int64), np.array(e, np.int64)


def mask_to_arrays(masked_tree):
    """Flatten {barcode|'ALL': IntervalTree} to arrays; barcode -1 = ALL."""
    s, e, w, b = [], [], [], []
    if masked_tree:
        for key, tree in masked_tree.items():
            for iv in tree:
                s.append(int(iv.begin)); e.append(int(iv.end)); w.append(float(iv.data))
                b.append(-1 if key == "ALL" else int(key))
    return (np.array(s, np.int64), np.array(e, np.int64), np.array(w, np.float64), np.array(b, np.int32))


def run_hot_path(case, genome_size, target_regions, masked_tree, tmp):
    """Call the reference's process_combination once per combination, capturing the read dict."""
    captured = {}
    orig_cm = counting.count_matches
    orig_gen = ro.generate_read_length_distribution

    def spy_count_matches(insertion_dict, read_dir, scaling, min_overlap):
        captured["reads"] = dict(read_dir)
        return orig_cm(insertion_dict, read_dir, scaling, min_overlap)

    def spy_gen(**kw):
        pool = orig_gen(**kw)
        captured["pool"] = np.array(pool)
        return pool

    cc.count_matches = spy_count_matches
    cc.generate_read_length_distribution = spy_gen
    out = []
    try:
        np.random.seed(case["seed"])
        for (mrl, cov) in case["combinations"]:
            captured.clear()
            detected, bdist = cc.process_combination(
                mrl, cov, genome_size, target_regions, case["iteration"],
                case.get("sequenced_data_path"), case["sigma"], case["min_read_length"],
                case["n_barcodes"], case["barcode_weights"], case["consuming"], masked_tree,
                tmp, case["scaling"], case["min_overlap"], True)
            reads = captured["reads"]
            rs = np.fromiter((v[0] for v in reads.values()), np.int64, len(reads))
            re_ = np.fromiter((int(v[1]) for v in reads.values()), np.int64, len(reads))
            lab = np.fromiter((label_code(k.split("_")[-1]) for k in reads), np.int16, len(reads))
            for n, k in enumerate(reads):
                assert k.split("_")[1] == str(n)
            meta_keys = ("coverage", "mean_read_length", "iteration", "N_bases")
            label_order = [k for k in bdist if k not in meta_keys]
            ov_flat, ov_off = [], [0]
            for d in detected:
                ov_flat.extend(int(x) for x in d["overlap"])
                ov_off.append(len(ov_flat))
            pool = captured.get("pool")
            out.append(dict(
                read_start=rs, read_end=re_, read_label=lab,
                covered_length=int(bdist["N_bases"]),
                label_order=label_order,
                label_counts=np.array([bdist[k] for k in label_order], np.int64),
                full=np.array([d["full_matches"] for d in detected], np.int64),
                partial=np.array([d["partial_matches"] for d in detected], np.int64),
                otb=np.array([int(d["on_target_bases"]) for d in detected], np.int64),
                target_region=[d["target_region"] for d in detected],
                overlap_flat=np.array(ov_flat, np.int64), overlap_off=np.array(ov_off, np.int64),
                pool_len=-1 if pool is None else int(pool.size),
                pool_head=np.zeros(0, np.int64) if pool is None else pool[:256].astype(np.int64),
                pool_sum=-1 if pool is None else int(pool.sum()),
                pool_sha=("" if pool is None else hashlib.sha256(pool.astype(np.int64).tobytes()).hexdigest()),
            ))
    finally:
        cc.count_matches = orig_cm
        cc.generate_read_length_distribution = orig_gen
    return out


def build_inputs(case, tmp):
    """Run the reference's own host setup (FASTA/BED -> genome size, targets, mask)."""
    fasta = os.path.join(tmp, "ref.fa")
    write_fasta(fasta, case["contigs"])
    blocked_path = None
    if case.get("blocked_rows"):
        blocked_path = os.path.join(tmp, "blocked.bed")
        write_bed(blocked_path, case.get("blocked_header"), case["blocked_rows"])
    np.random.seed(case["setup_seed"])
    if case["mode"] == "I":
        gsize, targets, masked, chromdir = create_barcoded_insertion_genome(
            1, fasta, None, blocked_path, case.get("chr_restriction"), case["insertion_length"],
            case["insertion_numbers"], None, case["n_barcodes"], os.path.join(tmp, "log.log"))
    else:
        roi_path = os.path.join(tmp, "roi.bed")
        write_bed(roi_path, ["chrom", "start", "end", "id"], case["roi_rows"])
        targets, _bed, masked, gsize, chromdir = create_barcoded_roi_genome(
            fasta, case.get("chr_restriction"), roi_path, case["n_barcodes"], blocked_path)
    tree = build_mask_tree(masked) if masked else masked
    if case.get("fastq_lengths") is not None:
        fq = os.path.join(tmp, "reads.fastq")
        write_fastq(fq, case["fastq_lengths"])
        case["sequenced_data_path"] = fq
    return gsize, targets, tree, chromdir


def direct_inputs(case):
    """Inputs given directly as dicts (adversarial geometry the BED path cannot express compactly)."""
    targets = {}
    for key, s, e in case["targets"]:
        targets[key] = [s, e] if case["mode"] == "I" else {"start": s, "end": e, "weight": 1, "barcode": []}
    tree = None
    if case.get("blocked"):
        masked = {f"m{i}": {"start": s, "end": e, "weight": w, "barcode": b}
                  for i, (s, e, w, b) in enumerate(case["blocked"])}
        tree = build_mask_tree(masked)
    return case["genome_size"], targets, tree, {}


def rng_cases():
    r = np.random.RandomState(20261019)
    cases = []

    # --- config-1 shape (BASELINE.json configs[0]): ROI, 3 contigs 10 Mb, 1 barcode, 10x, 20 ROIs ---
    contigs = [("chr1", 4_000_000), ("chr2", 3_000_000), ("chr3", 3_000_000)]
    rows = []
    for i in range(20):
        c = contigs[i % 3]
        ln = int(r.randint(1000, 50000))
        st = int(r.randint(1, c[1] - ln - 1))
        rows.append((c[0], st, st + ln, f"R{i}"))
    cases.append(dict(name="cfg1_roi_10mb", mode="ROI", contigs=contigs, roi_rows=rows, n_barcodes=1,
                      barcode_weights=None, combinations=[(15000, 10)], sigma=1, min_read_length=1,
                      consuming=False, scaling=1.0, min_overlap=1, seed=101, setup_seed=1, iteration=0))

    # --- I mode, 4 barcodes with weights, through the reference's insertion generator ---
    contigs2 = [("chr1", 900_000), ("chr2", 700_000), ("chrX", 400_000)]
    cases.append(dict(name="ins_b4_weights", mode="I", contigs=contigs2, insertion_length=1000,
                      insertion_numbers=8.0, n_barcodes=4, barcode_weights={"0": 10, "2": 3},
                      combinations=[(5000, 5)], sigma=1, min_read_length=1, consuming=False,
                      scaling=1.0, min_overlap=1, seed=202, setup_seed=2, iteration=3))

    # --- blocked regions: weight 1, all barcodes, not consuming / consuming ---
    hdr = ["chrom", "start", "end", "id", "barcode", "weight"]
    blocked = [("chr1", 100_000, 250_000, "cen1", "[]", 1), ("chr2", 0, 0, "allchr2", "[]", 1),
               ("chrX", 50_000, 60_000, "gapX", "[]", 1)]
    for consuming, nm in ((False, "blocked_w1_notconsuming"), (True, "blocked_w1_consuming")):
        cases.append(dict(name=nm, mode="I", contigs=contigs2, insertion_length=2000, insertion_numbers=6.0,
                          n_barcodes=2, barcode_weights=None, blocked_header=hdr, blocked_rows=blocked,
                          combinations=[(4000, 4)], sigma=1, min_read_length=1, consuming=consuming,
                          scaling=1.0, min_overlap=1, seed=303, setup_seed=3, iteration=1))

    # --- blocked: barcode-specific, fractional equal weights, overlapping intervals (multi-hit) ---
    blocked2 = [("chr1", 100_000, 400_000, "a", '["0"]', 0.5), ("chr1", 300_000, 500_000, "b", '["0","1"]', 0.5),
                ("chr1", 350_000, 450_000, "c", "[]", 0.5), ("chr2", 10_000, 600_000, "d", '["1"]', 0.5),
                ("chrX", 1, 399_999, "e", "[]", 0.5)]
    cases.append(dict(name="blocked_fractional_multihit", mode="ROI", contigs=contigs2,
                      roi_rows=[("chr1", 120_000, 180_000, "RA"), ("chr1", 340_000, 360_000, "RB"),
                                ("chr2", 300_000, 305_000, "RC"), ("chrX", 0, 0, "RX")],
                      n_barcodes=3, barcode_weights={"1": 2}, blocked_header=hdr, blocked_rows=blocked2,
                      combinations=[(3000, 6)], sigma=1, min_read_length=1, consuming=True,
                      scaling=1.0, min_overlap=1, seed=404, setup_seed=4, iteration=0))

    # --- min_overlap / min_read_length / sigma=2, two combinations sharing one stream ---
    cases.append(dict(name="minoverlap_sigma2_twocombos", mode="I", contigs=contigs2, insertion_length=3000,
                      insertion_numbers=10.0, n_barcodes=2, barcode_weights={"1": 4},
                      combinations=[(6000, 3), (2000, 2)], sigma=2, min_read_length=500, consuming=False,
                      scaling=1.0, min_overlap=1000, seed=505, setup_seed=5, iteration=7))

    # --- empirical FASTQ length source with min_read_length ---
    fq = np.clip(np.round(r.lognormal(np.log(3000), 0.9, 4000)), 50, 200000).astype(int).tolist()
    cases.append(dict(name="fastq_lengths", mode="ROI", contigs=contigs2,
                      roi_rows=[("chr1", 10_000, 40_000, "Q1"), ("chr2", 100_000, 101_000, "Q2"),
                                ("chrX", 200_000, 260_000, "Q3")],
                      n_barcodes=2, barcode_weights=None, fastq_lengths=fq,
                      combinations=[(int(np.mean([x for x in fq if x > 400])), 8)], sigma=1, min_read_length=400,
                      consuming=False, scaling=1.0, min_overlap=100, seed=606, setup_seed=6, iteration=2))

    # --- scaling < 1: one rand() per qualifying pair, target-major order (counting.py:46,51) ---
    cases.append(dict(name="scaling_half", mode="I", contigs=contigs2, insertion_length=5000,
                      insertion_numbers=5.0, n_barcodes=2, barcode_weights=None,
                      combinations=[(4000, 10)], sigma=1, min_read_length=1, consuming=False,
                      scaling=0.5, min_overlap=1, seed=707, setup_seed=7, iteration=0))

    # --- dense adversarial geometry: tiny genome, nested / touching / edge targets, 3 barcodes ---
    G = 60_000
    tg = [("E0_0", 0, 500), ("E1_0", G - 700, G), ("N0_0", 10_000, 20_000), ("N1_0", 12_000, 13_000),
          ("N2_0", 12_500, 12_600), ("T0_0", 20_000, 21_000), ("T1_0", 21_000, 22_000), ("W_0", 0, G),
          ("S_0", 30_000, 30_001)]
    for i in range(24):
        ln = int(r.randint(1, 4000)); st = int(r.randint(0, G - ln))
        tg.append((f"X{i}_{i % 3}", st, st + ln))
    for b in (1, 2):
        tg += [(f"N0_{b}", 10_000, 20_000), (f"W_{b}", 0, G)]
    cases.append(dict(name="dense_adversarial", mode="ROI", direct=True, genome_size=G, targets=tg,
                      blocked=[(5000, 5600, 1, []), (40_000, 41_000, "0.25", ["2"]), (40_500, 42_000, "0.25", ["2"])],
                      n_barcodes=3, barcode_weights={"0": 2}, combinations=[(800, 25)], sigma=1,
                      min_read_length=1, consuming=False, scaling=1.0, min_overlap=1, seed=808, iteration=0))
    cases.append(dict(name="dense_adversarial_minov", mode="ROI", direct=True, genome_size=G, targets=tg,
                      blocked=None, n_barcodes=3, barcode_weights=None, combinations=[(800, 25)], sigma=1,
                      min_read_length=100, consuming=False, scaling=1.0, min_overlap=250, seed=809, iteration=0))

    # --- genome larger than 2^32: 64-bit masked-rejection path of randint (Appendix B) ---
    G64 = 5_000_000_000
    tg64 = [(f"Barcode_0_insertion_{i}", int(s), int(s) + 200_000_000) for i, s in
            enumerate(r.randint(0, G64 - 200_000_000, size=6, dtype=np.int64))]
    cases.append(dict(name="genome_gt_u32", mode="I", direct=True, genome_size=G64, targets=tg64, blocked=None,
                      n_barcodes=1, barcode_weights=None, combinations=[(20000, 0.002)], sigma=1,
                      min_read_length=1, consuming=False, scaling=1.0, min_overlap=1, seed=909, iteration=0))

    # --- config-2 shape reduced (I mode, 1 barcode, 100 insertions of 5 kb), G = 5e7, 5x ---
    contigs5 = [(f"chr{i + 1}", 5_000_000) for i in range(10)]
    cases.append(dict(name="cfg2_reduced_50mb", mode="I", contigs=contigs5, insertion_length=5000,
                      insertion_numbers=100.0, n_barcodes=1, barcode_weights=None,
                      combinations=[(15000, 5)], sigma=1, min_read_length=1, consuming=False,
                      scaling=1.0, min_overlap=1, seed=1010, setup_seed=10, iteration=0))
    return cases


def main():
    outdir = HERE
    index = {}
    for case in rng_cases():
        with tempfile.TemporaryDirectory() as tmp:
            if case.get("direct"):
                gsize, targets, tree, chromdir = direct_inputs(case)
            else:
                gsize, targets, tree, chromdir = build_inputs(case, tmp)
            keys, ts, te = targets_to_arrays(targets)
            ms, me, mw, mb = mask_to_arrays(tree)
            combos = run_hot_path(case, gsize, targets, tree, tmp)
        arrays = dict(target_start=ts, target_end=te, mask_start=ms, mask_end=me, mask_weight=mw, mask_barcode=mb)
        if case.get("fastq_lengths") is not None:
            arrays["fastq_lengths"] = np.array(case["fastq_lengths"], np.int64)
        meta = dict(
            name=case["name"], mode=case["mode"], genome_size=int(gsize), target_keys=keys,
            chromosome_dir={k: [int(v[0]), int(v[1])] for k, v in chromdir.items()},
            n_barcodes=case["n_barcodes"], barcode_weights=case["barcode_weights"],
            combinations=[[m, c] for m, c in case["combinations"]], sigma=case["sigma"],
            min_read_length=case["min_read_length"], consuming=case["consuming"], scaling=case["scaling"],
            min_overlap=case["min_overlap"], seed=case["seed"], setup_seed=case.get("setup_seed"),
            iteration=case["iteration"], insertion_length=case.get("insertion_length"),
            insertion_numbers=case.get("insertion_numbers"), contigs=case.get("contigs"),
            roi_rows=case.get("roi_rows"), blocked_rows=case.get("blocked_rows"),
            blocked_header=case.get("blocked_header"), combos=[])
        for ci, c in enumerate(combos):
            meta["combos"].append(dict(covered_length=c["covered_length"], label_order=c["label_order"],
                                       target_region=c["target_region"], pool_len=c["pool_len"],
                                       pool_sum=c["pool_sum"], pool_sha=c["pool_sha"]))
            for k in ("read_start", "read_end", "read_label", "label_counts", "full", "partial", "otb",
                      "overlap_flat", "overlap_off", "pool_head"):
                arrays[f"c{ci}_{k}"] = c[k]
        arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(outdir, case["name"] + ".npz")
        np.savez_compressed(path, **arrays)
        index[case["name"]] = dict(reads=[int(c["read_start"].size) for c in combos],
                                   targets=len(keys), blocked=int(ms.size), bytes=os.path.getsize(path))
        print(case["name"], index[case["name"]], flush=True)
    with open(os.path.join(outdir, "INDEX.json"), "w") as fh:
        json.dump(dict(reference="aweich/esloco v1.0.6 (/root/reference)", numpy=np.__version__, cases=index), fh, indent=1)


if __name__ == "__main__":
    main()
