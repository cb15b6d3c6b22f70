"""End-to-end drop-in check on the GPU: `esloco --config` writes the reference's tables (format pinned by
tests/golden/cli_format.json, recorded from the unmodified reference CLI) and run_simulation_iteration keeps the
reference's return shape."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from test_host_setup import write_bed, write_fasta

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CONTIGS = [("chr1", 300_000), ("chr2", 200_000), ("chrX", 100_000), ("chr1_alt", 50_000), ("chrM", 16_000)]


def make_config(tmp, mode):
    fmt = json.load(open(os.path.join(HERE, "golden", "cli_format.json")))[mode]
    fasta = os.path.join(tmp, "ref.fa"); write_fasta(fasta, CONTIGS)
    write_bed(os.path.join(tmp, "blocked.bed"), ["chrom", "start", "end", "id", "barcode", "weight"], [("chr2", 1000, 90_000, "b1", "[]", 1)])
    write_bed(os.path.join(tmp, "roi.bed"), ["chrom", "start", "end", "id"],
              [("chr1", 1000, 5000, "RA"), ("chrX", 0, 0, "RX"), ("chr2", 100_000, 120_000, "RB")])
    os.makedirs(os.path.join(tmp, "out_" + mode), exist_ok=True)
    ini = os.path.join(tmp, mode + ".ini")
    open(ini, "w").write(fmt["ini"].replace("{tmp}", tmp))
    return ini, fmt


@pytest.mark.parametrize("mode", ["ROI", "I"])
def test_cli_tables_have_reference_format(tmp_path, mode):
    from esloco_b200.config_handler import parse_config
    from esloco_b200.esloco import run
    tmp = str(tmp_path)
    ini, fmt = make_config(tmp, mode)
    p = parse_config(ini)
    run(p)
    out = os.path.join(tmp, "out_" + mode)
    m = pd.read_csv(os.path.join(out, f"exp{mode}_matches_table.csv"), sep="\t", dtype=str)
    ref = fmt["matches_table.csv"]
    assert list(m.columns) == ref["columns"] and len(m) == ref["rows"]
    cols = ("target_region", "target", "barcode", "iteration", "mean_read_length", "coverage") + (("insertion",) if mode == "I" else ())
    for c in cols:
        assert m[c].tolist() == ref[c], c
    s = pd.read_csv(os.path.join(out, f"exp{mode}_summary.csv"), sep="\t", dtype=str)
    refs = fmt["summary.csv"]
    assert list(s.columns) == refs["columns"] and len(s) == refs["rows"]
    for c in ("target", "mean_read_length", "coverage"):
        assert s[c].tolist() == refs[c], c
    d = pd.read_csv(os.path.join(out, f"exp{mode}_barcode_distribution_table.csv"), sep="\t")
    refd = fmt["barcode_distribution_table.csv"]
    assert set(d.columns) == set(refd["columns"]) and len(d) == refd["rows"]  # label columns: barcode order here
    assert list(d.columns)[-5:] == refd["columns"][-5:]
    for c in ("coverage", "mean_read_length", "iteration"):
        assert d[c].astype(str).tolist() == refd[c], c
    G = 600_000
    assert (d["N_bases"] >= d["coverage"] * G).all()
    assert (d["total_Reads"] == d["0"] + d["1"]).all()
    assert (d["Blocked/NotConsuming"] > 0).all()
    share0 = d["0"].sum() / d["total_Reads"].sum()  # barcode_weights {"0": 3} -> 3/4 of the unblocked reads
    assert abs(share0 - 0.75) < 0.05
    # the summary agrees with a groupby over the matches table (what the reference computes)
    mm = pd.read_csv(os.path.join(out, f"exp{mode}_matches_table.csv"), sep="\t")
    g = mm.groupby(["target", "mean_read_length", "coverage"], as_index=False).agg(
        mean_bases_on_target=("on_target_bases", "mean"), sd_bases_on_target=("on_target_bases", "std"),
        mean_full_matches=("full_matches", "mean"), sd_partial_matches=("partial_matches", "std"))
    ss = pd.read_csv(os.path.join(out, f"exp{mode}_summary.csv"), sep="\t")
    for c in ("mean_bases_on_target", "sd_bases_on_target", "mean_full_matches", "sd_partial_matches"):
        assert np.allclose(ss[c], g[c], rtol=1e-9, equal_nan=True), c
    if mode == "I":
        b = pd.read_csv(os.path.join(out, "expI_insertion_locations.bed"), sep="\t")
        assert list(b.columns) == fmt["insertion_locations.bed"]["columns"]
    assert os.path.exists(os.path.join(out, f"exp{mode}_log.log"))


def test_run_simulation_iteration_shape_and_determinism(tmp_path):
    from esloco_b200.combined_calculations import run_simulation_batch, run_simulation_iteration
    from esloco_b200.synthetic import random_insertions
    from esloco_b200.utils import build_mask_tree
    G = 2_000_000
    targets = random_insertions(G, 6, 2000, 2, seed=1)
    tree = build_mask_tree({"m": {"start": 100_000, "end": 300_000, "weight": "0.5", "barcode": '["1"]'}})
    p = dict(n_barcodes=2, barcode_weights={"1": 2}, seed=42, scaling=1.0, combinations=[(4000, 3), (8000, 1.5)],
             sequenced_data_path=None, sigma=1, min_read_length=100, consuming=True, min_overlap_for_detection=50)
    results, dists = run_simulation_iteration(5, p, G, targets, tree, None)
    assert len(results) == 2 and len(dists) == 2
    keys = list(targets)
    for c, (mrl, cov) in enumerate(p["combinations"]):
        assert [r["target_region"] for r in results[c]] == [k + "_5" for k in keys]
        assert set(results[c][0]) == {"target_region", "full_matches", "partial_matches", "on_target_bases", "mean_read_length", "coverage"}
        assert dists[c]["coverage"] == cov and dists[c]["mean_read_length"] == mrl and dists[c]["iteration"] == 5
        assert dists[c]["N_bases"] >= cov * G
        assert "Blocked/Consuming" in dists[c] and "0" in dists[c] and "1" in dists[c]
    again, _ = run_simulation_iteration(5, p, G, targets, tree, None)
    assert again == results  # seed + iteration id fix the result, whatever the call pattern
    with_lists, _ = run_simulation_iteration(5, dict(p, overlap_lists=True), G, targets, tree, None)
    for a, b in zip(with_lists[0], results[0]):
        assert sum(a["overlap"]) == b["on_target_bases"] and len(a["overlap"]) == b["full_matches"] + b["partial_matches"]
        assert list(a) == ["target_region", "full_matches", "partial_matches", "overlap", "on_target_bases", "mean_read_length", "coverage"]
    batch, _ = run_simulation_batch(range(4, 7), p, G, targets, tree)
    assert [int(x) for x in batch[0].tallies[1, :, 2]] == [r["on_target_bases"] for r in results[0]]


def test_count_matches_mirror(tmp_path):
    from esloco_b200.counting import count_barcode_occurrences, count_matches
    from golden_util import Golden
    from oracle import oracle as O
    g = Golden("dense_adversarial")
    m = g.meta
    targets = {k: {"start": int(s), "end": int(e)} for k, s, e in zip(m["target_keys"], g["target_start"], g["target_end"])}
    names = {-1: "Blocked/Consuming", -2: "Blocked/NotConsuming"}
    reads = {f"Read_{n}_{names.get(int(l), 'Barcode_%d' % l)}": (int(s), int(e)) for n, (s, e, l) in
             enumerate(zip(g.combo(0, "read_start"), g.combo(0, "read_end"), g.combo(0, "read_label")))}
    out = count_matches(targets, reads, 1.0, m["min_overlap"], genome_size=m["genome_size"], n_barcodes=m["n_barcodes"])
    assert [d["target_region"] for d in out] == m["target_keys"]
    assert [d["full_matches"] for d in out] == g.combo(0, "full").tolist()
    assert [d["partial_matches"] for d in out] == g.combo(0, "partial").tolist()
    assert [d["on_target_bases"] for d in out] == g.combo(0, "otb").tolist()
    off, flat = g.combo(0, "overlap_off"), g.combo(0, "overlap_flat")
    assert [d["overlap"] for d in out] == [flat[a:b].tolist() for a, b in zip(off[:-1], off[1:])]
    assert list(out[0]) == ["target_region", "full_matches", "partial_matches", "overlap", "on_target_bases"]  # counting.py:34,54-57
    hist = count_barcode_occurrences(reads)
    assert list(hist) == m["combos"][0]["label_order"] and list(hist.values()) == g.combo(0, "label_counts").tolist()


def test_binned_coverage_track_matches_per_base_arrays(tmp_path):
    """Binned coverage (device) == the reference's per-base array + bin means (coverage.py:52-64,76-81), small genome."""
    from esloco_b200.coverage import binned_coverage
    from esloco_b200.session import get_session
    from esloco_b200.utils import build_mask_tree
    G, bin_size = 203_000, 10_000  # last bin is truncated
    sess = get_session(0)
    tree = build_mask_tree({"m": {"start": 50_000, "end": 80_000, "weight": 1, "barcode": []}})
    sess.invalidate()
    sess.configure(G, {"T_0": {"start": 10, "end": 20}}, tree, 2, {"0": 3})
    rng = np.random.RandomState(3)
    pool = rng.randint(200, 30_000, 5000)
    sess.set_pool(("covtest",), lambda: pool)
    eng = sess.engine
    res = eng.run_batch(11, 0, 0, 1, 8.0, consuming=True).cpu()
    n = int(res.n_reads[0])
    pos, raw, smoothed = binned_coverage(eng, 11, 0, 0, True, n, bin_size)
    s, ln, lab = (x.cpu().numpy() for x in eng.materialize_reads(11, 0, 0, True, n))
    names = {0: "0", 1: "1", -1: "Blocked/Consuming"}
    n_bins = -(-G // bin_size)
    for key in ["Combined"] + list(names.values()):
        cov = np.zeros(G)
        for a, l_, b in zip(s, ln, lab):
            if key == "Combined" or names[int(b)] == key:
                cov[a:a + l_] += 1
        ref = np.array([cov[i * bin_size:(i + 1) * bin_size].mean() for i in range(n_bins)])
        assert np.allclose(raw[key], ref, rtol=1e-12, atol=1e-12), key
    assert set(raw) == {"Combined", "0", "1", "Blocked/Consuming"}
    from scipy.ndimage import gaussian_filter1d
    assert np.allclose(smoothed["Combined"], gaussian_filter1d(raw["Combined"], 3))
    assert abs(raw["Combined"].mean() - 8.0) < 1.0


def test_generate_reads_and_process_combination_mirrors():
    """The reference-signature wrappers: reads dict from generate_reads_based_on_coverage fed to count_matches gives
    the tallies process_combination reports for the same seed / iteration / combination."""
    from esloco_b200.combined_calculations import process_combination
    from esloco_b200.counting import count_barcode_occurrences, count_matches
    from esloco_b200.read_operations import generate_reads_based_on_coverage
    from esloco_b200.synthetic import random_insertions
    from esloco_b200.utils import build_mask_tree
    G = 3_000_000
    targets = random_insertions(G, 5, 3000, 2, seed=9)
    tree = build_mask_tree({"m": {"start": 500_000, "end": 900_000, "weight": 1, "barcode": []}})
    pool = np.random.RandomState(0).randint(500, 20_000, 20_000)
    reads, covered = generate_reads_based_on_coverage(G, 4, pool, 2, {"0": 3}, False, tree, seed=5, iteration=2, combo=1)
    assert covered >= 4 * G and list(reads)[0].startswith("Read_0_")
    hist = count_barcode_occurrences(reads)
    assert set(hist) == {"0", "1", "Blocked/NotConsuming"} and sum(hist.values()) == len(reads)
    assert sum(e - s for k, (s, e) in reads.items() if not k.endswith("NotConsuming")) == covered
    by_reads = count_matches(targets, reads, 1.0, 10, genome_size=G, n_barcodes=2)
    # same combination through process_combination (needs the same pool: pass it as a FASTA-like file? use sigma path)
    from esloco_b200.session import get_session
    sess = get_session(0)
    sess.invalidate()
    sess.configure(G, targets, tree, 2, {"0": 3}, length_sampler="pool")  # the sampler generate_reads_based_on_coverage uses
    sess.set_pool(("explicit-test",), lambda: pool)
    res = sess.engine.run_batch(5, 1, 2, 1, 4, False, 10, 1.0).cpu()
    assert [int(x) for x in res.tallies[0, :, 2]] == [d["on_target_bases"] for d in by_reads]
    assert [int(x) for x in res.tallies[0, :, 0]] == [d["full_matches"] for d in by_reads]
    detected, dist = process_combination(5000, 2, G, targets, 3, None, 1, 100, 2, {"0": 3}, False, tree, "/tmp", 1.0, 10, True,
                                         seed=5, combo=0, overlap_lists=True)
    assert [d["target_region"] for d in detected] == [k + "_3" for k in targets]
    assert dist["N_bases"] >= 2 * G and dist["iteration"] == 3 and dist["coverage"] == 2
    assert all(sum(d["overlap"]) == d["on_target_bases"] for d in detected)


def test_cli_two_gpus_identical_tables(tmp_path):
    """`torchrun --nproc-per-node 2 -m esloco_b200 --config x.ini` (iterations sharded over two GPUs, NCCL reduce of
    the moments, gather of the tallies) writes byte-identical tables to the single-GPU run: the Philox counter
    carries the iteration id, so the split is invisible.  Skipped on a single-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(HERE)
    outs = {}
    for n in (1, 2):
        tmp = str(tmp_path / f"n{n}")
        os.makedirs(tmp)
        ini, _ = make_config(tmp, "I")
        text = open(ini).read().replace("iterations = 3", "iterations = 7")
        open(ini, "w").write(text)
        env = dict(os.environ, PYTHONPATH=root)
        if n == 1:
            cmd = [sys.executable, "-m", "esloco_b200", "--config", ini]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                   "127.0.0.1", "--master-port", "29533", "-m", "esloco_b200", "--config", ini]
        r = subprocess.run(cmd, env=env, cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[n] = {f: open(os.path.join(tmp, "out_I", f"expI_{f}")).read()
                   for f in ("matches_table.csv", "summary.csv", "barcode_distribution_table.csv", "insertion_locations.bed")}
    for f in outs[1]:
        assert outs[1][f] == outs[2][f], f
    assert outs[1]["matches_table.csv"].count("\n") == 1 + 7 * 2 * 8  # header + iterations x combinations x targets


def test_cli_streaming_chunks_are_invisible(tmp_path, monkeypatch):
    """The matches table is streamed to disk per chunk of iterations; chunk boundaries must not show."""
    import esloco_b200.esloco as cli
    from esloco_b200.config_handler import parse_config
    outs = []
    for k, step in enumerate((None, 2)):
        tmp = str(tmp_path / f"run{k}")
        os.makedirs(tmp)
        ini, _ = make_config(tmp, "ROI")
        text = open(ini).read().replace("iterations = 3", "iterations = 5").replace("[ROI]", "overlap_lists = True\n\n[ROI]")
        open(ini, "w").write(text)
        if step:
            monkeypatch.setattr(cli, "iteration_chunks", lambda its, T, C, **kw: (its[i:i + step] for i in range(0, len(its), step)))
        cli.run(parse_config(ini))
        outs.append(open(os.path.join(tmp, "out_ROI", "expROI_matches_table.csv")).read())
    assert outs[0] == outs[1]
    m = pd.read_csv(os.path.join(str(tmp_path / "run1"), "out_ROI", "expROI_matches_table.csv"), sep="\t")
    assert len(m) == 5 * 2 * 6
    import ast
    assert all(sum(ast.literal_eval(o)) == b for o, b in zip(m["overlap"], m["on_target_bases"]))


def test_cli_reference_label_columns(tmp_path):
    """`label_columns = reference`: the barcode-distribution table gets the reference's layout (esloco.py:161-180,
    counting.py:63-71) -- label columns in the order the labels first appear among the reads of the first row's
    iteration, total_Reads = the sum of the first n_barcodes columns.  The expected order comes from an independent
    route: the reads of that iteration as a reference-shaped dict, counted by the reference-style
    count_barcode_occurrences."""
    from esloco_b200.config_handler import parse_config
    from esloco_b200.counting import count_barcode_occurrences
    from esloco_b200.esloco import run
    from esloco_b200.read_operations import generate_reads_based_on_coverage
    from esloco_b200.session import lognormal_pool
    from esloco_b200.utils import build_mask_tree
    tmp = str(tmp_path)
    ini, fmt = make_config(tmp, "ROI")
    text = open(ini).read().replace("[ROI]", "label_columns = reference\nlength_sampler = pool\n\n[ROI]")
    open(ini, "w").write(text)
    p = parse_config(ini)
    dist_df, _summary = run(p)
    d = pd.read_csv(os.path.join(tmp, "out_ROI", "expROI_barcode_distribution_table.csv"), sep="\t")
    assert list(d.columns) == list(dist_df.columns)
    labels = list(d.columns[:list(d.columns).index("coverage")])
    assert set(labels) == {"0", "1", "Blocked/NotConsuming"} == set(fmt["barcode_distribution_table.csv"]["columns"][:3])
    assert list(d.columns)[3:] == fmt["barcode_distribution_table.csv"]["columns"][3:]
    # first row = iteration 0 of combination 0 (mean read length 3000, coverage 2): its reads, counted the reference's way
    G = 600_000
    tree = build_mask_tree({"b1": {"start": 300_000 + 1000, "end": 300_000 + 90_000, "weight": 1, "barcode": []}})
    pool = lognormal_pool(p["seed"], 0, 3000, p["sigma"], p["min_read_length"])
    reads, covered = generate_reads_based_on_coverage(G, 2, pool, 2, p["barcode_weights"], False, tree, seed=p["seed"], iteration=0, combo=0)
    hist = count_barcode_occurrences(reads)
    assert labels == list(hist)  # first-appearance order
    row0 = d.iloc[0]
    assert [int(row0[k]) for k in labels] == [hist[k] for k in labels] and int(row0["N_bases"]) == covered
    assert (d["total_Reads"] == d[labels[0]] + d[labels[1]]).all()  # the reference's rule: the first n_barcodes COLUMNS


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_session_cache_random_sequences(seed):
    """The session uploads only what changed since the last call (content-keyed, session.py).  A random sequence of calls
    in which any subset of the inputs changes between one call and the next -- genome size, target dict (also edited in
    place), blocked regions, barcode count / weights, length pool, sampler, nothing at all -- must give, call by call,
    exactly what a session that forgot everything gives."""
    import numpy as np
    from esloco_b200.combined_calculations import run_simulation_batch
    from esloco_b200.session import get_session
    from esloco_b200.utils import build_mask_tree
    rng = np.random.RandomState(40 + seed)
    sess = get_session(0)
    sess.invalidate()

    def new_targets(G, B, mode):
        n = int(rng.choice([1, 8, 60]))
        s = np.sort(rng.randint(0, G - 20_000, n))
        if mode == "I":
            return {f"Barcode_{int(rng.randint(0, B))}_insertion_{i}": [int(a), int(a) + 5000] for i, a in enumerate(s)}
        return {f"R{i}_{int(rng.randint(0, B))}": {"start": int(a), "end": int(a) + int(rng.randint(500, 20_000))} for i, a in enumerate(s)}

    def new_mask(G, B):
        if rng.rand() < 0.3:
            return None
        return {f"m{i}": {"start": int(a), "end": int(a) + int(rng.randint(1000, G // 10)), "weight": float(rng.choice([1, 0.5])),
                          "barcode": [] if rng.rand() < 0.5 else [str(int(rng.randint(0, B)))]}
                for i, a in enumerate(rng.randint(0, G - G // 10, int(rng.choice([1, 5]))))}

    G, B, weights, sampler = 5_000_000, 3, None, "quantile"
    targets, mask = new_targets(G, B, "I"), new_mask(G, B)
    pool = np.random.RandomState(1).randint(200, 6000, 50_000)
    for step in range(14):
        for what in rng.choice(["genome", "targets", "edit", "mask", "barcodes", "weights", "pool", "sampler", "none"],
                               size=int(rng.choice([1, 1, 2, 3])), replace=False):
            if what == "genome":
                G = int(rng.choice([3_000_000, 5_000_000, 40_000_000]))
                targets, mask = new_targets(G, B, "I"), new_mask(G, B)
            elif what == "targets":
                targets = new_targets(G, B, str(rng.choice(["I", "ROI"])))
            elif what == "edit":  # the caller's own dict, changed in place
                k = next(iter(targets))
                v = targets[k]
                if isinstance(v, list):
                    v[0] += 777; v[1] += 777
                else:
                    v["end"] += 333
            elif what == "mask":
                mask = new_mask(G, B)
            elif what == "barcodes":
                B = int(rng.choice([1, 3, 6]))
                targets, mask = new_targets(G, B, "I"), new_mask(G, B)
            elif what == "weights":
                weights = None if rng.rand() < 0.4 else {str(int(rng.randint(0, B))): int(rng.randint(2, 9))}
            elif what == "pool":
                pool = np.random.RandomState(int(rng.randint(1 << 30))).randint(100, int(rng.choice([3000, 30_000])), 50_000)
            elif what == "sampler":
                sampler = "pool" if sampler == "quantile" else "quantile"
        params = dict(n_barcodes=B, barcode_weights=weights, seed=11, scaling=1.0, combinations=[(3000, 4.0)], sequenced_data_path=None,
                      sigma=1, min_read_length=1, consuming=bool(step & 1), min_overlap_for_detection=1, length_sampler=sampler)
        tree = build_mask_tree(mask) if mask else None
        got = []
        for fresh in (False, True):
            if fresh:
                sess.invalidate()
            res, keys = run_simulation_batch(range(step, step + 2), params, G, targets, tree, length_pools={0: pool})
            got.append((res[0].tallies.numpy().copy(), res[0].label_counts.numpy().copy(), res[0].n_bases.numpy().copy(), list(keys)))
        for a, b in zip(got[0][:3], got[1][:3]):
            assert np.array_equal(a, b), (seed, step)
        assert got[0][3] == got[1][3] == list(targets)
        assert got[0][2].min() >= 4.0 * G
