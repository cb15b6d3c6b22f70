"""Host-side result assembly and the multi-rank logic (gloo, world_size 2) -- no GPU needed."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_tallies(n_iter, T, seed, big=False):
    r = np.random.RandomState(seed)
    t = np.zeros((n_iter, T, 3), np.int64)
    t[:, :, 0] = r.poisson(3, (n_iter, T)); t[:, :, 1] = r.poisson(5, (n_iter, T))
    t[:, :, 2] = r.randint(0, 7_500_000_000 if big else 200_000, (n_iter, T))
    return t


def reference_style_tables(keys, mode, combos, its, tallies):
    """What esloco.py:155-195 computes, written with the same pandas calls (rows -> DataFrame -> rsplit -> groupby)."""
    rows = []
    for i, it in enumerate(its):
        for c, (mrl, cov) in enumerate(combos):
            for j, k in enumerate(keys):
                rows.append({"target_region": f"{k}_{it}", "full_matches": tallies[c][i, j, 0], "partial_matches": tallies[c][i, j, 1],
                             "overlap": [], "on_target_bases": tallies[c][i, j, 2], "mean_read_length": mrl, "coverage": cov})
    df = pd.DataFrame(rows)
    if mode == "ROI":
        df[["target", "barcode", "iteration"]] = df["target_region"].str.rsplit("_", n=2, expand=True)
    else:
        sp = df["target_region"].str.rsplit("_", n=4, expand=True)
        df["target"] = sp.iloc[:, :4].agg("_".join, axis=1)
        df["barcode"] = sp.iloc[:, 1]
        df["insertion"] = sp.iloc[:, 2] + "_" + sp.iloc[:, 3]
        df["iteration"] = sp.iloc[:, 4]
    summ = df.groupby(["target", "mean_read_length", "coverage"], as_index=False).agg(
        mean_full_matches=("full_matches", "mean"), sd_full_matches=("full_matches", "std"),
        mean_partial_matches=("partial_matches", "mean"), sd_partial_matches=("partial_matches", "std"),
        mean_bases_on_target=("on_target_bases", "mean"), sd_bases_on_target=("on_target_bases", "std"),
        sem_bases_on_target=("on_target_bases", "sem"))
    summ["ci_lower_bases_on_target"] = summ["mean_bases_on_target"] - 1.96 * summ["sem_bases_on_target"]
    summ["ci_upper_bases_on_target"] = summ["mean_bases_on_target"] + 1.96 * summ["sem_bases_on_target"]
    return df, summ


@pytest.mark.parametrize("mode", ["I", "ROI"])
def test_tables_match_reference_style_pandas(mode):
    from esloco_b200.distributed import moments_from_tallies
    from esloco_b200.tables import matches_table, summary_table
    if mode == "I":
        keys = [f"Barcode_{b}_insertion_{i}" for b in range(3) for i in range(11)]
    else:
        keys = [f"R{i}_{b}" for b in range(3) for i in range(11)]
    combos = [(15000, 10), (15000, 30), (5000, 10)]
    its = list(range(7))
    tallies = [fake_tallies(len(its), len(keys), s, big=(s == 2)) for s in range(len(combos))]
    ref_df, ref_sum = reference_style_tables(keys, mode, combos, its, tallies)
    df = matches_table(keys, mode, combos, its, tallies)
    assert list(df.columns) == list(ref_df.columns)
    for col in df.columns:
        assert df[col].astype(str).tolist() == ref_df[col].astype(str).tolist(), col
    summ = summary_table(keys, mode, combos, [moments_from_tallies(t) for t in tallies])
    assert list(summ.columns) == list(ref_sum.columns)
    assert summ["target"].tolist() == ref_sum["target"].tolist()
    for col in summ.columns[1:]:
        assert np.allclose(summ[col].astype(float), ref_sum[col].astype(float), rtol=1e-9, atol=0), col


def test_barcode_distribution_table():
    from esloco_b200.tables import barcode_distribution_table
    lc = np.array([[5, 0, 7, 0, 2], [6, 1, 8, 0, 0]])
    df = barcode_distribution_table(3, [(1000, 2)], [0, 1], [lc], [np.array([111, 222])])
    assert list(df.columns) == ["0", "1", "2", "Blocked/NotConsuming", "coverage", "mean_read_length", "iteration", "N_bases", "total_Reads"]
    assert df["total_Reads"].tolist() == [12, 15] and df["N_bases"].tolist() == [111, 222]


def test_barcode_distribution_table_reference_layout():
    """label_columns="reference": what pandas makes of the reference's per-row dicts (esloco.py:161-180) -- label keys in
    first-appearance order, labels without reads absent, a label first seen in a later row behind the bookkeeping
    columns, total_Reads = sum of the first n_barcodes COLUMNS (Appendix C.5).  Expected frame written out by hand."""
    from esloco_b200.tables import barcode_distribution_table
    lc = np.array([[5, 7, 0, 0, 2], [6, 1, 8, 0, 0]])  # B = 3: barcode 2 has no read in iteration 0, nothing blocked in iteration 1
    orders = {0: ["1", "Blocked/NotConsuming", "0"], 1: ["2", "0", "1"]}
    asked = []

    def first_appearance(row, it, c):
        asked.append((row, it, c))
        return orders[it]

    df = barcode_distribution_table(3, [(1000, 2)], [0, 1], [lc], [np.array([111, 222])], label_columns="reference",
                                    first_appearance=first_appearance)
    exp = pd.DataFrame([
        {"1": 7, "Blocked/NotConsuming": 2, "0": 5, "coverage": 2, "mean_read_length": 1000, "iteration": 0, "N_bases": 111},
        {"2": 8, "0": 6, "1": 1, "coverage": 2, "mean_read_length": 1000, "iteration": 1, "N_bases": 222}])
    exp["total_Reads"] = exp.iloc[:, :3].sum(axis=1)
    assert list(df.columns) == ["1", "Blocked/NotConsuming", "0", "coverage", "mean_read_length", "iteration", "N_bases", "2", "total_Reads"]
    pd.testing.assert_frame_equal(df, exp)
    assert df["total_Reads"].tolist() == [14.0, 7.0]  # the quirk: 1 + Blocked + 0, barcode 2 left out
    assert asked == [(0, 0, 0), (1, 1, 0)]
    # a row that brings nothing new is not asked for its order
    asked.clear()
    lc2 = np.array([[5, 7, 1, 0, 0], [6, 1, 8, 0, 0]])
    barcode_distribution_table(3, [(1000, 2)], [0, 1], [lc2], [np.array([1, 2])], label_columns="reference",
                               first_appearance=lambda r, it, c: asked.append(r) or ["2", "0", "1"])
    assert asked == [0]


def test_shard_iterations_partition():
    from esloco_b200.distributed import shard_iterations
    for n in (0, 1, 7, 8, 1000):
        for ws in (1, 2, 3, 8):
            parts = [list(shard_iterations(n, r, ws)) for r in range(ws)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, ws, port, n_iter, T, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    import torch
    import torch.distributed as dist
    from esloco_b200.distributed import gather_iterations, moments_from_tallies, reduce_moments, shard_iterations
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    full = fake_tallies(n_iter, T, 99, big=True)
    mine = shard_iterations(n_iter, rank, ws)
    local = full[mine.start:mine.stop]
    acc = torch.from_numpy(moments_from_tallies(local))
    reduce_moments(acc, dst=0)
    gathered = gather_iterations(torch.from_numpy(local.copy()), n_iter, dst=0)
    if rank == 0:
        q.put((acc.numpy(), gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_iter", [9, 1])
def test_world_size_2_gloo_reduce_and_gather(n_iter):
    """Two ranks shard the iterations, reduce moment accumulators, gather tallies: rank 0 must end up with exactly
    the single-process result (an odd count and a rank with zero iterations included)."""
    import torch.multiprocessing as mp
    from esloco_b200.distributed import moments_from_tallies
    T = 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_iter, T, q)) for r in range(2)]
    for p in procs: p.start()
    acc, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = fake_tallies(n_iter, T, 99, big=True)
    assert np.array_equal(gathered, full)
    exp = moments_from_tallies(full).astype(object)
    got = acc.astype(object)
    assert np.array_equal(got[:, :6], exp[:, :6])
    assert [int(h) * 2**32 + int(l) for h, l in zip(got[:, 7], got[:, 6])] == [int(h) * 2**32 + int(l) for h, l in zip(exp[:, 7], exp[:, 6])]


def test_gaussian_filter_matches_scipy():
    from scipy.ndimage import gaussian_filter1d as ref
    from esloco_b200.coverage import gaussian_filter1d
    for n in (1, 4, 13, 500, 3100):
        x = np.random.RandomState(n).rand(n) * 30
        assert np.allclose(gaussian_filter1d(x, 3), ref(x, 3))


@pytest.mark.parametrize("mode", ["I", "ROI"])
def test_native_row_writer_equals_pandas_bytes(tmp_path, mode):
    """esl_write_match_rows (host helper of the library) writes the matches table byte for byte as pandas.to_csv does:
    two combinations (an int and a float coverage), chunked appends, a part file without header, keys without a
    barcode token, 13-digit tallies; keys that would need CSV quoting fall back to pandas."""
    from esloco_b200.tables import MatchesWriter
    rng = np.random.RandomState(3)
    if mode == "I":
        keys = [f"Barcode_{i % 5}_insertion_{i // 5}" for i in range(400)]
    else:
        keys = [f"R{i // 3}_{i % 3}" for i in range(300)] + ["weird", "a_b_c", "x_"]
    T, its, combos = len(keys), [0, 1, 2, 10, 11], [(15000, 30), (8000, 0.5)]
    tal = [rng.randint(0, 10**13, (len(its), T, 3)).astype(np.int64) for _ in combos]
    tal[0][0, 0] = 0
    out = {}
    for native in (True, False):
        for header in (True, False):
            p = str(tmp_path / f"m_{native}_{header}.csv")
            w = MatchesWriter(p, keys, mode, combos, write_header=header, native=native)
            assert (w._blobs is not None) == native
            w.append(its[:2], [x[:2] for x in tal]); w.append([], [x[:0] for x in tal]); w.append(its[2:], [x[2:] for x in tal])
            out[native, header] = open(p, "rb").read()
    assert out[True, True] == out[False, True] and out[True, False] == out[False, False]
    assert out[True, True].count(b"\n") == 1 + len(its) * len(combos) * T
    bad = keys[:3] + ['has"quote_1', "has\ttab_2"]
    w = MatchesWriter(str(tmp_path / "q.csv"), bad, mode, combos)
    assert w._blobs is None  # pandas path: it knows how to quote
    w.append([0], [x[:1, :len(bad)] for x in tal])
    assert b'"' in open(w.path, "rb").read()


def test_native_row_writer_errors(tmp_path):
    """No exception crosses the ABI: a path that cannot be opened or a NULL argument gives a negative status, and the
    Python writer turns it into OSError."""
    from esloco_b200 import _native
    from esloco_b200.tables import MatchesWriter
    lib = _native.load()
    off = np.zeros(2, np.int64); off[1] = 2
    t = np.zeros((1, 3), np.int64)
    bad = str(tmp_path / "no_such_dir" / "m.csv").encode()
    assert lib.esl_write_match_rows(bad, 1, b"k_", off.ctypes.data, b"\t\t", off.ctypes.data, b"[]", t.ctypes.data, 1, b"0") < 0
    assert lib.esl_write_match_rows(None, 1, b"k_", off.ctypes.data, b"\t\t", off.ctypes.data, b"[]", t.ctypes.data, 1, b"0") < 0
    ok = str(tmp_path / "m.csv").encode()
    assert lib.esl_write_match_rows(ok, 0, b"k_", off.ctypes.data, b"\t\t", off.ctypes.data, b"[]", t.ctypes.data, 1, b"7") == len(b"k_7\t0\t0\t[]\t0\t\t7\n")
    assert open(ok, "rb").read() == b"k_7\t0\t0\t[]\t0\t\t7\n"
    w = MatchesWriter(str(tmp_path / "w.csv"), ["a_0"], "ROI", [(1, 1)])
    w.path = str(tmp_path / "gone" / "w.csv")
    with pytest.raises(OSError):
        w.append([0], [np.zeros((1, 1, 3), np.int64)])


def _seed_worker(rank, ws, port, ini, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    import numpy as np
    import torch.distributed as dist
    from esloco_b200.config_handler import parse_config
    from esloco_b200.create_insertion_genome import add_insertions_to_genome_sequence_with_bed
    from esloco_b200.distributed import sync_seed
    p = parse_config(ini)  # no `seed` in the INI: every process draws its own (config_handler.py:53-57,102)
    own = p["seed"]
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    common = sync_seed(p)
    np.random.seed(p["seed"])  # what esloco.run does next (esloco.py:100): host set-up draws from this stream
    sites = add_insertions_to_genome_sequence_with_bed(10_000_000, 500, 25, {"chr1": [0, 10_000_000]}, None, None)
    q.put((rank, own, common, sorted((k, tuple(v)) for k, v in sites.items())))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_seedless_ini_gets_one_seed(tmp_path):
    """An INI without `seed` (the reference's config.ini default) under a 2-rank launch: both ranks must end up with
    rank 0's seed and therefore build identical insertion sites -- otherwise rank 0 would sum moments of other ranks'
    target sets under its own names."""
    import torch.multiprocessing as mp
    fa = tmp_path / "g.fa"
    fa.write_text(">chr1\n" + "N" * 1000 + "\n")
    ini = tmp_path / "c.ini"
    ini.write_text(f"[COMMON]\nmode = I\nreference_genome_path = {fa}\noutput_path = {tmp_path}/out/\n[I]\ninsertion_numbers = 5\n")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, str(ini), q)) for r in range(2)]
    for p in procs: p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, own0, common0, sites0), (_, own1, common1, sites1) = got
    assert own0 != own1  # 2 independent draws from [0, 2^32) -- the situation the broadcast is for
    assert common0 == common1 == own0
    assert sites0 == sites1 and len(sites0) == 25


def test_target_columns_equal_dict_conversion():
    from esloco_b200.utils import TargetColumns, targets_to_arrays
    d = {f"Barcode_{i % 3}_insertion_{i}": [100 * i, 100 * i + 50] for i in range(30)}
    d["odd_key"] = [5, 9]
    keys, s, e, b = targets_to_arrays(d, 3)
    tc = TargetColumns.from_dict(d, 3)
    k2, s2, e2, b2 = targets_to_arrays(tc, 3)
    assert k2 == keys and np.array_equal(s, s2) and np.array_equal(e, e2) and np.array_equal(b, b2)
    assert b[-1] == -1 and len(tc) == 31
    with pytest.raises(ValueError):
        TargetColumns([1, 2], [3], [0, 0])
