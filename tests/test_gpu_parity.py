"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Replay check (BASELINE.json north_star): fed the reference's own numpy draws, per-target integer tallies,
label counts, N_bases and the number of reads placed must be BIT-EXACT; derived coverage
(on_target_bases / target length) within 1e-12 relative (identical integers => identical doubles).
"""
import numpy as np
import pytest

from golden_util import Golden, case_names
from oracle import native_np as NN
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = case_names()
REPLAYABLE = [c for c in CASES if c != "scaling_half"]  # scaling < 1 consumes the reference's stream per pair


@pytest.fixture(scope="module")
def eng():
    from esloco_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def configure(eng, g):
    from esloco_b200.read_operations import barcode_cdf, barcode_probabilities
    from esloco_b200.utils import targets_to_arrays
    m = g.meta
    eng.set_genome(m["genome_size"])
    eng.set_barcode_cdf(barcode_cdf(barcode_probabilities(m["n_barcodes"], m["barcode_weights"])))
    targets = {k: [int(s), int(e)] for k, s, e in zip(m["target_keys"], g["target_start"], g["target_end"])}
    keys, ts, te, tb = targets_to_arrays(targets, m["n_barcodes"])
    assert np.array_equal(tb, O.target_barcodes(m["target_keys"], m["n_barcodes"]))
    eng.set_targets(ts, te, tb)
    if g.mask is None:
        eng.set_blocked()
    else:
        eng.set_blocked(*g.mask)
    return tb


def golden_label_counts(g, ci, n_barcodes):
    names = O.label_names(n_barcodes)
    out = np.zeros(n_barcodes + 2, np.int64)
    for k, v in zip(g.meta["combos"][ci]["label_order"], g.combo(ci, "label_counts")):
        out[names.index(k)] = v
    return out


def assert_matches_golden(res, i, g, ci, n_reads):
    m = g.meta
    t = res.tallies[i].numpy()
    assert np.array_equal(t[:, 0], g.combo(ci, "full"))
    assert np.array_equal(t[:, 1], g.combo(ci, "partial"))
    assert np.array_equal(t[:, 2], g.combo(ci, "otb"))
    assert np.array_equal(res.label_counts[i].numpy(), golden_label_counts(g, ci, m["n_barcodes"]))
    assert int(res.n_bases[i]) == m["combos"][ci]["covered_length"]
    assert int(res.n_reads[i]) == n_reads
    tlen = np.maximum(g["target_end"] - g["target_start"], 1).astype(np.float64)
    cov_gpu, cov_ref = t[:, 2] / tlen, g.combo(ci, "otb") / tlen
    assert np.allclose(cov_gpu, cov_ref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", REPLAYABLE)
def test_replay_reads_bit_exact(eng, name):
    """Level 1: the reference's recorded reads -> tallies (count_matches + count_barcode_occurrences)."""
    g = Golden(name)
    configure(eng, g)
    m = g.meta
    starts, lens, labs, seg = [], [], [], [0]
    for ci in range(len(m["combinations"])):  # all combinations as iterations of one batch
        rs, re_ = g.combo(ci, "read_start"), g.combo(ci, "read_end")
        starts.append(rs); lens.append(re_ - rs); labs.append(g.combo(ci, "read_label").astype(np.int32))
        seg.append(seg[-1] + rs.size)
    res = eng.replay_reads(np.concatenate(starts), np.concatenate(lens), np.concatenate(labs), seg,
                           min_overlap=m["min_overlap"]).cpu()
    for ci in range(len(m["combinations"])):
        assert_matches_golden(res, ci, g, ci, seg[ci + 1] - seg[ci])
        assert int(res.status[ci]) == 0


@pytest.mark.parametrize("name", REPLAYABLE)
def test_overlap_lists_match_reference(eng, name):
    """The per-target `overlap` lists (counting.py:49,53,56), rebuilt from the device's hit records, equal the
    reference's lists element for element and in the reference's order -- from a level-1 replay and from a level-2
    replay whose bulk range overshoots the cut-off (records of taken-back draws must be dropped)."""
    g = Golden(name)
    configure(eng, g)
    m = g.meta
    tb = O.target_barcodes(m["target_keys"], m["n_barcodes"])
    ci = 0
    rs, re_ = g.combo(ci, "read_start"), g.combo(ci, "read_end")
    eng.set_hit_buffer(1 << 18)
    try:
        res = eng.replay_reads(rs, re_ - rs, g.combo(ci, "read_label").astype(np.int32), [0, rs.size], min_overlap=m["min_overlap"])
        h1 = eng.take_hits(res.n_reads)
        rng = O.Rng(m["seed"])
        mrl, cov = m["combinations"][ci]
        out = O.process_combination(rng, mrl, cov, m["genome_size"], g["target_start"], g["target_end"], tb, m["n_barcodes"],
                                    m["barcode_weights"], m["consuming"], g.mask, m["sigma"], m["min_read_length"],
                                    m["scaling"], m["min_overlap"], source_lengths=g.source_lengths, n_extra=3000, record_draws=True)
        pl = out["placement"]
        n_total = pl["start"].size
        res2 = eng.replay_draws(pl["u_barcode"], pl["length"], pl["start"], [0, n_total], cov, consuming=m["consuming"],
                                min_overlap=m["min_overlap"], ublock_off=pl["ublock_off"] if g.mask is not None else None,
                                ublock=pl["ublock"] if g.mask is not None else None, bulk_draws=n_total)
        h2 = eng.take_hits(res2.n_reads)
    finally:
        eng.set_hit_buffer(0)
    off, flat = g.combo(ci, "overlap_off"), g.combo(ci, "overlap_flat")
    for h in (h1, h2):
        assert h.shape[0] == flat.size
        assert np.array_equal(h[:, 3], flat)  # sorted by (row, draw index) == target-major, read order
        assert np.array_equal(np.bincount(h[:, 1], minlength=len(tb)), np.diff(off))
    assert np.array_equal(res.tallies.cpu()[0, :, 2].numpy(), np.add.reduceat(np.append(flat, 0), off[:-1]) * (np.diff(off) > 0))


def test_replay_scaling_below_one_bit_exact(eng):
    """scaling = 0.5 golden: the reference draws one uniform per qualifying (target, read) pair, target-major in read
    order (counting.py:46,51).  The device finds the qualifying pairs (scaling-1 replay + hit records), the recorded
    uniforms are applied to them in the reference's order on the device: tallies equal the reference's."""
    g = Golden("scaling_half")
    configure(eng, g)
    m = g.meta
    tb = O.target_barcodes(m["target_keys"], m["n_barcodes"])
    rng = O.Rng(m["seed"])
    mrl, cov = m["combinations"][0]
    src = O.lognormal_pool(rng, mrl, m["sigma"], m["min_read_length"])
    pool = O.choice(rng, src, int(cov * m["genome_size"] / mrl) * 10)
    cdf = O.barcode_cdf(O.barcode_probabilities(m["n_barcodes"], m["barcode_weights"]))
    pl = O.place(rng, m["genome_size"], cov, pool, m["n_barcodes"], cdf, m["consuming"], g.mask)
    n = pl["n_placed"]
    rs, ln, rl = pl["start"][:n], pl["length"][:n], pl["label"][:n]
    assert np.array_equal(rs, g.combo(0, "read_start"))
    orc = O.count_matches(rng, g["target_start"], g["target_end"], tb, rs, rs + ln, rl, m["scaling"], m["min_overlap"], want_pair_draws=True)
    assert np.array_equal(orc["otb"], g.combo(0, "otb"))  # the oracle's own thinning is the reference's
    eng.set_hit_buffer(1 << 18)
    try:
        res = eng.replay_reads(rs, ln, rl, [0, n], min_overlap=m["min_overlap"])
        hits = eng.take_hits(res.n_reads, keep_kind=True)
    finally:
        eng.set_hit_buffer(0)
    assert hits.shape[0] == orc["pair_u"].size
    t = eng.replay_thin(hits, orc["pair_u"], m["scaling"], 1).cpu().numpy()[0]
    assert np.array_equal(t[:, 0], g.combo(0, "full")) and np.array_equal(t[:, 1], g.combo(0, "partial")) and np.array_equal(t[:, 2], g.combo(0, "otb"))
    assert t[:, :2].sum() < res.tallies.cpu().numpy()[0, :, :2].sum()  # something was thinned away


@pytest.mark.parametrize("bulk_mode", ["tail_only", "bulk_half", "bulk_overshoot"])
@pytest.mark.parametrize("name", REPLAYABLE)
def test_replay_draws_bit_exact(eng, name, bulk_mode):
    """Level 2: the reference's RAW draws (oversampled past the cut-off) -> CDF lookup, blocked search,
    prefix-sum cut-off and tally on the device.  The raw draws come from the oracle's stream, which is pinned to
    reproduce the reference from the seed; the expected results are the reference's own (golden)."""
    g = Golden(name)
    configure(eng, g)
    m = g.meta
    tb = O.target_barcodes(m["target_keys"], m["n_barcodes"])

    def oracle_combo(rng, ci, **kw):
        mrl, cov = m["combinations"][ci]
        return O.process_combination(rng, mrl, cov, m["genome_size"], g["target_start"], g["target_end"], tb,
                                     m["n_barcodes"], m["barcode_weights"], m["consuming"], g.mask, m["sigma"],
                                     m["min_read_length"], m["scaling"], m["min_overlap"],
                                     source_lengths=g.source_lengths, **kw)

    for ci, (mrl, cov) in enumerate(m["combinations"]):
        rng = O.Rng(m["seed"])
        for prev in range(ci):  # earlier combinations advance the shared stream exactly as in the reference
            oracle_combo(rng, prev)
        out = oracle_combo(rng, ci, n_extra=3000, record_draws=True)
        pl = out["placement"]
        n_placed = pl["n_placed"]
        assert n_placed == g.combo(ci, "read_start").size
        n_total = pl["start"].size
        bulk = {"tail_only": 0, "bulk_half": (n_placed // 2), "bulk_overshoot": n_total}[bulk_mode]
        res = eng.replay_draws(pl["u_barcode"], pl["length"], pl["start"], [0, n_total], cov,
                               consuming=m["consuming"], min_overlap=m["min_overlap"],
                               ublock_off=pl["ublock_off"] if g.mask is not None else None,
                               ublock=pl["ublock"] if g.mask is not None else None, bulk_draws=bulk).cpu()
        assert_matches_golden(res, 0, g, ci, n_placed)
        st = int(res.status[0])
        assert not st & 2
        if bulk_mode == "bulk_overshoot" and n_total >= 1024:
            assert st & 1  # the cut-off fell inside the bulk range and was taken back on the device
        if bulk_mode == "tail_only":
            assert st == 0


def test_replay_draws_exhausted_flag(eng):
    g = Golden("cfg1_roi_10mb")
    configure(eng, g)
    rs, re_ = g.combo(0, "read_start"), g.combo(0, "read_end")
    n = rs.size // 2  # half the draws: coverage cannot be reached
    res = eng.replay_draws(np.full(n, 0.5), (re_ - rs)[:n], rs[:n], [0, n], 10.0).cpu()
    assert int(res.status[0]) & 2
    assert int(res.n_reads[0]) == n


def test_replay_batch_of_iterations_and_empty_iteration(eng):
    """Ragged batch: several iterations of different sizes in one call, one of them empty."""
    g = Golden("dense_adversarial")
    configure(eng, g)
    m = g.meta
    rs, re_, rl = g.combo(0, "read_start"), g.combo(0, "read_end"), g.combo(0, "read_label").astype(np.int32)
    cuts = [0, 700, 700, 1500, rs.size]
    res = eng.replay_reads(rs, re_ - rs, rl, cuts, min_overlap=m["min_overlap"]).cpu()
    tb = O.target_barcodes(m["target_keys"], m["n_barcodes"])
    total = np.zeros((len(m["target_keys"]), 3), np.int64)
    for i in range(4):
        a, b = cuts[i], cuts[i + 1]
        exp = O.count_matches(None, g["target_start"], g["target_end"], tb, rs[a:b], re_[a:b], rl[a:b], 1.0, m["min_overlap"])
        t = res.tallies[i].numpy()
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        assert int(res.n_reads[i]) == b - a
        total += t
    assert np.array_equal(total[:, 2], g.combo(0, "otb"))  # linearity: tallies of a partition add up
    assert int(res.n_reads[1]) == 0 and not res.tallies[1].any()


# ---------------------------------------------------------------- native (Philox) mode

def native_expected(g, ci, seed, combo, iteration, pool, consuming=None, min_overlap=None, coverage=None):
    m = g.meta
    cov = m["combinations"][ci][1] if coverage is None else coverage
    cdf = O.barcode_cdf(O.barcode_probabilities(m["n_barcodes"], m["barcode_weights"]))
    it = NN.native_iteration(seed, combo, iteration, m["genome_size"], cov, pool, cdf,
                             consuming=m["consuming"] if consuming is None else consuming, mask=g.mask)
    tb = O.target_barcodes(m["target_keys"], m["n_barcodes"])
    exp = O.count_matches(None, g["target_start"], g["target_end"], tb, it["start"], it["start"] + it["length"],
                          it["label"], 1.0, m["min_overlap"] if min_overlap is None else min_overlap)
    counts, _ = O.label_hist(it["label"], m["n_barcodes"])
    return it, exp, counts


@pytest.mark.parametrize("name", ["cfg1_roi_10mb", "ins_b4_weights", "blocked_w1_notconsuming", "blocked_w1_consuming",
                                  "blocked_fractional_multihit", "dense_adversarial", "genome_gt_u32", "cfg2_reduced_50mb"])
def test_native_mode_matches_cpu_restatement(eng, name):
    """Native Philox mode, bit-exact: the CPU regenerates the same counter-based draws (oracle/native_np.py),
    applies the reference's placement rules and the oracle's tally; the GPU must give identical integers."""
    g = Golden(name)
    configure(eng, g)
    m = g.meta
    mrl, cov = m["combinations"][0]
    pool = O.lognormal_pool(O.Rng(5), mrl, m["sigma"], m["min_read_length"], n=200000)
    eng.set_length_table(pool)
    seed, combo, it0, n_iter = 0xDEADBEEF12345678, 3, 5, 3
    res = eng.run_batch(seed, combo, it0, n_iter, cov, m["consuming"], m["min_overlap"], 1.0).cpu()
    for i in range(n_iter):
        it, exp, counts = native_expected(g, 0, seed, combo, it0 + i, pool)
        t = res.tallies[i].numpy()
        assert int(res.n_reads[i]) == it["n_placed"]
        assert int(res.n_bases[i]) == it["covered"]
        assert np.array_equal(res.label_counts[i].numpy(), counts)
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        assert not int(res.status[i]) & 2
    # batching / sharding invariance: one iteration alone equals its slot in the batch
    solo = eng.run_batch(seed, combo, it0 + 1, 1, cov, m["consuming"], m["min_overlap"], 1.0).cpu()
    assert np.array_equal(solo.tallies[0].numpy(), res.tallies[1].numpy())
    assert int(solo.n_reads[0]) == int(res.n_reads[1])


@pytest.mark.parametrize("consuming", [False, True])
def test_native_masked_second_bulk_pass(eng, consuming):
    """Large enough for the bulk kernel to run, 35 % of the genome blocked: with consuming=False the cut-off moves
    out by the blocked fraction and the device-planned second bulk pass covers most of the extension.  Exact
    against the CPU restatement of the same Philox draws."""
    G = 20_000_000
    rng = np.random.RandomState(12)
    ts = np.sort(rng.randint(0, G - 5000, 60)); te = ts + rng.randint(200, 5000, 60)
    tb = (np.arange(60) % 2).astype(np.int32)
    mask = (np.array([1_000_000, 9_000_000, 9_500_000, 15_000_000]), np.array([5_000_000, 10_000_000, 11_000_000, 17_000_000]),
            np.array([1.0, 0.5, 1.0, 1.0]), np.array([-1, 0, -1, 1], np.int32))
    eng.set_genome(G)
    cdf = O.barcode_cdf(O.barcode_probabilities(2, {"0": 2}))
    eng.set_barcode_cdf(cdf)
    eng.set_targets(ts, te, tb)
    eng.set_blocked(*mask)
    pool = O.lognormal_pool(O.Rng(21), 2000, 1, 50, n=100000)
    eng.set_length_table(pool)
    cov, seed = 6.0, 4242
    n1 = eng.bulk_reads(cov)
    assert n1 > 20000
    res = eng.run_batch(seed, 1, 10, 2, cov, consuming, 1, 1.0).cpu()
    for i in range(2):
        it = NN.native_iteration(seed, 1, 10 + i, G, cov, pool, cdf, consuming=consuming, mask=mask)
        exp = O.count_matches(None, ts, te, tb, it["start"], it["start"] + it["length"], it["label"], 1.0, 1)
        counts, _ = O.label_hist(it["label"], 2)
        t = res.tallies[i].numpy()
        assert int(res.n_reads[i]) == it["n_placed"] and int(res.n_bases[i]) == it["covered"]
        assert np.array_equal(res.label_counts[i].numpy(), counts)
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        if not consuming:
            assert it["n_placed"] > 1.2 * n1  # the second pass had real work to do
    assert not (res.status.numpy() & 2).any()


def test_native_materialize_reads(eng):
    g = Golden("blocked_fractional_multihit")
    configure(eng, g)
    m = g.meta
    mrl, cov = m["combinations"][0]
    pool = O.lognormal_pool(O.Rng(6), mrl, 1, 1, n=50000)
    eng.set_length_table(pool)
    res = eng.run_batch(77, 0, 0, 1, cov, m["consuming"], 1, 1.0).cpu()
    n = int(res.n_reads[0])
    s, ln, lab = (x.cpu().numpy() for x in eng.materialize_reads(77, 0, 0, m["consuming"], n))
    it, _, _ = native_expected(g, 0, 77, 0, 0, pool)
    assert np.array_equal(s, it["start"]) and np.array_equal(ln, it["length"]) and np.array_equal(lab, it["label"])


def test_native_scaling_thins_pairs(eng):
    g = Golden("scaling_half")
    configure(eng, g)
    m = g.meta
    mrl, cov = m["combinations"][0]
    pool = O.lognormal_pool(O.Rng(8), mrl, 1, 1, n=100000)
    eng.set_length_table(pool)
    full = eng.run_batch(9, 0, 0, 40, cov, False, 1, 1.0).cpu()
    half = eng.run_batch(9, 0, 0, 40, cov, False, 1, 0.5).cpu()
    assert np.array_equal(full.n_reads.numpy(), half.n_reads.numpy())  # thinning does not touch placement
    a = full.tallies[:, :, :2].sum().item(); b = half.tallies[:, :, :2].sum().item()
    assert abs(b / a - 0.5) < 4 * np.sqrt(0.25 / a)  # binomial thinning, 4 sigma
    # exact: recompute the kept pairs on the CPU from the same counter-based uniforms
    it = NN.native_iteration(9, 0, 0, m["genome_size"], cov, pool, O.barcode_cdf(O.barcode_probabilities(2, None)))
    tb = O.target_barcodes(m["target_keys"], 2)
    exp = np.zeros((len(tb), 3), np.int64)
    rs, re_ = it["start"], it["start"] + it["length"]
    for row, (s, e, b_) in enumerate(zip(g["target_start"], g["target_end"], tb)):
        for n in np.nonzero((it["label"] == b_) & (rs < e) & (re_ > s))[0]:
            ov = min(e, re_[n]) - max(s, rs[n])
            if ov >= 1 and NN.scale_u(9, 0, 0, int(n), row) <= 0.5:
                exp[row, 0 if (rs[n] <= s and e <= re_[n]) else 1] += 1
                exp[row, 2] += ov
    assert np.array_equal(half.tallies[0].numpy(), exp)


def test_statistical_check_vs_reference_golden_and_analytic(eng):
    """Statistical check (north_star): with the native stream, per-target mean coverage and detection rate agree
    with the expectation within 3 standard errors.  Expectation: analytic (uniform placement => E[on-target bases]
    = coverage * p_barcode * target length away from the genome ends, README.md:36) and the reference's own
    single-iteration golden values must lie inside the native distribution."""
    g = Golden("cfg2_reduced_50mb")
    configure(eng, g)
    m = g.meta
    mrl, cov = m["combinations"][0]
    pool = O.lognormal_pool(O.Rng(11), mrl, 1, 1, n=1000000)
    eng.set_length_table(pool)
    n_iter = 400
    res = eng.run_batch(2026, 0, 0, n_iter, cov, False, 1, 1.0).cpu()
    t = res.tallies.numpy().astype(np.float64)
    tlen = (g["target_end"] - g["target_start"]).astype(np.float64)
    depth = t[:, :, 2] / tlen  # [iter, target] derived coverage
    mean, se = depth.mean(0), depth.std(0, ddof=1) / np.sqrt(n_iter)
    z = (mean - cov) / se
    assert np.abs(z).max() < 4.5 and (np.abs(z) < 3).mean() > 0.97, z  # 100 targets: allow the expected few > 3 SE
    pooled = depth.mean()
    assert abs(pooled - cov) < 3 * depth.mean(1).std(ddof=1) / np.sqrt(n_iter)
    # the reference's golden iteration is one draw from the same distribution
    ref_depth = g.combo(0, "otb") / tlen
    zz = (ref_depth - mean) / depth.std(0, ddof=1)
    assert np.abs(zz).max() < 5 and np.abs(zz.mean()) < 3 / np.sqrt(zz.size) * 1.5
    det = (t[:, :, 0] + t[:, :, 1]).mean(0)  # detections per target per iteration
    ref_det = (g.combo(0, "full") + g.combo(0, "partial")).astype(np.float64)
    assert abs(ref_det.mean() - det.mean()) < 3 * np.sqrt(det.mean() / ref_det.size + det.mean() / (ref_det.size * n_iter)) * 1.5
    # reads placed and N_bases: cut-off semantics (covered >= coverage*G, last read included)
    G = m["genome_size"]
    assert (res.n_bases.numpy() >= cov * G).all()
    assert (res.n_bases.numpy() - cov * G < pool.max()).all()
    assert not (res.status.numpy() & 2).any()


def test_moments_kernel(eng):
    g = Golden("ins_b4_weights")
    configure(eng, g)
    pool = O.lognormal_pool(O.Rng(3), 5000, 1, 1, n=50000)
    eng.set_length_table(pool)
    res = eng.run_batch(1, 0, 0, 7, 5, False, 1, 1.0)
    acc = eng.moments(res.tallies)
    acc = eng.moments(res.tallies, acc).cpu().numpy()  # accumulates
    t = res.tallies.cpu().numpy().astype(object)
    assert (acc[:, 0] == 14).all()
    assert np.array_equal(acc[:, 1], 2 * t[:, :, 0].sum(0)) and np.array_equal(acc[:, 2], 2 * (t[:, :, 0] ** 2).sum(0))
    assert np.array_equal(acc[:, 3], 2 * t[:, :, 1].sum(0)) and np.array_equal(acc[:, 4], 2 * (t[:, :, 1] ** 2).sum(0))
    assert np.array_equal(acc[:, 5], 2 * t[:, :, 2].sum(0))
    sq = (t[:, :, 2] ** 2).sum(0)  # python ints: sum b^2 is carried as hi * 2^32 + lo
    assert [int(hi) * 2**32 + int(lo) for hi, lo in zip(acc[:, 7], acc[:, 6])] == [2 * int(x) for x in sq]


def test_error_behaviour(eng):
    from esloco_b200.engine import EngineError
    g = Golden("cfg1_roi_10mb")
    configure(eng, g)
    eng.set_length_table(np.array([20_000_000], np.int64))  # longer than the 10 Mb genome: numpy raises low >= high
    with pytest.raises(EngineError):
        eng.run_batch(1, 0, 0, 1, 1.0)
    with pytest.raises(EngineError):
        eng.set_targets([-5], [10], [0])
    with pytest.raises(EngineError):
        eng.set_barcode_cdf([0.5, 0.2])


def test_full_size_config2_consistency(eng):
    """BASELINE.json configs[1] at FULL size (hg38-size genome, 30x, 100 insertions, ~6.2 M reads per iteration).
    Size-independent properties: the native kernels, a level-1 replay of the materialised reads and the oracle's
    indexed CPU tally of the same reads agree bit for bit; the cut-off is tight; batch splitting is invisible."""
    from esloco_b200 import synthetic as S
    from esloco_b200.utils import targets_to_arrays
    G, _ = S.chromosome_dir(S.HG38)
    targets = S.random_insertions(G, 100, 5000, 1, seed=2)
    keys, ts, te, tb = targets_to_arrays(targets, 1)
    pool = S.lognormal_lengths(15000, 1, 1, seed=1)
    eng.set_genome(G)
    eng.set_barcode_cdf([1.0])
    eng.set_targets(ts, te, tb)
    eng.set_blocked()
    eng.set_length_table(pool)
    cov, seed = 30, 20261019
    res = eng.run_batch(seed, 0, 7, 3, cov).cpu()
    thr = cov * G
    nb, nr = res.n_bases.numpy(), res.n_reads.numpy()
    assert (nb >= thr).all() and (nb - thr < pool.max()).all() and (res.status.numpy() == 0).all()
    assert abs(nr.mean() - thr / pool.mean()) < 6 * pool.std() / pool.mean() * np.sqrt(thr / pool.mean())
    assert np.array_equal(res.label_counts[:, 0].numpy(), nr)
    solo = eng.run_batch(seed, 0, 8, 1, cov).cpu()
    assert np.array_equal(solo.tallies[0].numpy(), res.tallies[1].numpy()) and int(solo.n_bases[0]) == int(nb[1])
    # the same reads through the other two implementations
    s, ln, lab = eng.materialize_reads(seed, 0, 8, False, int(nr[1]))
    assert int(ln.sum().item()) == int(nb[1])
    rep = eng.replay_reads(s, ln, lab, [0, int(nr[1])]).cpu()
    assert np.array_equal(rep.tallies[0].numpy(), res.tallies[1].numpy())
    sh, lh, labh = s.cpu().numpy(), ln.cpu().numpy(), lab.cpu().numpy()
    exp = O.count_matches_indexed(ts, te, tb, 1, sh, sh + lh, labh, 1)
    t = res.tallies[1].numpy()
    assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
    depth = t[:, 2] / 5000.0
    assert abs(depth.mean() - cov) < 5 * depth.std(ddof=1) / np.sqrt(depth.size)  # ~30x on the targets


@pytest.mark.parametrize("B", [7, 96])
def test_native_many_barcodes_exact(eng, B):
    """Barcode selection through the 1024-slice start table + walk, B up to 96 with uneven weights: exact against the
    CPU restatement; every barcode's read count is inside 5 sigma of its binomial expectation."""
    G = 30_000_000
    rng = np.random.RandomState(B)
    T = 3 * B
    ts = np.sort(rng.randint(0, G - 9000, T)); te = ts + rng.randint(500, 9000, T)
    tb = (np.arange(T) % B).astype(np.int32)
    weights = {str(k): int(w) for k, w in zip(range(0, B, 3), rng.randint(2, 40, B))}
    cdf = O.barcode_cdf(O.barcode_probabilities(B, weights))
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    pool = O.lognormal_pool(O.Rng(B), 3000, 1, 10, n=100000)
    eng.set_length_table(pool)
    res = eng.run_batch(99, 0, 0, 2, 8.0).cpu()
    for i in range(2):
        it = NN.native_iteration(99, 0, i, G, 8.0, pool, cdf)
        exp = O.count_matches(None, ts, te, tb, it["start"], it["start"] + it["length"], it["label"], 1.0, 1)
        counts, _ = O.label_hist(it["label"], B)
        assert np.array_equal(res.label_counts[i].numpy(), counts)
        t = res.tallies[i].numpy()
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
    p = np.diff(np.concatenate(([0.0], cdf)))
    n = res.n_reads[0].item()
    z = (res.label_counts[0, :B].numpy() - n * p) / np.sqrt(n * p * (1 - p))
    assert np.abs(z).max() < 5


def test_edge_cases(eng):
    """Empty target set, zero coverage, a single-entry length table, targets beyond the genome end, empty batch slot."""
    G = 1_000_000
    eng.set_genome(G); eng.set_barcode_cdf([0.5, 1.0]); eng.set_blocked()
    eng.set_length_table(np.array([1000]))
    eng.set_targets([], [], [])
    res = eng.run_batch(1, 0, 0, 2, 3.0).cpu()
    assert res.tallies.shape == (2, 0, 3) and (res.n_bases.numpy() == 3 * G).all() and (res.n_reads.numpy() == 3000).all()
    eng.set_targets([0, G - 10, 2 * G, 500_000, 500_000], [10, G + 500, 3 * G, 500_000, 400_000], [0, 1, 0, 1, 0])
    res = eng.run_batch(1, 0, 0, 1, 0.0).cpu()  # coverage 0: the reference's while-loop never runs
    assert int(res.n_reads[0]) == 0 and int(res.n_bases[0]) == 0 and not res.tallies.any()
    res = eng.run_batch(1, 0, 0, 3, 50.0).cpu()  # 50 k reads of 1 kb per iteration
    t = res.tallies.numpy()
    assert not t[:, 2].any() and not t[:, 3].any() and not t[:, 4].any()  # beyond the genome / empty / inverted: never hit
    assert (t[:, 1, 0] == 0).all()  # a target sticking out of the genome can never be fully covered
    assert t[:, :2, 2].sum() > 0  # the two edge targets are hit now and then (p ~ 1e-5 per read, 150 k reads)
    it = NN.native_iteration(1, 0, 1, G, 50.0, np.array([1000]), np.array([0.5, 1.0]))
    exp = O.count_matches(None, np.array([0, G - 10, 2 * G, 500_000, 500_000]), np.array([10, G + 500, 3 * G, 500_000, 400_000]),
                          np.array([0, 1, 0, 1, 0], np.int32), it["start"], it["start"] + it["length"], it["label"], 1.0, 1)
    assert np.array_equal(t[1, :, 0], exp["full"]) and np.array_equal(t[1, :, 1], exp["partial"]) and np.array_equal(t[1, :, 2], exp["otb"])


def test_native_u64_coordinates_bulk_path(eng):
    """Genome > 2^32: every kernel runs its 64-bit instantiation, with enough reads for the bulk kernel (the golden
    genome_gt_u32 case only reaches the cut-off kernel)."""
    G = 6_000_000_000
    rng = np.random.RandomState(64)
    T = 300
    ts = np.sort(rng.randint(0, G - 3_000_000, T, dtype=np.int64)); te = ts + rng.randint(1000, 3_000_000, T)
    tb = np.zeros(T, np.int32)
    eng.set_genome(G); eng.set_barcode_cdf([1.0]); eng.set_targets(ts, te, tb); eng.set_blocked()
    pool = O.lognormal_pool(O.Rng(64), 20000, 1, 1, n=200000)
    eng.set_length_table(pool)
    cov = 0.5
    assert eng.bulk_reads(cov) > 100_000
    res = eng.run_batch(7, 2, 3, 2, cov).cpu()
    for i in range(2):
        it = NN.native_iteration(7, 2, 3 + i, G, cov, pool, np.array([1.0]))
        assert it["start"].max() > 2**32
        exp = O.count_matches_indexed(ts, te, tb, 1, it["start"], it["start"] + it["length"], it["label"], 1)
        t = res.tallies[i].numpy()
        assert int(res.n_reads[i]) == it["n_placed"] and int(res.n_bases[i]) == it["covered"]
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        assert t[:, 2].sum() > 0


def test_native_more_barcodes_than_shared_descriptors(eng):
    """B = 1500 > 1024: first-layer descriptors are read from global memory instead of shared memory."""
    G, B = 40_000_000, 1500
    rng = np.random.RandomState(15)
    ts = np.sort(rng.randint(0, G - 20000, 2 * B)); te = ts + rng.randint(2000, 20000, 2 * B)
    tb = (np.arange(2 * B) % B).astype(np.int32)
    cdf = O.barcode_cdf(O.barcode_probabilities(B, {"7": 50, "900": 30}))
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    pool = O.lognormal_pool(O.Rng(15), 4000, 1, 10, n=100000)
    eng.set_length_table(pool)
    res = eng.run_batch(5, 0, 0, 1, 12.0).cpu()
    it = NN.native_iteration(5, 0, 0, G, 12.0, pool, cdf)
    exp = O.count_matches_indexed(ts, te, tb, B, it["start"], it["start"] + it["length"], it["label"], 1)
    counts, _ = O.label_hist(it["label"], B)
    t = res.tallies[0].numpy()
    assert np.array_equal(res.label_counts[0].numpy(), counts)
    assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
    assert t[:, 2].sum() > 0


@pytest.mark.parametrize("name", ["cfg1_roi_10mb", "ins_b4_weights", "blocked_w1_notconsuming", "blocked_fractional_multihit",
                                  "dense_adversarial", "cfg2_reduced_50mb"])
def test_staged_baseline_equals_fused(eng, name):
    """The staged pipeline (generate -> block-scan cut-off -> radix sort -> per-target sweep) and the fused kernels are
    two independent implementations of the same draws: every output must agree bit for bit."""
    g = Golden(name)
    configure(eng, g)
    m = g.meta
    mrl, cov = m["combinations"][0]
    pool = O.lognormal_pool(O.Rng(17), mrl, m["sigma"], m["min_read_length"], n=150000)
    eng.set_length_table(pool)
    fused = eng.run_batch(4711, 2, 9, 3, cov, m["consuming"], m["min_overlap"], 1.0).cpu()
    staged = eng.run_batch_staged(4711, 2, 9, 3, cov, m["consuming"], m["min_overlap"]).cpu()
    assert np.array_equal(staged.n_reads.numpy(), fused.n_reads.numpy())
    assert np.array_equal(staged.n_bases.numpy(), fused.n_bases.numpy())
    assert np.array_equal(staged.label_counts.numpy(), fused.label_counts.numpy())
    assert np.array_equal(staged.tallies.numpy(), fused.tallies.numpy())
    assert not (staged.status.numpy() & 2).any() and fused.tallies.numpy().sum() > 0


def test_staged_baseline_larger_run(eng):
    """~0.5 M reads per iteration, 24 barcodes, 30 % of the genome blocked: radix sort over many CTAs, both bulk passes."""
    G, B = 60_000_000, 24
    rng = np.random.RandomState(77)
    T = 40 * B
    ts = np.sort(rng.randint(0, G - 30000, T)); te = ts + rng.randint(500, 30000, T)
    tb = (np.arange(T) % B).astype(np.int32)
    mask = (np.array([2_000_000, 30_000_000]), np.array([12_000_000, 38_000_000]), np.array([1.0, 1.0]), np.array([-1, 3], np.int32))
    cdf = O.barcode_cdf(O.barcode_probabilities(B, {"0": 10, "1": 5}))
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked(*mask)
    pool = O.lognormal_pool(O.Rng(78), 2500, 1, 1, n=300000)
    eng.set_length_table(pool)
    fused = eng.run_batch(3, 0, 0, 2, 20.0, False, 100, 1.0).cpu()
    staged = eng.run_batch_staged(3, 0, 0, 2, 20.0, False, 100).cpu()
    assert int(fused.n_reads[0]) > 400_000
    for a, b in ((staged.n_reads, fused.n_reads), (staged.n_bases, fused.n_bases), (staged.label_counts, fused.label_counts),
                 (staged.tallies, fused.tallies)):
        assert np.array_equal(a.numpy(), b.numpy())


def test_full_size_config2_replay_of_reference_stream(eng):
    """BASELINE.json configs[1] at FULL size, replay check: the oracle (pinned to the reference by the golden files)
    produces the reference's numpy-stream draws for one hg38-size iteration from a seed -- 1 M log-normal pool, 62 M
    resample, ~6.2 M placed reads, oversampled by 50 k draws; the device does the cut-off and the tally from the raw
    draws, and must match the oracle's (faithful O(T*R)) tally, N_bases and read count bit for bit."""
    from esloco_b200 import synthetic as S
    from esloco_b200.utils import targets_to_arrays
    G, _ = S.chromosome_dir(S.HG38)
    np.random.seed(20261019)
    from esloco_b200.create_insertion_genome import add_insertions_to_genome_sequence_with_bed
    _, cdir = S.chromosome_dir(S.HG38)
    targets = add_insertions_to_genome_sequence_with_bed(G, 5000, 100, {f"Barcode_0_{k}": v for k, v in cdir.items()})
    keys, ts, te, tb = targets_to_arrays(targets, 1)
    rng = O.Rng(424242)
    out = O.process_combination(rng, 15000, 30, G, ts, te, tb, 1, None, False, None, 1, 1, 1.0, 1, n_extra=50_000, record_draws=True)
    pl = out["placement"]
    n_placed, n_total = pl["n_placed"], pl["start"].size
    assert 6_000_000 < n_placed < 6_400_000 and n_total == n_placed + 50_000
    eng.set_genome(G); eng.set_barcode_cdf([1.0]); eng.set_targets(ts, te, tb); eng.set_blocked()
    for bulk in (0, n_placed - 30_000, n_total):  # cut-off kernel only / normal split / overshoot + take-back
        res = eng.replay_draws(pl["u_barcode"], pl["length"], pl["start"], [0, n_total], 30, bulk_draws=bulk).cpu()
        t = res.tallies[0].numpy()
        assert int(res.n_reads[0]) == n_placed and int(res.n_bases[0]) == pl["covered_length"]
        assert np.array_equal(t[:, 0], out["full"]) and np.array_equal(t[:, 1], out["partial"]) and np.array_equal(t[:, 2], out["otb"])
        assert int(res.label_counts[0, 0]) == n_placed and not int(res.status[0]) & 2
    assert out["otb"].sum() > 0


def test_native_deep_nesting_fills_candidate_queues(eng):
    """General variant of the bulk kernel (too many first-layer targets for the shared pre-filter): 200 nested
    genome-spanning targets on barcode 0 (the index tolerates 16 open records per layer, so they need 12 layers) make
    every read of that barcode a candidate in every layer, so the per-warp candidate queues overflow and are drained
    mid-tile; barcode 1 has a single layer.  Exact against the CPU restatement, with and without thinning (the
    thinning uniform is keyed by the draw index the queue item carries)."""
    G = 5_000_000
    rng = np.random.RandomState(12)
    n_small = 3000
    ss = np.sort(rng.randint(0, G - 3000, n_small)); se = ss + rng.randint(200, 1500, n_small)
    ns = np.arange(200) * 1000 + 5; ne = G - np.arange(200) * 1000 - 7
    ts = np.concatenate([ns, ss, ss + 3]); te = np.concatenate([ne, se, se + 3])
    tb = np.concatenate([np.zeros(200 + n_small), np.ones(n_small)]).astype(np.int32)
    cdf = np.array([0.5, 1.0])
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    pool = O.lognormal_pool(O.Rng(12), 2500, 1, 10, n=50000)
    eng.set_length_table(pool)
    cov = 4.0
    assert eng.bulk_reads(cov) > 3000
    res = eng.run_batch(21, 1, 5, 2, cov).cpu()
    assert eng.target_layers() >= 12  # (the index is built by the first run after set_targets)
    for i in range(2):
        it = NN.native_iteration(21, 1, 5 + i, G, cov, pool, cdf)
        exp = O.count_matches_indexed(ts, te, tb, 2, it["start"], it["start"] + it["length"], it["label"], 1)
        t = res.tallies[i].numpy()
        assert int(res.n_reads[i]) == it["n_placed"] and int(res.n_bases[i]) == it["covered"]
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        assert (t[:200, 1] > 0.35 * it["n_placed"]).all()
    thin = eng.run_batch(21, 1, 5, 1, cov, False, 1, 0.6).cpu().tallies[0].numpy()
    it = NN.native_iteration(21, 1, 5, G, cov, pool, cdf)
    rs, re_ = it["start"], it["start"] + it["length"]
    for row in (0, 199, 200, 200 + n_small):  # outermost / innermost nested, one small target per barcode
        s, e, b_ = ts[row], te[row], tb[row]
        exp = np.zeros(3, np.int64)
        for n in np.nonzero((it["label"] == b_) & (rs < e) & (re_ > s))[0]:
            if NN.scale_u(21, 1, 5, int(n), row) <= 0.6:
                exp[0 if (rs[n] <= s and e <= re_[n]) else 1] += 1
                exp[2] += min(e, re_[n]) - max(s, rs[n])
        assert np.array_equal(thin[row], exp)


def test_full_size_config4_consistency(eng):
    """BASELINE.json configs[3] at FULL size: hg38-size genome, 30x, 50 k ROIs of 1-100 kb (nested all over: the index
    keeps them in one layer with dead records), 824 blocked intervals of weight 1 not consuming coverage (two bulk
    passes), min_overlap 1000, ~13.6 M draws per iteration.  The native kernels, a level-1 replay of the materialised
    reads and the oracle's indexed CPU tally of the same reads agree bit for bit; the blocked labels equal an
    independent numpy interval test; the cut-off is tight on the counted bases."""
    from esloco_b200 import synthetic as S
    from esloco_b200.utils import build_mask_tree, mask_tree_to_arrays, targets_to_arrays
    G, _ = S.chromosome_dir(S.HG38)
    keys, ts, te, tb = targets_to_arrays(S.barcode_rois(S.random_rois(S.HG38, 50000, 1000, 100000, seed=4), 1), 1)
    ms, me, mw, mb = (np.asarray(x) for x in mask_tree_to_arrays(build_mask_tree(S.gap_mask(S.HG38, 800, seed=5)), 1))
    pool = S.lognormal_lengths(8000, 1, 1, seed=1)
    eng.set_genome(G); eng.set_barcode_cdf([1.0]); eng.set_targets(ts, te, tb); eng.set_blocked(ms, me, mw, mb)
    eng.set_length_table(pool)
    cov, seed, min_ov = 30, 4, 1000
    res = eng.run_batch(seed, 0, 3, 2, cov, False, min_ov).cpu()
    assert eng.target_layers() == 1 and not eng.index_built_on_device()  # nested ROIs: the device builder hands them to the host
    thr = cov * G
    nb, nr = res.n_bases.numpy(), res.n_reads.numpy()
    assert (nb >= thr).all() and (nb - thr < pool.max()).all() and not (res.status.numpy() & 2).any()
    s, ln, lab = eng.materialize_reads(seed, 0, 4, False, int(nr[1]))
    sh, lh, labh = s.cpu().numpy(), ln.cpu().numpy(), lab.cpu().numpy()
    # blocked <=> the read overlaps some blocked interval (all weights are 1): merged intervals + two bisections
    o = np.argsort(ms); a, b = ms[o], np.maximum.accumulate(me[o])
    first = np.searchsorted(b, sh, side="right")          # first interval whose running max of ends passes the start
    blocked = (first < a.size) & (a[np.minimum(first, a.size - 1)] < sh + lh)
    assert np.array_equal(labh == -2, blocked) and ((labh == 0) | (labh == -2)).all()
    assert int(lh[~blocked].sum()) == int(nb[1]) and 0.10 < blocked.mean() < 0.20
    assert np.array_equal(res.label_counts[1].numpy(), [int((~blocked).sum()), 0, int(blocked.sum())])
    rep = eng.replay_reads(s, ln, lab, [0, int(nr[1])], min_ov).cpu()
    assert np.array_equal(rep.tallies[0].numpy(), res.tallies[1].numpy())
    exp = O.count_matches_indexed(ts, te, tb, 1, sh, sh + lh, labh, min_ov)
    t = res.tallies[1].numpy()
    assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
    assert t[:, 2].sum() > 0 and t[:, 0].sum() > 0


def test_native_hot_targets_private_tallies(eng):
    """Targets a read hits with probability >= 1/512 ("hot": whole-genome / whole-chromosome ROIs) are tallied in
    per-CTA shared-memory sums that the bulk kernel flushes every 64 tiles and at iteration boundaries; 75 such
    targets exceed the 64 slots, so some take the plain global-atomic path.  12 M reads per iteration make every CTA
    pass several flush points.  Exact against the CPU restatement; a level-1 replay of the same reads agrees."""
    G = 20_000_000
    rng = np.random.RandomState(5)
    big_s = np.array([0, 0, 5_000_000, 17_000_000]); big_e = np.array([G, G // 2, 9_000_000, G + 50])
    mid_s = np.sort(rng.randint(0, G - 300_000, 71)); mid_e = mid_s + rng.randint(60_000, 300_000, 71)  # p_hit >= 1/330
    sm_s = np.sort(rng.randint(0, G - 5000, 400)); sm_e = sm_s + rng.randint(100, 5000, 400)
    ts = np.concatenate([big_s, mid_s, sm_s]); te = np.concatenate([big_e, mid_e, sm_e])
    tb = (np.arange(ts.size) % 2).astype(np.int32)
    cdf = np.array([0.7, 1.0])
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    pool = O.lognormal_pool(O.Rng(5), 1000, 1, 50, n=100000)
    eng.set_length_table(pool)
    cov = 600.0
    res = eng.run_batch(77, 0, 0, 2, cov).cpu()
    nr = res.n_reads.numpy()
    assert nr.min() > 10_000_000
    it = NN.native_iteration(77, 0, 1, G, cov, pool, cdf)
    assert int(nr[1]) == it["n_placed"] and int(res.n_bases[1]) == it["covered"]
    exp = O.count_matches_indexed(ts, te, tb, 2, it["start"], it["start"] + it["length"], it["label"], 1)
    t = res.tallies[1].numpy()
    assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
    assert t[0, 1] > 8_000_000 and t[0, 2] > 2**32  # the whole-genome target: every read of barcode 0, > 32 bits of bases
    s, ln, lab = eng.materialize_reads(77, 0, 1, False, int(nr[1]))
    rep = eng.replay_reads(s, ln, lab, [0, int(nr[1])]).cpu()
    assert np.array_equal(rep.tallies[0].numpy(), t)


def test_native_longer_pool_rebuilds_prefilter(eng):
    """Small target sets are pre-filtered in shared memory with per-bucket bounds that assume a longest read; a new
    length table with longer reads must rebuild them (no stale bound may hide an overlap).  Exact against the CPU
    restatement before and after the switch, targets untouched in between."""
    G = 40_000_000
    rng = np.random.RandomState(31)
    ts = np.sort(rng.randint(0, G - 20000, 60)); te = ts + rng.randint(500, 20000, 60)
    tb = (np.arange(60) % 3).astype(np.int32)
    cdf = np.array([0.2, 0.5, 1.0])
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    for mean, cov in ((800, 3.0), (150000, 25.0)):  # the second pool's reads are ~200x longer than the first's
        pool = O.lognormal_pool(O.Rng(mean), mean, 1, 10, n=50000)
        eng.set_length_table(pool)
        assert eng.bulk_reads(cov) > 3000
        res = eng.run_batch(8, 0, 2, 1, cov).cpu()
        it = NN.native_iteration(8, 0, 2, G, cov, pool, cdf)
        exp = O.count_matches_indexed(ts, te, tb, 3, it["start"], it["start"] + it["length"], it["label"], 1)
        t = res.tallies[0].numpy()
        assert int(res.n_reads[0]) == it["n_placed"]
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        assert t[:, 2].sum() > 0


def test_full_size_config3_consistency(eng):
    """BASELINE.json configs[2] at FULL size -- the north star's headline shape: hg38-size genome, 30x, 24 barcodes with
    barcode_weights {"0": 10, "1": 5}, 10 k insertions of 5 kb per barcode (240 k targets), ~6.2 M reads per iteration.
    This shape takes the GENERAL kernel variant (global grids + per-warp candidate queues), which the toy-size tests
    only brush.  Size-independent properties, all bit-exact (counting.py:19-57, read_operations.py:94-103):
      * native tallies == level-1 replay of the materialised reads == the oracle's indexed CPU tally of those reads;
      * the materialised reads, label counts, N_bases and reads placed == the numpy restatement of the native stream
        (oracle/native_np.py) with the reference's cut-off rule;
      * an iteration is the same alone and inside a batch; the cut-off is tight."""
    from esloco_b200 import synthetic as S
    from esloco_b200.read_operations import barcode_cdf, barcode_probabilities
    from esloco_b200.utils import targets_to_arrays
    G, _ = S.chromosome_dir(S.HG38)
    B = 24
    keys, ts, te, tb = targets_to_arrays(S.random_insertions(G, 10000, 5000, B, seed=3), B)
    assert ts.size == 240_000 and np.array_equal(np.unique(tb), np.arange(B))
    cdf = barcode_cdf(barcode_probabilities(B, {"0": 10, "1": 5}))
    pool = S.lognormal_lengths(15000, 1, 1, seed=1)
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    eng.set_length_table(pool)
    cov, seed = 30, 20261019
    res = eng.run_batch(seed, 0, 40, 3, cov).cpu()
    assert eng.index_built_on_device()  # 240 k equal-length targets: the device builder's case
    thr = cov * G
    nb, nr = res.n_bases.numpy(), res.n_reads.numpy()
    assert (nb >= thr).all() and (nb - thr < pool.max()).all() and (res.status.numpy() == 0).all()
    assert np.array_equal(res.label_counts.numpy().sum(1), nr) and (res.label_counts[:, B:].numpy() == 0).all()
    solo = eng.run_batch(seed, 0, 41, 1, cov).cpu()
    assert np.array_equal(solo.tallies[0].numpy(), res.tallies[1].numpy()) and int(solo.n_bases[0]) == int(nb[1])
    for i in (1, 2):
        # the CPU restatement of the native stream: same reads, same cut-off, same label histogram
        it = NN.native_iteration(seed, 0, 40 + i, G, cov, pool, cdf)
        assert it["n_placed"] == int(nr[i]) and it["covered"] == int(nb[i])
        assert np.array_equal(np.bincount(it["label"], minlength=B), res.label_counts[i, :B].numpy())
        s, ln, lab = eng.materialize_reads(seed, 0, 40 + i, False, int(nr[i]))
        sh, lh, labh = s.cpu().numpy(), ln.cpu().numpy(), lab.cpu().numpy()
        assert np.array_equal(sh, it["start"]) and np.array_equal(lh, it["length"]) and np.array_equal(labh, it["label"])
        # the same reads through the other two implementations
        rep = eng.replay_reads(s, ln, lab, [0, int(nr[i])]).cpu()
        t = res.tallies[i].numpy()
        assert np.array_equal(rep.tallies[0].numpy(), t)
        assert np.array_equal(rep.label_counts[0].numpy(), res.label_counts[i].numpy())
        exp = O.count_matches_indexed(ts, te, tb, B, it["start"], it["start"] + it["length"], it["label"], 1)
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        # weighted barcodes: barcode 0 gets 10/37 of the reads, barcode 1 5/37, the others 1/37 each (README.md:204-210)
        p = np.diff(np.concatenate(([0.0], cdf)))
        depth = t[:, 2] / 5000.0
        for b in (0, 1, 7, 23):
            d = depth[tb == b]
            assert abs(d.mean() - cov * p[b]) < 5 * d.std(ddof=1) / np.sqrt(d.size), (b, d.mean(), cov * p[b])


@pytest.mark.parametrize("wide", [False, True])
def test_quantile_length_sampler_exact_and_close_to_the_pool(eng, wide):
    """ESL_SAMPLER_QUANTILE (the north star's inverse-CDF look-up in a shared-memory-staged table): the sorted pool and
    its 2048-bin quantile table are built on the device; 11 bits pick the bin, 13 bits interpolate, the top bin is an
    exact pick.  (1) Exact: reads, labels, cut-off and tallies equal the numpy restatement of the same rule, for 32-
    and 64-bit coordinates.  (2) Close to the reference's distribution (a uniform pick from the pool): Kolmogorov
    distance between the sampled lengths and the pool below 1/2048 + sampling noise, mean within 0.2 %, no read longer
    than the pool's longest.  (3) Switching back restores the pool pick bit for bit."""
    G = 6_000_000_000 if wide else 80_000_000
    rng = np.random.RandomState(12)
    T = 90
    ts = np.sort(rng.randint(0, G - 200_000, T).astype(np.int64) if not wide else (rng.rand(T) * (G - 200_000)).astype(np.int64))
    te = ts + rng.randint(500, 150_000, T)
    tb = (np.arange(T) % 3).astype(np.int32)
    cdf = np.array([0.5, 0.8, 1.0])
    pool = O.lognormal_pool(O.Rng(77), 9000, 1, 20, n=300000)
    cov = 0.4 if wide else 12.0
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    eng.set_length_table(pool)
    base = eng.run_batch(5, 1, 3, 1, cov).cpu()
    try:
        eng.set_length_sampler("quantile")
        res = eng.run_batch(5, 1, 3, 2, cov).cpu()
        it = NN.native_iteration(5, 1, 4, G, cov, pool, cdf, sampler="quantile")
        assert int(res.n_reads[1]) == it["n_placed"] and int(res.n_bases[1]) == it["covered"]
        exp = O.count_matches_indexed(ts, te, tb, 3, it["start"], it["start"] + it["length"], it["label"], 1)
        t = res.tallies[1].numpy()
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
        assert t[:, 2].sum() > 0
        s, ln, lab = eng.materialize_reads(5, 1, 4, False, int(res.n_reads[1]))
        lh = ln.cpu().numpy()
        assert np.array_equal(lh, it["length"]) and np.array_equal(s.cpu().numpy(), it["start"])
        # distribution of the sampled lengths against the pool
        assert lh.max() <= pool.max() and lh.min() >= pool.min()
        sp = np.sort(pool)
        ks = np.abs(np.searchsorted(sp, np.sort(lh), side="right") / sp.size - np.arange(1, lh.size + 1) / lh.size).max()
        assert ks < 1 / 2048 + 1.63 / np.sqrt(lh.size), ks  # table resolution + the 1 % point of the KS statistic
        assert abs(lh.mean() / pool.mean() - 1) < 0.002 + 4 * pool.std() / pool.mean() / np.sqrt(lh.size)
    finally:
        eng.set_length_sampler("pool")
    again = eng.run_batch(5, 1, 3, 1, cov).cpu()
    assert np.array_equal(again.tallies.numpy(), base.tallies.numpy()) and int(again.n_reads[0]) == int(base.n_reads[0])


@pytest.mark.parametrize("B", [6, 1300])
def test_device_index_builder_equals_host_builder(eng, B):
    """esl_prepare builds the target index of large, non-nesting sets on the device (radix sort by (barcode, start) +
    two parallel passes).  20 k targets over B barcodes (the last one without targets; 1300 barcodes take the bounds
    kernel through two chunks and the bulk kernel through descriptors in global memory), with targets that match no
    barcode, empty ones, ones starting beyond the genome end, exact duplicates, and a cluster of 300 short targets
    inside one grid bucket (the crowded-bucket bisect): device-built and host-built indexes give the same integers as
    the CPU restatement; a nested pair sends the device builder back to the host."""
    G = 400_000_000
    rng = np.random.RandomState(44)
    n = 20_000
    ts = np.sort(rng.randint(0, G - 3000, n)).astype(np.int64)
    te = ts + 2500
    tb = rng.randint(0, B - 1, n).astype(np.int32)  # the last barcode gets nothing
    tb[::97] = -1                               # matches no barcode
    te[5::211] = ts[5::211]                     # empty
    ts[-3:] = G + np.array([0, 10, 5000]); te[-3:] = ts[-3:] + 2500  # start at / beyond the genome end
    ts[1000], te[1000], tb[1000] = ts[999], te[999], tb[999]     # exact duplicate
    c0 = 123_456_789
    ts[2000:2300] = c0 + np.arange(300) * 30; te[2000:2300] = ts[2000:2300] + 50; tb[2000:2300] = 2  # 300 targets in 9 kb
    order = np.argsort(ts, kind="stable"); ts, te, tb = ts[order], te[order], tb[order]
    cdf = np.cumsum(rng.rand(B) + 0.2); cdf /= cdf[-1]; cdf[2] = max(cdf[2], 0.2); cdf = np.maximum.accumulate(cdf); cdf[-1] = 1.0
    pool = O.lognormal_pool(O.Rng(9), 6000, 1, 50, n=200000)
    cov = 3.0
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    eng.set_length_table(pool)
    it = NN.native_iteration(31, 2, 5, G, cov, pool, cdf)
    exp = O.count_matches_indexed(ts, te, tb, B, it["start"], it["start"] + it["length"], it["label"], 1)
    try:
        got = {}
        for mode in ("device", "host"):
            eng.set_index_builder(mode)
            res = eng.run_batch(31, 2, 5, 1, cov).cpu()
            assert eng.index_built_on_device() == (mode == "device")
            t = res.tallies[0].numpy()
            assert int(res.n_reads[0]) == it["n_placed"]
            assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"]), mode
            got[mode] = t
        assert got["device"][:, 2].sum() > 0 and got["device"][tb == 2, 2].sum() > 0
        # one nested target (it ends before its predecessor in the same barcode does): the device builder declines, the
        # result stays exact
        k = next(i for i in range(5000, n) if tb[i - 1] >= 0 and te[i - 1] > ts[i - 1] + 100 and ts[i] > ts[i - 1])
        ts2, te2, tb2 = ts.copy(), te.copy(), tb.copy()
        ts2[k] = ts[k - 1] + 1; te2[k] = ts2[k] + 10; tb2[k] = tb[k - 1]
        eng.set_index_builder("device")
        eng.set_targets(ts2, te2, tb2)
        res = eng.run_batch(31, 2, 5, 1, cov).cpu()
        assert not eng.index_built_on_device()
        exp2 = O.count_matches_indexed(ts2, te2, tb2, B, it["start"], it["start"] + it["length"], it["label"], 1)
        assert np.array_equal(res.tallies[0, :, 2].numpy(), exp2["otb"])
    finally:
        eng.set_index_builder("auto")


def random_geometry(seed):
    """A seeded random workload that mixes everything the index and the kernels special-case: clustered, nested,
    duplicate, empty, inverted and out-of-genome targets, unknown barcodes, hot (very long) targets, weighted barcodes,
    blocked intervals with weights 1 / fractional / 0 for one barcode or all, either coordinate width."""
    rng = np.random.RandomState(7000 + seed)
    wide = seed % 8 == 7
    G = 5_000_000_000 if wide else int(rng.choice([300_000, 2_000_000, 40_000_000]))
    B = int(rng.choice([1, 2, 5, 40]))
    weights = {str(int(k)): int(rng.randint(2, 30)) for k in rng.choice(B, size=rng.randint(0, min(B, 4) + 1), replace=False)} or None
    mean_len = int(rng.choice([300, 2000, 12000]))
    mean_len = min(mean_len, G // 200)
    pool = O.lognormal_pool(O.Rng(900 + seed), mean_len, 1, 10, n=50000)
    pool = pool[pool < G // 8]
    n_reads = int(rng.choice([30_000, 60_000, 110_000]))
    cov = n_reads * float(pool.mean()) / G
    T = int(rng.choice([0, 1, 40, 400, 2500]))
    ts = np.zeros(T, np.int64); ln = np.zeros(T, np.int64)
    if T:
        centres = rng.randint(0, G, 6, dtype=np.int64)
        clustered = rng.rand(T) < 0.4
        ts = np.where(clustered, centres[rng.randint(0, 6, T)] + (rng.randn(T) * 20 * mean_len).astype(np.int64),
                      (rng.rand(T) * G).astype(np.int64))
        kind = rng.choice(4, T, p=[0.25, 0.55, 0.15, 0.05])
        ln = np.select([kind == 0, kind == 1, kind == 2], [rng.randint(1, 50, T), rng.randint(mean_len // 4 + 1, 4 * mean_len + 2, T),
                                                          rng.randint(10 * mean_len, 60 * mean_len, T)], 0).astype(np.int64)
        n_hot = min(T, int(rng.choice([0, 0, 2, 5])))
        ln[:n_hot] = rng.randint(G // 400, G // 6, n_hot)  # long enough to be "hot" (>= G/512 - mean read length)
        for k in range(1, T, 7):  # nest a target inside its predecessor, duplicate another
            ts[k] = ts[k - 1] + ln[k - 1] // 4; ln[k] = max(1, ln[k - 1] // 3)
        for k in range(3, T, 11):
            ts[k] = ts[k - 2]; ln[k] = ln[k - 2]
        ts = np.clip(ts, 0, None)
        if T > 10:
            ln[5] = 0; ln[6] = -min(500, int(ts[6]))  # empty / inverted (coordinates stay >= 0: the engine rejects negative ones)
            ts[7] = G - 10; ts[8] = G + 1000  # sticking out of / beyond the genome
    te = ts + ln
    tb = rng.randint(0, B, T).astype(np.int32)
    if T > 10:
        tb[9] = -1  # a key whose second token is no barcode number
    K = int(rng.choice([0, 2, 12, 40]))
    mask = None
    if K:
        ms = (rng.rand(K) * G).astype(np.int64)
        me = ms + (rng.rand(K) * G / rng.choice([30, 200, 2000])).astype(np.int64) + 1
        mw = rng.choice([1.0, 0.35, 0.0], K, p=[0.6, 0.3, 0.1])
        mb = np.where(rng.rand(K) < 0.5, -1, rng.randint(0, B, K)).astype(np.int32)
        mask = (ms, me, mw, mb)
    if T >= 400:  # twenty 60-base targets of one barcode 3 bases apart: a crowded grid bucket (the bisect inside a bucket)
        ts[20:40] = G // 3 + 3 * np.arange(20); te[20:40] = ts[20:40] + 60; tb[20:40] = tb[20]
    if K >= 12:   # twelve blocked intervals of mixed weights inside 100 bases: crowded bucket + several uniforms per read
        ms[:12] = G // 2 + 8 * np.arange(12); me[:12] = ms[:12] + 5; mb[:12] = -1
        mw[:12] = np.tile([0.35, 0.0, 0.35, 1.0], 3)
    return dict(G=G, B=B, weights=weights, pool=pool, cov=cov, ts=ts, te=te, tb=tb, mask=mask, consuming=bool(rng.rand() < 0.5),
                min_overlap=int(rng.choice([1, 25, max(1, mean_len // 2)])), sampler=str(rng.choice(["pool", "quantile"])),
                scaling=float(rng.choice([1.0, 1.0, 0.6])) if T <= 400 else 1.0, seed=int(rng.randint(1, 2**31)),
                combo=int(rng.randint(0, 5)), it0=int(rng.randint(0, 1000)))


def fuzz_seeds(default):
    """Seeds of the random-workload tests; ESL_FUZZ_SEEDS=lo:hi widens the sweep (e.g. 24:300 for a one-off hunt)."""
    import os
    lo, hi = (int(x) for x in os.environ.get("ESL_FUZZ_SEEDS", f"0:{default}").split(":"))
    return range(lo, hi)


@pytest.mark.parametrize("seed", fuzz_seeds(24))
def test_native_random_geometries_exact(eng, seed):
    """Seeded random workloads (random_geometry) through the native path, six iterations each: reads placed, N_bases,
    label counts and every tally bit-exact against the numpy restatement of the draws + the oracle's faithful tally,
    including pair thinning (scaling < 1) recomputed from the same counter-based uniforms."""
    w = random_geometry(seed)
    G, B, ts, te, tb = w["G"], w["B"], w["ts"], w["te"], w["tb"]
    cdf = O.barcode_cdf(O.barcode_probabilities(B, w["weights"]))
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb)
    eng.set_blocked(*(w["mask"] or ()))
    eng.set_length_table(w["pool"])
    try:
        eng.set_length_sampler(w["sampler"])
        n_iter = 6  # more iterations than CTA groups: the first and the last against the CPU, the others against solo runs
        res_dev = eng.run_batch(w["seed"], w["combo"], w["it0"], n_iter, w["cov"], w["consuming"], w["min_overlap"], w["scaling"])
        res = res_dev.cpu()
        if ts.size:  # the moment kernel (operand of the multi-GPU reduce) on the same tallies, hot targets' big squares included
            acc = eng.moments(res_dev.tallies).cpu().numpy()
            tt = res_dev.tallies.cpu().numpy().astype(object)
            assert (acc[:, 0] == n_iter).all()
            for col, k in ((1, 0), (3, 1), (5, 2)):
                assert np.array_equal(acc[:, col], tt[:, :, k].sum(0)), (seed, col)
            assert np.array_equal(acc[:, 2], (tt[:, :, 0] ** 2).sum(0)) and np.array_equal(acc[:, 4], (tt[:, :, 1] ** 2).sum(0))
            assert [int(h) * 2**32 + int(l) for h, l in zip(acc[:, 7], acc[:, 6])] == [int(x) for x in (tt[:, :, 2] ** 2).sum(0)]
        for i in range(1, n_iter - 1):
            solo = eng.run_batch(w["seed"], w["combo"], w["it0"] + i, 1, w["cov"], w["consuming"], w["min_overlap"], w["scaling"]).cpu()
            assert np.array_equal(solo.tallies[0].numpy(), res.tallies[i].numpy()), (seed, i)
            assert np.array_equal(solo.label_counts[0].numpy(), res.label_counts[i].numpy()), (seed, i)
            assert int(solo.n_reads[0]) == int(res.n_reads[i]) and int(solo.n_bases[0]) == int(res.n_bases[i]), (seed, i)
    finally:
        eng.set_length_sampler("pool")
    assert not (res.status.numpy() & 2).any()
    n_draws = int(w["cov"] * G / w["pool"].mean() * 4) + 20000
    for i in (0, n_iter - 1):
        it = NN.native_iteration(w["seed"], w["combo"], w["it0"] + i, G, w["cov"], w["pool"], cdf, consuming=w["consuming"],
                                 mask=w["mask"], n_draws=n_draws, sampler=w["sampler"])
        rs, re_, lab = it["start"], it["start"] + it["length"], it["label"]
        counts, _ = O.label_hist(lab, B)
        assert int(res.n_reads[i]) == it["n_placed"] and int(res.n_bases[i]) == it["covered"], seed
        assert np.array_equal(res.label_counts[i].numpy(), counts), seed
        t = res.tallies[i].numpy()
        pairs = []  # expected hit records of this iteration: (target row, overlap), target-major, reads in draw order
        if w["scaling"] >= 1.0:
            exp = O.count_matches(None, ts, te, tb, rs, re_, lab, 1.0, w["min_overlap"], want_overlaps=True)
            pairs = np.stack([np.repeat(np.arange(ts.size), np.diff(exp["overlap_off"])), exp["overlap_flat"]], 1)
            exp = np.stack([exp["full"], exp["partial"], exp["otb"]], 1)
        else:  # the kept pairs from the same per-pair uniforms (Philox block (draw, target row), slot SCALE)
            exp = np.zeros((ts.size, 3), np.int64)
            k0, k1 = w["seed"] & 0xFFFFFFFF, (w["seed"] >> 32) & 0xFFFFFFFF
            for row in range(ts.size):
                s, e = int(ts[row]), int(te[row])
                ov = np.minimum(e, re_) - np.maximum(s, rs)
                # (tb = -1 is a key without a barcode number: it matches no read, least of all the blocked ones labelled -1)
                n = np.nonzero((tb[row] >= 0) & (lab == tb[row]) & (np.maximum(s, rs) < np.minimum(e, re_)) & (ov >= w["min_overlap"]))[0]
                if n.size == 0:
                    continue
                x, y, _, _ = NN.philox4x32_10(n.astype(np.uint64), np.uint64(row), w["it0"] + i, (w["combo"] << 4) | NN.SLOT_SCALE, k0, k1)
                n = n[NN.u53(x, y) <= w["scaling"]]
                full = (rs[n] <= s) & (e <= re_[n])
                exp[row] = [int(full.sum()), int((~full).sum()), int(ov[n].sum())]
                pairs.append(np.stack([np.full(n.size, row), ov[n]], 1))
            pairs = np.concatenate(pairs) if pairs else np.zeros((0, 2), np.int64)
        assert np.array_equal(t, exp), (seed, np.nonzero((t != exp).any(1))[0][:10])
        if i and len(pairs) <= 400_000:
            # the per-target `overlap` lists (counting.py:56) from device hit records of a solo run of this iteration; the
            # buffer is a little short of the bulk kernel's overshoot on purpose when there are many pairs: overflow must be
            # reported, not written past the end
            try:
                eng.set_length_sampler(w["sampler"])
                eng.set_hit_buffer(len(pairs) + 64)
                solo = eng.run_batch(w["seed"], w["combo"], w["it0"] + i, 1, w["cov"], w["consuming"], w["min_overlap"], w["scaling"])
                try:
                    hits = eng.take_hits(solo.n_reads)
                except Exception as e:  # pairs of draws beyond the cut-off did not fit: once more with room for them
                    assert getattr(e, "code", None) == -5, e
                    eng.set_hit_buffer(4 * len(pairs) + 100_000)
                    solo = eng.run_batch(w["seed"], w["combo"], w["it0"] + i, 1, w["cov"], w["consuming"], w["min_overlap"], w["scaling"])
                    hits = eng.take_hits(solo.n_reads)
            finally:
                eng.set_hit_buffer(0)
                eng.set_length_sampler("pool")
            assert np.array_equal(solo.tallies[0].cpu().numpy(), t), seed  # recording hits does not change the tallies
            assert np.array_equal(hits[:, 1], pairs[:, 0]) and np.array_equal(hits[:, 3], pairs[:, 1]), seed


@pytest.mark.parametrize("B,owner", [(40, 20), (16, 15), (5, 3)])
def test_barcodes_without_targets_before_one_with(eng, B, owner):
    """Regression: barcodes without targets get two reject-all grid entries behind the real grids; their space must not
    shift the real grids' offsets (it once did, and the reject-all entries then overwrote the top buckets of the last
    real grid: a lone target of barcode 20 of 40 was never hit).  Native and replay paths, host index builder."""
    G = 2_000_000
    ts, te, tb = np.array([1_582_638, 40_000]), np.array([1_792_432, 47_000]), np.array([owner, owner], np.int32)
    cdf = O.barcode_cdf(O.barcode_probabilities(B, None))
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    pool = O.lognormal_pool(O.Rng(3), 2000, 1, 10, n=50000)
    eng.set_length_table(pool)
    res = eng.run_batch(11, 0, 0, 1, 100.0).cpu()
    it = NN.native_iteration(11, 0, 0, G, 100.0, pool, cdf)
    exp = O.count_matches(None, ts, te, tb, it["start"], it["start"] + it["length"], it["label"], 1.0, 1)
    t = res.tallies[0].numpy()
    assert int(res.n_reads[0]) == it["n_placed"]
    assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"])
    assert t[0, 2] > 0 and t[1, 2] > 0
    rep = eng.replay_reads(it["start"], it["length"], it["label"], [0, it["n_placed"]]).cpu()
    assert np.array_equal(rep.tallies[0].numpy(), t)


@pytest.mark.parametrize("seed", fuzz_seeds(16))
def test_replay_random_geometries_exact(eng, seed):
    """The same random workloads through the REFERENCE stream: the oracle (MT19937 port, pinned by the goldens) runs
    process_combination on them and records its raw draws; the device does CDF look-up, blocked test with the recorded
    uniforms, cut-off and tally from those draws (level-2 replay) with the bulk / cut-off split in three places.
    Everything the reference reports must come out bit-exact."""
    w = random_geometry(seed)
    G, B, ts, te, tb = w["G"], w["B"], w["ts"], w["te"], w["tb"]
    mrl = max(20, int(w["pool"].mean()))
    rng = O.Rng(3000 + seed)
    try:
        out = O.process_combination(rng, mrl, w["cov"], G, ts, te, tb, B, w["weights"], w["consuming"], w["mask"], 1, 10, 1.0,
                                    w["min_overlap"], n_extra=4000, record_draws=True)
    except ValueError:  # (seen once in a 1000-seed sweep, at G = 300 kb)
        pytest.skip("the 1 M-entry log-normal pool holds a length >= the genome size: the reference raises here too")
    pl = out["placement"]
    n_placed, n_total = pl["n_placed"], pl["start"].size
    eng.set_genome(G); eng.set_barcode_cdf(O.barcode_cdf(O.barcode_probabilities(B, w["weights"]))); eng.set_targets(ts, te, tb)
    eng.set_blocked(*(w["mask"] or ()))
    for bulk in (0, n_placed // 2, n_total):
        res = eng.replay_draws(pl["u_barcode"], pl["length"], pl["start"], [0, n_total], w["cov"], consuming=w["consuming"],
                               min_overlap=w["min_overlap"], ublock_off=pl["ublock_off"] if w["mask"] is not None else None,
                               ublock=pl["ublock"] if w["mask"] is not None else None, bulk_draws=bulk).cpu()
        assert not int(res.status[0]) & 2
        assert int(res.n_reads[0]) == n_placed and int(res.n_bases[0]) == pl["covered_length"], (seed, bulk)
        assert np.array_equal(res.label_counts[0].numpy(), out["label_counts"]), (seed, bulk)
        t = res.tallies[0].numpy()
        assert np.array_equal(t[:, 0], out["full"]) and np.array_equal(t[:, 1], out["partial"]) and np.array_equal(t[:, 2], out["otb"]), (seed, bulk)


@pytest.mark.parametrize("seed", range(10))
def test_device_index_builder_random_flat_sets(eng, seed):
    """Random non-nesting target sets (one target length per case, so ends ascend with starts) forced through the
    device index builder and through the host builder: 1-300 barcodes of which a random subset owns all targets (empty
    barcodes in front of, between and behind the used ones), overlapping and duplicate targets, unknown barcodes,
    targets at / beyond the genome end, 1-40 k targets.  Both indexes give the CPU restatement's integers."""
    rng = np.random.RandomState((900 if seed < 5 else 1300) + seed)
    G = int(rng.choice([3_000_000, 80_000_000, 900_000_000]))
    B = int(rng.choice([1, 7, 64, 300]))
    used = rng.choice(B, size=max(1, int(B * rng.choice([0.1, 0.5, 1.0]))), replace=False)
    n = int(rng.choice([1, 50, 3000, 40_000]))
    tlen = int(rng.choice([40, 3000, 30_000]))
    ts = (rng.rand(n) * G).astype(np.int64)
    if n > 20:
        ts[3] = ts[2]; ts[10] = G - 5; ts[11] = G; ts[12] = G + 999
    te = ts + tlen
    tb = used[rng.randint(0, used.size, n)].astype(np.int32)
    if n > 20:
        tb[3] = tb[2]; tb[15] = -1
    cdf = np.cumsum(rng.rand(B) + 0.05); cdf /= cdf[-1]; cdf[-1] = 1.0
    mean_len = int(min(rng.choice([500, 8000]), G // 300))
    pool = O.lognormal_pool(O.Rng(seed), mean_len, 1, 10, n=50000)
    cov = 80_000 * float(pool.mean()) / G
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb); eng.set_blocked()
    eng.set_length_table(pool)
    it = NN.native_iteration(77, 1, seed, G, cov, pool, cdf)
    exp = O.count_matches_indexed(ts, te, tb, B, it["start"], it["start"] + it["length"], it["label"], 1)
    hot = tlen >= G / 512 - pool.mean()  # hot targets are the host builder's business
    try:
        for mode in ("device", "host"):
            eng.set_index_builder(mode)
            res = eng.run_batch(77, 1, seed, 1, cov).cpu()
            assert eng.index_built_on_device() == (mode == "device" and not hot), (seed, mode)
            t = res.tallies[0].numpy()
            assert int(res.n_reads[0]) == it["n_placed"]
            assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"]), (seed, mode)
    finally:
        eng.set_index_builder("auto")


def extreme_geometry(seed):
    """Seeded workloads at the edges of the parameter space: genomes of 2 kb and right below / above the 32-bit
    coordinate limit, a length pool with one entry or with reads nearly as long as the genome, a handful of reads per
    iteration (cut-off kernel only), up to 1300 barcodes with weight ratios of 10^4, no targets / tens of thousands of
    them (device index builder) in one barcode or spread, a blocked interval over the whole genome for one barcode."""
    rng = np.random.RandomState(91000 + seed)
    G = int(rng.choice([2_000, 50_000, 3_000_000, 4_294_967_000, 4_294_967_400, 9_000_000_000]))
    B = int(rng.choice([1, 3, 257, 1300]))
    heavy = rng.choice(B, size=min(B, int(rng.choice([0, 1, 3]))), replace=False)
    weights = {str(int(k)): int(rng.choice([2, 100, 10_000])) for k in heavy} or None
    kind = int(rng.choice(4))
    if kind == 0:
        pool = np.array([int(rng.choice([1, 7, max(2, G // 1000)]))], np.int64)          # one length
    elif kind == 1:
        pool = np.array([1, max(2, G // 3), max(3, G - 1)], np.int64)                    # 1 base ... the whole genome but one base
        pool = np.minimum(pool, 0xFFFFFFFF)                                              # (lengths are 32-bit)
    else:
        pool = O.lognormal_pool(O.Rng(seed), max(30, G // int(rng.choice([100, 5000, 200_000]))), 1, 1, n=20000)
        pool = pool[pool < min(G, 1 << 32)]  # (the reference raises for a read as long as the genome; lengths are 32-bit)
    n_reads = int(rng.choice([4, 300, 40_000]))
    cov = max(n_reads * float(pool.mean()) / G, 1e-9)
    T = int(rng.choice([0, 3, 500, 20_000, 70_000]))
    tlen = int(rng.choice([1, max(1, int(pool.mean())), max(2, G // 50)]))
    ts = (rng.rand(T) * G).astype(np.int64)
    te = ts + tlen                                                                       # one length: no nesting, ends may pass G
    tb = (np.full(T, rng.randint(0, B)) if rng.rand() < 0.3 else rng.randint(0, B, T)).astype(np.int32)
    mask = None
    mk = int(rng.choice(3))
    if mk == 1 and B > 1:   # every read of one barcode is blocked
        mask = (np.array([0], np.int64), np.array([G], np.int64), np.array([1.0]), np.array([int(tb[0]) if T else 0], np.int32))
    elif mk == 2:           # a third of the genome blocked for everybody with probability 1/2, plus one base
        mask = (np.array([G // 3, G - 1], np.int64), np.array([2 * (G // 3), G], np.int64), np.array([0.5, 1.0]), np.array([-1, -1], np.int32))
    consuming = bool(rng.rand() < 0.5)
    if mk == 1 and B > 1 and weights and str(int(mask[3][0])) in weights:
        consuming = True  # (nearly every read blocked AND not counted would need 10^4 times the draws)
    return dict(G=G, B=B, weights=weights, pool=pool, cov=cov, ts=ts, te=te, tb=tb, mask=mask, consuming=consuming,
                min_overlap=int(rng.choice([1, 2, max(1, tlen)])), sampler=str(rng.choice(["pool", "quantile"])),
                seed=int(rng.randint(1, 2**31)) | (int(rng.randint(0, 2**31)) << 32), combo=int(rng.randint(0, 16)),
                it0=int(rng.choice([0, 999_999, 2**31 - 10])))


@pytest.mark.parametrize("seed", fuzz_seeds(16))
def test_native_extreme_geometries_exact(eng, seed):
    """extreme_geometry through the native path, three iterations per batch: everything the engine returns bit-exact
    against the numpy restatement + the oracle's indexed tally."""
    w = extreme_geometry(seed)
    G, B, ts, te, tb = w["G"], w["B"], w["ts"], w["te"], w["tb"]
    cdf = O.barcode_cdf(O.barcode_probabilities(B, w["weights"]))
    eng.set_genome(G); eng.set_barcode_cdf(cdf); eng.set_targets(ts, te, tb)
    eng.set_blocked(*(w["mask"] or ()))
    eng.set_length_table(w["pool"])
    try:
        eng.set_length_sampler(w["sampler"])
        res = eng.run_batch(w["seed"], w["combo"], w["it0"], 3, w["cov"], w["consuming"], w["min_overlap"], 1.0).cpu()
    finally:
        eng.set_length_sampler("pool")
    assert not (res.status.numpy() & 2).any()
    n_draws = int(w["cov"] * G / w["pool"].mean() * 6) + 20000
    for i in (0, 2):
        it = NN.native_iteration(w["seed"], w["combo"], w["it0"] + i, G, w["cov"], w["pool"], cdf, consuming=w["consuming"],
                                 mask=w["mask"], n_draws=n_draws, sampler=w["sampler"])
        counts, _ = O.label_hist(it["label"], B)
        assert int(res.n_reads[i]) == it["n_placed"] and int(res.n_bases[i]) == it["covered"], seed
        assert np.array_equal(res.label_counts[i].numpy(), counts), seed
        exp = O.count_matches_indexed(ts, te, tb, B, it["start"], it["start"] + it["length"], it["label"], w["min_overlap"])
        t = res.tallies[i].numpy()
        assert np.array_equal(t[:, 0], exp["full"]) and np.array_equal(t[:, 1], exp["partial"]) and np.array_equal(t[:, 2], exp["otb"]), seed


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_engine_state_sequences_equal_fresh_context(eng, seed):
    """One context driven through a random sequence of set-up calls (genome size across the 32 / 64-bit coordinate
    switch, targets between 0 and 30 k across the device-builder threshold and the shared pre-filter limit, blocked
    intervals, barcode tables, length tables with longer and shorter reads, sampler, index builder) gives, after every
    change, exactly what a context created for those inputs gives: no set-up call leaves anything stale behind."""
    from esloco_b200.engine import Engine
    rng = np.random.RandomState(600 + seed)
    st = dict(G=30_000_000, cdf=np.array([1.0]), ts=np.zeros(0, np.int64), te=np.zeros(0, np.int64), tb=np.zeros(0, np.int32),
              mask=(), pool=np.array([2000, 3000]), sampler="pool", builder="auto")

    def apply(e, only=None):
        for k in (only or ["G", "cdf", "targets", "mask", "pool", "sampler", "builder"]):
            if k == "G": e.set_genome(st["G"])
            elif k == "cdf": e.set_barcode_cdf(st["cdf"])
            elif k == "targets": e.set_targets(st["ts"], st["te"], st["tb"])
            elif k == "mask": e.set_blocked(*st["mask"])
            elif k == "pool": e.set_length_table(st["pool"])
            elif k == "sampler": e.set_length_sampler(st["sampler"])
            elif k == "builder": e.set_index_builder(st["builder"])

    def new_targets():
        B, G = st["cdf"].size, st["G"]
        n = int(rng.choice([0, 5, 300, 30_000]))
        tlen = int(rng.choice([200, 5000, G // 300]))
        st["ts"] = (rng.rand(n) * G).astype(np.int64)
        st["te"] = st["ts"] + (tlen if rng.rand() < 0.5 else rng.randint(1, 2 * tlen + 2, n))
        st["tb"] = rng.randint(0, B, n).astype(np.int32)

    try:
        apply(eng)
        for step in range(12):
            changed = []
            for what in rng.choice(["G", "cdf", "targets", "mask", "pool", "sampler", "builder"], size=int(rng.choice([1, 2, 3])), replace=False):
                if what == "G":
                    st["G"] = int(rng.choice([30_000_000, 400_000_000, 4_294_967_200, 4_294_967_400, 7_000_000_000]))
                elif what == "cdf":
                    B = int(rng.choice([1, 2, 40, 1100]))
                    c = np.cumsum(rng.rand(B) + 0.1); st["cdf"] = c / c[-1]; st["cdf"][-1] = 1.0
                    new_targets(); changed.append("targets")
                    st["mask"] = (); changed.append("mask")
                elif what == "targets":
                    new_targets()
                elif what == "mask":
                    k = int(rng.choice([0, 1, 6]))
                    ms = (rng.rand(k) * st["G"]).astype(np.int64)
                    st["mask"] = (ms, ms + rng.randint(1, st["G"] // 20, k), rng.choice([1.0, 0.4], k),
                                  rng.randint(-1, st["cdf"].size, k).astype(np.int32)) if k else ()
                elif what == "pool":
                    st["pool"] = O.lognormal_pool(O.Rng(int(rng.randint(1 << 30))), int(rng.choice([800, 20_000, 300_000])), 1, 5, n=20_000)
                elif what == "sampler":
                    st["sampler"] = "quantile" if st["sampler"] == "pool" else "pool"
                elif what == "builder":
                    st["builder"] = str(rng.choice(["auto", "host", "device"]))
                changed.append(what)
            # (targets / mask refer to genome and barcode count: re-send them when those change, as a caller would)
            if "G" in changed or "cdf" in changed:
                st["ts"], st["te"] = np.minimum(st["ts"], st["G"] - 1), np.minimum(st["te"], st["G"])
                if st["mask"]:
                    st["mask"] = (np.minimum(st["mask"][0], st["G"] - 1), np.minimum(st["mask"][1], st["G"]), st["mask"][2], st["mask"][3])
                changed += ["targets", "mask"]
            apply(eng, list(dict.fromkeys(changed)))
            cov = 30_000 * float(st["pool"].mean()) / st["G"]
            got = eng.run_batch(77, 1, step, 2, cov, bool(step & 1), 1, 1.0).cpu()
            fresh = Engine(0)
            try:
                apply(fresh)
                want = fresh.run_batch(77, 1, step, 2, cov, bool(step & 1), 1, 1.0).cpu()
            finally:
                fresh.close()
            for f in ("tallies", "label_counts", "n_bases", "n_reads", "status"):
                assert np.array_equal(getattr(got, f).numpy(), getattr(want, f).numpy()), (seed, step, f, changed)
    finally:
        eng.set_length_sampler("pool"); eng.set_index_builder("auto")
