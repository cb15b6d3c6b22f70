"""Pin the oracle: from the numpy seed alone it must reproduce every output the unmodified reference
produced for the committed fixtures (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from golden_util import Golden, case_names
from oracle import oracle as O

CASES = case_names()


def test_fixtures_present():
    assert len(CASES) >= 12


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_from_seed(name):
    g = Golden(name)
    m = g.meta
    rng = O.Rng(m["seed"])
    tb = O.target_barcodes(m["target_keys"])
    for ci, (mrl, cov) in enumerate(m["combinations"]):
        out = O.process_combination(rng, mrl, cov, m["genome_size"], g["target_start"], g["target_end"], tb,
                                    m["n_barcodes"], m["barcode_weights"], m["consuming"], g.mask, m["sigma"],
                                    m["min_read_length"], m["scaling"], m["min_overlap"],
                                    source_lengths=g.source_lengths, want_overlaps=True)
        gm = m["combos"][ci]
        pl = out["placement"]
        n = pl["n_placed"]
        assert n == g.combo(ci, "read_start").size
        assert np.array_equal(pl["start"][:n], g.combo(ci, "read_start"))
        assert np.array_equal(pl["start"][:n] + pl["length"][:n], g.combo(ci, "read_end"))
        assert np.array_equal(pl["label"][:n], g.combo(ci, "read_label").astype(np.int32))
        assert pl["covered_length"] == gm["covered_length"]
        if gm["pool_len"] >= 0:  # log-normal source pool (read_operations.py:8-27)
            assert out["source"].size == gm["pool_len"]
            assert int(out["source"].sum()) == gm["pool_sum"]
            assert np.array_equal(out["source"][:256], g.combo(ci, "pool_head"))
        assert out["label_order"] == gm["label_order"]
        names = O.label_names(m["n_barcodes"])
        counts = [int(out["label_counts"][names.index(k)]) for k in gm["label_order"]]
        assert counts == g.combo(ci, "label_counts").tolist()
        assert np.array_equal(out["full"], g.combo(ci, "full"))
        assert np.array_equal(out["partial"], g.combo(ci, "partial"))
        assert np.array_equal(out["otb"], g.combo(ci, "otb"))
        assert np.array_equal(out["overlap_off"], g.combo(ci, "overlap_off"))
        assert np.array_equal(out["overlap_flat"], g.combo(ci, "overlap_flat"))
        assert [k + "_" + str(m["iteration"]) for k in m["target_keys"]] == gm["target_region"]


@pytest.mark.parametrize("name", [c for c in CASES if c != "scaling_half"])
def test_indexed_tally_equals_faithful_loop(name):
    g = Golden(name)
    m = g.meta
    tb = O.target_barcodes(m["target_keys"])
    for ci in range(len(m["combinations"])):
        rs, re_, rl = g.combo(ci, "read_start"), g.combo(ci, "read_end"), g.combo(ci, "read_label").astype(np.int32)
        a = O.count_matches(None, g["target_start"], g["target_end"], tb, rs, re_, rl, 1.0, m["min_overlap"])
        b = O.count_matches_indexed(g["target_start"], g["target_end"], tb, m["n_barcodes"], rs, re_, rl, m["min_overlap"])
        for k in ("full", "partial", "otb"):
            assert np.array_equal(a[k], b[k])
            assert np.array_equal(a[k], g.combo(ci, k))
