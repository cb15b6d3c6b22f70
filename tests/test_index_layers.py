"""Host logic of the target index (no GPU): how esl_plan_target_layers splits a barcode's targets into layers, checked
against the rules DESIGN.md section 3 states and against a brute-force overlap search that walks the layers the way
the device does (first record whose running max of ends passes the read start, then forward while starts < read end)."""
import ctypes as C

import numpy as np
import pytest

from esloco_b200 import _native
from esloco_b200 import synthetic as S

K_DEAD = 16  # kDeadScan in csrc/esl_api.cu


def plan(G, ts, te):
    lib = _native.load()
    ts = np.ascontiguousarray(ts, np.int64); te = np.ascontiguousarray(te, np.int64)
    out = np.full(ts.size, -7, np.int32)
    n = lib.esl_plan_target_layers(int(G), ts.ctypes.data, te.ctypes.data, ts.size, out.ctypes.data)
    assert n >= 0
    return n, out


def layer_arrays(ts, te, lay, l):
    idx = np.nonzero(lay == l)[0]
    o = np.lexsort((idx, te[idx], ts[idx]))
    return ts[idx][o], te[idx][o], idx[o]


def check_rules(G, ts, te, n, lay):
    assert ((lay >= 0) == (te > ts)).all() and lay.max(initial=-1) == n - 1
    for l in range(n):
        s, e, _ = layer_arrays(ts, te, lay, l)
        assert s.size > 0 and (np.diff(s) >= 0).all()
        pmax = np.maximum.accumulate(e)
        dead = np.concatenate(([0], np.maximum(pmax[:-1] - e[1:], 0)))
        assert dead.sum() <= 0.25 * G
        if l >= 64:
            assert (np.diff(e) >= 0).all()  # the strict layers behind the relaxed ones
        else:  # a nested target lies at most K_DEAD records behind the oldest record that outlives it
            nested = np.nonzero(e[1:] < pmax[:-1])[0] + 1
            oldest = np.searchsorted(pmax, e[nested], side="left")  # first record whose running max reaches its end
            assert (nested - oldest <= K_DEAD).all()


def walk_overlaps(ts, te, lay, n, rs, re):
    """Rows a device walk over the layers would count for the read [rs, re)."""
    rows = []
    for l in range(n):
        s, e, row = layer_arrays(ts, te, lay, l)
        pmax = np.maximum.accumulate(e)
        j = int(np.searchsorted(pmax, rs, side="right"))
        while j < s.size and s[j] < re:
            if e[j] > rs:
                rows.append(int(row[j]))
            j += 1
    return sorted(rows)


def test_plain_shapes():
    G = 1_000_000
    n, lay = plan(G, [10, 500, 900, 40], [20, 800, 950, 40])  # one empty target
    assert n == 1 and list(lay) == [0, 0, 0, -1]
    n, lay = plan(G, [], [])
    assert n == 0
    rng = np.random.RandomState(0)
    s = np.sort(rng.randint(0, G - 5000, 3000))
    n, lay = plan(G, s, s + 5000)  # equal lengths never nest (I mode)
    assert n == 1
    assert _native.load().esl_plan_target_layers(0, None, None, 0, None) < 0


def test_config4_rois_fit_one_layer():
    from esloco_b200.utils import targets_to_arrays
    G, _ = S.chromosome_dir(S.HG38)
    keys, ts, te, tb = targets_to_arrays(S.barcode_rois(S.random_rois(S.HG38, 50000, 1000, 100000, seed=4), 1), 1)
    n, lay = plan(G, ts, te)
    assert n == 1
    check_rules(G, ts, te, n, lay)


def test_whole_chromosomes_over_small_rois_need_two_layers():
    G, cdir = S.chromosome_dir(S.HG38)
    rng = np.random.RandomState(1)
    ss = np.sort(rng.randint(0, G - 5000, 20000)); se = ss + rng.randint(500, 5000, 20000)
    ts = np.concatenate([[v[0] for v in cdir.values()], ss]); te = np.concatenate([[v[1] for v in cdir.values()], se])
    n, lay = plan(G, ts, te)
    assert n == 2
    check_rules(G, ts, te, n, lay)
    assert (lay[:24] == 0).all() and (lay[24:] == 1).mean() > 0.98


@pytest.mark.parametrize("depth,layers", [(200, 12), (5000, 64 + 5000 - 64 * 17)])
def test_mutual_nesting(depth, layers):
    G = 50_000_000
    ts = np.arange(depth) * 1000 + 5; te = G - np.arange(depth) * 1000 - 7
    n, lay = plan(G, ts, te)
    assert n == layers
    check_rules(G, ts, te, n, lay)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_walk_finds_exactly_the_overlapping_targets(seed):
    """Dense adversarial geometry: nested, touching, duplicate, 1-bp and genome-spanning targets."""
    rng = np.random.RandomState(seed)
    G = 200_000
    ts = rng.randint(0, G, 600); te = ts + rng.choice([1, 2, 50, 500, 5000, 60000], 600)
    ts = np.concatenate([ts, [0, 0, 1000, 1000, 1000, G - 1]]); te = np.concatenate([te, [G, G, 3000, 3000, 1001, G + 10]])
    n, lay = plan(G, ts, te)
    check_rules(G, ts, te, n, lay)
    for _ in range(400):
        rs = int(rng.randint(0, G - 1)); re = rs + int(rng.choice([1, 30, 900, 20000]))
        brute = sorted(np.nonzero((ts < re) & (te > rs) & (te > ts))[0].tolist())
        assert walk_overlaps(ts, te, lay, n, rs, re) == brute
