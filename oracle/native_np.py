"""numpy restatement of the engine's NATIVE (Philox) draw stream -- TEST INFRASTRUCTURE ONLY.

The native mode has no counterpart stream in the reference (which draws from numpy's MT19937), so besides the
statistical check the tests pin it exactly: this module regenerates on the CPU the draws the CUDA kernels make
(esloco_b200/csrc/philox.cuh, kernels.cuh PhiloxSrc) and applies the reference's placement rules
(read_operations.py:100-121) to them; the oracle's tally of those reads must equal the GPU tallies bit for bit.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
SLOT_PLACE, SLOT_BLOCK, SLOT_SCALE = 0, 1, 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; counters are uint64 arrays holding 32-bit values, keys python ints."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 &= 0xFFFFFFFF; k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1 = p1 & MASK
        c3 = p0 & MASK
        c0, c2 = n0, n2
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u53(a, b):
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def mulhi64(a, r):
    """High 64 bits of the 128-bit product of uint64 arrays a and r."""
    a0, a1 = a & MASK, a >> np.uint64(32)
    r0, r1 = r & MASK, r >> np.uint64(32)
    p00, p01, p10, p11 = a0 * r0, a0 * r1, a1 * r0, a1 * r1
    mid = (p00 >> np.uint64(32)) + (p01 & MASK) + (p10 & MASK)
    return p11 + (p01 >> np.uint64(32)) + (p10 >> np.uint64(32)) + (mid >> np.uint64(32))


def cdf_thresholds(cdf):
    v = np.ceil(np.asarray(cdf, np.float64) * 4294967296.0)
    return np.where(v >= 4294967295.0, 0xFFFFFFFF, v).astype(np.uint64)


QUANT_BINS = 2048


def quantile_table(pool):
    """(sorted pool, quantile table q[0..QUANT_BINS]) as esl_prepare builds them for ESL_SAMPLER_QUANTILE:
    q[k] = sorted[k * n // QUANT_BINS], q[QUANT_BINS] = the longest read."""
    sp = np.sort(np.asarray(pool, np.uint64))
    n = sp.size
    q = np.empty(QUANT_BINS + 1, np.uint64)
    q[:QUANT_BINS] = sp[(np.arange(QUANT_BINS, dtype=np.uint64) * np.uint64(n)) // np.uint64(QUANT_BINS)]
    q[QUANT_BINS] = sp[-1]
    return sp, q


def lengths_from_bits(u24, pool, sampler):
    """Read lengths from 24 uniform bits (kernels.cuh PhiloxSrc::length_of)."""
    u24 = np.asarray(u24, np.uint64)
    pool = np.asarray(pool, np.uint64)
    if sampler == "pool":
        return pool[((u24 * np.uint64(pool.size)) >> np.uint64(24)).astype(np.int64)]
    sp, q = quantile_table(pool)
    n = sp.size
    k, f = (u24 >> np.uint64(13)).astype(np.int64), u24 & np.uint64(0x1FFF)
    top_lo = (QUANT_BINS - 1) * n // QUANT_BINS
    top = sp[np.minimum(top_lo + ((f * np.uint64(n - top_lo)) >> np.uint64(13)).astype(np.int64), n - 1)]
    kk = np.minimum(k, QUANT_BINS - 1)
    inner = q[kk] + (((q[kk + 1] - q[kk]) * f) >> np.uint64(13))
    return np.where(k == QUANT_BINS - 1, top, inner)


def native_draws(seed, combo, iteration, n, genome_size, pool, cdf, sampler="pool"):
    """Draws 0..n-1 of one iteration: (length, start, barcode) as int64 arrays (kernels.cuh PhiloxSrc).
    32-bit coordinates (genome_size <= 2^32 - 16): draw n takes the 64 bits (a, b) = words (x, y) / (z, w) of block
    (n >> 1, 0) for even / odd n -- length = pool[(b >> 8) * n_pool >> 24], start = floor((a * 2^32 + (b & 0xff) * 2^24)
    * (G - L) / 2^64) -- and its barcode uniform is word n & 3 of block (n >> 2, 1).  64-bit coordinates: draw n owns
    block (n, 0): (x, y) = start, z = length pick, w = barcode."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    idx = np.arange(n, dtype=np.uint64)
    pool = np.asarray(pool, np.uint64)
    slot = (combo << 4) | SLOT_PLACE
    if genome_size <= 0xFFFFFFF0:
        x, y, z, w = philox4x32_10(idx >> np.uint64(1), 0, iteration, slot, k0, k1)
        odd = (idx & np.uint64(1)).astype(bool)
        a, b = np.where(odd, z, x), np.where(odd, w, y)
        L = lengths_from_bits(b >> np.uint64(8), pool, sampler)
        start = mulhi64((a << np.uint64(32)) | ((b & np.uint64(0xFF)) << np.uint64(24)), np.uint64(genome_size) - L)
        q = philox4x32_10(idx >> np.uint64(2), 1, iteration, slot, k0, k1)
        u = np.choose((idx & np.uint64(3)).astype(np.int64), q)
    else:
        x, y, z, u = philox4x32_10(idx, 0, iteration, slot, k0, k1)
        L = (pool[((z * np.uint64(pool.size)) >> np.uint64(32)).astype(np.int64)] if sampler == "pool"
             else lengths_from_bits(z >> np.uint64(8), pool, sampler))
        start = mulhi64((x << np.uint64(32)) | y, np.uint64(genome_size) - L)
    thr = cdf_thresholds(cdf)
    B = thr.size
    bc = np.minimum(np.searchsorted(thr[:B - 1], u, side="right"), B - 1) if B > 1 else np.zeros(n, np.int64)
    return L.astype(np.int64), start.astype(np.int64), bc.astype(np.int64)


def block_u(seed, combo, iteration, n, j):
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    x, y, z, w = philox4x32_10(np.uint64(n), np.uint64(j >> 1), iteration, (combo << 4) | SLOT_BLOCK, k0, k1)
    return float(u53(z, w)) if (j & 1) else float(u53(x, y))


def scale_u(seed, combo, iteration, n, row):
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    x, y, _, _ = philox4x32_10(np.uint64(n), np.uint64(row), iteration, (combo << 4) | SLOT_SCALE, k0, k1)
    return float(u53(x, y))


def native_iteration(seed, combo, iteration, genome_size, coverage, pool, cdf, consuming=False, mask=None, n_draws=None,
                     sampler="pool"):
    """Reads of one native iteration after blocked test and coverage cut-off (reference rules on native draws).
    mask = (start, end, weight, barcode) arrays, barcode -1 = ALL.  Returns dict(start, length, label, covered)."""
    pool = np.asarray(pool, np.int64)
    thr = int(np.ceil(coverage * float(genome_size)))
    if n_draws is None:
        n_draws = int(coverage * genome_size / max(1.0, pool.mean()) * 1.5) + 20000
    L, start, bc = native_draws(seed, combo, iteration, n_draws, genome_size, pool, cdf, sampler)
    label = bc.copy()
    counted = L.copy()
    if mask is not None and len(mask[0]):
        ms, me, mw, mb = (np.asarray(a) for a in mask)
        order = np.lexsort((me, ms))
        ms, me, mw, mb = ms[order], me[order], mw[order], mb[order]
        end = start + L
        cand = np.zeros(n_draws, bool)
        for s, e in zip(ms, me):  # reads touching any interval (barcode filter applied in the loop below)
            cand |= (start < e) & (end > s)
        for n in np.nonzero(cand)[0]:
            j = 0
            blocked = False
            for want in (int(bc[n]), -1):
                for k in range(ms.size):
                    if mb[k] != want or not (ms[k] < end[n] and me[k] > start[n]):
                        continue
                    u = block_u(seed, combo, iteration, int(n), j); j += 1
                    if u < mw[k]:
                        blocked = True
                        break
                if blocked:
                    break
            if blocked:
                label[n] = -1 if consuming else -2
                counted[n] = L[n] if consuming else 0
    prefix = np.concatenate(([0], np.cumsum(counted)[:-1]))
    included = prefix < thr
    n_placed = int(included.sum())
    if n_placed == n_draws and prefix[-1] + counted[-1] < thr:
        raise RuntimeError("n_draws too small: coverage not reached")
    return dict(start=start[:n_placed], length=L[:n_placed], label=label[:n_placed].astype(np.int32),
                covered=int(counted[:n_placed].sum()), n_placed=n_placed)
