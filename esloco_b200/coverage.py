"""Binned coverage track of the first iteration (the data behind the reference's coverage plot, coverage.py:33-136)
-- SURVEY.md section 8(f) row 2.  The reference allocates one genome-length float64 array per barcode, which is
impossible at hg38 size; here the reads of the iteration are regenerated on the device (same Philox counters as
the simulation) and accumulated per bin.  Plotting itself stays with the reference's plotting layer."""
import numpy as np

from .counting import LABELS_BLOCKED


def gaussian_filter1d(x, sigma=3, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d(x, sigma) with its defaults (mode='reflect'), numpy only."""
    x = np.asarray(x, np.float64)
    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * k ** 2)
    w /= w.sum()
    pad = x
    while pad.size < x.size + 2 * radius:  # scipy's 'reflect' repeats the edge sample: d c b a | a b c d | d c b a
        need = (x.size + 2 * radius - pad.size + 1) // 2
        pad = np.pad(pad, min(need, pad.size), mode="symmetric")
    off = (pad.size - x.size) // 2 - radius
    return np.convolve(pad, w[::-1], mode="valid")[off:off + x.size]


def binned_coverage(engine, seed, combo, iteration, consuming, n_reads, bin_size=1000000, smooth_sigma=3):
    """-> (bin positions, {label: mean coverage per bin}, {label: smoothed}) for one native-mode iteration; labels are
    "Combined", the barcode numbers and the blocked labels that occur (coverage.py:52-81).  n_reads is the
    iteration's reads-placed count from the simulation."""
    start, length, label = engine.materialize_reads(seed, combo, iteration, consuming, n_reads)
    bases = engine.coverage_bins(start, length, label, bin_size).cpu().numpy().astype(np.float64)
    G, B = engine.genome_size, engine.n_barcodes
    pos = np.arange(0, G, bin_size)
    width = np.minimum(pos + bin_size, G) - pos  # np.mean over a truncated last bin (coverage.py:24-30)
    names = [str(k) for k in range(B)] + list(LABELS_BLOCKED)
    raw = {"Combined": bases.sum(0) / width}
    for k, name in enumerate(names):
        if bases[k].any():
            raw[name] = bases[k] / width
    return pos, raw, {k: gaussian_filter1d(v, smooth_sigma) for k, v in raw.items()}


def write_coverage_track(path, pos, raw, smoothed):
    import pandas as pd
    df = pd.DataFrame({"bin_start": pos})
    for k in raw:
        df[f"{k}"] = raw[k]
        df[f"{k}_smoothed"] = smoothed[k]
    df.to_csv(path, sep="\t", header=True, index=False)
