"""Target tally on the GPU behind the reference's count_matches signature (src/esloco/counting.py)."""
import numpy as np

from .read_operations import dict_to_reads
from .session import get_session
from .utils import targets_to_arrays

LABELS_BLOCKED = ("Blocked/Consuming", "Blocked/NotConsuming")


def count_matches(insertion_dict, read_dir, scaling, min_overlap, genome_size=None, n_barcodes=None, device=0):
    """count_matches(insertion_dict, read_dir, scaling, min_overlap) (counting.py:7-61) computed by the CUDA
    engine from a reference-shaped read dict (replay level 1).  Returns the reference's list of dicts including the
    per-read `overlap` lists; `scaling` < 1 is rejected because the reference thins pairs with draws from its own
    numpy stream (counting.py:46,51)."""
    if not isinstance(scaling, (int, float)):
        scaling = 1  # counting.py:16-17
    if scaling < 1:
        raise NotImplementedError("count_matches replay with scaling < 1 draws from the reference's numpy stream")
    start, length, label = dict_to_reads(read_dir)
    if n_barcodes is None:
        n_barcodes = int(label.max()) + 1 if label.size and label.max() >= 0 else 1
    keys, ts, te, tb = targets_to_arrays(insertion_dict, n_barcodes)
    if genome_size is None:
        genome_size = int(max(int((start + length).max()) if start.size else 1, int(te.max()) if te.size else 1))
    sess = get_session(device)
    eng = sess.engine
    eng.set_genome(genome_size)
    eng.set_barcode_cdf(np.arange(1, n_barcodes + 1) / n_barcodes)
    eng.set_targets(ts, te, tb)
    eng.set_blocked()
    sess.invalidate()
    eng.set_hit_buffer(max(1 << 16, 8 * start.size))
    try:
        res = eng.replay_reads(start, length, label, [0, start.size], min_overlap=min_overlap).cpu()
        hits = eng.take_hits(res.n_reads)
    finally:
        eng.set_hit_buffer(0)
    overlaps = [[] for _ in keys]
    for _it, row, _n, ov in hits.tolist():
        overlaps[row].append(ov)
    t = res.tallies[0].numpy()
    return [{"target_region": k, "full_matches": int(t[i, 0]), "partial_matches": int(t[i, 1]), "overlap": overlaps[i],
             "on_target_bases": int(t[i, 2])} for i, k in enumerate(keys)]


def count_barcode_occurrences(dictionary):
    """Histogram of read labels (counting.py:63-71), first-appearance key order."""
    counts = {}
    for key in dictionary.keys():
        suffix = key.split("_")[-1]
        counts[suffix] = counts.get(suffix, 0) + 1
    return counts


def label_counts_to_dict(counts, n_barcodes):
    """Engine label histogram row -> the reference's dict (labels with zero reads are absent, as in the
    reference); keys in barcode order then the blocked labels."""
    out = {}
    for k in range(n_barcodes):
        if counts[k]:
            out[str(k)] = int(counts[k])
    for j, name in enumerate(LABELS_BLOCKED):
        if counts[n_barcodes + j]:
            out[name] = int(counts[n_barcodes + j])
    return out


def label_first_appearance(engine, seed, combo, iteration, consuming, n_reads):
    """Labels of one native iteration in the order they first appear among its reads -- the key order of the
    reference's count_barcode_occurrences (counting.py:63-71).  The reads are regenerated on the device (counter-based
    stream) and reduced there; only B + 2 positions come back."""
    import torch
    n_reads = int(n_reads)
    B = engine.n_barcodes
    names = [str(k) for k in range(B)] + list(LABELS_BLOCKED)
    if n_reads == 0:
        return []
    _s, _l, lab = engine.materialize_reads(seed, combo, iteration, consuming, n_reads)
    slot = torch.where(lab >= 0, lab, B - 1 - lab).long()  # -1 -> B (Blocked/Consuming), -2 -> B + 1
    first = torch.full((B + 2,), n_reads, dtype=torch.int64, device=lab.device)
    first.scatter_reduce_(0, slot, torch.arange(n_reads, device=lab.device), reduce="amin")
    first = first.cpu().tolist()
    return [names[k] for k in sorted(range(B + 2), key=lambda k: first[k]) if first[k] < n_reads]
