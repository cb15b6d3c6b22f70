"""The drop-in seam: run_simulation_iteration / process_combination with the reference's signatures
(src/esloco/combined_calculations.py:11-126), executed by the CUDA engine, plus a batched form that keeps
many iterations in one launch set."""
import logging

import numpy as np

from .config_handler import seq_read_data
from .counting import label_counts_to_dict
from .session import get_session, lognormal_pool


def _length_pool_key_and_factory(seed, combo, mean_read_length, sequenced_data_path, sigma, min_read_length):
    """Length source of one combination (combined_calculations.py:36-49): FASTA/FASTQ lengths > min_read_length
    when the file parses, else the log-normal pool."""
    if sequenced_data_path:
        try:
            lengths = seq_read_data(sequenced_data_path, distribution=True, min_read_length=min_read_length)
            logging.info("Custom FASTA/FASTQ data provided.")
            return ("file", sequenced_data_path, min_read_length), (lambda: lengths)
        except Exception:
            pass  # the reference falls back silently as well (combined_calculations.py:42)
    logging.info("No custom read length distribution provided... Generating artificial one...")
    return (("lognormal", seed, combo, mean_read_length, sigma, min_read_length),
            lambda: lognormal_pool(seed, combo, mean_read_length, sigma, min_read_length))


def group_hits(hits, n_iter, n_targets):
    """Sorted hit records [k, 4] -> per-iteration list of per-target overlap lists (counting.py:56 `overlap`)."""
    out = [[[] for _ in range(n_targets)] for _ in range(n_iter)]
    for it, row, _n, ov in hits.tolist():
        out[it][row].append(ov)
    return out


def run_simulation_batch(iteration_ids, param_dictionary, genome_size, target_regions, masked_tree, device=0,
                         to_host=True, length_pools=None, overlap_lists=False, hit_capacity=None, reuse_host_buffers=False):
    """All combinations for a contiguous run of iterations.  Returns one BatchResult per combination (on the
    host when to_host) plus the target keys; iteration i of the batch is iteration_ids[0] + i.
    length_pools optionally maps a combination index to a ready-made host array of read lengths (already
    filtered by min_read_length), bypassing the FASTQ scan / log-normal draw.  With overlap_lists=True each
    BatchResult gets an `overlaps` attribute: per iteration, per target, the list of overlaps in read order (the
    reference's `overlap` column); hit_capacity bounds the number of (target, read) pairs per combination.
    Host copies go through page-locked staging tensors owned by the session; with reuse_host_buffers=True the
    returned tensors ARE those staging tensors (valid until the next call) instead of private copies."""
    iteration_ids = list(iteration_ids)
    first, n_iter = iteration_ids[0], len(iteration_ids)
    if iteration_ids != list(range(first, first + n_iter)):
        raise ValueError("iteration_ids must be contiguous")
    p = param_dictionary
    sess = get_session(device)
    sess.configure(genome_size, target_regions, masked_tree, p["n_barcodes"], p["barcode_weights"])
    seed = int(p.get("seed", 0))
    scaling = p["scaling"] if isinstance(p["scaling"], (int, float)) else 1
    out = []
    for combo, (mean_read_length, coverage) in enumerate(p["combinations"]):
        if length_pools is not None and combo in length_pools:
            key, factory = ("given", id(length_pools[combo])), (lambda c=combo: length_pools[c])
        else:
            key, factory = _length_pool_key_and_factory(seed, combo, mean_read_length, p.get("sequenced_data_path"),
                                                        p["sigma"], p["min_read_length"])
        sess.set_pool(key, factory)
        if overlap_lists:
            sess.engine.set_hit_buffer(hit_capacity or max(1 << 20, 64 * len(target_regions) * n_iter))
        try:
            res = sess.engine.run_batch(seed, combo, first, n_iter, coverage, p["consuming"],
                                        p["min_overlap_for_detection"], scaling)
            overlaps = group_hits(sess.engine.take_hits(res.n_reads), n_iter, len(target_regions)) if overlap_lists else None
        finally:
            if overlap_lists:
                sess.engine.set_hit_buffer(0)
        if to_host:
            staged = res.cpu(pinned=sess.pinned)
            res = staged if (reuse_host_buffers and len(p["combinations"]) == 1) else type(staged)(*(t.clone() for t in (
                staged.tallies, staged.label_counts, staged.n_bases, staged.n_reads, staged.status)))
        res.overlaps = overlaps
        out.append(res)
    return out, sess.keys


def format_iteration(res, i, keys, iteration_id, mean_read_length, coverage, n_barcodes):
    """One iteration of a BatchResult -> the reference's (detected, barcode_distribution) pair
    (combined_calculations.py:76-89)."""
    t = res.tallies[i].numpy()
    suffix = "_" + str(iteration_id)
    ov = getattr(res, "overlaps", None)
    detected = []
    for j, k in enumerate(keys):
        d = {"target_region": k + suffix, "full_matches": int(t[j, 0]), "partial_matches": int(t[j, 1])}
        if ov is not None:
            d["overlap"] = ov[i][j]
        d.update(on_target_bases=int(t[j, 2]), mean_read_length=mean_read_length, coverage=coverage)
        detected.append(d)
    dist = label_counts_to_dict(res.label_counts[i].numpy(), n_barcodes)
    dist["coverage"] = coverage
    dist["mean_read_length"] = mean_read_length
    dist["iteration"] = iteration_id
    dist["N_bases"] = int(res.n_bases[i])
    return detected, dist


def process_combination(mean_read_length, coverage, genome_size, target_regions, iteration, sequenced_data_path, sigma,
                        min_read_length, n_barcodes, barcode_weights, consuming, masked_tree, output_path, scaling,
                        min_overlap_for_detection, no_cov_plots, seed=0, combo=0, device=0, overlap_lists=False):
    """process_combination with the reference's positional signature (combined_calculations.py:11-28) -> (detected,
    barcode_distribution) for ONE combination of ONE iteration.  `seed` and `combo` key the Philox stream (the
    reference takes its randomness from numpy's global state instead); output_path / no_cov_plots are accepted for
    signature compatibility -- the coverage track is written by the CLI, not here."""
    p = dict(n_barcodes=n_barcodes, barcode_weights=barcode_weights, seed=seed, scaling=scaling,
             combinations=[None] * combo + [(mean_read_length, coverage)], sequenced_data_path=sequenced_data_path,
             sigma=sigma, min_read_length=min_read_length, consuming=consuming,
             min_overlap_for_detection=min_overlap_for_detection)
    sess = get_session(device)
    sess.configure(genome_size, target_regions, masked_tree, n_barcodes, barcode_weights)
    key, factory = _length_pool_key_and_factory(seed, combo, mean_read_length, sequenced_data_path, sigma, min_read_length)
    sess.set_pool(key, factory)
    sc = scaling if isinstance(scaling, (int, float)) else 1
    if overlap_lists:
        sess.engine.set_hit_buffer(max(1 << 20, 64 * len(target_regions)))
    try:
        res = sess.engine.run_batch(seed, combo, iteration, 1, coverage, consuming, min_overlap_for_detection, sc)
        overlaps = group_hits(sess.engine.take_hits(res.n_reads), 1, len(target_regions)) if overlap_lists else None
    finally:
        if overlap_lists:
            sess.engine.set_hit_buffer(0)
    res = res.cpu()
    res.overlaps = overlaps
    if int(res.status[0]) & 2:
        raise RuntimeError("coverage not reached")
    return format_iteration(res, 0, sess.keys, iteration, mean_read_length, coverage, n_barcodes)


def run_simulation_iteration(iteration_id, param_dictionary, genome_size, target_regions, masked_tree, log_file=None,
                             device=0, overlap_lists=None):
    """run_simulation_iteration(iteration_id, param_dictionary, genome_size, target_regions, masked_tree,
    log_file) -> (results, barcode_distributions) with the reference's shapes (combined_calculations.py:97-126):
    results[c] is the per-target list of dicts of combination c, barcode_distributions[c] its label histogram.
    The per-read `overlap` list of each dict is filled when overlap_lists (or param_dictionary["overlap_lists"]) is
    true; it is off by default because it is the only output whose size grows with the number of hits."""
    if overlap_lists is None:
        overlap_lists = bool(param_dictionary.get("overlap_lists", False))
    batches, keys = run_simulation_batch([iteration_id], param_dictionary, genome_size, target_regions, masked_tree, device,
                                         overlap_lists=overlap_lists)
    results, dists = [], []
    for res, (mean_read_length, coverage) in zip(batches, param_dictionary["combinations"]):
        if int(res.status[0]) & 2:
            raise RuntimeError("coverage not reached")
        detected, dist = format_iteration(res, 0, keys, iteration_id, mean_read_length, coverage,
                                          param_dictionary["n_barcodes"])
        results.append(detected)
        dists.append(dist)
    return results, dists
