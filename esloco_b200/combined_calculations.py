"""The drop-in seam: run_simulation_iteration / process_combination with the reference's signatures
(src/esloco/combined_calculations.py:11-126), executed by the CUDA engine, plus a batched form that keeps
many iterations in one launch set."""
import logging
from dataclasses import dataclass

import numpy as np

from .config_handler import seq_read_data
from .counting import label_counts_to_dict
from .engine import BatchResult, EngineError
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


@dataclass
class MomentsResult:
    """`results="moments"` output of one combination: everything summary.csv and the barcode-distribution table need
    (esloco.py:161-195) without the per-iteration tallies -- 64 bytes per target instead of 24 per target and iteration."""
    moments: "torch.Tensor"       # [T, 8] int64: n, sum f, sum f^2, sum p, sum p^2, sum b, lo32(sum b^2), hi(sum b^2)
    label_counts: "torch.Tensor"  # [n_iter, B+2]
    n_bases: "torch.Tensor"       # [n_iter]
    n_reads: "torch.Tensor"       # [n_iter]
    status: "torch.Tensor"        # [n_iter] int32
    tallies = None
    overlaps = None


TALLY_BUDGET_BYTES = 768 << 20  # device bytes of per-iteration tallies held at once in moments mode
HIT_BUDGET_RECORDS = 48 << 20   # hit records (32 B each) per chunk of iterations: 1.5 GB


def expected_pairs_per_iteration(sess, coverage, n_barcodes, barcode_weights):
    """Expected counted (target, read) pairs of one iteration: a read of barcode b overlaps target t with probability
    ~ (len_t + mean read length) / G, and coverage * G / mean read length reads are placed."""
    from .read_operations import barcode_probabilities
    ts, te, tb = sess._targets
    p = np.asarray(barcode_probabilities(n_barcodes, barcode_weights), np.float64)
    p = p / p.sum()
    ok = tb >= 0
    mean_len = max(1.0, sess.pool_mean or 1.0)
    return float(coverage * ((np.maximum(te[ok] - ts[ok], 0) + mean_len) / mean_len * p[tb[ok]]).sum())


def _chunks(first, n_iter, step):
    for c0 in range(0, n_iter, step):
        yield first + c0, min(step, n_iter - c0)


def _run_with_hits(eng, n_targets, pairs_per_iter, hit_capacity, run_chunk, first, n_iter):
    """Native batches with hit records (the reference's per-target `overlap` lists, counting.py:56).  The buffer is
    sized from the expected number of pairs (x1.5 + slack), iterations are split so it fits HIT_BUDGET_RECORDS, and a
    chunk whose pairs still overflow it is re-run with a larger buffer."""
    per_iter = max(1.0, pairs_per_iter * 1.5 + 4096)
    step = n_iter if hit_capacity else max(1, min(n_iter, int(HIT_BUDGET_RECORDS // per_iter)))
    parts, overlaps = [], []
    for it0, n in _chunks(first, n_iter, step):
        cap = int(hit_capacity or min(max(1 << 16, per_iter * n), max(HIT_BUDGET_RECORDS, per_iter)))
        while True:
            eng.set_hit_buffer(cap)
            try:
                res = run_chunk(it0, n)
                try:
                    hits = eng.take_hits(res.n_reads)
                except EngineError as e:
                    if e.code != -5 or hit_capacity:
                        raise
                    cap *= 2  # more pairs than planned: same chunk again (deterministic stream) with more room
                    continue
            finally:
                eng.set_hit_buffer(0)
            break
        parts.append(res)
        overlaps.extend(group_hits(hits, n, n_targets))
    return parts, overlaps


def run_simulation_batch(iteration_ids, param_dictionary, genome_size, target_regions, masked_tree, device=0,
                         to_host=True, length_pools=None, overlap_lists=False, hit_capacity=None, reuse_host_buffers=False,
                         results="tallies"):
    """All combinations for a contiguous run of iterations.  Returns one result per combination (on the host when
    to_host) plus the target keys; iteration i of the batch is iteration_ids[0] + i.
    target_regions is the reference's dict or a utils.TargetColumns (arrays; no per-call dict walk).
    results="tallies": BatchResult with the per-iteration tallies [n_iter, T, 3] (the rows of matches_table.csv).
    results="moments": MomentsResult -- the iterations run in chunks whose tallies stay on the device and are folded
    into the [T, 8] moment accumulators behind summary.csv (esloco.py:183-195); only those, the label counts,
    N_bases and reads placed come back.  For 240 k targets that is 15 MB instead of 5.8 MB per iteration.
    length_pools optionally maps a combination index to a ready-made host array of read lengths (already
    filtered by min_read_length), bypassing the FASTQ scan / log-normal draw.  With overlap_lists=True each
    BatchResult gets an `overlaps` attribute: per iteration, per target, the list of overlaps in read order (the
    reference's `overlap` column); hit_capacity fixes the record buffer (default: sized from the expected pairs).
    Host copies go through page-locked staging tensors owned by the session; with reuse_host_buffers=True the
    returned tensors ARE those staging tensors (valid until the next call) instead of private copies."""
    import torch
    iteration_ids = list(iteration_ids)
    first, n_iter = iteration_ids[0], len(iteration_ids)
    if iteration_ids != list(range(first, first + n_iter)):
        raise ValueError("iteration_ids must be contiguous")
    if results not in ("tallies", "moments"):
        raise ValueError("results must be 'tallies' or 'moments'")
    if results == "moments" and overlap_lists:
        raise ValueError("overlap lists need the per-iteration results (results='tallies')")
    p = param_dictionary
    sess = get_session(device)
    sess.configure(genome_size, target_regions, masked_tree, p["n_barcodes"], p["barcode_weights"],
                   p.get("length_sampler") or "quantile")
    eng = sess.engine
    T = eng.n_targets
    seed = int(p.get("seed", 0))
    scaling = p["scaling"] if isinstance(p["scaling"], (int, float)) else 1
    out = []
    for combo, (mean_read_length, coverage) in enumerate(p["combinations"]):
        if length_pools is not None and combo in length_pools:
            key, factory = None, (lambda c=combo: length_pools[c])
        else:
            key, factory = _length_pool_key_and_factory(seed, combo, mean_read_length, p.get("sequenced_data_path"),
                                                        p["sigma"], p["min_read_length"])
        sess.set_pool(key, factory)

        def run_chunk(it0, n, out_buf=None):
            return eng.run_batch(seed, combo, it0, n, coverage, p["consuming"], p["min_overlap_for_detection"], scaling, out=out_buf)

        overlaps = None
        if results == "moments":
            step = max(1, min(n_iter, TALLY_BUDGET_BYTES // max(1, T * 24)))
            okey = (step, T, eng.n_barcodes)
            if okey not in sess.outputs:
                sess.outputs.clear()
                sess.outputs[okey] = eng.alloc_outputs(step)
            buf = sess.outputs[okey]
            res = MomentsResult(torch.zeros((T, 8), dtype=torch.int64, device=eng.tdev),
                                torch.empty((n_iter, eng.n_barcodes + 2), dtype=torch.int64, device=eng.tdev),
                                torch.empty((n_iter,), dtype=torch.int64, device=eng.tdev),
                                torch.empty((n_iter,), dtype=torch.int64, device=eng.tdev),
                                torch.empty((n_iter,), dtype=torch.int32, device=eng.tdev))
            for it0, n in _chunks(first, n_iter, step):
                part = type(buf)(buf.tallies[:n], buf.label_counts[:n], buf.n_bases[:n], buf.n_reads[:n], buf.status[:n])
                run_chunk(it0, n, part)
                eng.moments(part.tallies, res.moments)
                o = it0 - first
                res.label_counts[o:o + n].copy_(part.label_counts); res.n_bases[o:o + n].copy_(part.n_bases)
                res.n_reads[o:o + n].copy_(part.n_reads); res.status[o:o + n].copy_(part.status)
            fields = ("moments", "label_counts", "n_bases", "n_reads", "status")
        elif overlap_lists:
            pairs = expected_pairs_per_iteration(sess, coverage, p["n_barcodes"], p["barcode_weights"])
            parts, overlaps = _run_with_hits(eng, T, pairs, hit_capacity, run_chunk, first, n_iter)
            res = parts[0] if len(parts) == 1 else BatchResult(*(torch.cat([getattr(x, f) for x in parts]) for f in (
                "tallies", "label_counts", "n_bases", "n_reads", "status")))
            fields = ("tallies", "label_counts", "n_bases", "n_reads", "status")
        else:
            res = run_chunk(first, n_iter)
            fields = ("tallies", "label_counts", "n_bases", "n_reads", "status")
        if to_host:
            res = _to_host(res, fields, sess.pinned, private=not (reuse_host_buffers and len(p["combinations"]) == 1))
        res.overlaps = overlaps
        out.append(res)
    return out, sess.keys


def _to_host(res, fields, pinned, private):
    """Device -> host through the session's page-locked staging tensors (PCIe speed; pageable copies crawl once the
    results are hundreds of MB); private=True returns copies that stay valid after the next call."""
    import torch
    host = []
    for f in fields:
        t = getattr(res, f)
        key = (f, tuple(t.shape), t.dtype)
        if key not in pinned:
            pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pinned[key].copy_(t, non_blocking=True)
        host.append(pinned[key])
    torch.cuda.current_stream(getattr(res, fields[0]).device).synchronize()
    if private:
        host = [t.clone() for t in host]
    return type(res)(*host)


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
                        min_overlap_for_detection, no_cov_plots, seed=0, combo=0, device=0, overlap_lists=False,
                        length_sampler="quantile"):
    """process_combination with the reference's positional signature (combined_calculations.py:11-28) -> (detected,
    barcode_distribution) for ONE combination of ONE iteration.  `seed` and `combo` key the Philox stream (the
    reference takes its randomness from numpy's global state instead); output_path / no_cov_plots are accepted for
    signature compatibility -- the coverage track is written by the CLI, not here."""
    p = dict(n_barcodes=n_barcodes, barcode_weights=barcode_weights, seed=seed, scaling=scaling,
             combinations=[None] * combo + [(mean_read_length, coverage)], sequenced_data_path=sequenced_data_path,
             sigma=sigma, min_read_length=min_read_length, consuming=consuming,
             min_overlap_for_detection=min_overlap_for_detection)
    sess = get_session(device)
    sess.configure(genome_size, target_regions, masked_tree, n_barcodes, barcode_weights, length_sampler)
    key, factory = _length_pool_key_and_factory(seed, combo, mean_read_length, sequenced_data_path, sigma, min_read_length)
    sess.set_pool(key, factory)
    sc = scaling if isinstance(scaling, (int, float)) else 1
    eng = sess.engine

    def run_chunk(it0, n, out_buf=None):
        return eng.run_batch(seed, combo, it0, n, coverage, consuming, min_overlap_for_detection, sc, out=out_buf)

    overlaps = None
    if overlap_lists:
        pairs = expected_pairs_per_iteration(sess, coverage, n_barcodes, barcode_weights)
        parts, overlaps = _run_with_hits(eng, eng.n_targets, pairs, None, run_chunk, iteration, 1)
        res = parts[0]
    else:
        res = run_chunk(iteration, 1)
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
