#!/usr/bin/env python3
"""`esloco --config <ini>` drop-in (mirrors src/esloco/esloco.py): same CLI, INI schema and output tables; the
Monte-Carlo iterations run on the B200 engine.  Launch under torchrun to shard iterations over several GPUs."""
import logging
import os
import sys
import time

import numpy as np

from .bed_operations import global_to_chromosome_coordinates
from .config_handler import parse_config
from .genome_generation import create_barcoded_insertion_genome, create_barcoded_roi_genome
from .utils import build_mask_tree

BANNER = r"""
             __
  ___  _____/ /___  _________
 / _ \/ ___/ / __ \/ ___/ __ \     b200
/  __(__  ) / /_/ / /__/ /_/ /
\___/____/_/\____/\___/\____/
"""


def print_help():
    print("")
    print("Usage: ")
    print("         esloco --config <config_file>")
    print("")
    print("Run a simulation for local coverage estimation.")
    print("")
    print("Options:")
    print("     --config  <config_file>    Path to the configuration file.")
    print("     --help                     Show this help message.")
    print("")
    sys.exit(0)


def setup_logging(filename):
    root = logging.getLogger()
    for handler in root.handlers[:]:
        root.removeHandler(handler)
    logging.basicConfig(filename=filename, level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s", filemode="a")


def iteration_chunks(iterations, n_targets, n_combos, budget_bytes=2 << 30, max_rows=4_000_000):
    """Split a range of iterations into batches whose tallies fit a device-memory budget and whose table rows
    (iterations x combinations x targets) stay below max_rows per streamed chunk."""
    per_iter = max(1, n_targets * 24 * n_combos)
    step = max(1, min(len(iterations), budget_bytes // per_iter, max_rows // max(1, n_targets * n_combos)))
    for i in range(0, len(iterations), step):
        yield iterations[i:i + step]


def run(param_dictionary):
    """Everything after config parsing; returns (barcode distribution, summary) DataFrames on rank 0 (None elsewhere).
    The per-iteration matches table is streamed to disk and not kept in memory."""
    import torch
    import torch.distributed as dist
    from .combined_calculations import run_simulation_batch
    from .distributed import gather_iterations, reduce_moments, shard_iterations, sync_seed, world
    from .session import get_session
    from .tables import MatchesWriter, barcode_distribution_table, summary_table, write_tables

    p = param_dictionary
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sync_seed(p)  # one seed for all ranks, also when the INI leaves it to chance
    mode, output_path, experiment_name = p["mode"], p["output_path"], p["experiment_name"]
    log_file = os.path.join(output_path, f"{experiment_name}_log.log")
    setup_logging(log_file)
    logging.info(f"Starting simulation: {experiment_name}")
    logging.info(f"Running in mode: {mode}")
    start_time = time.time()
    np.random.seed(p["seed"])  # host set-up (insertion sites) draws from numpy's stream like the reference (esloco.py:100)
    logging.info(f"Seed for reproducibility: {p['seed']}")
    insertion_bed = None
    if mode == "I":
        genome_size, target_regions, masked_regions, chromosome_dir = create_barcoded_insertion_genome(
            p["parallel_jobs"], p["reference_genome_path"], p["bedpath"], p["blocked_regions_bedpath"], p["chr_restriction"],
            p["insertion_length"], p["insertion_numbers"], p["insertion_number_distribution"], p["n_barcodes"], log_file)
        logging.info(f"Number of insertions: {len(target_regions)}")
        insertion_bed = global_to_chromosome_coordinates(chromosome_dir, target_regions)
    elif mode == "ROI":
        target_regions, bed, masked_regions, genome_size, chromosome_dir = create_barcoded_roi_genome(
            p["reference_genome_path"], p["chr_restriction"], p["roi_bedpath"], p["n_barcodes"], p["blocked_regions_bedpath"])
        logging.info(f"Chromosome borders: {chromosome_dir}")
    else:
        logging.error("Error: Invalid mode selected.")
        sys.exit(1)
    masked_tree = build_mask_tree(masked_regions) if masked_regions else masked_regions

    combos = p["combinations"]
    mine = shard_iterations(p["iterations"], rank, ws)
    sess = get_session(local)
    dev = torch.device("cuda", local)
    n_targets = len(target_regions)
    sess.configure(genome_size, target_regions, masked_tree, p["n_barcodes"], p["barcode_weights"],
                   p.get("length_sampler") or "quantile")
    keys = sess.keys
    moments = [torch.zeros((n_targets, 8), dtype=torch.int64, device=dev) for _ in combos]
    labels = [[torch.zeros((0, p["n_barcodes"] + 2), dtype=torch.int64, device=dev)] for _ in combos]
    nbases = [[torch.zeros((0,), dtype=torch.int64, device=dev)] for _ in combos]
    statuses = []
    # per-iteration rows are streamed to disk chunk by chunk; every rank writes the block of iterations it owns
    base = f"{output_path}{experiment_name}"
    part = f"{base}_matches_table.csv" if ws == 1 else f"{base}_matches_table.part{rank}.csv"
    writer = MatchesWriter(part, keys, mode, combos, write_header=rank == 0)
    want_lists = bool(p.get("overlap_lists", False))
    t_gpu = time.time()
    # Software pipeline over chunks of iterations: the GPU work of chunk k+1 (asynchronous launches) and the copy of its
    # tallies into one of two page-locked staging buffers are issued BEFORE the host formats the rows of chunk k, so
    # device time and PCIe time hide behind the row writer (which is what the wall clock of a large run is made of).
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [{}, {}]
    pending = None

    def flush(item):
        if item is None:
            return
        chunk, host, done, overlaps = item
        done.synchronize()
        writer.append(list(chunk), [h.numpy() for h in host], overlaps)

    for k, chunk in enumerate(iteration_chunks(mine, n_targets, len(combos)) if len(mine) else []):
        batches, keys = run_simulation_batch(chunk, p, genome_size, target_regions, masked_tree, device=local, to_host=False,
                                             overlap_lists=want_lists)
        for c, res in enumerate(batches):
            if chunk[0] == 0 and not p["no_cov_plots"]:
                # first iteration of every combination (combined_calculations.py:66-74): the binned track behind the
                # reference's coverage plot, written as a table; drawing it is left to the plotting layer
                from .coverage import binned_coverage, write_coverage_track
                from .combined_calculations import _length_pool_key_and_factory
                mrl, cov = combos[c]
                key, factory = _length_pool_key_and_factory(int(p.get("seed", 0)), c, mrl, p.get("sequenced_data_path"),
                                                            p["sigma"], p["min_read_length"])
                sess.set_pool(key, factory)
                track = binned_coverage(sess.engine, int(p.get("seed", 0)), c, 0, p["consuming"], int(res.n_reads[0]))
                write_coverage_track(os.path.join(p["output_path_plots"], f"{mrl}_{cov}_coverage.tsv"), *track)
            sess.engine.moments(res.tallies, moments[c])
            labels[c].append(res.label_counts); nbases[c].append(res.n_bases); statuses.append(res.status)
        overlaps = None
        if want_lists:  # rows: iteration -> combination -> target
            overlaps = [batches[c].overlaps[i][j] for i in range(len(chunk)) for c in range(len(combos)) for j in range(n_targets)]
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        host = []
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            for c, res in enumerate(batches):
                slot = staging[k & 1]
                hkey = (c, tuple(res.tallies.shape))
                if hkey not in slot:
                    slot[hkey] = torch.empty(res.tallies.shape, dtype=torch.int64, pin_memory=True)
                slot[hkey].copy_(res.tallies, non_blocking=True)
                res.tallies.record_stream(copy_stream)
                host.append(slot[hkey])
            done = torch.cuda.Event()
            done.record(copy_stream)
        flush(pending)  # rows of the previous chunk, while the GPU runs this one
        pending = (chunk, host, done, overlaps)
    flush(pending)
    if statuses and (torch.cat(statuses) & 2).any().item():
        raise RuntimeError("coverage not reached")
    for c in range(len(combos)):
        reduce_moments(moments[c])  # the one collective of the data path
        labels[c] = gather_iterations(torch.cat(labels[c]), p["iterations"])
        nbases[c] = gather_iterations(torch.cat(nbases[c]), p["iterations"])
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()  # all part files are complete
    logging.info(f"GPU iterations and row streaming: {time.time() - t_gpu:.2f} seconds on {ws} GPU(s).")
    out = None
    if rank == 0:
        its = list(range(p["iterations"]))
        labels_host = [x.cpu().numpy() for x in labels]

        def first_appearance(_row, iteration_id, c):
            # the reference's label order needs the reads of that iteration: regenerate them (any rank can, the stream
            # is keyed by the iteration id) with the combination's length table resident
            from .combined_calculations import _length_pool_key_and_factory
            from .counting import label_first_appearance
            mrl, _cov = combos[c]
            key, factory = _length_pool_key_and_factory(int(p.get("seed", 0)), c, mrl, p.get("sequenced_data_path"),
                                                        p["sigma"], p["min_read_length"])
            sess.set_pool(key, factory)
            return label_first_appearance(sess.engine, int(p.get("seed", 0)), c, iteration_id, p["consuming"],
                                          int(labels_host[c][iteration_id].sum()))

        dist_df = barcode_distribution_table(p["n_barcodes"], combos, its, labels_host, [x.cpu().numpy() for x in nbases],
                                             label_columns=p.get("label_columns") or "barcode", first_appearance=first_appearance)
        summary = summary_table(keys, mode, combos, [m.cpu().numpy() for m in moments])
        try:
            if ws > 1:
                parts = [f"{base}_matches_table.part{r}.csv" for r in range(ws)]
                MatchesWriter.merge_parts(f"{base}_matches_table.csv", parts)
                for f in parts:
                    os.remove(f)
            write_tables(output_path, experiment_name, None, dist_df, summary, insertion_bed if mode == "I" else None)
            if mode == "I":
                logging.info("Insertion locations saved.")
        except Exception as e:
            logging.error(f"Error writing output files: {e}")
        logging.info("Simulation Parameters:")
        for param, value in p.items():
            logging.info(f"{param} = {value} ({type(value)})")
        logging.info(f"Total execution time: {time.time() - start_time:.2f} seconds.")
        logging.info("Simulation complete!")
        out = (dist_df, summary)
    if ws > 1:
        dist.barrier()
    return out


def main():
    print(BANNER)
    if len(sys.argv) != 3 or sys.argv[1] == "--help" or sys.argv[1] != "--config":
        print_help()
        sys.exit(1)
    config_file = sys.argv[2]
    if not os.path.exists(config_file):
        print(f"Configuration file '{config_file}' does not exist.")
        sys.exit(1)
    print("Checking config...")
    try:
        param_dictionary = parse_config(config_file)
    except Exception as e:
        print(f"Error parsing config: {e}")
        sys.exit(1)
    print(f"Starting simulation: {param_dictionary.get('experiment_name')}")
    print(f"Running in mode: {param_dictionary.get('mode')}")
    run(param_dictionary)
    print("Simulation complete!")
    sys.exit()


if __name__ == "__main__":
    try:
        main()
    except ValueError as e:
        print("Configuration error:", e)
        sys.exit(1)
