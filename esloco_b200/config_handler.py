"""INI schema and sequencing-data length scan (mirrors src/esloco/config_handler.py)."""
import ast
import configparser
import gzip
import itertools
import json
import logging
import os
import sys

import numpy as np

FASTQ_EXTS = (".fastq", ".fastq.gz", ".fq", ".fq.gz")
FASTA_EXTS = (".fasta", ".fasta.gz", ".fa", ".fa.gz")


def _fasta_lengths(handle):
    n, seen = 0, False
    for line in handle:
        if line.startswith(">"):
            if seen:
                yield n
            seen, n = True, 0
        elif seen:
            n += len(line.strip())
    if seen:
        yield n


def _fastq_lengths(handle, path):
    while True:
        head = handle.readline()
        if not head:
            return
        seq = handle.readline().rstrip("\n")
        plus = handle.readline()
        qual = handle.readline().rstrip("\n")
        if not head.startswith("@") or not plus.startswith("+") or (len(seq) and not len(qual)):
            raise ValueError(f"This is not a correct FASTQ file {path}. Provide FASTA or FASTQ.")
        yield len(seq)


def _scan_native(fasta_file, is_fastq, min_read_length):
    """One streaming pass in the native library (esl_scan_read_lengths); None when the library is not built."""
    import ctypes as C
    try:
        from . import _native
        lib = _native.load()
    except Exception:
        return None
    mean = C.c_double()
    n = lib.esl_scan_read_lengths(fasta_file.encode(), int(is_fastq), int(min_read_length), None, 0, C.byref(mean))
    if n == -6:
        raise ValueError(f"This is not a correct FASTQ file {fasta_file}. Provide FASTA or FASTQ.")
    if n < 0:
        raise OSError(f"cannot read {fasta_file}")
    out = np.empty(n, np.uint32)
    lib.esl_scan_read_lengths(fasta_file.encode(), int(is_fastq), int(min_read_length), out.ctypes.data, n, C.byref(mean))
    return out.astype(np.int64)


def seq_read_data(fasta_file, distribution=False, min_read_length=0):
    """Read lengths of a FASTA/FASTQ file, plain or gzipped (config_handler.py:16-51): lengths > min_read_length
    as an array (distribution=True) or their mean.  Streams the file; sequences are never kept.  Uses the native
    scanner of libesloco_b200.so when it is built, else a pure-Python pass with the same result."""
    if not os.path.exists(fasta_file):
        raise FileNotFoundError(fasta_file)
    if fasta_file.endswith(FASTQ_EXTS) or fasta_file.endswith(FASTA_EXTS):
        lengths = _scan_native(fasta_file, fasta_file.endswith(FASTQ_EXTS), min_read_length)
        if lengths is not None:
            if lengths.size == 0:
                raise ValueError(f"No reads found in {fasta_file}")
            return lengths if distribution else np.mean(lengths)
    open_func = gzip.open if fasta_file.endswith(".gz") else open
    with open_func(fasta_file, "rt") as handle:
        if fasta_file.endswith(FASTQ_EXTS):
            it = _fastq_lengths(handle, fasta_file)
        elif fasta_file.endswith(FASTA_EXTS):
            it = _fasta_lengths(handle)
        else:
            raise ValueError(f"Unsupported file format for {fasta_file}. Only FASTA and FASTQ are allowed.")
        lengths = np.fromiter((n for n in it if n > min_read_length), dtype=np.int64)
    if lengths.size == 0:
        raise ValueError(f"No reads found in {fasta_file}")
    return lengths if distribution else np.mean(lengths)


def random_seed():
    return np.random.randint(0, 2**32 - 1)


def parse_config(config_file):
    """[COMMON]/[ROI]/[I] INI -> parameter dict with the reference's keys and defaults (config_handler.py:59-152)."""
    try:
        config = configparser.ConfigParser()
        config.read(config_file)
        mode = config.get("COMMON", "mode", fallback=None)
        reference_genome_path = config.get("COMMON", "reference_genome_path", fallback=None)
        if mode is None:
            print("Missing required parameter: mode")
            raise ValueError("Missing required parameter: mode")
        if reference_genome_path is None:
            print("Missing required parameter: reference_genome_path")
            raise ValueError("Missing required parameter: reference_genome_path")
        sequenced_data_path = config.get("COMMON", "sequenced_data_path", fallback=None)
        sigma = config.getint("COMMON", "sigma", fallback=1)
        output_path = config.get("COMMON", "output_path", fallback="./output/")
        experiment_name = config.get("COMMON", "experiment_name", fallback="default_experiment")
        output_path_plots = config.get("COMMON", "output_path_plots", fallback=output_path)
        min_overlap_for_detection = config.getint("COMMON", "min_overlap_for_detection", fallback=1)
        chr_restriction = config.get("COMMON", "chr_restriction", fallback=None)
        barcode_weights = config.get("COMMON", "barcode_weights", fallback="None")
        barcode_weights = ast.literal_eval(barcode_weights) if barcode_weights != "None" else None
        n_barcodes = config.getint("COMMON", "n_barcodes", fallback=1)
        iterations = config.getint("COMMON", "iterations", fallback=1)
        scaling = config.getfloat("COMMON", "scaling", fallback=1.0)
        min_read_length = config.getint("COMMON", "min_read_length", fallback=1)
        no_cov_plots = config.getboolean("COMMON", "no_cov_plots", fallback=False)
        parallel_jobs = config.getint("COMMON", "parallel_jobs", fallback=1)
        seed = config.getint("COMMON", "seed", fallback=random_seed())
        for path in (output_path, output_path_plots):
            if not os.path.exists(path):
                os.makedirs(path)
        coverages = json.loads(config.get("COMMON", "coverages", fallback="[1]"))
        if not isinstance(coverages, list):
            coverages = [coverages]
        mean_read_lengths = json.loads(config.get("COMMON", "mean_read_lengths", fallback="[10000]"))
        if not isinstance(mean_read_lengths, list):
            mean_read_lengths = [mean_read_lengths]
        if sequenced_data_path:
            print("Sequencing data provided. Calculating the mean read length may take a while...")
            logging.info("Sequencing data provided. Calculating the mean read length may take a while...")
            mrl = int(seq_read_data(sequenced_data_path, min_read_length=min_read_length))
            logging.info(f"Mean read length set to: {mrl}")
            mean_read_lengths = [mrl]
        overlap_lists = config.getboolean("COMMON", "overlap_lists", fallback=False)  # extension: fill the `overlap` column
        blocked_regions_bedpath = config.get("COMMON", "blocked_regions_bedpath", fallback=None)
        consuming = config.getboolean("COMMON", "consuming", fallback=False)
        if mode == "ROI":
            roi_bedpath = config.get("ROI", "roi_bedpath", fallback=None)
        elif mode == "I":
            insertion_length = config.getint("I", "insertion_length", fallback=1000)
            insertion_number_distribution = config.get("I", "insertion_number_distribution", fallback=None)
            bedpath = config.get("I", "bedpath", fallback=None)
            insertion_numbers = config.getfloat("I", "insertion_numbers", fallback=5.0)
        else:
            print("Invalid mode selected. Exiting.")
            raise ValueError("Invalid mode selected. Allowed values: 'ROI' or 'I'.")
        combinations = list(itertools.product(mean_read_lengths, coverages))
        return {k: v for k, v in locals().items() if k != "config" and k != "path" and not k.startswith("_")}
    except Exception as e:
        print("Configuration parsing failed: ", e)
        sys.exit(1)
