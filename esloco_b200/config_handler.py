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
    from . import _native
    try:
        lib = _native.load()
    except _native.NativeLibraryError:  # library not built yet: the Python pass below does the same job, slower
        return None
    mean = C.c_double()
    # one pass when the guess holds (long reads: at least ~1 kb of file per record), a second one with the exact count else
    try:
        guess = max(1 << 16, os.path.getsize(fasta_file) // 1000)
    except OSError:
        guess = 1 << 16
    out = np.empty(guess, np.uint32)
    n = lib.esl_scan_read_lengths(fasta_file.encode(), int(is_fastq), int(min_read_length), out.ctypes.data, guess, C.byref(mean))
    if n == -6:
        raise ValueError(f"This is not a correct FASTQ file {fasta_file}. Provide FASTA or FASTQ.")
    if n < 0:
        raise OSError(f"cannot read {fasta_file}")
    if n > guess:
        out = np.empty(n, np.uint32)
        lib.esl_scan_read_lengths(fasta_file.encode(), int(is_fastq), int(min_read_length), out.ctypes.data, n, C.byref(mean))
    return out[:n].astype(np.int64)


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


# [COMMON] options: (key, reader, default) in the order the reference's parameter dictionary lists them
# (config_handler.py:70-129); `None` reader = plain string.
_COMMON = (
    ("sequenced_data_path", None, None), ("sigma", "int", 1), ("output_path", None, "./output/"),
    ("experiment_name", None, "default_experiment"), ("output_path_plots", None, "<output_path>"),
    ("min_overlap_for_detection", "int", 1), ("chr_restriction", None, None), ("barcode_weights", "literal", None),
    ("n_barcodes", "int", 1), ("iterations", "int", 1), ("scaling", "float", 1.0), ("min_read_length", "int", 1),
    ("no_cov_plots", "bool", False), ("parallel_jobs", "int", 1), ("seed", "int", "<random>"),
    ("coverages", "json_list", "[1]"), ("mean_read_lengths", "json_list", "[10000]"),
    ("overlap_lists", "bool", False),  # extension: fill the per-read `overlap` column
    ("length_sampler", None, "quantile"),
    ("label_columns", None, "barcode"),  # extension: "reference" = the reference's first-appearance columns and total_Reads rule  # extension: "quantile" (shared-memory inverse-CDF table) | "pool" (exact uniform pick)
    ("blocked_regions_bedpath", None, None), ("consuming", "bool", False),
)
_MODE = {
    "ROI": (("roi_bedpath", None, None),),
    "I": (("insertion_length", "int", 1000), ("insertion_number_distribution", None, None), ("bedpath", None, None),
          ("insertion_numbers", "float", 5.0)),
}


def _read(config, section, key, kind, default):
    if kind == "int":
        return config.getint(section, key, fallback=default)
    if kind == "float":
        return config.getfloat(section, key, fallback=default)
    if kind == "bool":
        return config.getboolean(section, key, fallback=default)
    if kind == "literal":  # Python literal, the string "None" meaning absent (barcode_weights)
        text = config.get(section, key, fallback="None")
        return ast.literal_eval(text) if text != "None" else None
    if kind == "json_list":  # JSON scalar or list -> list
        value = json.loads(config.get(section, key, fallback=default))
        return value if isinstance(value, list) else [value]
    return config.get(section, key, fallback=default)


def parse_config(config_file):
    """[COMMON]/[ROI]/[I] INI -> parameter dict with the reference's keys, types and defaults
    (config_handler.py:59-152, README.md:76-114): creates the output directories, replaces mean_read_lengths by the
    integer mean of the sequencing data when sequenced_data_path is set, adds `combinations`.  Any problem prints
    "Configuration parsing failed" and exits with status 1, as the reference does."""
    try:
        config = configparser.ConfigParser()
        config.read(config_file)
        params = {"config_file": config_file}
        for required in ("mode", "reference_genome_path"):
            params[required] = config.get("COMMON", required, fallback=None)
            if params[required] is None:
                print(f"Missing required parameter: {required}")
                raise ValueError(f"Missing required parameter: {required}")
        if params["mode"] not in _MODE:
            print("Invalid mode selected. Exiting.")
            raise ValueError("Invalid mode selected. Allowed values: 'ROI' or 'I'.")
        for key, kind, default in _COMMON:
            if default == "<output_path>":
                default = params["output_path"]
            elif default == "<random>":
                default = random_seed()
            params[key] = _read(config, "COMMON", key, kind, default)
        for path in (params["output_path"], params["output_path_plots"]):
            if not os.path.exists(path):
                os.makedirs(path)
        if params["sequenced_data_path"]:
            print("Sequencing data provided. Calculating the mean read length may take a while...")
            logging.info("Sequencing data provided. Calculating the mean read length may take a while...")
            mrl = int(seq_read_data(params["sequenced_data_path"], min_read_length=params["min_read_length"]))
            logging.info(f"Mean read length set to: {mrl}")
            params["mean_read_lengths"] = [mrl]
            params["mrl"] = mrl  # the reference's dictionary is its locals(), so this helper shows up there too
        for key, kind, default in _MODE[params["mode"]]:
            params[key] = _read(config, params["mode"], key, kind, default)
        params["combinations"] = list(itertools.product(params["mean_read_lengths"], params["coverages"]))
        return params
    except Exception as e:
        print("Configuration parsing failed: ", e)
        sys.exit(1)
