"""Insertion target generator (mirrors src/esloco/create_insertion_genome.py) -- SURVEY.md section 8(f) row 4.

Same draws, in the same order, from numpy's global legacy stream as the reference, so a given `seed` yields the
same insertion coordinates; the O(n^2) Python dict walk of the reference (create_insertion_genome.py:60-64,76-80)
is replaced by a vectorised shift over an array of starts with identical sequential semantics."""
import logging

import numpy as np


def _chromosome_of(positions, chromosome_dir):
    """get_chromosome (utils.py:46-54) for many positions: the FIRST contig with start <= pos <= end."""
    names = list(chromosome_dir.keys())
    ends = np.array([v[1] for v in chromosome_dir.values()], np.int64)
    starts = np.array([v[0] for v in chromosome_dir.values()], np.int64)
    idx = np.searchsorted(ends, positions, side="left")  # first contig whose end >= pos
    idx = np.minimum(idx, len(names) - 1)
    ok = (starts[idx] <= positions) & (positions <= ends[idx])
    return [names[i] if good else None for i, good in zip(idx, ok)]


def _shifted_starts(positions, insertion_length):
    """Every new insertion pushes the earlier insertions that start at or after its position right by
    insertion_length (create_insertion_genome.py:76-80).  Runs in the native host helper esl_shift_insertions; the
    numpy loop below is the same recurrence and only serves a checkout whose library has not been built yet."""
    positions = np.ascontiguousarray(positions, np.int64)
    starts = np.empty(len(positions), np.int64)
    try:
        from . import _native
        lib = _native.load()
    except _native.NativeLibraryError:
        lib = None
    if lib is not None:
        if lib.esl_shift_insertions(positions.ctypes.data, positions.size, int(insertion_length), starts.ctypes.data) != 0:
            raise ValueError("esl_shift_insertions rejected its arguments")
        return starts
    for i, p in enumerate(positions):
        head = starts[:i]
        head[head >= p] += insertion_length
        starts[i] = p
    return starts


def add_insertions_to_genome_sequence_with_bed(reference_sequence, insertion_length, num_insertions, chromosome_dir,
                                               insertion_number_distribution=None, bed_df=None):
    """-> {"<prefix>insertion_<i>": [start, end]} with prefix = contig name up to "chr"
    (create_insertion_genome.py:8-87).  reference_sequence is the genome LENGTH."""
    if insertion_number_distribution == "poisson":
        num_insertions = np.random.poisson(num_insertions)
        logging.info(f"Number of insertions drawn from Poisson distribution: {num_insertions}")
    else:
        logging.info(f"Using exactly {num_insertions}.")
    num_insertions = int(num_insertions)
    if bed_df is not None:
        logging.info("BED guided insertion pattern...")
        weights = bed_df["weight"].astype(float)
        if np.mean(weights) != 1.0:
            probabilities = weights
            if probabilities.sum() != 1:
                probabilities = probabilities / probabilities.sum()
        else:
            lengths = bed_df["end"] - bed_df["start"]
            probabilities = lengths / lengths.sum()
        positions, chroms = [], []
        for i in range(num_insertions):
            if len(bed_df.index) == num_insertions:
                region = bed_df.iloc[i]
            else:
                region = bed_df.iloc[np.random.choice(bed_df.index, p=probabilities)]
            local = np.random.randint(region["start"], region["end"])
            chroms.append(region["chrom"])
            positions.append(chromosome_dir[region["chrom"]][0] + local)
        positions = np.array(positions, np.int64)
    else:
        if num_insertions > 0:
            positions = np.random.randint(0, reference_sequence, size=num_insertions).astype(np.int64)
        else:
            positions = np.zeros(0, np.int64)
        chroms = _chromosome_of(positions, chromosome_dir)
    starts = _shifted_starts(positions, insertion_length)
    position = {}
    for i, (chrom, s) in enumerate(zip(chroms, starts.tolist())):
        position[chrom.split("chr")[0] + f"insertion_{i}"] = [s, s + insertion_length]
    return position
