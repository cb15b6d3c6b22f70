"""Synthetic inputs of the shapes BASELINE.json names (no network, no real genome needed: the hot path only ever
sees contig LENGTHS, never sequence)."""
import numpy as np

from .create_insertion_genome import _shifted_starts

# GRCh38 primary assembly contig lengths, chr1..chr22, chrX, chrY (sum 3,088,269,832)
HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
        ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
        ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


def chromosome_dir(contigs):
    """{contig: [global start, global end]} and total length, as pseudo_fasta_coordinates returns them
    (fasta_operations.py:3-22)."""
    out, pos = {}, 0
    for name, n in contigs:
        out[name] = [pos, pos + n]
        pos += n
    return pos, out


def lognormal_lengths(mean_read_length, sigma=1, min_read_length=1, n=1000000, seed=0):
    rng = np.random.default_rng(seed)
    x = np.round(rng.lognormal(np.log(mean_read_length) - 0.5, sigma, n)).astype(np.int64)
    return x[x > min_read_length]


def random_insertions(genome_size, n_insertions, insertion_length, n_barcodes, seed=0):
    """Insertion targets with the reference's sequential coordinate-shift rule (create_insertion_genome.py:67-83):
    each new insertion pushes every earlier insertion at or after its position right by insertion_length."""
    rng = np.random.default_rng(seed)
    targets = {}
    for b in range(n_barcodes):
        pos = rng.integers(0, genome_size, size=n_insertions)
        starts = _shifted_starts(pos, insertion_length).tolist()
        for i in range(n_insertions):
            targets[f"Barcode_{b}_insertion_{i}"] = [starts[i], starts[i] + insertion_length]
    return targets


def random_rois(contigs, n_rois, min_len, max_len, seed=0):
    """Sorted ROI set in global coordinates: {id: {"start","end","weight","barcode"}} (bed_operations.py:80-105)."""
    rng = np.random.default_rng(seed)
    total, cdir = chromosome_dir(contigs)
    lens = rng.integers(min_len, max_len, size=n_rois)
    starts = np.sort(rng.integers(0, total - max_len, size=n_rois))
    return {f"R{i}": {"start": int(s), "end": int(s + ln), "weight": 1, "barcode": []} for i, (s, ln) in enumerate(zip(starts, lens))}


def barcode_rois(rois, n_barcodes):
    """ROI barcoding `{id}_{b}` (genome_generation.py:85-88)."""
    return {f"{k}_{b}": v for b in range(n_barcodes) for k, v in rois.items()}


def gap_mask(contigs, n_gaps, seed=0):
    """Centromere-like 3 Mb interval per contig plus random gaps, weight 1, all barcodes (config 4)."""
    rng = np.random.default_rng(seed)
    total, cdir = chromosome_dir(contigs)
    masked = {}
    for name, (s, e) in cdir.items():
        mid = (s + e) // 2
        masked[f"cen_{name}"] = {"start": mid - 1_500_000, "end": mid + 1_500_000, "weight": 1, "barcode": []}
    for i in range(n_gaps):
        ln = int(rng.integers(10_000, 1_000_000))
        st = int(rng.integers(0, total - ln))
        masked[f"gap{i}"] = {"start": st, "end": st + ln, "weight": 1, "barcode": []}
    return masked
