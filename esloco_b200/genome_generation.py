"""Host set-up of both modes (mirrors src/esloco/genome_generation.py): genome size, targets, blocked regions."""
import logging

from .bed_operations import chromosome_to_global_coordinates, readbed
from .create_insertion_genome import add_insertions_to_genome_sequence_with_bed
from .fasta_operations import pseudo_fasta_coordinates


def barcode_genome(chromosome_dir, barcode):
    return {f"Barcode_{barcode}_{k}": v for k, v in chromosome_dir.items()}  # utils.py:59-64


def _masked_regions(blocked_regions_bedpath, chromosome_dir):
    if not blocked_regions_bedpath or blocked_regions_bedpath.lower() == "none":
        logging.info("No regions provided for masking...")
        return None
    logging.info(f"Masking regions as defined in {blocked_regions_bedpath}")
    return chromosome_to_global_coordinates(readbed(blocked_regions_bedpath, chromosome_dir.keys()), chromosome_dir)


def create_barcoded_insertion_genome(parallel_jobs, reference_genome_path, bedpath, blocked_regions_bedpath, restriction,
                                     insertion_length, insertion_numbers, insertion_number_distribution, n_barcodes,
                                     log_file=None):
    """I mode (genome_generation.py:25-68) -> (genome size, {Barcode_b_insertion_i: [s, e]}, masked regions,
    chromosome dir).  Barcodes are processed in order in this process, so the numpy seed fixes the insertion sites
    for any parallel_jobs (the reference loses that with parallel_jobs > 1)."""
    logging.info("Create Barcoded Insertion Genome...")
    genome_size, chromosome_dir = pseudo_fasta_coordinates(reference_genome_path, restriction)
    masked_regions = _masked_regions(blocked_regions_bedpath, chromosome_dir)
    insertions = {}
    for b in range(n_barcodes):
        barcoded = barcode_genome(chromosome_dir, b)
        if not bedpath or bedpath.lower() == "none":
            bed_df = None
        else:
            logging.info(f"Guided insertion placement within regions defined in {bedpath}.")
            bed_df = readbed(bedpath, barcoded.keys(), barcoding=n_barcodes >= 1)
        insertions.update(add_insertions_to_genome_sequence_with_bed(genome_size, insertion_length, insertion_numbers,
                                                                     barcoded, insertion_number_distribution, bed_df))
    return genome_size, insertions, masked_regions, chromosome_dir


def create_barcoded_roi_genome(reference_genome_path, restriction, roi_bedpath, n_barcodes, blocked_regions_bedpath):
    """ROI mode (genome_generation.py:70-98) -> (barcoded ROI dict "{id}_{b}", bed DataFrame, masked regions,
    genome size, chromosome dir)."""
    logging.info("Create Barcoded ROI Genome...")
    genome_size, chromosome_dir = pseudo_fasta_coordinates(reference_genome_path, restriction)
    bed = readbed(roi_bedpath, chromosome_dir.keys())
    roi_dict = chromosome_to_global_coordinates(bed, chromosome_dir)
    barcoded = {f"{key}_{b}": value for b in range(n_barcodes) for key, value in roi_dict.items()}
    masked_regions = _masked_regions(blocked_regions_bedpath, chromosome_dir)
    return barcoded, bed, masked_regions, genome_size, chromosome_dir
