"""BED handling for targets and blocked regions (mirrors src/esloco/bed_operations.py)."""
import logging
import math
import sys

import pandas as pd

ALL_COLS = ["chrom", "start", "end", "id", "barcode", "weight"]


def bedcheck(bed):
    """chrN / integer / integer in the first three columns (bed_operations.py:6-16)."""
    ok = (bed.iloc[:, 0].astype(str).str.startswith("chr").all()
          and bed.iloc[:, 1].apply(lambda x: isinstance(x, int)).all()
          and bed.iloc[:, 2].apply(lambda x: isinstance(x, int)).all())
    if not ok:
        logging.info("The file does not seem to be in BED format. Make sure the data looks like: chrN integer integer")
    return bool(ok)


def _is_missing(x):
    return x in ("", None) or (isinstance(x, float) and math.isnan(x))


def readbed(bedpath, list_of_chromosomes_in_reference, barcoding=False):
    """Flexible BED reader (bed_operations.py:28-78): optional header (first cell contains "chr", second is
    "start"), missing id / barcode / weight columns get defaults (id = chrom_start_end_weight, barcode = [],
    weight = 1), rows on contigs absent from the reference are dropped.  With barcoding=True the contig names are
    prefixed like the barcoded chromosome dict ("Barcode_k_")."""
    try:
        bed = pd.read_csv(bedpath, sep="\t", header=None, dtype=str)
        if "chr" in bed.iloc[0, 0].lower() and bed.iloc[0, 1].lower() == "start":
            bed = pd.read_csv(bedpath, sep="\t", header=0, dtype=str)
        else:
            logging.info(f"No header in {bedpath}. Extracting columns 1-3 and assigning default column names.")
            bed = bed.iloc[:, :3].copy()
            bed.columns = ["chrom", "start", "end"]
        bed["start"] = pd.to_numeric(bed["start"], errors="coerce")
        bed["end"] = pd.to_numeric(bed["end"], errors="coerce")
        if not bedcheck(bed):
            return None
        bed.columns = bed.columns.str.lower()
        missing = sorted(set(ALL_COLS) - set(bed.columns), reverse=True)  # weight before id: the id uses it
        logging.info(f"Your BED is missing the following columns:{missing}. Trying to fill in default values where possible.")
        for col in missing:
            if col == "weight":
                bed[col] = 1
            elif col == "id":
                bed[col] = (bed["chrom"].astype(str) + "_" + bed["start"].astype(str) + "_" + bed["end"].astype(str)
                            + "_" + bed["weight"].astype(str))
            else:
                bed[col] = [[]] * len(bed)
        bed["weight"] = bed["weight"].apply(lambda x: 1 if _is_missing(x) else x)
        bed["barcode"] = bed["barcode"].apply(lambda x: [] if _is_missing(x) else x)
        if barcoding:
            prefix = "_".join(list(list_of_chromosomes_in_reference)[0].split("_")[:-1])
            bed["chrom"] = prefix + "_" + bed["chrom"].astype(str)
        bed = bed[bed["chrom"].isin(list_of_chromosomes_in_reference)]
    except Exception as e:
        logging.info(f"Not defined or error reading the BED file: {e}")
        if bedpath is not None:
            logging.info(f"Exiting. Please check the file at: {bedpath}")
            print(f"Exiting. Please check the file at: {bedpath}")
            sys.exit()
        return None
    return bed


def chromosome_to_global_coordinates(beddf, input_chromosome_dict):
    """Per-contig BED rows -> {id: {start, end, weight, barcode}} in global coordinates; start == end == 0 selects
    the whole contig (bed_operations.py:80-105)."""
    try:
        out = {}
        for row in beddf.itertuples(index=False):
            row = row._asdict()
            cstart, cend = input_chromosome_dict[row["chrom"]]
            if row["start"] == row["end"] == 0:
                logging.info(f"Start and End coordinates in are both 0. Full {row['chrom']} used for {row['id']}.")
                start, end = cstart, cend
            else:
                start, end = row["start"] + cstart, row["end"] + cstart
            out[row["id"]] = {"start": start, "end": end, "weight": row["weight"], "barcode": row["barcode"]}
        return out
    except Exception:
        logging.warning("BED coordinates could not be updated.")
        return None


def global_to_chromosome_coordinates(chromosome_dict, items_dict):
    """{item: [global start, global end]} -> BED DataFrame chrom/start/end/item; an item is reported on the contig
    whose borders strictly contain its start (bed_operations.py:107-119)."""
    rows = []
    for item, (gs, ge) in items_dict.items():
        for chrom, (cs, ce) in chromosome_dict.items():
            if cs < gs < ce:
                rows.append([chrom, gs - cs, ge - cs, item])
    return pd.DataFrame(rows, columns=["chrom", "start", "end", "item"])
