"""BED handling for targets and blocked regions (mirrors src/esloco/bed_operations.py)."""
import logging
import math
import sys

import pandas as pd

def bedcheck(bed):
    """chrN / integer / integer in the first three columns (bed_operations.py:6-16)."""
    ok = (bed.iloc[:, 0].astype(str).str.startswith("chr").all()
          and bed.iloc[:, 1].apply(lambda x: isinstance(x, int)).all()
          and bed.iloc[:, 2].apply(lambda x: isinstance(x, int)).all())
    if not ok:
        logging.info("The file does not seem to be in BED format. Make sure the data looks like: chrN integer integer")
    return bool(ok)


def _blank(cell):
    return cell is None or cell == "" or (isinstance(cell, float) and math.isnan(cell))


# Optional BED columns in the order they are completed (the default id is built from the weight, so weight first):
#   name -> (value of a column the file does not have, given the columns so far; replacement for a blank cell or None)
_OPTIONAL = (
    ("weight", lambda cols, n: [1] * n, lambda: 1),
    ("id", lambda cols, n: [f"{c}_{s}_{e}_{w}" for c, s, e, w in zip(cols["chrom"], cols["start"], cols["end"], cols["weight"])], None),
    ("barcode", lambda cols, n: [[] for _ in range(n)], lambda: []),
)


def _parse_bed_text(bedpath):
    """Tab-separated text -> (column names or None, list of rows of str cells); blank lines are skipped."""
    with open(bedpath, "r") as fh:
        rows = [line.rstrip("\r\n").split("\t") for line in fh if line.strip()]
    if not rows:
        raise ValueError("the file holds no rows")
    head = rows[0]
    if len(head) > 1 and "chr" in head[0].lower() and head[1].lower() == "start":
        return head, rows[1:]
    return None, rows


def readbed(bedpath, list_of_chromosomes_in_reference, barcoding=False):
    """Flexible BED reader with the reference's rules (bed_operations.py:28-78) -> DataFrame with at least
    chrom / start / end / id / barcode / weight, or None.
      * header row: first cell contains "chr", second cell is "start" (any case); column names are lower-cased.
        Without a header only the first three columns count (a header-less file loses its id column).
      * start / end must be integers and every chrom must begin with "chr", else the file is not BED: None.
      * columns the file lacks, and blank cells, get defaults: weight 1, barcode [], id "chrom_start_end_weight".
      * barcoding=True prefixes contig names like the barcoded chromosome dict ("Barcode_k_chrN").
      * rows on contigs absent from the reference are dropped.
    A path that cannot be read or parsed ends the program (the reference does the same); bedpath None returns None."""
    if bedpath is None:
        logging.info("No BED path given.")
        return None
    try:
        names, rows = _parse_bed_text(bedpath)
        if names is None:
            logging.info(f"{bedpath} has no header row: using its first three columns as chrom, start, end.")
            names, rows = ["chrom", "start", "end"], [r[:3] for r in rows]
        width = len(names)
        if any(len(r) > width for r in rows):
            raise ValueError("a row has more cells than the header")
        rows = [r + [None] * (width - len(r)) for r in rows]
        if names[1] != "start" or names[2] != "end":
            raise KeyError("the second and third columns must be named start and end")
        cols = {nm: [r[k] if r[k] != "" else None for r in rows] for k, nm in enumerate(names)}
        n = len(rows)
        for nm in ("start", "end"):
            try:
                cols[nm] = [int(c) for c in cols[nm]]
            except (TypeError, ValueError):
                logging.info("The file does not seem to be in BED format. Make sure the data looks like: chrN integer integer")
                return None
        if not all(str(c).startswith("chr") for c in cols[names[0]]):
            logging.info("The file does not seem to be in BED format. Make sure the data looks like: chrN integer integer")
            return None
        cols = {nm.lower(): v for nm, v in cols.items()}
        if "chrom" not in cols:
            raise KeyError("no chrom column")
        absent = [nm for nm, _, _ in _OPTIONAL if nm not in cols]
        logging.info(f"{bedpath}: columns filled with defaults: {absent or 'none'}.")
        for nm, make, blank_value in _OPTIONAL:
            if nm not in cols:
                cols[nm] = make(cols, n)
            elif blank_value is not None:
                cols[nm] = [blank_value() if _blank(c) else c for c in cols[nm]]
        if barcoding:
            prefix = "_".join(list(list_of_chromosomes_in_reference)[0].split("_")[:-1])
            cols["chrom"] = [f"{prefix}_{c}" for c in cols["chrom"]]
        keep = set(list_of_chromosomes_in_reference)
        bed = pd.DataFrame({nm: pd.Series(v, dtype="int64" if nm in ("start", "end") else object) for nm, v in cols.items()})
        return bed[bed["chrom"].isin(keep)]
    except Exception as e:
        logging.info(f"Could not read the BED file {bedpath}: {e}")
        print(f"Stopping: the BED file {bedpath} could not be read ({e}).")
        sys.exit()


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
