"""Result assembly and table writing (mirrors src/esloco/esloco.py:155-216) -- SURVEY.md section 8(f) row 1.

The reference builds one Python dict per (iteration, target) and splits strings per row with pandas; here the
per-target name columns are derived once from the T keys and tiled over iterations, and the summary comes from the
per-target moment accumulators (device kernel + NCCL reduce) instead of a groupby over all rows."""
import ctypes as C

import numpy as np
import pandas as pd

from .counting import LABELS_BLOCKED
from .distributed import summary_from_moments


_SPLIT_CACHE = [None]  # (keys object, mode, columns) of the last call: the CLI splits the same 240 k keys twice


def split_target_keys(keys, mode):
    """target / barcode / insertion columns from the target keys, exactly the rsplit logic of esloco.py:165-177
    applied to "<key>_<iteration>"."""
    hit = _SPLIT_CACHE[0]
    if hit is not None and hit[0] is keys and hit[1] == mode and len(hit[2]["target"]) == len(keys):
        return hit[2]
    if mode == "ROI":  # {ROI}_{barcode}[_{iteration}]
        parts = [k.rsplit("_", 1) for k in keys]
        cols = {"target": [p[0] for p in parts], "barcode": [p[1] if len(p) > 1 else None for p in parts]}
    else:  # Barcode_{b}_insertion_{i}[_{iteration}]: rsplit n=4 on key+iteration == rsplit n=3 on key
        parts = [k.rsplit("_", 3) for k in keys]
        if all(len(p) == 4 for p in parts):  # the generator's keys: joining the four parts gives the key back
            cols = {"target": list(keys), "barcode": [p[1] for p in parts], "insertion": [p[2] + "_" + p[3] for p in parts]}
        else:
            parts = [[None] * (4 - len(p)) + p for p in parts]
            cols = {"target": ["_".join(str(x) for x in p) for p in parts], "barcode": [p[1] for p in parts],
                    "insertion": [f"{p[2]}_{p[3]}" for p in parts]}
    _SPLIT_CACHE[0] = (keys, mode, cols)
    return cols


def matches_table(keys, mode, combos, iteration_ids, tallies_per_combo, overlaps=None):
    """DataFrame of `{exp}_matches_table.csv`: rows ordered iteration -> combination -> target (esloco.py:155-160).
    tallies_per_combo[c] is an int64 array [n_iter, T, 3]."""
    T, n_iter, n_combo = len(keys), len(iteration_ids), len(combos)
    kcols = split_target_keys(keys, mode)
    keys_arr = np.array(keys, dtype=object)
    blocks = np.stack([np.asarray(t) for t in tallies_per_combo], axis=1)  # [n_iter, n_combo, T, 3]
    flat = blocks.reshape(n_iter * n_combo * T, 3)
    it_col = np.repeat(np.array([str(i) for i in iteration_ids], dtype=object), n_combo * T)
    df = pd.DataFrame({"target_region": np.tile(keys_arr, n_iter * n_combo) + "_" + it_col,
                       "full_matches": flat[:, 0], "partial_matches": flat[:, 1]})
    if overlaps is None:
        df["overlap"] = [[] for _ in range(len(df))]
    else:
        df["overlap"] = overlaps
    df["on_target_bases"] = flat[:, 2]
    df["mean_read_length"] = np.tile(np.repeat(np.array([c[0] for c in combos], dtype=object), T), n_iter)
    df["coverage"] = np.tile(np.repeat(np.array([c[1] for c in combos], dtype=object), T), n_iter)
    for name, vals in kcols.items():
        df[name] = np.tile(np.array(vals, dtype=object), n_iter * n_combo)
    df["iteration"] = it_col
    return df


MATCH_COLUMNS = ["target_region", "full_matches", "partial_matches", "overlap", "on_target_bases", "mean_read_length",
                 "coverage", "target", "barcode", "insertion", "iteration"]
_NEEDS_QUOTING = ("\t", '"', "\n", "\r")


def _field(v):
    """Text pandas.to_csv writes for an object-column value."""
    return "" if v is None else str(v)


class MatchesWriter:
    """Streams `{exp}_matches_table.csv` to disk chunk by chunk (rows: iteration -> combination -> target), so a
    run with 10^8+ rows never holds more than one chunk of iterations in memory.  Each rank of a multi-GPU run writes
    its own contiguous block of iterations to a part file; `merge_parts` concatenates them in rank order.

    Rows go through the native writer (`esl_write_match_rows`): the text of a row that does not change between
    iterations is prepared once per target here, the library formats the three integers next to it straight from
    the tally array (tens of millions of rows per second instead of pandas' ~10^5).  The pandas path of
    `matches_table` remains for per-read overlap lists and for keys that would need CSV quoting; both produce the
    same bytes (tests)."""

    def __init__(self, path, keys, mode, combos, write_header=True, native=True):
        self.path, self.keys, self.mode, self.combos = path, keys, mode, combos
        self.header = write_header
        self._blobs = None
        open(path, "w").close()
        if native:
            self._prepare_native()

    def _prepare_native(self):
        kcols = split_target_keys(self.keys, self.mode)
        names = [c for c in ("target", "barcode", "insertion") if c in kcols]
        texts = [list(map(_field, kcols[c])) for c in names]
        everything = "".join(self.keys) + "".join("".join(col) for col in texts)
        if any(ch in everything for ch in _NEEDS_QUOTING):
            return  # some field needs quoting: let pandas write
        pre = [(k + "_").encode() for k in self.keys]
        pre_off = np.zeros(len(pre) + 1, np.int64)
        np.cumsum([len(b) for b in pre], out=pre_off[1:])
        mids = []
        for mrl, cov in self.combos:
            head = f"\t{_field(mrl)}\t{_field(cov)}\t"
            rows = [(head + "\t".join(vals) + "\t").encode() for vals in zip(*texts)]
            off = np.zeros(len(rows) + 1, np.int64)
            np.cumsum([len(b) for b in rows], out=off[1:])
            mids.append((b"".join(rows), off))
        self._blobs = (b"".join(pre), pre_off, mids)
        self._columns = [c for c in MATCH_COLUMNS if c != "insertion" or "insertion" in kcols]

    def append(self, iteration_ids, tallies_per_combo, overlaps=None):
        if len(iteration_ids) == 0:
            return
        if self._blobs is None or overlaps is not None:
            df = matches_table(self.keys, self.mode, self.combos, iteration_ids, tallies_per_combo, overlaps)
            df.to_csv(self.path, sep="\t", header=self.header, index=False, mode="a")
            self.header = False
            return
        from . import _native
        lib = _native.load()
        if self.header:
            with open(self.path, "a") as fh:
                fh.write("\t".join(self._columns) + "\n")
            self.header = False
        pre_blob, pre_off, mids = self._blobs
        T = len(self.keys)
        blocks = [np.ascontiguousarray(t, dtype=np.int64) for t in tallies_per_combo]  # [n_iter, T, 3] each
        ids = np.ascontiguousarray(iteration_ids, dtype=np.int64)
        n = len(mids)
        # one library call for the whole chunk: it keeps a single open file and loops over iterations itself
        ptr = C.c_void_p * n
        mid_blobs = (C.c_char_p * n)(*[m[0] for m in mids])
        mid_offs = ptr(*[m[1].ctypes.data for m in mids])
        tall = ptr(*[b.ctypes.data for b in blocks])
        rc = lib.esl_write_match_rows_batch(self.path.encode(), 1, pre_blob, pre_off.ctypes.data, n, mid_blobs, mid_offs,
                                            b"[]", tall, T, ids.size, ids.ctypes.data)
        if rc < 0:
            raise OSError(f"esl_write_match_rows_batch failed on {self.path} (status {rc})")

    @staticmethod
    def merge_parts(final_path, part_paths):
        import shutil
        with open(final_path, "wb") as out:
            for part in part_paths:
                with open(part, "rb") as fh:
                    shutil.copyfileobj(fh, out, 16 << 20)


def barcode_distribution_table(n_barcodes, combos, iteration_ids, label_counts_per_combo, n_bases_per_combo,
                               label_columns="barcode", first_appearance=None):
    """DataFrame of `{exp}_barcode_distribution_table.csv` (esloco.py:161-180): one row per (iteration,
    combination); label columns, coverage, mean_read_length, iteration, N_bases, total_Reads.

    label_columns="barcode" (default): all barcode columns, in barcode order, are always present and total_Reads is the
    number of unblocked reads; blocked columns appear only when such reads exist.
    label_columns="reference": the reference's layout, quirks included.  Every row is a dict
    whose label keys come in the order the labels first appear among that iteration's reads (counting.py:63-71) with
    labels that have no read left out; pandas then orders the columns by first appearance across the rows, so a label
    first seen in a later row lands behind the bookkeeping columns, and total_Reads is the sum of the FIRST n_barcodes
    COLUMNS (esloco.py:168,180) whatever they hold.  `first_appearance(row_index, iteration_id, combination)` must
    return that iteration's labels in first-appearance order; it is only asked for rows that bring a label no earlier
    row had (the first row always does) -- for the others the order of the keys cannot influence the columns."""
    names = [str(k) for k in range(n_barcodes)] + list(LABELS_BLOCKED)
    rows = []
    lc = [np.asarray(x) for x in label_counts_per_combo]
    if label_columns == "reference":
        if first_appearance is None:
            raise ValueError("the reference layout needs the first-appearance order of the labels")
        known = []
        for i, it in enumerate(iteration_ids):
            for c, (mrl, cov) in enumerate(combos):
                present = [names[k] for k in range(n_barcodes + 2) if lc[c][i, k]]
                if any(nm not in known for nm in present):
                    order = [nm for nm in first_appearance(len(rows), it, c) if nm in present]
                    if sorted(order) != sorted(present):
                        raise ValueError("first-appearance order does not cover the labels of its iteration")
                    known += [nm for nm in order if nm not in known]
                else:
                    order = [nm for nm in known if nm in present]
                row = {nm: int(lc[c][i, names.index(nm)]) for nm in order}
                row.update(coverage=cov, mean_read_length=mrl, iteration=it, N_bases=int(np.asarray(n_bases_per_combo[c])[i]))
                rows.append(row)
        df = pd.DataFrame(rows)
        df["total_Reads"] = df.iloc[:, :n_barcodes].sum(axis=1)  # esloco.py:168,180
        return df
    if label_columns != "barcode":
        raise ValueError("label_columns must be 'barcode' or 'reference'")
    any_blocked = [any(x[:, n_barcodes + j].any() for x in lc) for j in range(2)]
    for i, it in enumerate(iteration_ids):
        for c, (mrl, cov) in enumerate(combos):
            row = {str(k): int(lc[c][i, k]) for k in range(n_barcodes)}
            for j, name in enumerate(LABELS_BLOCKED):
                if any_blocked[j]:
                    row[name] = int(lc[c][i, n_barcodes + j])
            row.update(coverage=cov, mean_read_length=mrl, iteration=it, N_bases=int(np.asarray(n_bases_per_combo[c])[i]))
            rows.append(row)
    df = pd.DataFrame(rows)
    df["total_Reads"] = df.iloc[:, :n_barcodes].sum(axis=1)
    return df


def summary_table(keys, mode, combos, moments_per_combo):
    """DataFrame of `{exp}_summary.csv` from per-target moment accumulators [T, 8] per combination
    (esloco.py:183-195): groups are (target, mean_read_length, coverage), sorted like pandas groupby; in ROI mode a
    target's rows of all barcodes fall into one group, so their accumulators are added."""
    names = split_target_keys(keys, mode)["target"]
    frames = []
    index = {}
    for nm in names:
        if nm not in index:
            index[nm] = len(index)
    order = list(index)
    for (mrl, cov), mom in zip(combos, moments_per_combo):
        mom = np.asarray(mom).astype(object)
        if len(order) == len(names):  # every target its own group (I mode): nothing to add up
            grouped = mom
        else:
            grouped = np.zeros((len(order), 8), dtype=object)
            np.add.at(grouped, np.fromiter((index[nm] for nm in names), np.int64, len(names)), mom)
            lo_hi = grouped[:, 7] * (1 << 32) + grouped[:, 6]  # re-normalise the split square sum after adding rows
            grouped[:, 6] = [int(x) & 0xFFFFFFFF for x in lo_hi]
            grouped[:, 7] = [int(x) >> 32 for x in lo_hi]
        cols = summary_from_moments(grouped)
        df = pd.DataFrame({"target": order, "mean_read_length": mrl, "coverage": cov})
        for k in ("mean_full_matches", "sd_full_matches", "mean_partial_matches", "sd_partial_matches",
                  "mean_bases_on_target", "sd_bases_on_target", "sem_bases_on_target",
                  "ci_lower_bases_on_target", "ci_upper_bases_on_target"):
            df[k] = cols[k]
        frames.append(df)
    out = pd.concat(frames, ignore_index=True)
    return out.sort_values(["target", "mean_read_length", "coverage"], kind="stable").reset_index(drop=True)


def write_tables(output_path, experiment_name, matches=None, barcode_distribution=None, summary=None, insertion_bed=None):
    """Tab-separated, header, no index (esloco.py:198-216)."""
    base = f"{output_path}{experiment_name}"
    if insertion_bed is not None:
        insertion_bed.to_csv(f"{base}_insertion_locations.bed", sep="\t", header=True, index=False)
    if barcode_distribution is not None:
        barcode_distribution.to_csv(f"{base}_barcode_distribution_table.csv", sep="\t", header=True, index=False)
    if matches is not None:
        matches.to_csv(f"{base}_matches_table.csv", sep="\t", header=True, index=False)
    if summary is not None:
        summary.to_csv(f"{base}_summary.csv", sep="\t", header=True, index=False)
