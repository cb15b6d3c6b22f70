"""Host helpers of the hot path (mirrors src/esloco/utils.py of the reference where the hot path needs it)."""
import ast
import logging
from collections import namedtuple

import numpy as np

Interval = namedtuple("Interval", "begin end data")


class IntervalSet:
    """Minimal stand-in for intervaltree.IntervalTree (the reference's dependency, utils.py:100): holds the
    blocked intervals of one barcode tree.  The engine only needs to iterate it; `overlap` is provided with the
    package's half-open rule so host code written against the reference keeps working."""

    def __init__(self):
        self._ivs = []

    def addi(self, begin, end, data=None):
        self._ivs.append(Interval(begin, end, data))

    def overlap(self, begin, end):
        return {iv for iv in self._ivs if iv.begin < end and iv.end > begin}

    def __iter__(self):
        return iter(self._ivs)

    def __len__(self):
        return len(self._ivs)


def build_mask_tree(masked_regions):
    """masked_regions dict -> {barcode string | "ALL": IntervalSet} (utils.py:89-101): the BED `barcode` column
    is a Python-literal list of barcode STRINGS, an empty list means every barcode, `weight` rides in .data."""
    trees = {}
    for values in masked_regions.values():
        barcodes = values["barcode"]
        if isinstance(barcodes, str):
            barcodes = ast.literal_eval(barcodes)
        if not barcodes:
            barcodes = ["ALL"]
        for bc in barcodes:
            trees.setdefault(bc, IntervalSet()).addi(values["start"], values["end"], values.get("weight"))
    return trees


def mask_tree_to_arrays(masked_tree, n_barcodes):
    """Flatten a mask tree (ours or a real intervaltree dict) into the arrays esl_set_blocked takes.
    Keys are compared the way the reference does (read_operations.py:61,105): only the STRING form of a barcode
    number, or "ALL", ever matches; anything else is unreachable and dropped."""
    s, e, w, b = [], [], [], []
    if masked_tree:
        for key, tree in masked_tree.items():
            if key == "ALL":
                code = -1
            elif isinstance(key, str) and key.isdigit() and str(int(key)) == key and int(key) < n_barcodes:
                code = int(key)
            else:
                logging.info(f"Blocked-region barcode key {key!r} matches no barcode; ignored.")
                continue
            for iv in tree:
                s.append(int(iv.begin)); e.append(int(iv.end)); w.append(float(iv.data)); b.append(code)
    return (np.array(s, np.int64), np.array(e, np.int64), np.array(w, np.float64), np.array(b, np.int32))


def targets_to_arrays(target_regions, n_barcodes):
    """target dict (I: [start, end] lists; ROI: dicts with start/end, counting.py:24-30) -> keys, start, end and
    the barcode a read must carry: key.split("_")[1] compared as a string with the read label (counting.py:37)."""
    keys = list(target_regions.keys())
    n = len(keys)
    start = np.empty(n, np.int64); end = np.empty(n, np.int64); bc = np.full(n, -1, np.int32)
    for i, k in enumerate(keys):
        v = target_regions[k]
        if isinstance(v, list):
            start[i], end[i] = int(v[0]), int(v[1])
        else:
            start[i], end[i] = int(v["start"]), int(v["end"])
        parts = k.split("_")
        if len(parts) > 1 and parts[1].isdigit() and str(int(parts[1])) == parts[1] and int(parts[1]) < n_barcodes:
            bc[i] = int(parts[1])
    return keys, start, end, bc
