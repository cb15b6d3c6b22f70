"""Host helpers of the hot path (mirrors src/esloco/utils.py of the reference where the hot path needs it)."""
import ast
import logging
import itertools
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


def _target_barcode(token, n_barcodes):
    """Barcode a target key's second `_` token stands for, or -1 (counting.py:37 compares it with the read label as a
    string, so only the canonical decimal spelling of a barcode number can ever match)."""
    if token is not None and token.isdigit() and str(int(token)) == token and int(token) < n_barcodes:
        return int(token)
    return -1


class TargetColumns:
    """Columnar form of the target dict for large target sets (config 3: 240 k insertions).  The reference's dict
    `{key: [start, end]}` / `{key: {"start", "end", ...}}` (counting.py:19-30) costs one Python object walk per call;
    callers that already hold arrays hand them over as they are: start / end = half-open interval per target in row
    order, barcode = the barcode a read must carry to match (-1: matches nothing), keys = the dict keys (optional, only
    needed where names are printed).  `from_dict` converts a reference-shaped dict once."""

    def __init__(self, start, end, barcode, keys=None):
        self.start = np.ascontiguousarray(start, np.int64)
        self.end = np.ascontiguousarray(end, np.int64)
        self.barcode = np.ascontiguousarray(barcode, np.int32)
        if not (self.start.shape == self.end.shape == self.barcode.shape and self.start.ndim == 1):
            raise ValueError("start, end and barcode must be 1-D arrays of one length")
        if keys is not None and len(keys) != self.start.size:
            raise ValueError("keys and columns differ in length")
        self.keys = keys

    @classmethod
    def from_dict(cls, target_regions, n_barcodes):
        keys, s, e, b = targets_to_arrays(target_regions, n_barcodes)
        return cls(s, e, b, keys)

    def __len__(self):
        return int(self.start.size)


def targets_to_arrays(target_regions, n_barcodes):
    """target dict (I: [start, end] lists; ROI: dicts with start/end, counting.py:24-30) -> keys, start, end and
    the barcode a read must carry: key.split("_")[1] compared as a string with the read label (counting.py:37).
    Written for dicts of 10^5..10^6 targets (config 3: 240 k): one pass per column, the barcode rule evaluated once per
    distinct token."""
    if isinstance(target_regions, TargetColumns):
        return target_regions.keys, target_regions.start, target_regions.end, target_regions.barcode
    keys = list(target_regions.keys())
    n = len(keys)
    vals = list(target_regions.values())
    se = None
    if n and set(map(type, vals)) == {list} and set(map(len, vals)) == {2}:
        try:  # I mode: every value is [start, end] -- one C-level pass over all of them
            se = np.fromiter(itertools.chain.from_iterable(vals), np.int64, 2 * n).reshape(n, 2)
        except (TypeError, ValueError, OverflowError):
            se = None
    if se is None:
        se = np.empty((n, 2), np.int64)
        for i, v in enumerate(vals):
            if isinstance(v, list):
                se[i, 0], se[i, 1] = int(v[0]), int(v[1])
            else:
                se[i, 0], se[i, 1] = int(v["start"]), int(v["end"])
    start, end = np.ascontiguousarray(se[:, 0]), np.ascontiguousarray(se[:, 1])
    tokens = [k.split("_", 2)[1] if "_" in k else None for k in keys]
    rule = {t: _target_barcode(t, n_barcodes) for t in set(tokens)}
    bc = np.fromiter((rule[t] for t in tokens), np.int32, n)
    return keys, start, end, bc
