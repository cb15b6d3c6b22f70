"""Loader for the fixtures written by tests/golden/make_golden.py."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.arrays = {k: z[k] for k in z.files}
        self.meta = json.loads(bytes(self.arrays.pop("meta_json")).decode())
        self.name = name

    def __getitem__(self, k):
        return self.arrays[k]

    def combo(self, ci, key):
        return self.arrays[f"c{ci}_{key}"]

    @property
    def mask(self):
        if self["mask_start"].size == 0:
            return None
        return (self["mask_start"], self["mask_end"], self["mask_weight"], self["mask_barcode"])

    @property
    def source_lengths(self):
        return self.arrays.get("fastq_lengths")
