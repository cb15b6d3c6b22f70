"""Per-process, per-GPU session: one Engine plus the identity of the inputs currently uploaded, so that
run_simulation_iteration can be called once per iteration (the reference's calling pattern, esloco.py:144-152)
without re-uploading targets, mask and length table every time."""
import numpy as np

from .engine import Engine
from .read_operations import barcode_cdf, barcode_probabilities, generate_read_length_distribution
from .utils import mask_tree_to_arrays, targets_to_arrays

_sessions = {}


class Session:
    def __init__(self, device=0):
        self.engine = Engine(device)
        self.pinned = {}  # page-locked staging tensors for device -> host result copies (BatchResult.cpu)
        self.invalidate()

    def invalidate(self):
        self._targets_key = None
        self._mask_key = None
        self._bc_key = None
        self._pool_key = None
        self.keys = None

    def configure(self, genome_size, target_regions, masked_tree, n_barcodes, barcode_weights):
        eng = self.engine
        if eng.genome_size != genome_size:
            eng.set_genome(genome_size)
        bc_key = (n_barcodes, repr(barcode_weights))
        if bc_key != self._bc_key:
            eng.set_barcode_cdf(barcode_cdf(barcode_probabilities(n_barcodes, barcode_weights)))
            self._bc_key = bc_key
            self._targets_key = self._mask_key = None
        tkey = (id(target_regions), len(target_regions))
        if tkey != self._targets_key:
            self.keys, ts, te, tb = targets_to_arrays(target_regions, n_barcodes)
            eng.set_targets(ts, te, tb)
            self._targets_key = tkey
            self._targets_ref = target_regions  # keep alive: id() is the cache key
        mkey = (id(masked_tree), len(masked_tree) if masked_tree else 0)
        if mkey != self._mask_key:
            eng.set_blocked(*mask_tree_to_arrays(masked_tree, n_barcodes))
            self._mask_key = mkey
            self._mask_ref = masked_tree

    def set_pool(self, key, make_pool):
        """Upload the length table unless `key` is the one already resident."""
        if key != self._pool_key:
            self.engine.set_length_table(make_pool())
            self._pool_key = key


def get_session(device=0):
    if device not in _sessions:
        _sessions[device] = Session(device)
    return _sessions[device]


def lognormal_pool(seed, combo, mean_read_length, sigma, min_read_length, n=1000000):
    """1 M-entry log-normal length pool (read_operations.py:8-27 via combined_calculations.py:44-49), drawn once
    per (seed, combination) on the host.  The reference redraws it every iteration; the pool is the population the
    device samples from, so fixing it changes per-iteration mean read length by <= sd/sqrt(1e6) (~0.13 %)."""
    rng = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, 0x65736C6F636F0000 + combo]))
    return generate_read_length_distribution(n, mean_read_length, min_read_length, sigma, "lognormal", rng)
