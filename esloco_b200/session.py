"""Per-process, per-GPU session: one Engine plus a copy of the inputs currently uploaded, so that
run_simulation_iteration can be called once per iteration (the reference's calling pattern, esloco.py:144-152)
without re-uploading targets, mask and length table every time.

Inputs are compared by CONTENT (the arrays that would be uploaded against the arrays that were), never by object
identity: a caller that edits its target dict in place gets the new targets."""
import logging

import numpy as np

from .engine import Engine
from .read_operations import barcode_cdf, barcode_probabilities, generate_read_length_distribution
from .utils import mask_tree_to_arrays, targets_to_arrays

_sessions = {}


def _same(a, b):
    return a is not None and len(a) == len(b) and all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))


class Session:
    def __init__(self, device=0):
        self.engine = Engine(device)
        self.pinned = {}  # page-locked staging tensors for device -> host result copies (BatchResult.cpu)
        self.outputs = {}  # device output buffers reused between chunks of iterations (run_simulation_batch)
        self._sampler = None
        self.invalidate()

    def invalidate(self):
        """Forget what is uploaded: the next configure / set_pool sends everything again."""
        self._targets = None
        self._targets_valid = False
        self._mask = None
        self._bc_key = None
        self._pool_key = None
        self._pool_copy = None
        self.pool_mean = None
        self.keys = None

    def configure(self, genome_size, target_regions, masked_tree, n_barcodes, barcode_weights, length_sampler="quantile"):
        """Upload whatever differs from what the engine holds.  target_regions: the reference's dict or
        utils.TargetColumns.  length_sampler: how native runs turn uniform bits into read lengths -- "quantile" (inverse
        CDF of the length table through a 2048-bin quantile table in shared memory; default, the faster one) or "pool"
        (uniform pick from the table itself: exactly the reference's distribution, read_operations.py:102)."""
        eng = self.engine
        if length_sampler not in ("quantile", "pool"):
            raise ValueError("length_sampler must be 'quantile' or 'pool'")
        if length_sampler != self._sampler:
            eng.set_length_sampler(length_sampler)
            self._sampler = length_sampler
        if eng.genome_size != genome_size:
            eng.set_genome(genome_size)
            self._pool_key = None  # the pool is filtered against the genome size
        bc_key = (n_barcodes, repr(barcode_weights))
        if bc_key != self._bc_key:
            eng.set_barcode_cdf(barcode_cdf(barcode_probabilities(n_barcodes, barcode_weights)))
            self._bc_key = bc_key
            self._targets_valid = False
            self._mask = None
        keys, ts, te, tb = targets_to_arrays(target_regions, n_barcodes)
        self.keys = keys
        # (the library keeps its own copy of the targets and compares against it: no second copy on this side)
        if not (self._targets_valid and eng.targets_equal(ts, te, tb)):
            eng.set_targets(ts, te, tb)
            self._targets_valid = True
        self._targets = (ts, te, tb)  # this call's arrays, for expected_pairs_per_iteration
        mask = mask_tree_to_arrays(masked_tree, n_barcodes)
        if not _same(self._mask, mask):
            eng.set_blocked(*mask)
            self._mask = mask
        eng.prepare()

    def set_pool(self, key, make_pool):
        """Upload the length table unless `key` is the one already resident (key None: the table itself is compared with
        the resident one).  Entries that are not shorter than the
        genome are dropped with a warning: the reference fails (numpy's randint raises "low >= high",
        read_operations.py:103) only when such a length is actually drawn, which for a 1 M-entry log-normal pool on a
        small genome is rare per iteration but certain per pool -- dropping them lets those runs complete, as they
        almost always do in the reference.  The C library itself stays strict (esl_run_batch rejects such a table)."""
        pool = None
        if key is None:
            pool = np.asarray(make_pool())
            same = self._pool_copy is not None and pool.shape == self._pool_copy.shape and np.array_equal(pool, self._pool_copy)
            key = self._pool_key if same and self._pool_key is not None else ("given", object())
        if key != self._pool_key:
            self._pool_copy = pool.copy() if pool is not None else None
            pool = np.asarray(make_pool()) if pool is None else pool
            eng = self.engine
            eng.set_length_table(pool)
            mean, _sd, longest = eng.pool_stats()  # reduced on the device during the upload: no host pass over the table
            G = eng.genome_size
            if G and longest >= G:
                keep = pool < G
                logging.warning(f"{int((~keep).sum())} of {pool.size} read lengths are not shorter than the genome "
                                f"({G}) and were dropped from the length table.")
                eng.set_length_table(pool[keep])
                mean = eng.pool_stats()[0]
            eng.prepare()
            self.pool_mean = mean
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
