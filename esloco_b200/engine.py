"""Per-GPU engine: owns one esl_ctx and the torch-allocated device buffers passed through the C ABI.

PyTorch is used for device memory and streams only; every computation happens in libesloco_b200.so.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"esloco_b200 error {code}: {msg}")
        self.code = code


@dataclass
class BatchResult:
    """Device-resident outputs of one batch of iterations (int64 torch tensors)."""
    tallies: "torch.Tensor"       # [n_iter, T, 3] full, partial, on_target_bases
    label_counts: "torch.Tensor"  # [n_iter, B+2]
    n_bases: "torch.Tensor"       # [n_iter]
    n_reads: "torch.Tensor"       # [n_iter]
    status: "torch.Tensor"        # [n_iter] int32
    overlaps = None               # optional: per iteration, per target, list of overlaps (run_simulation_batch)

    def cpu(self, pinned=None):
        """Host copy.  `pinned`: optional dict used as a cache of page-locked staging tensors (keyed by field and
        shape); with it the copies run at PCIe speed instead of through pageable memory -- it matters once the
        tallies are hundreds of MB (240 k targets x 125 iterations = 720 MB)."""
        import torch
        fields = ("tallies", "label_counts", "n_bases", "n_reads", "status")
        if pinned is None or not self.tallies.is_cuda:
            return BatchResult(*(getattr(self, f).cpu() for f in fields))
        out = []
        for f in fields:
            t = getattr(self, f)
            key = (f, tuple(t.shape), t.dtype)
            if key not in pinned:
                pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            pinned[key].copy_(t, non_blocking=True)
            out.append(pinned[key])
        torch.cuda.current_stream(self.tallies.device).synchronize()
        return BatchResult(*out)


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    def __init__(self, device=0):
        import torch  # device buffers / streams only
        self.torch = torch
        self.lib = _native.load()
        if not torch.cuda.is_available():
            raise EngineError(-3, "no CUDA device visible to torch; esloco_b200 has no CPU fallback")
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        ctx = C.c_void_p()
        rc = self.lib.esl_create(C.byref(ctx), self.device)
        if rc != 0:
            raise EngineError(rc, self.lib.esl_last_error(None).decode())
        self.ctx = ctx
        self.n_targets = 0
        self.n_barcodes = 0
        self.genome_size = 0
        self._ws = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.esl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError(rc, self.lib.esl_last_error(self.ctx).decode())

    # ---- one-off inputs (host arrays) ----------------------------------------------------------------
    def set_genome(self, genome_size):
        self._check(self.lib.esl_set_genome(self.ctx, int(genome_size)))
        self.genome_size = int(genome_size)

    def set_barcode_cdf(self, cdf):
        cdf = _np(cdf, np.float64)
        self._check(self.lib.esl_set_barcode_cdf(self.ctx, cdf.ctypes.data, cdf.size))
        self.n_barcodes = int(cdf.size)

    def set_targets(self, start, end, barcode):
        s, e, b = _np(start, np.int64), _np(end, np.int64), _np(barcode, np.int32)
        if not (s.size == e.size == b.size):
            raise ValueError("target arrays differ in length")
        self._check(self.lib.esl_set_targets(self.ctx, s.ctypes.data, e.ctypes.data, b.ctypes.data, s.size))
        self.n_targets = int(s.size)

    def targets_equal(self, start, end, barcode):
        """True when the arrays equal the targets the context already holds (compared inside the library)."""
        s, e, b = _np(start, np.int64), _np(end, np.int64), _np(barcode, np.int32)
        if not (s.size == e.size == b.size):
            return False
        return bool(self.lib.esl_targets_equal(self.ctx, s.ctypes.data, e.ctypes.data, b.ctypes.data, s.size))

    def set_blocked(self, start=(), end=(), weight=(), barcode=()):
        s, e, w, b = _np(start, np.int64), _np(end, np.int64), _np(weight, np.float64), _np(barcode, np.int32)
        if not (s.size == e.size == w.size == b.size):
            raise ValueError("blocked arrays differ in length")
        self._check(self.lib.esl_set_blocked(self.ctx, s.ctypes.data, e.ctypes.data, w.ctypes.data, b.ctypes.data, s.size))

    def set_length_table(self, lengths):
        a = np.asarray(lengths)
        if a.dtype != np.uint32:  # (uint32 input is range-checked by the library itself: no extra passes over it)
            if a.size and (a.min() < 1 or a.max() > 0xFFFFFFFF):
                raise ValueError("read lengths must be in [1, 2^32)")
        a = _np(a, np.uint32)
        self._check(self.lib.esl_set_length_table(self.ctx, a.ctypes.data, a.size, self._stream()))

    def set_length_sampler(self, mode):
        """"pool" (uniform pick from the length table, the reference's exact distribution) or "quantile" (inverse CDF
        through a 2048-bin quantile table in shared memory); native mode only, takes effect with the next run."""
        code = {"pool": _native.ESL_SAMPLER_POOL, "quantile": _native.ESL_SAMPLER_QUANTILE}[mode]
        self._check(self.lib.esl_set_length_sampler(self.ctx, code))

    def set_index_builder(self, mode):
        """"auto" (device for >= 16384 targets with 32-bit coordinates, host otherwise), "host" or "device"."""
        self._check(self.lib.esl_set_index_builder(self.ctx, {"auto": 0, "host": 1, "device": 2}[mode]))

    def index_built_on_device(self):
        self.prepare()
        return bool(self.lib.esl_index_built_on_device(self.ctx))

    def pool_stats(self):
        """(mean, sd, longest) of the resident length table, reduced on the device when it was uploaded."""
        m, sd, mx = C.c_double(), C.c_double(), C.c_uint32()
        self._check(self.lib.esl_pool_stats(self.ctx, C.byref(m), C.byref(sd), C.byref(mx)))
        return m.value, sd.value, mx.value

    def prepare(self):
        """Build / upload the device index for the inputs set so far (no-op when nothing changed).  The run methods
        call it themselves, so the C-level rule "run calls never allocate or block" costs Python callers nothing."""
        self._check(self.lib.esl_prepare(self.ctx, self._stream()))

    # ---- device buffers ------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.tdev)
        return self._ws

    def alloc_outputs(self, n_iter):
        t = self.torch
        return BatchResult(
            t.empty((n_iter, self.n_targets, 3), dtype=t.int64, device=self.tdev),
            t.empty((n_iter, self.n_barcodes + 2), dtype=t.int64, device=self.tdev),
            t.empty((n_iter,), dtype=t.int64, device=self.tdev),
            t.empty((n_iter,), dtype=t.int64, device=self.tdev),
            t.empty((n_iter,), dtype=t.int32, device=self.tdev))

    def _dev(self, a, dtype):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a.to(device=self.tdev, dtype=dtype).contiguous()
        return t.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.tdev)

    # ---- hot path ------------------------------------------------------------------------------------
    def run_batch(self, seed, combo, iter_begin, n_iter, coverage, consuming=False, min_overlap=1, scaling=1.0, out=None):
        """Native Philox mode: iterations [iter_begin, iter_begin+n_iter) of one combination.  Asynchronous."""
        self.prepare()
        out = out or self.alloc_outputs(n_iter)
        need = self.lib.esl_workspace_bytes(self.ctx, n_iter, float(coverage), 0)
        ws = self._workspace(max(need, 1))
        self._check(self.lib.esl_run_batch(
            self.ctx, int(seed) & 0xFFFFFFFFFFFFFFFF, int(combo), int(iter_begin), int(n_iter), float(coverage),
            int(bool(consuming)), int(min_overlap), float(scaling), out.tallies.data_ptr(), out.label_counts.data_ptr(),
            out.n_bases.data_ptr(), out.n_reads.data_ptr(), out.status.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    def run_batch_staged(self, seed, combo, iter_begin, n_iter, coverage, consuming=False, min_overlap=1, out=None):
        """The staged baseline pipeline (sort + sweep); same draws and outputs as run_batch, slower by design."""
        self.prepare()
        out = out or self.alloc_outputs(n_iter)
        need = self.lib.esl_staged_workspace_bytes(self.ctx, float(coverage))
        if need < 0:
            self._check(int(need))
        ws = self._workspace(max(need, 1))
        self._check(self.lib.esl_run_batch_staged(
            self.ctx, int(seed) & 0xFFFFFFFFFFFFFFFF, int(combo), int(iter_begin), int(n_iter), float(coverage),
            int(bool(consuming)), int(min_overlap), out.tallies.data_ptr(), out.label_counts.data_ptr(),
            out.n_bases.data_ptr(), out.n_reads.data_ptr(), out.status.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    def replay_reads(self, start, length, label, seg_offsets, min_overlap=1, out=None):
        """Replay level 1: recorded reads, iteration i = reads [seg_offsets[i], seg_offsets[i+1])."""
        t = self.torch
        seg = np.asarray(seg_offsets, np.int64)
        n_iter = seg.size - 1
        self.prepare()
        out = out or self.alloc_outputs(n_iter)
        ds, dl, dlab = self._dev(start, t.int64), self._dev(length, t.int64), self._dev(label, t.int32)
        dseg = self._dev(seg, t.int64)
        self._keep = (ds, dl, dlab, dseg)
        max_reads = int(np.max(np.diff(seg))) if n_iter else 0
        ws = self._workspace(max(self.lib.esl_workspace_bytes(self.ctx, n_iter, 0.0, max(max_reads, 1)), 1))
        self._check(self.lib.esl_replay_reads(
            self.ctx, ds.data_ptr(), dl.data_ptr(), dlab.data_ptr(), dseg.data_ptr(), int(seg[-1]), max_reads, n_iter,
            int(min_overlap), out.tallies.data_ptr(), out.label_counts.data_ptr(), out.n_bases.data_ptr(),
            out.n_reads.data_ptr(), out.status.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    def replay_draws(self, u_barcode, length, start, seg_offsets, coverage, consuming=False, min_overlap=1,
                     ublock_off=None, ublock=None, bulk_draws=0, out=None):
        """Replay level 2: recorded raw draws of an oversampled batch."""
        t = self.torch
        seg = np.asarray(seg_offsets, np.int64)
        n_iter = seg.size - 1
        max_draws = int(np.max(np.diff(seg))) if n_iter else 0
        self.prepare()
        out = out or self.alloc_outputs(n_iter)
        du, dl, ds = self._dev(u_barcode, t.float64), self._dev(length, t.int64), self._dev(start, t.int64)
        dseg = self._dev(seg, t.int64)
        duo = self._dev(ublock_off, t.int64) if ublock_off is not None else None
        dub = self._dev(ublock if ublock is not None and len(ublock) else np.zeros(1), t.float64) if ublock_off is not None else None
        self._keep = (du, dl, ds, dseg, duo, dub)
        need = self.lib.esl_workspace_bytes(self.ctx, n_iter, float(coverage), max(max_draws, 1))
        ws = self._workspace(max(need, 1))
        self._check(self.lib.esl_replay_draws(
            self.ctx, du.data_ptr(), dl.data_ptr(), ds.data_ptr(), duo.data_ptr() if duo is not None else None,
            dub.data_ptr() if dub is not None else None, dseg.data_ptr(), max_draws, int(bulk_draws), n_iter,
            float(coverage), int(bool(consuming)), int(min_overlap), out.tallies.data_ptr(), out.label_counts.data_ptr(),
            out.n_bases.data_ptr(), out.n_reads.data_ptr(), out.status.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return out

    def materialize_reads(self, seed, combo, iteration, consuming, n_reads):
        t = self.torch
        n_reads = int(n_reads)
        self.prepare()
        s = t.empty(n_reads, dtype=t.int64, device=self.tdev)
        ln = t.empty(n_reads, dtype=t.int64, device=self.tdev)
        lab = t.empty(n_reads, dtype=t.int32, device=self.tdev)
        self._check(self.lib.esl_materialize_reads(self.ctx, int(seed) & 0xFFFFFFFFFFFFFFFF, int(combo), int(iteration),
                                                   int(bool(consuming)), n_reads, s.data_ptr(), ln.data_ptr(),
                                                   lab.data_ptr(), self._stream()))
        return s, ln, lab

    def set_hit_buffer(self, capacity):
        """Collect {iteration, row, draw index, overlap} records of the following batches (None/0 = off)."""
        t = self.torch
        if not capacity:
            self._hits = self._hit_count = None
            self._check(self.lib.esl_set_hit_buffer(self.ctx, None, 0, None))
            return
        self._hits = t.empty((int(capacity), 4), dtype=t.int64, device=self.tdev)
        self._hit_count = t.zeros(1, dtype=t.int64, device=self.tdev)
        self._check(self.lib.esl_set_hit_buffer(self.ctx, self._hits.data_ptr(), int(capacity), self._hit_count.data_ptr()))

    def take_hits(self, n_reads, keep_kind=False):
        """Records collected since the last call, draws beyond each iteration's cut-off removed, sorted by
        (iteration, row, draw index): int64 array [k, 4].  n_reads = the batch's reads-placed counts."""
        count = int(self._hit_count.item())
        if count > self._hits.shape[0]:
            raise EngineError(-5, f"hit buffer overflow: {count} pairs, capacity {self._hits.shape[0]}")
        h = self._hits[:count].cpu().numpy()
        self._hit_count.zero_()
        nr = np.asarray(n_reads.cpu() if hasattr(n_reads, "cpu") else n_reads, dtype=np.int64)
        h = h[h[:, 2] < nr[h[:, 0]]]
        h = h[np.lexsort((h[:, 2], h[:, 1], h[:, 0]))]
        if not keep_kind:
            h[:, 3] &= 0x7FFFFFFFFFFFFFFF  # the sign bit of the overlap column marks full matches
        return h

    def replay_thin(self, hits, uniforms, scaling, n_iter):
        """Tallies [n_iter, T, 3] of the pairs the reference keeps for scaling < 1: hits = take_hits(..., keep_kind=True)
        of a scaling-1 replay, uniforms = the reference's per-pair draws in that order."""
        t = self.torch
        dh = self._dev(hits, t.int64)
        du = self._dev(uniforms, t.float64)
        out = t.empty((n_iter, self.n_targets, 3), dtype=t.int64, device=self.tdev)
        self._check(self.lib.esl_replay_thin(self.ctx, dh.data_ptr(), du.data_ptr(), int(dh.shape[0]), float(scaling),
                                             int(n_iter), out.data_ptr(), self._stream()))
        return out

    def coverage_bins(self, start, length, label, bin_size):
        """[n_barcodes+2, n_bins] int64 read bases per bin and label slot for device-resident reads."""
        t = self.torch
        n_bins = -(-self.genome_size // int(bin_size))
        bins = t.empty((self.n_barcodes + 2, n_bins), dtype=t.int64, device=self.tdev)
        self._check(self.lib.esl_coverage_bins(self.ctx, start.data_ptr(), length.data_ptr(), label.data_ptr(),
                                               int(start.numel()), int(bin_size), n_bins, bins.data_ptr(), self._stream()))
        return bins

    def moments(self, tallies, acc=None):
        """acc[T, 8] += per-target moment sums over the iterations in `tallies` (operand of the NCCL reduce)."""
        t = self.torch
        if acc is None:
            acc = t.zeros((self.n_targets, _native.ESL_MOMENT_COLS), dtype=t.int64, device=self.tdev)
        self._check(self.lib.esl_moments(self.ctx, tallies.data_ptr(), int(tallies.shape[0]), acc.data_ptr(), self._stream()))
        return acc

    # ---- introspection -------------------------------------------------------------------------------
    def set_profiling(self, on=True):
        self._check(self.lib.esl_set_profiling(self.ctx, int(bool(on))))

    def last_timing_ms(self):
        """(bulk kernel ms, cut-off kernel ms) of the last profiled batch; waits for it to finish."""
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.esl_get_timing(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return int(self.lib.esl_launch_count(self.ctx))

    def target_layers(self):
        self.prepare()
        return int(self.lib.esl_target_layers(self.ctx))

    def bulk_reads(self, coverage):
        return int(self.lib.esl_bulk_reads(self.ctx, float(coverage)))
