"""ctypes binding of libesloco_b200.so (include/esloco_b200.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESL_LIB") or os.path.join(_HERE, "libesloco_b200.so")  # ESL_LIB: tuning builds only

ESL_OK = 0
ESL_ITER_CUT_IN_BULK = 1
ESL_ITER_EXHAUSTED = 2
ESL_MOMENT_COLS = 8
ESL_SAMPLER_POOL, ESL_SAMPLER_QUANTILE = 0, 1
LABEL_BLOCKED_CONSUMING = -1
LABEL_BLOCKED_NOTCONSUMING = -2

# every symbol include/esloco_b200.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = {
    "esl_abi_version": (C.c_int, []),
    "esl_last_error": (C.c_char_p, [C.c_void_p]),
    "esl_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "esl_destroy": (C.c_int, [C.c_void_p]),
    "esl_set_genome": (C.c_int, [C.c_void_p, C.c_uint64]),
    "esl_set_barcode_cdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "esl_set_targets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "esl_targets_equal": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "esl_set_blocked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "esl_set_length_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "esl_set_index_builder": (C.c_int, [C.c_void_p, C.c_int32]),
    "esl_index_built_on_device": (C.c_int32, [C.c_void_p]),
    "esl_set_length_sampler": (C.c_int, [C.c_void_p, C.c_int32]),
    "esl_pool_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32)]),
    "esl_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esl_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_double, C.c_int64]),
    "esl_run_batch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_double, C.c_int32,
                                C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "esl_staged_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_double]),
    "esl_run_batch_staged": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_double, C.c_int32,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "esl_replay_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                   C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "esl_replay_draws": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "esl_materialize_reads": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esl_coverage_bins": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
    "esl_replay_thin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "esl_moments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "esl_scan_read_lengths": (C.c_int64, [C.c_char_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_double)]),
    "esl_scan_fasta_contigs": (C.c_int64, [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "esl_shift_insertions": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "esl_write_match_rows": (C.c_int64, [C.c_char_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_char_p]),
    "esl_write_match_rows_batch": (C.c_int64, [C.c_char_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                               C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "esl_plan_target_layers": (C.c_int32, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "esl_set_hit_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "esl_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "esl_get_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "esl_launch_count": (C.c_int64, [C.c_void_p]),
    "esl_target_layers": (C.c_int32, [C.c_void_p]),
    "esl_bulk_reads": (C.c_int64, [C.c_void_p, C.c_double]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load the CUDA library.  No fallback: a missing library is an error, never a silent CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m esloco_b200.build` "
            "(or __graft_entry__.build()); esloco_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = library / header mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
