/*
 * esloco_b200.h -- C ABI of the B200-native engine for esloco's Monte-Carlo hot path
 * (read placement + per-target / per-barcode tally).
 *
 * The reference (aweich/esloco v1.0.6) is pure Python and has no FFI of its own; the seam this
 * library sits behind is the Python call
 *     run_simulation_iteration(iteration_id, param_dictionary, genome_size, target_regions,
 *                              masked_tree, log_file)          src/esloco/combined_calculations.py:97-126
 * and its two inner calls
 *     generate_reads_based_on_coverage(...)                     src/esloco/read_operations.py:86-123
 *     count_matches(insertion_dict, read_dir, scaling, min_overlap)   src/esloco/counting.py:7-61
 * Each entry point below cites the reference lines whose work it takes over.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every call returns 0 (ESL_OK) or a negative esl_status; no exceptions cross;
 *     esl_last_error(ctx) gives the message of the last failing call on that context.
 *   - SET-UP calls: esl_set_* take HOST pointers (small, one-off inputs) and only record them (esl_set_length_table
 *     also uploads and reduces the table); esl_prepare builds the device-resident index from what was set.  Set-up
 *     calls may allocate device memory and may block on the stream they are given.
 *   - RUN calls: esl_run_batch / esl_replay_* / esl_moments / esl_materialize_reads / esl_coverage_bins take DEVICE
 *     pointers for bulk data and outputs, are asynchronous on the given CUDA stream (a cudaStream_t passed as
 *     void*), never allocate and never block.  They fail with ESL_ERR_STATE when an input changed since the last
 *     esl_prepare.  Scratch is a caller-provided workspace of esl_workspace_bytes().  (The staged baseline
 *     esl_run_batch_staged is a measuring stick, not a product call: its first use uploads the targets once more
 *     in caller order.)
 *   - one host thread per context at a time.
 *   - there is no CPU fallback: esl_create fails if no sm_100 device is present.
 *
 * Read labels: barcode k >= 0, ESL_LABEL_BLOCKED_CONSUMING, ESL_LABEL_BLOCKED_NOTCONSUMING
 * (the reference's "Blocked/Consuming" / "Blocked/NotConsuming" dict-key suffixes, read_operations.py:113,116).
 *
 * Output layouts (row-major, int64):
 *   tallies      [n_iter][n_targets][3] = {full_matches, partial_matches, on_target_bases}  counting.py:54-57
 *   label_counts [n_iter][n_barcodes+2] = reads per barcode, then Blocked/Consuming, Blocked/NotConsuming
 *                                                                                 counting.py:63-71
 *   n_bases      [n_iter]  covered_length ("N_bases", combined_calculations.py:81)
 *   n_reads      [n_iter]  reads placed = entries of the reference's read dict (blocked ones included)
 *   status       [n_iter]  int32 bit mask, see ESL_ITER_*
 */
#ifndef ESLOCO_B200_H
#define ESLOCO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct esl_ctx esl_ctx;

typedef enum {
    ESL_OK = 0,
    ESL_ERR_INVALID = -1,    /* bad argument (message says which) */
    ESL_ERR_CUDA = -2,       /* a CUDA runtime call failed */
    ESL_ERR_NO_DEVICE = -3,  /* no CUDA device, or not compute capability 10.x */
    ESL_ERR_STATE = -4,      /* a required esl_set_* call is missing */
    ESL_ERR_WORKSPACE = -5,  /* workspace too small */
    ESL_ERR_UNSUPPORTED = -6 /* valid in the reference but not supported here (message says what) */
} esl_status;

#define ESL_LABEL_BLOCKED_CONSUMING (-1)
#define ESL_LABEL_BLOCKED_NOTCONSUMING (-2)

/* per-iteration status bits */
#define ESL_ITER_CUT_IN_BULK 1 /* informational: coverage was reached inside the bulk range; corrected on device */
#define ESL_ITER_EXHAUSTED 2   /* replay only: the recorded draws ran out before coverage was reached */

/* Version of this ABI (bumped on any signature change). */
int esl_abi_version(void);

/* Message of the last error raised by esl_create (ctx == NULL) or by a call on ctx. */
const char *esl_last_error(const esl_ctx *ctx);

/* One context per GPU.  Fails with ESL_ERR_NO_DEVICE when `device` is not a compute-capability-10.x GPU. */
int esl_create(esl_ctx **ctx, int device);
int esl_destroy(esl_ctx *ctx);

/* genome_size of the concatenated coordinate space (fasta_operations.py:3-22 -> esloco.py:107,124).
 * Genomes >= 2^32 switch every kernel to 64-bit coordinates. */
int esl_set_genome(esl_ctx *ctx, uint64_t genome_size);

/* Barcode CDF exactly as numpy's choice(p=) builds it from read_operations.py:96-97 (cumsum, /= last). */
int esl_set_barcode_cdf(esl_ctx *ctx, const double *cdf, int32_t n_barcodes);

/* Targets in the reference's dict order (row order of `tallies`): half-open [start, end), and the barcode a
 * read must carry to match (key.split("_")[1], counting.py:37); barcode < 0 = matches nothing. */
int esl_set_targets(esl_ctx *ctx, const int64_t *start, const int64_t *end, const int32_t *barcode,
                    int64_t n_targets);

/* 1 when the arrays equal the targets the context holds (a caller that is handed the same target set again, as the
 * reference's per-iteration call pattern does, can skip esl_set_targets and the index rebuild it triggers). */
int32_t esl_targets_equal(const esl_ctx *ctx, const int64_t *start, const int64_t *end, const int32_t *barcode, int64_t n_targets);

/* Blocked intervals of build_mask_tree (utils.py:89-101): barcode >= 0 for that barcode's tree, -1 for "ALL".
 * A read [s, s+L) hits interval k iff start[k] < s+L and end[k] > s (read_operations.py:63); each visited hit
 * draws u and blocks when u < weight[k] (:65).  n = 0 clears the mask. */
int esl_set_blocked(esl_ctx *ctx, const int64_t *start, const int64_t *end, const double *weight,
                    const int32_t *barcode, int64_t n);

/* Read-length pool: the source distribution after the `> min_read_length` filter (read_operations.py:24-25,
 * config_handler.py:44).  A read length is a uniform pick from it (combined_calculations.py:57 +
 * read_operations.py:102 collapse to that).  Host memory; copied to the device and reduced there (mean, sd, max).
 * The device index of small target sets carries bounds that assume the longest read, so a table whose longest
 * read exceeds every earlier one (rounded up to a power of two) makes the index stale: call esl_prepare again.
 * Upload and reduction are ordered on `stream`; the call waits for them (the bulk range is planned on the host
 * from mean / sd / max). */
int esl_set_length_table(esl_ctx *ctx, const uint32_t *lengths, int64_t n, void *stream);

/* How native-mode runs turn 24 uniform bits into a read length (the north star's "inverse-CDF look-up into a
 * shared-memory-staged table"; replay modes are fed their lengths and ignore this).
 *   ESL_SAMPLER_POOL      uniform pick from the length table -- exactly the reference's distribution
 *                         (read_operations.py:102 on combined_calculations.py:57); one random 4-byte gather per read.
 *   ESL_SAMPLER_QUANTILE  inverse CDF of the table through a 2048-bin quantile table staged in shared memory:
 *                         11 bits pick the bin, 13 bits interpolate linearly between its two quantiles; the top bin
 *                         is picked exactly from the sorted table.  The sampled CDF is within 1/2048 of the table's
 *                         everywhere (tests: KS distance and mean), and no read is longer than the table's longest.
 * Takes effect with the next esl_prepare (which sorts the resident table on the device and builds the quantiles). */
#define ESL_SAMPLER_POOL 0
#define ESL_SAMPLER_QUANTILE 1
int esl_set_length_sampler(esl_ctx *ctx, int32_t mode);

/* Who builds the target index in esl_prepare.  ESL_INDEX_AUTO (default): the device for sets of >= 16384 targets with
 * 32-bit coordinates -- a radix sort by (barcode, start) and two parallel passes write the records and bucket grids,
 * no host work beyond the upload of the raw targets -- the host otherwise.  The device builder handles one layer per
 * barcode; when it meets nesting (an end below its predecessor's) or hot targets it hands the set to the host
 * builder, which handles every geometry.  ESL_INDEX_HOST / ESL_INDEX_DEVICE force one (the latter still falls back).
 * Both builders produce indexes that give identical tallies. */
#define ESL_INDEX_AUTO 0
#define ESL_INDEX_HOST 1
#define ESL_INDEX_DEVICE 2
int esl_set_index_builder(esl_ctx *ctx, int32_t mode);
/* 1 when the current target index came from the device builder (after esl_prepare). */
int32_t esl_index_built_on_device(const esl_ctx *ctx);

/* Statistics of the resident length table as the device reduced them (any pointer may be NULL). */
int esl_pool_stats(const esl_ctx *ctx, double *mean, double *sd, uint32_t *longest);

/* Build and upload the device index (target layers + bucket grids per barcode, blocked groups, CDF tables) for the
 * inputs set so far; a no-op when nothing changed since the last call.  This is where build_mask_tree's trees
 * (utils.py:89-101) and the target dict of count_matches (counting.py:19-30) become device structures.  Uploads are
 * ordered on `stream`, so run calls issued afterwards on the same stream see the new index.  Required before the
 * first run call and after any esl_set_* that changes genome, CDF, targets or blocked intervals. */
int esl_prepare(esl_ctx *ctx, void *stream);

/* Scratch bytes esl_run_batch / esl_replay_draws need for n_iter iterations at `coverage`
 * (for replay pass the largest per-iteration draw count in max_draws_per_iter, else 0). */
int64_t esl_workspace_bytes(const esl_ctx *ctx, int32_t n_iter, double coverage, int64_t max_draws_per_iter);

/* Native mode: run iterations [iter_begin, iter_begin + n_iter) of combination `combo` with the counter-based
 * Philox4x32-10 stream keyed by `seed`; results do not depend on how iterations are split into batches or
 * across GPUs.  Replaces run_simulation_iteration's per-combination body (combined_calculations.py:56-84):
 * placement until covered_length >= coverage * genome_size (read_operations.py:100), blocked-region test,
 * label histogram and target tally.  scaling < 1 thins qualifying (target, read) pairs (counting.py:46,51).
 * Outputs are zeroed by the call. */
int esl_run_batch(esl_ctx *ctx, uint64_t seed, uint32_t combo, uint32_t iter_begin, int32_t n_iter,
                  double coverage, int32_t consuming, int64_t min_overlap, double scaling, int64_t *d_tallies,
                  int64_t *d_label_counts, int64_t *d_n_bases, int64_t *d_n_reads, int32_t *d_status,
                  void *d_workspace, size_t workspace_bytes, void *stream);

/* The STAGED pipeline of the north star as a measured baseline (csrc/staged.cuh): generate -> block-scan cut-off ->
 * LSD radix sort of read starts -> per-target binary-search sweep with warp-shuffle reduction.  Same draws, same
 * outputs as esl_run_batch (tests compare them bit for bit); 32-bit coordinates, scaling = 1 only.  Iterations are
 * processed one after the other.  Workspace: esl_staged_workspace_bytes. */
int64_t esl_staged_workspace_bytes(const esl_ctx *ctx, double coverage);
int esl_run_batch_staged(esl_ctx *ctx, uint64_t seed, uint32_t combo, uint32_t iter_begin, int32_t n_iter,
                         double coverage, int32_t consuming, int64_t min_overlap, int64_t *d_tallies,
                         int64_t *d_label_counts, int64_t *d_n_bases, int64_t *d_n_reads, int32_t *d_status,
                         void *d_workspace, size_t workspace_bytes, void *stream);

/* Replay level 1: recorded reads (already cut off), iteration i owns reads [seg_offsets[i], seg_offsets[i+1]).
 * count_matches + count_barcode_occurrences over them (counting.py:7-71).  seg_offsets is a DEVICE array of
 * n_iter+1 int64; total_reads = seg_offsets[n_iter] and max_reads_per_iter = the largest segment (passed so the
 * host need not read them back; max_reads_per_iter <= 0 means "unknown", which only costs empty tiles).
 * Workspace: esl_workspace_bytes(ctx, n_iter, 0, max_reads_per_iter) (total_reads when that is unknown). */
int esl_replay_reads(esl_ctx *ctx, const int64_t *d_start, const int64_t *d_length, const int32_t *d_label,
                     const int64_t *d_seg_offsets, int64_t total_reads, int64_t max_reads_per_iter, int32_t n_iter,
                     int64_t min_overlap,
                     int64_t *d_tallies, int64_t *d_label_counts, int64_t *d_n_bases, int64_t *d_n_reads,
                     int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream);

/* Replay level 2: recorded RAW draws of an oversampled batch (read_operations.py:101-103,65): per draw the
 * barcode uniform, the picked length and the start; per visited blocked hit one uniform
 * (d_ublock[d_ublock_off[n] + j] for the j-th visited hit of draw n, hits visited in (start, end) order, the
 * read's barcode tree first, then "ALL").  The device does the CDF lookup, the blocked search, the
 * prefix-sum cut-off and the tally.  d_ublock_off may be NULL when no mask is set.  `bulk_draws` is how many
 * leading draws of each iteration go through the order-free bulk kernel before the exact cut-off kernel
 * takes over (any value gives the same result; 0 = everything through the cut-off kernel). */
int esl_replay_draws(esl_ctx *ctx, const double *d_u_barcode, const int64_t *d_length, const int64_t *d_start,
                     const int64_t *d_ublock_off, const double *d_ublock, const int64_t *d_seg_offsets,
                     int64_t max_draws_per_iter, int64_t bulk_draws, int32_t n_iter, double coverage,
                     int32_t consuming, int64_t min_overlap, int64_t *d_tallies, int64_t *d_label_counts,
                     int64_t *d_n_bases, int64_t *d_n_reads, int32_t *d_status, void *d_workspace,
                     size_t workspace_bytes, void *stream);

/* Materialise the first n_reads draws of ONE native-mode iteration in draw order (start, length, label), for
 * consumers that need the reads themselves (coverage track coverage.py:52-81, overlap lists counting.py:56).
 * n_reads is the iteration's d_n_reads value from esl_run_batch with the same seed / combo / consuming. */
int esl_materialize_reads(esl_ctx *ctx, uint64_t seed, uint32_t combo, uint32_t iteration, int32_t consuming,
                          int64_t n_reads, int64_t *d_start, int64_t *d_length, int32_t *d_label, void *stream);

/* Binned coverage track of materialised reads, replacing the genome-length float64 arrays of the first-iteration
 * coverage plot (coverage.py:52-64,76-81), which cannot be allocated at hg38 size: d_bins [n_barcodes+2][n_bins]
 * receives the number of read bases per bin and label slot (barcodes, Blocked/Consuming, Blocked/NotConsuming);
 * bin b covers [b*bin_size, (b+1)*bin_size).  Zeroes d_bins first. */
int esl_coverage_bins(esl_ctx *ctx, const int64_t *d_start, const int64_t *d_length, const int32_t *d_label,
                      int64_t n_reads, int64_t bin_size, int64_t n_bins, int64_t *d_bins, void *stream);

/* Per-target moment accumulators over iterations, the operand of the multi-GPU sum-reduce that replaces the
 * joblib result gather + pandas groupby (esloco.py:149-158,183-195):
 *   d_moments [n_targets][ESL_MOMENT_COLS] += {n, sum f, sum f^2, sum p, sum p^2, sum b, lo32(sum b^2), hi(sum b^2)}
 * with f/p/b = full / partial / on_target_bases and sum b^2 split as hi * 2^32 + lo so every column stays an
 * exact int64 under summation.  Accumulates (does not zero). */
#define ESL_MOMENT_COLS 8
int esl_moments(esl_ctx *ctx, const int64_t *d_tallies, int32_t n_iter, int64_t *d_moments, void *stream);

/* Optional hit records behind the reference's per-target `overlap` list column (counting.py:49,53,56): while a
 * buffer is set, every counted (target, read) pair of esl_run_batch / esl_replay_* appends
 * {iteration in batch, target row, draw index, overlap} (4 x int64) at an atomically claimed slot; *d_count is
 * the number of pairs seen (the caller zeroes it; if it exceeds `capacity` the surplus was dropped).  Records
 * whose draw index is >= that iteration's d_n_reads lie beyond the coverage cut-off and must be ignored.
 * The overlap column carries the match kind in its sign bit (set = full match); mask it off for the length.
 * Sorting by (iteration, row, draw index) gives the reference's list order.  NULL / 0 / NULL switches it off. */
int esl_set_hit_buffer(esl_ctx *ctx, int64_t *d_hits, int64_t capacity, int64_t *d_count);

/* Host helper (no GPU): read lengths of a FASTA / FASTQ file, plain or gzip, in one streaming pass
 * (config_handler.py:16-51).  Lengths > min_read_length are written to out[0..capacity) (out may be NULL to count
 * only); returns how many were kept (may exceed capacity), or a negative esl_status (ESL_ERR_INVALID: cannot
 * open; ESL_ERR_UNSUPPORTED: malformed FASTQ).  *mean_out receives their mean. */
int64_t esl_scan_read_lengths(const char *path, int32_t is_fastq, int64_t min_read_length, uint32_t *out,
                              int64_t capacity, double *mean_out);

/* Host helper (no GPU): contig ids and lengths of a FASTA, plain or gzip, in one streaming pass -- all that
 * pseudo_fasta_coordinates (fasta_operations.py:3-22) keeps of the sequences it loads.  A record starts at a line
 * beginning with '>'; its id (first whitespace-delimited token) is written NUL-terminated behind the previous one into
 * names[0..names_cap), its length (characters of the following lines without line breaks, blanks, tabs) to
 * lengths[i], i < max_records.  Returns the number of records and *names_bytes = the bytes all ids need (either may
 * exceed the capacities: call again with room), or a negative esl_status (ESL_ERR_INVALID: cannot open). */
int64_t esl_scan_fasta_contigs(const char *path, char *names, int64_t names_cap, int64_t *lengths, int64_t max_records,
                               int64_t *names_bytes);

/* Host helper (no GPU): final start coordinates of n insertions placed one after the other, each pushing the earlier
 * insertions at or after its position right by insertion_length (create_insertion_genome.py:60-64,76-80). */
int esl_shift_insertions(const int64_t *positions, int64_t n, int64_t insertion_length, int64_t *starts);

/* Host helper (no GPU): append the rows of one (iteration, combination) block of `{exp}_matches_table.csv`
 * (esloco.py:155-160,198-205) to `path` (append = 0 truncates first).  Row k is
 *   pre[k] iteration \t full \t partial \t overlap_field \t on_target_bases mid[k] iteration \n
 * where pre[k] = pre_blob[pre_off[k] .. pre_off[k+1]) is "<target key>_" and mid[k] the tab-framed columns that do
 * not change between iterations (mean_read_length, coverage, target, barcode[, insertion]); tallies = host
 * [n_targets][3] int64.  Returns the bytes written or a negative esl_status.  The caller guarantees that no field
 * needs CSV quoting. */
int64_t esl_write_match_rows(const char *path, int32_t append, const char *pre_blob, const int64_t *pre_off,
                             const char *mid_blob, const int64_t *mid_off, const char *overlap_field,
                             const int64_t *tallies, int64_t n_targets, const char *iteration);

/* Host helper (no GPU): the same rows for n_iter iterations x n_combos combinations in ONE call -- one open file, one
 * buffer sized to the text it will hold -- in the table's row order iteration -> combination -> target.
 * mid_blobs / mid_offs / tallies are arrays of n_combos pointers; tallies[c] = host [n_iter][n_targets][3] int64;
 * iteration_ids[i] is printed as the iteration column and key suffix of block i. */
int64_t esl_write_match_rows_batch(const char *path, int32_t append, const char *pre_blob, const int64_t *pre_off,
                                   int32_t n_combos, const char *const *mid_blobs, const int64_t *const *mid_offs,
                                   const char *overlap_field, const int64_t *const *tallies, int64_t n_targets,
                                   int64_t n_iter, const int64_t *iteration_ids);

/* Host helper (no GPU): how the library would split one barcode's n targets into index layers over a genome of
 * size genome_size (a layer is scanned front to back, nesting is tolerated while it costs
 * reads at most a few wasted record loads).  layer_out[i] receives the layer of target i (-1 for an empty target,
 * end <= start, which is never indexed); returns the number of layers, or a negative esl_status.  For tests and for
 * sizing: the device index esl_prepare builds uses exactly this split per barcode. */
int32_t esl_plan_target_layers(uint64_t genome_size, const int64_t *start, const int64_t *end, int64_t n, int32_t *layer_out);

/* Replay of the reference's pair thinning for scaling < 1 (counting.py:46,51: one rand() per qualifying pair, kept
 * iff u <= scaling, pairs visited target-major in read order).  d_hits = the qualifying pairs of a replay run with
 * scaling = 1 (esl_set_hit_buffer records, sorted by iteration, row, draw index, cut-off applied), d_u = the
 * recorded uniforms in that order.  Writes tallies [n_iter][T][3] of the kept pairs (zeroed first). */
int esl_replay_thin(esl_ctx *ctx, const int64_t *d_hits, const double *d_u, int64_t n_hits, double scaling,
                    int32_t n_iter, int64_t *d_tallies, void *stream);

/* Per-kernel device timing for bench.py's roofline: when on, the next esl_run_batch / esl_replay_* records CUDA
 * events on its stream around the bulk kernel and around the cut-off kernel; esl_get_timing waits for them and
 * returns the two durations of the last such call in milliseconds. */
int esl_set_profiling(esl_ctx *ctx, int32_t on);
int esl_get_timing(esl_ctx *ctx, float *bulk_ms, float *cutoff_ms);

/* Introspection (for bench.py / DESIGN.md): number of kernel launches issued by this context so far. */
int64_t esl_launch_count(const esl_ctx *ctx);
/* Target index shape: total layers over all barcodes (1 per barcode unless targets nest deeply).  The index is
 * built by esl_prepare, so query it after that. */
int32_t esl_target_layers(const esl_ctx *ctx);
/* Reads per iteration the bulk kernel processes before the exact cut-off kernel takes over. */
int64_t esl_bulk_reads(const esl_ctx *ctx, double coverage);

#ifdef __cplusplus
}
#endif
#endif /* ESLOCO_B200_H */
