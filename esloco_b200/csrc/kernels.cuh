// Device code of the esloco hot path for sm_100a.
//
// One placed read is resolved by the same inlined pipeline in every kernel:
//   K1 Philox draw (or a recorded draw)            -> barcode uniform, length pick, start
//   K2 length table gather                          read_operations.py:102 / combined_calculations.py:57
//   K3 barcode CDF search + uniform placement       read_operations.py:96-103
//   K4 blocked-interval search                      read_operations.py:59-67
//   K6 target overlap search (read-centric, over start-sorted target layers) counting.py:36-53
//   K7 integer tallies via red.global               counting.py:54-57, :63-71
// and K5, the coverage cut-off (read_operations.py:100), is split in two kernels:
//   k_place_tally_bulk  : the first n_bulk draws of every iteration, which lie before the cut-off with
//                         overwhelming probability; no ordering between reads is needed, so it is a flat
//                         persistent grid.  Per-tile counted-length sums are kept.
//   k_cutoff_tail       : one CTA per iteration continues from n_bulk in draw order with a block-scan
//                         prefix sum until covered >= coverage*G, includes exactly the reads the
//                         reference's while-loop would have drawn, and -- if the bulk range did overshoot
//                         (checked exactly from the tile sums) -- regenerates the overshoot and subtracts
//                         it again, so the result never depends on the statistical bound.
// All counters are integers accumulated with atomics, so results are independent of scheduling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "philox.cuh"

namespace esl {

typedef unsigned long long u64;
typedef int64_t i64;

#ifndef ESL_BULK_BLOCK
#define ESL_BULK_BLOCK 256
#endif
constexpr int kBulkBlock = ESL_BULK_BLOCK; // threads per CTA of the bulk kernel
// Reads per thread and tile: FOUR CONSECUTIVE draws (n0 .. n0+3, n0 = base + 4 * tid).  With 32-bit coordinates one
// Philox block serves two draws and one more block the barcodes of four (PhiloxSrc), so a thread spends 2 (+1) blocks
// on its four reads instead of 4.
constexpr int kBulkIpt = 4;
constexpr int kBulkTile = kBulkBlock * kBulkIpt;
#ifndef ESL_BULK_CLAIM
#define ESL_BULK_CLAIM 16
#endif
constexpr int kBulkClaim = ESL_BULK_CLAIM; // tiles a CTA claims from an iteration's tile counter at a time
#ifndef ESL_BULK_MIN_BLOCKS
#define ESL_BULK_MIN_BLOCKS 4
#endif
constexpr int kBulkMinBlocks = ESL_BULK_MIN_BLOCKS; // register cap: 65536 / (256 * min blocks)
#ifndef ESL_BULK_MIN_BLOCKS_GENERAL
#define ESL_BULK_MIN_BLOCKS_GENERAL 4
#endif
// The general variant of the kernel (candidate queues, any number of targets and layers) needs 62 registers to stay
// free of spills: 4 CTAs/SM.  Measured on config 3: queue @4 96 G reads/s, queue @5 with 120 B of spills 83 G,
// direct walks @5 85 G, direct walks @4 81 G.  64-bit coordinates (genomes above 4 Gb) cost one CTA/SM more.
constexpr int kBulkMinBlocksGeneral = ESL_BULK_MIN_BLOCKS_GENERAL;
template <typename C, bool kShared>
constexpr int bulk_min_blocks() { return (kShared ? kBulkMinBlocks : kBulkMinBlocksGeneral) - (sizeof(C) == 8 ? 1 : 0); }
constexpr int kTailBlock = 1024;
constexpr int kTailIpt = 4; // consecutive draws per thread: one scan orders a chunk of 4096 draws, and draw4 serves them
constexpr int kTailChunk = kTailBlock * kTailIpt;
constexpr int kMaxBarcodes = 4096;

constexpr int kLabelBlockedConsuming = -1;
constexpr int kLabelBlockedNotConsuming = -2;
constexpr int kLabelInvalid = -1000; // slot of a tile that holds no draw

// Bounds checks for a debug build (-DESL_DEBUG_BOUNDS): compute-sanitizer is not available on the GPU pool, so
// every index into the grids, records, pool, tallies and tile sums can be checked in-kernel instead; a violation
// prints its site and traps.  Compiled out by default.
#ifdef ESL_DEBUG_BOUNDS
#define ESL_CHK(cond, what)                                                                    \
    do {                                                                                       \
        if (!(cond)) { printf("esl bounds violation: %s (%s:%d)\n", what, __FILE__, __LINE__); __trap(); } \
    } while (0)
#else
#define ESL_CHK(cond, what) ((void)0)
#endif

// One sorted run of intervals (a target layer, or one blocked group) plus its bucket grid.
// Grid entry b (b = position >> shift) holds j = first index of the run whose (running-max) end exceeds
// b << shift, together with the START of that interval.  For a read [rs, re) with bucket b = rs >> shift:
//   * every interval before j ends at or before the bucket start <= rs, so it cannot overlap;
//   * if start[j] >= re nothing overlaps at all (starts ascend) -- the common case, decided by ONE 8-byte load;
//   * otherwise the candidates are j, j+1, ... while start < re, each checked exactly (end > rs).
// A bucket normally holds at most a few interval ends; for crowded runs (flag) the candidate scan first
// bisects inside [j, j of the next bucket].
struct __align__(16) SegDesc {
    int lo, hi;   // index range of the run
    int grid_off; // offset of the run's grid
    int flags;    // bits 0-7 log2(bucket width), bit 8 crowded, bits 16-31 number of EXTRA layers (targets only)
};
constexpr int kCrowdedBucket = 8;

template <typename C> struct GridEntry;
template <> struct __align__(8) GridEntry<uint32_t> { int j; uint32_t s; };
template <> struct __align__(16) GridEntry<uint64_t> { int j; int pad; uint64_t s; };

// One indexed target: start, end, original row and the start of the next target of the layer (all ones at the
// end), so a candidate walk needs one 16-byte load per candidate and none to find out that it is over.
template <typename C> struct TargetRec;
// (row: original row of the target; bit 31 set = "hot" target, see tally_layer)
template <> struct __align__(16) TargetRec<uint32_t> { uint32_t s, e; int row; uint32_t s_next; };
template <> struct __align__(16) TargetRec<uint64_t> { uint64_t s, e, s_next; int row; int pad; };
__device__ __forceinline__ TargetRec<uint32_t> load_rec(const TargetRec<uint32_t> *p)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    return TargetRec<uint32_t>{v.x, v.y, (int)v.z, v.w};
}
__device__ __forceinline__ TargetRec<uint64_t> load_rec(const TargetRec<uint64_t> *p)
{
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(p) + 1);
    return TargetRec<uint64_t>{a.x, a.y, b.x, (int)(uint32_t)b.y, 0};
}

// One blocked interval: start, end and blocking probability in one 16-byte (u32) record.
template <typename C> struct MaskRec;
template <> struct __align__(16) MaskRec<uint32_t> { uint32_t s, e; double w; };
template <> struct __align__(16) MaskRec<uint64_t> { uint64_t s, e; double w; double pad; };
__device__ __forceinline__ MaskRec<uint32_t> load_rec(const MaskRec<uint32_t> *p)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    return MaskRec<uint32_t>{v.x, v.y, __hiloint2double((int)v.w, (int)v.z)};
}
__device__ __forceinline__ MaskRec<uint64_t> load_rec(const MaskRec<uint64_t> *p)
{
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    const double w = __ldg(reinterpret_cast<const double *>(p) + 2);
    return MaskRec<uint64_t>{a.x, a.y, w, 0.0};
}

__device__ __forceinline__ SegDesc load_desc(const SegDesc *p)
{
    const int4 v = __ldg(reinterpret_cast<const int4 *>(p)); // one LDG.128 through the read-only path
    return SegDesc{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ GridEntry<uint32_t> load_entry(const GridEntry<uint32_t> *p)
{
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    return GridEntry<uint32_t>{(int)v.x, v.y};
}
__device__ __forceinline__ GridEntry<uint64_t> load_entry(const GridEntry<uint64_t> *p)
{
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    return GridEntry<uint64_t>{(int)(uint32_t)v.x, 0, v.y};
}

// Everything a kernel needs to resolve one read; passed by value.
template <typename C>
struct DevCtx {
    u64 G;           // genome size
    i64 thr;         // ceil(coverage * G): covered < coverage*G  <=>  covered < thr for integer covered
    i64 min_overlap; // min_overlap_for_detection
    double scaling;  // >= 1: no thinning draw
    int n_barcodes;
    int consuming;
    int has_mask;
    int mask_mode; // bit 0: some barcode has a tree of its own, bit 1: the "ALL" tree holds intervals
    // coarse map over the genome, staged in shared memory by the bulk kernel, two bits per bucket b = position >>
    // mbits_shift.  Bit 0: ANY blocked interval (of any tree) touches the bucket -- a read spanning at most two buckets
    // whose bits are clear is certainly not blocked.  Bit 1: the bucket lies entirely inside "ALL" intervals of weight >= 1
    // -- a read spanning at most two such buckets is certainly blocked (u < 1 <= weight for every uniform, whichever tree
    // is asked first).  Neither needs a global load; everything else takes the exact test.
    const uint32_t *__restrict__ mbits;
    int mbits_shift, mbits_words;
    uint32_t n_pool;
    const uint32_t *__restrict__ pool;    // length table (K2)
    // quantile sampler (esl_set_length_sampler): qtab[k] = sorted pool entry at rank k * n_pool / kQuantBins, k <= kQuantBins
    // (null: lengths are uniform picks from `pool`); the top bin is picked exactly from pool_sorted[q_top_lo ..]
    const uint32_t *__restrict__ qtab;
    const uint32_t *__restrict__ pool_sorted;
    uint32_t q_top_lo, q_top_n;
    const double *__restrict__ cdf;       // numpy-style cdf, replay path
    const uint32_t *__restrict__ cdf_u32; // ceil(cdf * 2^32) saturated, native path
    const uint2 *__restrict__ cdf_lut;    // [kCdfLut] per slice of 2^22 draws: {barcode k of the slice's lowest draw | (k < B-1) << 16, threshold cdf_u32[k]}
    int cdf_walk;                         // most thresholds inside one slice (steps the walk after the table needs)
    uint32_t min_ov32;                    // min_overlap clamped to [0, 2^32-1] (an overlap never exceeds a read length)
    int thin;                             // scaling < 1: draw a thinning uniform per qualifying pair
    // target index: per barcode a few layers sorted by start whose ends are (nearly) ascending: the targets a read
    // overlaps lie in one short contiguous range per layer, with at most a few nested "dead" records in between
    // (host: build_layers).  seg[b], b < B, is barcode b's first layer; its extra layers (deep nesting only) are
    // seg[B + extra_off[b] ...].
    const TargetRec<C> *__restrict__ trec; // targets in layer order
    const C *__restrict__ te;              // running max of their ends inside each layer (bisect key of crowded buckets)
    const SegDesc *__restrict__ seg;
    const GridEntry<C> *__restrict__ tgrid;
    const int32_t *__restrict__ extra_off; // [B]
    int any_extra;                         // some barcode has more than one layer
    const int32_t *__restrict__ hot_row;   // [n_hot] row of each hot target's slot
    int n_hot;
    int n_seg, n_tgrid, n_trec, n_mgrid, n_mrec; // array sizes (debug bounds checks)
    // small target sets: for every first layer a coarser grid holding per bucket the pair {START of the bucket's
    // first candidate, largest END a read starting in the bucket can meet} (sgrid_n pairs per barcode, bucket width
    // 2^sgrid_shift), staged in shared memory by the bulk kernel: "start >= read end" or "end <= read start" settles
    // the common no-overlap case without any global load, everything else goes through the exact global grid.
    // NULL when the first layers do not fit the shared-memory budget.
    const C *__restrict__ sgrid;
    int sgrid_n, sgrid_shift;
    i64 n_targets;
    // blocked index: group g < B is barcode g's tree, group B is "ALL"; inside a group sorted by
    // (start, end) with a running max of ends (pmax) as the grid key.
    const MaskRec<C> *__restrict__ mrec;
    const C *__restrict__ mpmax;
    const SegDesc *__restrict__ grp;      // [B + 1]
    const GridEntry<C> *__restrict__ mgrid;
};
constexpr int kCdfLut = 1024;
constexpr int kQuantBins = 2048; // bins of the quantile length sampler: 11 bits pick the bin, 13 bits the place inside it

// Look-up tables a draw reads: barcode thresholds + start table (shared memory) and, with the quantile sampler, the
// quantile table (shared memory in the bulk kernel, global elsewhere).
struct DrawTabs {
    const uint32_t *cdf;
    const uint2 *lut;
    const uint32_t *qtab;
};
constexpr int kSmemDescs = 1024; // first-layer descriptors are staged in shared memory up to this many barcodes

struct DevOut {
    u64 *tallies;      // [n_iter][T][3]
    u64 *label_counts; // [n_iter][B+2]
    u64 *n_bases;      // [n_iter]
    u64 *n_reads;      // [n_iter]
    int *status;       // [n_iter]
    u64 *tile_sum;     // [n_iter][tiles_per_iter] counted-length sum of each bulk tile (zeroed by the host call)
    int tiles_per_iter;
    // optional hit records for the reference's per-target `overlap` lists (counting.py:49,53,56): one
    // {iteration in batch, target row, draw index, overlap} per counted pair, in no particular order
    i64 *hits;         // [hit_cap][4] or NULL
    u64 *hit_count;    // number of records appended (may exceed hit_cap: the host reports the overflow)
    i64 hit_cap;
};

// ---------------------------------------------------------------- draw sources

// Native stream.  Counter layout (philox.cuh): c2 = iteration, c3 = combination << 4 | slot, key = seed.
//   32-bit coordinates (G < 2^32): draw n takes 64 bits (a, b) of block (c0 = n >> 1, c1 = 0) -- words (x, y) for even
//     n, (z, w) for odd n:
//        length = pool[ floor((b >> 8) * n_pool / 2^24) ]                     24 bits: uniform pick from the pool
//        start  = floor( (a * 2^32 + (b & 0xff) * 2^24) * (G - L) / 2^64 )    40 bits: U{0 .. G-L-1}
//     (a position's probability deviates from uniform by < (G - L) / 2^40 = 0.3 % of itself at hg38 size, a window of
//     w positions by 1/w of that; a pool entry's by < n_pool / 2^24) and, with more than one barcode, word n & 3 of
//     block (c0 = n >> 2, c1 = 1) for searchsorted(cdf, u, 'right').
//   64-bit coordinates: draw n owns block (c0 = n, c1 = 0): (x, y) = 64-bit start, z = length pick, w = barcode.
struct PhiloxSrc {
    uint32_t k0, k1, combo, iter_begin;
    PhiloxKeys keys; // round keys of (k0, k1), filled by the host
    static constexpr bool kHasLabel = false;
    static constexpr bool kBounded = false;
    static constexpr bool kPoolLengths = true; // no read is longer than the pool's longest entry
    __device__ __forceinline__ i64 count(int) const { return 0x7fffffffffffffffLL; }

    __device__ __forceinline__ Philox4 block(uint32_t c0, uint32_t c1, int it) const
    {
        return philox4x32_10(c0, c1, iter_begin + (uint32_t)it, (combo << 4) | ESL_SLOT_PLACE, keys);
    }
    template <typename C>
    static __device__ __forceinline__ int barcode_of(const DevCtx<C> &cx, const DrawTabs &tb, uint32_t w)
    {
        // searchsorted(cdf, u, side='right') on integer thresholds: table start + short walk.  A slice of the start
        // table holds at most cx.cdf_walk thresholds: one branch-free step covers almost every CDF; the loop only
        // exists for more than ~1000 barcodes or pathological weights
        const uint2 e = tb.lut[w >> 22]; // one 8-byte load: start barcode, whether a step is possible, and its threshold
        int k = (int)(e.x & 0xffffu) + (int)((e.x >> 16) & (e.y <= w ? 1u : 0u));
        if (cx.cdf_walk > 1)
            while (k < cx.n_barcodes - 1 && tb.cdf[k] <= w) ++k;
        return k;
    }
    // Read length from 24 uniform bits u.  Default: uniform pick from the pool, pool[floor(u * n_pool / 2^24)] (one random
    // 4-byte gather from the L2-resident pool per read).  Quantile sampler: inverse-CDF look-up in the kQuantBins-bin
    // quantile table -- 11 bits pick the bin, 13 bits interpolate between its two quantiles; the top bin, where a
    // long-tailed distribution is anything but linear, is picked exactly from the sorted pool.
    template <typename C>
    static __device__ __forceinline__ uint32_t length_of(const DevCtx<C> &cx, const DrawTabs &tb, uint32_t u24)
    {
        if (tb.qtab) {
            const uint32_t k = u24 >> 13, f = u24 & 0x1FFFu;
            if (k == kQuantBins - 1) return __ldg(cx.pool_sorted + cx.q_top_lo + (uint32_t)(((u64)f * cx.q_top_n) >> 13));
            const uint32_t q0 = tb.qtab[k], q1 = tb.qtab[k + 1];
            return q0 + (uint32_t)(((u64)(q1 - q0) * f) >> 13);
        }
        return __ldg(cx.pool + __umulhi(u24 << 8, cx.n_pool)); // index < n_pool by construction
    }
    static __device__ __forceinline__ void place32(const DevCtx<uint32_t> &cx, const DrawTabs &tb, uint32_t a, uint32_t b, uint32_t &L, uint32_t &start)
    {
        L = length_of(cx, tb, b >> 8);
        ESL_CHK((u64)L < cx.G, "read not shorter than the genome");
        const uint32_t rg = (uint32_t)(cx.G - (u64)L); // two 32x32->64 products instead of a 64x64 high multiply
        start = (uint32_t)(((u64)a * rg + (((u64)(b << 24) * rg) >> 32)) >> 32);
    }
    static __device__ __forceinline__ void place64(const DevCtx<uint64_t> &cx, const DrawTabs &tb, const Philox4 &r, uint32_t &L, uint64_t &start)
    {
        L = tb.qtab ? length_of(cx, tb, r.z >> 8) : __ldg(cx.pool + __umulhi(r.z, cx.n_pool));
        ESL_CHK((u64)L < cx.G, "read not shorter than the genome");
        start = __umul64hi(((u64)r.x << 32) | r.y, cx.G - (u64)L);
    }

    // One draw (cut-off kernel, partial tiles, materialisation).
    template <typename C>
    __device__ __forceinline__ void draw(const DevCtx<C> &cx, const DrawTabs &tb, int it, i64 n,
                                         uint32_t &L, C &start, int &bc) const
    {
        bc = 0;
        if constexpr (sizeof(C) == 4) {
            const Philox4 r = block((uint32_t)n >> 1, 0u, it);
            place32(cx, tb, (n & 1) ? r.z : r.x, (n & 1) ? r.w : r.y, L, start);
            if (cx.n_barcodes > 1) {
                const Philox4 q = block((uint32_t)n >> 2, 1u, it);
                const uint32_t w = (n & 2) ? ((n & 1) ? q.w : q.z) : ((n & 1) ? q.y : q.x);
                bc = barcode_of(cx, tb, w);
            }
        } else {
            const Philox4 r = block((uint32_t)n, 0u, it);
            place64(cx, tb, r, L, start);
            if (cx.n_barcodes > 1) bc = barcode_of(cx, tb, r.w);
        }
    }
    // Four consecutive draws n0 .. n0+3 (n0 a multiple of 4): the bulk kernel's full tiles.
    template <typename C>
    __device__ __forceinline__ void draw4(const DevCtx<C> &cx, const DrawTabs &tb, int it, uint32_t n0,
                                          uint32_t (&L)[4], C (&start)[4], int (&bc)[4]) const
    {
        if constexpr (sizeof(C) == 4) {
            const Philox4 r0 = block(n0 >> 1, 0u, it), r1 = block((n0 >> 1) + 1u, 0u, it);
            place32(cx, tb, r0.x, r0.y, L[0], start[0]);
            place32(cx, tb, r0.z, r0.w, L[1], start[1]);
            place32(cx, tb, r1.x, r1.y, L[2], start[2]);
            place32(cx, tb, r1.z, r1.w, L[3], start[3]);
            bc[0] = bc[1] = bc[2] = bc[3] = 0;
            if (cx.n_barcodes > 1) {
                const Philox4 q = block(n0 >> 2, 1u, it);
                bc[0] = barcode_of(cx, tb, q.x); bc[1] = barcode_of(cx, tb, q.y);
                bc[2] = barcode_of(cx, tb, q.z); bc[3] = barcode_of(cx, tb, q.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) draw(cx, tb, it, (i64)(n0 + i), L[i], start[i], bc[i]);
        }
    }
    __device__ __forceinline__ double block_u(int it, i64 n, int j) const
    {
        const Philox4 r = philox4x32_10((uint32_t)n, (uint32_t)(j >> 1), iter_begin + (uint32_t)it,
                                        (combo << 4) | ESL_SLOT_BLOCK, k0, k1);
        return (j & 1) ? u53(r.z, r.w) : u53(r.x, r.y);
    }
    __device__ __forceinline__ double scale_u(int it, i64 n, int row) const
    {
        const Philox4 r = philox4x32_10((uint32_t)n, (uint32_t)row, iter_begin + (uint32_t)it,
                                        (combo << 4) | ESL_SLOT_SCALE, k0, k1);
        return u53(r.x, r.y);
    }
};

// Recorded raw draws (replay level 2).
struct DrawSrc {
    const double *__restrict__ ubc;
    const i64 *__restrict__ length;
    const i64 *__restrict__ start;
    const i64 *__restrict__ uoff;
    const double *__restrict__ ublock;
    const i64 *__restrict__ seg;
    static constexpr bool kHasLabel = false;
    static constexpr bool kBounded = true;
    static constexpr bool kPoolLengths = false; // recorded lengths: the pool's maximum says nothing about them
    __device__ __forceinline__ i64 count(int it) const { return seg[it + 1] - seg[it]; }

    template <typename C>
    __device__ __forceinline__ void draw(const DevCtx<C> &cx, const DrawTabs &, int it, i64 n, uint32_t &L,
                                         C &st, int &bc) const
    {
        const i64 g = seg[it] + n;
        L = (uint32_t)length[g];
        st = (C)start[g];
        const double u = ubc[g];
        int k = 0;
        while (k < cx.n_barcodes - 1 && cx.cdf[k] <= u) ++k;
        bc = k;
    }
    template <typename C>
    __device__ __forceinline__ void draw4(const DevCtx<C> &cx, const DrawTabs &tb, int it, uint32_t n0,
                                          uint32_t (&L)[4], C (&st)[4], int (&bc)[4]) const
    {
#pragma unroll
        for (int i = 0; i < 4; ++i) draw(cx, tb, it, (i64)(n0 + i), L[i], st[i], bc[i]);
    }
    __device__ __forceinline__ double block_u(int it, i64 n, int j) const { return ublock[uoff[seg[it] + n] + j]; }
    __device__ __forceinline__ double scale_u(int, i64, int) const { return 0.0; }
};

// Recorded reads (replay level 1): the label is given, no CDF / blocked search.
struct ReadSrc {
    const i64 *__restrict__ start;
    const i64 *__restrict__ length;
    const int32_t *__restrict__ label;
    const i64 *__restrict__ seg;
    static constexpr bool kHasLabel = true;
    static constexpr bool kBounded = true;
    static constexpr bool kPoolLengths = false; // recorded lengths: the pool's maximum says nothing about them
    __device__ __forceinline__ i64 count(int it) const { return seg[it + 1] - seg[it]; }

    template <typename C>
    __device__ __forceinline__ void draw(const DevCtx<C> &, const DrawTabs &, int it, i64 n, uint32_t &L,
                                         C &st, int &bc) const
    {
        const i64 g = seg[it] + n;
        L = (uint32_t)length[g];
        st = (C)start[g];
        bc = label[g];
    }
    template <typename C>
    __device__ __forceinline__ void draw4(const DevCtx<C> &cx, const DrawTabs &tb, int it, uint32_t n0,
                                          uint32_t (&L)[4], C (&st)[4], int (&bc)[4]) const
    {
#pragma unroll
        for (int i = 0; i < 4; ++i) draw(cx, tb, it, (i64)(n0 + i), L[i], st[i], bc[i]);
    }
    __device__ __forceinline__ double block_u(int, i64, int) const { return 0.0; }
    __device__ __forceinline__ double scale_u(int, i64, int) const { return 0.0; }
};

// ---------------------------------------------------------------- per-read pipeline

// First candidate of a run for a read starting at rs: one descriptor load (L1-resident) + one grid-entry load.
template <typename C>
struct Probe {
    int a; // first candidate index
    C s;   // its start (all ones when there is none)
};
template <typename C>
__device__ __forceinline__ Probe<C> probe(const SegDesc *descs, int idx, const GridEntry<C> *grid, C rs)
{
    const int2 d = __ldg(reinterpret_cast<const int2 *>(descs + idx) + 1); // {grid_off, flags}: half a descriptor, half the registers
    const GridEntry<C> e = load_entry(grid + d.x + (int)((u64)rs >> (d.y & 0xff)));
    ESL_CHK(e.j >= descs[idx].lo && e.j <= descs[idx].hi, "grid entry outside its run");
    return Probe<C>{e.j, e.s};
}

// Index of the first interval in [a, next bucket's j] whose key exceeds rs, for crowded runs only.
template <typename C>
__device__ __forceinline__ int skip_crowded(const SegDesc &d, const GridEntry<C> *grid, const C *__restrict__ keys, int a, C rs)
{
    if (!(d.flags & 0x100)) return a;
    int z = load_entry(grid + d.grid_off + (int)((u64)rs >> (d.flags & 0xff)) + 1).j;
    if (z - a <= kCrowdedBucket) return a;
    while (a < z) {
        const int mid = (a + z) >> 1;
        if (__ldg(keys + mid) <= rs) a = mid + 1; else z = mid;
    }
    return a;
}

// K4: is [rs, re) blocked for barcode bc?  Trees: the barcode's own, then "ALL"; hits in (start, end)
// order; one uniform per visited hit; u < weight blocks (read_operations.py:59-67).
template <typename C, typename Src>
__device__ __forceinline__ bool is_blocked(const DevCtx<C> &cx, const Src &src, int it, i64 n, int bc, C rs, C re)
{
    int j = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int g = pass == 0 ? bc : cx.n_barcodes;
        const Probe<C> p = probe(cx.grp, g, cx.mgrid, rs);
        if (p.s >= re) continue;
        const SegDesc d = load_desc(cx.grp + g);
        for (int k = skip_crowded(d, cx.mgrid, cx.mpmax, p.a, rs); k < d.hi; ++k) {
            ESL_CHK(k >= 0 && k < cx.n_mrec, "blocked record index");
            const MaskRec<C> r = load_rec(cx.mrec + k);
            if (r.s >= re) break;
            if (r.e > rs) {
                if (!Src::kBounded && r.w >= 1.0) return true; // u < 1 <= w for every u: no need to draw it
                const double u = src.block_u(it, n, j++);
                if (u < r.w) return true;
            }
        }
    }
    return false;
}

constexpr int kHotSlots = 64;               // "hot" targets with a private shared-memory tally per CTA
constexpr int kHotWords = kHotSlots * 4 + kHotSlots; // {full, partial, overlap & 0xffff, overlap >> 16} per slot, then the slots' rows
static_assert(kHotWords % 4 == 0, "the hot-target block must keep 16-byte alignment behind it");
constexpr int kHotFlushTiles = 32;          // 32 tiles x 1024 reads x 0xffff < 2^32: the 32-bit shared sums cannot overflow
static_assert((unsigned long long)kHotFlushTiles * kBulkTile * 0xffffull < (1ull << 32), "hot-target sums must fit 32 bits between flushes");

// Rare paths of tally_hit, kept out of line: they would otherwise be compiled into every one of its call sites (the
// bulk kernel of a masked run with many targets was 100 KB of code, and 14 % of its stall samples waited for
// instructions).
template <typename Src>
__device__ __noinline__ bool thin_keeps(const Src &src, int it, i64 n, int row, double scaling)
{
    return src.scale_u(it, n, row) <= scaling; // counting.py:46,51
}
__device__ __noinline__ void record_hit(i64 *hits, u64 *hit_count, i64 hit_cap, int it, int row, i64 n, i64 ov, int kind)
{
    const u64 slot_h = atomicAdd(hit_count, 1ull);
    if ((i64)slot_h < hit_cap) {
        i64 *h = hits + slot_h * 4;
        h[0] = it; h[1] = row; h[2] = n;
        h[3] = ov | (kind == 0 ? (i64)0x8000000000000000ULL : 0); // sign bit = full match
    }
}

// K6 + K7 for ONE target record a read [rs, re) was found to overlap (r.s < re and r.e > rs): add (sign = +1) or take
// back (sign = -1) the read's contribution (counting.py:40-57).  s_hot = the CTA's private tallies of the hot targets
// (bulk kernel) or null (cut-off kernel: plain atomics, it sees 0.4 % of the reads).
// kCompact: the rare paths (pair thinning, hit records) are calls instead of inline code -- for the queue rounds, whose
// code is compiled into several places of the largest kernel; the walk-in-place callers keep them inline (as calls they
// cost the small-target-set kernel 8 registers and 5 % of its throughput).
template <typename C, typename Src, bool kCompact = false>
__device__ __forceinline__ void tally_hit(const DevCtx<C> &cx, const Src &src, const DevOut &out, u64 *__restrict__ tallies_it, int it, i64 n,
                                          const TargetRec<C> &r, C rs, C re, i64 sign, uint32_t *s_hot)
{
    const C ov = (r.e < re ? r.e : re) - (r.s > rs ? r.s : rs);
    // "hot" target (flagged by the host: a read hits it with probability >= 1/512, e.g. a whole-chromosome
    // ROI): the record carries a slot number instead of the row
    int row = r.row & 0x7fffffff;
    const int slot = row;
    if (r.row < 0) row = s_hot ? (int)s_hot[kHotSlots * 4 + slot] : __ldg(cx.hot_row + slot);
    ESL_CHK(row < cx.n_targets, "target row");
    if ((u64)ov >= (u64)cx.min_ov32 &&
        !(cx.thin && !(kCompact ? thin_keeps(src, it, n, row, cx.scaling) : src.scale_u(it, n, row) <= cx.scaling))) {
        const int kind = (rs <= r.s && r.e <= re) ? 0 : 1;
        if (r.row < 0 && s_hot) {
            // bulk kernel: hits go to the CTA's private 32-bit sums in shared memory (flushed to the global
            // tallies every kHotFlushTiles tiles) -- thousands of atomics per iteration on one global address
            // serialise in L2, and combining lanes first (match.any + REDUX per group) cost more than it saved
            uint32_t *h = s_hot + slot * 4;
            atomicAdd(h + kind, 1u);
            atomicAdd(h + 2, (uint32_t)ov & 0xffffu); // ov <= read length < 2^32
            atomicAdd(h + 3, (uint32_t)((u64)ov >> 16));
        } else {
            u64 *t = tallies_it + (i64)row * 3;
            atomicAdd(t + kind, (u64)sign);
            atomicAdd(t + 2, (u64)(sign * (i64)ov));
        }
        if (out.hits && sign > 0) { // draws beyond the cut-off are dropped on the host by their index
            if (kCompact) {
                record_hit(out.hits, out.hit_count, out.hit_cap, it, row, n, (i64)ov, kind);
            } else {
                const u64 slot_h = atomicAdd(out.hit_count, 1ull);
                if ((i64)slot_h < out.hit_cap) {
                    i64 *h = out.hits + slot_h * 4;
                    h[0] = it; h[1] = row; h[2] = n;
                    h[3] = (i64)ov | (kind == 0 ? (i64)0x8000000000000000ULL : 0); // sign bit = full match
                }
            }
        }
    }
}

// One layer whose first candidate starts before the read's end, walked where it was found (cut-off kernel, and the
// rare candidate of the small-target-set path).
template <typename C, typename Src>
__device__ __forceinline__ void tally_layer(const DevCtx<C> &cx, const Src &src, const DevOut &out, u64 *__restrict__ tallies_it, int it, i64 n,
                                         int sg, int a, C rs, C re, i64 sign, uint32_t *s_hot = nullptr)
{
    const SegDesc d = load_desc(cx.seg + sg);
    for (int j = skip_crowded(d, cx.tgrid, cx.te, a, rs); j < d.hi; ++j) {
        ESL_CHK(j >= 0 && j < cx.n_trec, "target record index");
        const TargetRec<C> r = load_rec(cx.trec + j);
        if (r.s >= re) break; // starts ascend: nothing further overlaps
        if (r.e > rs) tally_hit(cx, src, out, tallies_it, it, n, r, rs, re, sign, s_hot); // (an end before the read does not overlap)
        if (r.s_next >= re) break; // known without touching the next record
    }
}

// Add the CTA's private sums of the hot targets to the global tallies of iteration `it` and clear them.  Called by
// all threads of the CTA, between two block barriers.
__device__ __forceinline__ void flush_hot(uint32_t *s_hot, int n_hot, u64 *__restrict__ tallies_it)
{
    for (int i = threadIdx.x; i < n_hot; i += blockDim.x) {
        uint32_t *h = s_hot + i * 4;
        const uint32_t f = h[0], p = h[1], lo = h[2], hi = h[3];
        if (f | p | lo | hi) {
            h[0] = h[1] = h[2] = h[3] = 0;
            u64 *t = tallies_it + (i64)s_hot[kHotSlots * 4 + i] * 3;
            if (f) atomicAdd(t, (u64)f);
            if (p) atomicAdd(t + 1, (u64)p);
            atomicAdd(t + 2, (u64)lo + ((u64)hi << 16));
        }
    }
}

// Extra layers of barcode bc (only when its targets nest).
template <typename C, typename Src>
__device__ __forceinline__ void tally_extra_layers(const DevCtx<C> &cx, const Src &src, const DevOut &out, u64 *__restrict__ tallies_it, int it,
                                                i64 n, int bc, int n_extra, C rs, C re, i64 sign, uint32_t *s_hot = nullptr)
{
    const int s0 = cx.n_barcodes + __ldg(cx.extra_off + bc);
    for (int sg = s0; sg < s0 + n_extra; ++sg) {
        const Probe<C> p = probe(cx.seg, sg, cx.tgrid, rs);
        if (p.s < re) tally_layer(cx, src, out, tallies_it, it, n, sg, p.a, rs, re, sign, s_hot);
    }
}

// First layer of the read's barcode through the exact global grid.
template <typename C, typename Src>
__device__ __forceinline__ void tally_first_layer(const DevCtx<C> &cx, const Src &src, const DevOut &out, u64 *__restrict__ tallies_it,
                                                  int it, i64 n, C rs, C re, int bc, i64 sign, uint32_t *s_hot = nullptr)
{
    const Probe<C> p = probe(cx.seg, bc, cx.tgrid, rs);
    if (p.s < re) tally_layer(cx, src, out, tallies_it, it, n, bc, p.a, rs, re, sign, s_hot);
}

// All layers of the read's barcode: probe each layer, walk its candidates when the first one starts before
// the read's end (cut-off kernel).
template <typename C, typename Src>
__device__ __forceinline__ void tally_read(const DevCtx<C> &cx, const Src &src, const DevOut &out, u64 *__restrict__ tallies_it, int it,
                                        i64 n, C rs, C re, int bc, i64 sign)
{
    const SegDesc d = load_desc(cx.seg + bc);
    const GridEntry<C> e = load_entry(cx.tgrid + d.grid_off + (int)((u64)rs >> (d.flags & 0xff)));
    if (e.s < re) tally_layer(cx, src, out, tallies_it, it, n, bc, e.j, rs, re, sign);
    const int n_extra = (int)((uint32_t)d.flags >> 16);
    if (n_extra) tally_extra_layers(cx, src, out, tallies_it, it, n, bc, n_extra, rs, re, sign);
}

// Resolve label + counted length of one drawn read (read_operations.py:104-121).
template <typename C, typename Src>
__device__ __forceinline__ void resolve_read(const DevCtx<C> &cx, const Src &src, int it, i64 n, uint32_t L, C st,
                                             int bc, int &label, uint32_t &counted)
{
    label = bc;
    counted = L;
    if (Src::kHasLabel) {
        if (bc == kLabelBlockedNotConsuming) counted = 0;
        return;
    }
    if (cx.has_mask && is_blocked(cx, src, it, n, bc, st, (C)(st + L))) {
        label = cx.consuming ? kLabelBlockedConsuming : kLabelBlockedNotConsuming;
        counted = cx.consuming ? L : 0u;
    }
}

__device__ __forceinline__ int label_slot(int label, int B) { return label >= 0 ? label : (label == -1 ? B : B + 1); }

// Shared-memory label histogram update, called by ALL lanes of a warp (slot < 0 = nothing to add); n_slots = how
// many labels exist (block-uniform).  Few labels: lanes with the same slot are combined with match.any so one
// 32-bit ATOMS per distinct label is issued -- a masked single-barcode run sends every lane to the same counter and
// plain per-lane atomics serialise 32-fold there (config 4 lost a third of its throughput).  Many labels (>= 18):
// lanes rarely collide and a plain ATOMS per lane costs fewer issue slots than combining first (config 3: +3 %).
// (64-bit shared atomics are CAS loops, hence 32-bit bins flushed to the 64-bit global counters.)
__device__ __forceinline__ void hist_add(int *s_hist, int slot, int delta, int n_slots = 0)
{
    if (n_slots >= 18) {
        if (slot >= 0) atomicAdd(s_hist + slot, delta);
        return;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    if (slot >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(s_hist + slot, delta * __popc(peers));
}

__device__ __forceinline__ u64 warp_sum(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Warp sum of values below 2^52 with two REDUX instead of ten shuffles: 26-bit halves cannot overflow 32 bits
// when 32 lanes are added.
__device__ __forceinline__ u64 warp_sum_redux(u64 v)
{
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0x3ffffffu));
    const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(v >> 26));
    return (u64)lo + ((u64)hi << 26);
}
__device__ __forceinline__ u64 warp_incl_scan(u64 v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---------------------------------------------------------------- per-warp work queues
//
// Only a few lanes of a warp have a target candidate for a given read (config 3: 6 %, config 4: about half), and a
// candidate walk visits one to a handful of records, so walking where the candidate is found runs at a few lanes'
// occupancy.  Instead every (read, layer, candidate record) is an ITEM in a queue in shared memory owned by the
// warp, and the warp works the queue in ROUNDS: 32 items are popped, every lane visits exactly ONE record, and an
// item whose walk goes on is pushed back with the next record.  Rounds only run while 32 items are there (what is
// left waits for the next tile), so every round has all lanes busy; the queue is emptied when the CTA moves to
// another iteration.  The same queue type carries the reads whose blocked-interval probe said "maybe" (drained once
// per tile: the labels are needed before the targets are looked at).
template <typename C>
struct WalkItem {
    C st;         // read start
    uint32_t L;   // read length
    int a;        // record index; bit 30 = the run is crowded and the walk has not yet skipped ahead (first visit)
    uint32_t sg;  // run (layer) index            | mask items: first candidate in the "ALL" tree (or -1)
    uint32_t n;   // draw index in the iteration  | mask items: barcode | slot in the warp's tile row << 16
};
constexpr int kQueueCap = 96;      // target items per warp: < 32 carried over + <= 64 pushed (a round pops 32 before it pushes back)
constexpr int kMaskQueueCap = 64;  // mask items per warp; drained early when the next 32 pushes may not fit
constexpr int kQueueSegBits = 22;  // the host keeps the number of runs below 2^22
constexpr int kCrowdedBit = 1 << 30;

// A warp's queue in shared memory: one 16-byte record per item (a single LDS.128 / STS.128) plus a side word that
// is only touched when somebody needs it (the draw index: hit records and pair thinning; mask items: always).
template <typename C> struct QueueRef;
template <> struct QueueRef<uint32_t> { // 20 bytes per item
    uint4 *main;    // {st, L, a, sg}
    uint32_t *side; // n
    __device__ __forceinline__ QueueRef(void *base, int cap) : main((uint4 *)base), side((uint32_t *)((uint4 *)base + cap)) {}
    __device__ __forceinline__ void put(int i, const WalkItem<uint32_t> &w, bool with_n) const
    {
        main[i] = make_uint4(w.st, w.L, (uint32_t)w.a, w.sg);
        if (with_n) side[i] = w.n;
    }
    __device__ __forceinline__ WalkItem<uint32_t> get(int i, bool with_n) const
    {
        const uint4 v = main[i];
        return WalkItem<uint32_t>{v.x, v.y, (int)v.z, v.w, with_n ? side[i] : 0u};
    }
};
template <> struct QueueRef<uint64_t> { // 24 bytes per item
    uint4 *main; // {st low, st high, L, a}
    uint2 *side; // {sg, n}
    __device__ __forceinline__ QueueRef(void *base, int cap) : main((uint4 *)base), side((uint2 *)((uint4 *)base + cap)) {}
    __device__ __forceinline__ void put(int i, const WalkItem<uint64_t> &w, bool) const
    {
        main[i] = make_uint4((uint32_t)w.st, (uint32_t)(w.st >> 32), w.L, (uint32_t)w.a);
        side[i] = make_uint2(w.sg, w.n);
    }
    __device__ __forceinline__ WalkItem<uint64_t> get(int i, bool) const
    {
        const uint4 v = main[i];
        const uint2 t = side[i];
        return WalkItem<uint64_t>{(uint64_t)v.x | (uint64_t)v.y << 32, v.z, (int)v.w, t.x, t.y};
    }
};

// Called by all 32 lanes; n (the queue length) is warp-uniform.
template <typename C>
__device__ __forceinline__ void queue_push(const QueueRef<C> &q, int &n, bool want, C st, uint32_t L, int a, uint32_t sg, uint32_t dn, bool with_n)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (want) q.put(n + __popc(m & ((1u << (threadIdx.x & 31)) - 1u)), WalkItem<C>{st, L, a, sg, dn}, with_n);
    n += __popc(m);
}

// One round: pop up to 32 items, one record visit per lane, push the walks that go on; returns the new queue length.
// (Compiled in line at every call site: as a real call it spills 300 bytes and config 4 loses a third.)
template <typename C, typename Src>
__device__ __forceinline__ int queue_round(const DevCtx<C> &cx, const Src &src, const DevOut &out, u64 *__restrict__ tallies_it,
                                            int it, const QueueRef<C> q, int n, bool with_n, uint32_t *s_hot)
{
    const int lane = threadIdx.x & 31;
    const int take = n < 32 ? n : 32;
    const bool on = lane < take;
    WalkItem<C> w{};
    if (on) w = q.get(n - take + lane, with_n);
    __syncwarp();
    n -= take;
    bool more = false;
    if (on) {
        const C rs = w.st, re = (C)(w.st + w.L);
        int j = w.a & ~kCrowdedBit;
        bool live = true;
        if (w.a & kCrowdedBit) { // crowded run, first visit: bisect to the first record that can overlap
            const SegDesc d = load_desc(cx.seg + w.sg);
            j = skip_crowded(d, cx.tgrid, cx.te, j, rs);
            live = j < d.hi;
        }
        if (live) {
            ESL_CHK(j >= 0 && j < cx.n_trec, "target record index");
            const TargetRec<C> r = load_rec(cx.trec + j);
            if (r.s < re) { // else: starts ascend, nothing further overlaps
                if (r.e > rs) tally_hit<C, Src, true>(cx, src, out, tallies_it, it, (i64)w.n, r, rs, re, 1, s_hot);
                more = r.s_next < re; // known without touching the next record (all ones behind a run's last record)
            }
            w.a = j + 1;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, more);
    if (more) q.put(n + __popc(m & ((1u << lane) - 1u)), w, with_n);
    n += __popc(m);
    __syncwarp();
    return n;
}

// Blocked test of a queued read, starting from the first candidates its probes found (a < 0: that tree has none).
// Trees: the barcode's own, then "ALL"; hits in (start, end) order; one uniform per visited hit; u < weight blocks
// (read_operations.py:59-67).  Every group of blocked records ends with a sentinel whose start is all ones.
template <typename C, typename Src>
__device__ __forceinline__ bool blocked_from(const DevCtx<C> &cx, const Src &src, int it, i64 n, int bc, C rs, C re, int a_bc, int a_all)
{
    int j = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        int k = pass == 0 ? a_bc : a_all;
        if (k < 0) continue;
        if (k & kCrowdedBit) {
            const SegDesc d = load_desc(cx.grp + (pass == 0 ? bc : cx.n_barcodes));
            k = skip_crowded(d, cx.mgrid, cx.mpmax, k & ~kCrowdedBit, rs);
        }
        for (;; ++k) {
            ESL_CHK(k >= 0 && k < cx.n_mrec, "blocked record index");
            const MaskRec<C> r = load_rec(cx.mrec + k);
            if (r.s >= re) break; // (the sentinel ends every group)
            if (r.e > rs) {
                if (!Src::kBounded && r.w >= 1.0) return true; // u < 1 <= w for every u: no need to draw it
                const double u = src.block_u(it, n, j++);
                if (u < r.w) return true;
            }
        }
    }
    return false;
}

// Shared-memory layout of the bulk kernel, in 8-byte words, one definition for host (size) and device (offsets).
// The arrays of compile-time size come first, so the hot ones sit at constant addresses.
struct BulkSmem {
    size_t queue, mqueue, mflags, hot, lut, sched, acc, sgrid, desc, hist, cdf, mbits, qtab, total; // word offsets; total in BYTES
};
__host__ __device__ inline BulkSmem bulk_smem_layout(size_t coord_bytes, bool mask, bool shared, bool hot, int B, int sgrid_n, int mbits_words, bool quant)
{
    const size_t item = coord_bytes == 4 ? 20 : 24, warps = kBulkBlock / 32;
    BulkSmem o;
    size_t w = 0;
    o.queue = w; w += shared ? 0 : (warps * kQueueCap * item + 7) / 8;
    o.mqueue = w; w += mask ? (warps * kMaskQueueCap * item + 7) / 8 : 0;
    o.mflags = w; w += mask ? warps * kBulkIpt * 4 / 8 : 0;
    o.hot = w; w += hot ? kHotWords / 2 : 0;
    o.lut = w; w += B > 1 ? kCdfLut : 0; // (one barcode: no table)
    o.sched = w; w += 2;
    o.acc = w; w += (warps + 2) & ~(size_t)1; // (even: what follows keeps its 16-byte alignment)
    o.sgrid = w; w += ((size_t)(shared ? 2 * B * sgrid_n : 0) * coord_bytes + 15) / 16 * 2;
    o.desc = w; w += (size_t)(B <= kSmemDescs ? B : 0) * 2;
    o.hist = w; w += ((size_t)(B + 2) * 4 + 7) / 8;
    o.cdf = w; w += ((size_t)B * 4 + 7) / 8;
    o.mbits = w; w += mask ? ((size_t)mbits_words * 4 + 7) / 8 : 0;
    o.qtab = w; w += quant ? ((size_t)(kQuantBins + 1) * 4 + 7) / 8 : 0;
    o.total = w * 8;
    return o;
}

// ---------------------------------------------------------------- bulk kernel

// One launch covers the absolute tiles [tile0, tile0 + tiles_launch) of every iteration; inside that window iteration
// `it` owns the draws [lo(it), hi(it)) (lo is a multiple of the tile size).  Draw indices are < 2^32 (checked on the
// host).  Scheduling: every iteration has a tile counter in global memory (`sched[it]`, zeroed by the host call) from
// which CTAs claim kBulkClaim tiles at a time.  The CTAs are split into kIterSlots groups; group s works through the
// iterations s, s + kIterSlots, ... one after the other, so only a handful of iterations are in flight at any time:
// the tally rows their atomics hit (config 3: 5.8 MB per iteration) stay in L2 instead of being read-modified-written
// in HBM, a CTA changes iteration (flushes its histogram, empties its queue) only every few hundred tiles, and the
// dynamic claims leave no CTA with a long tail.  A group that runs out of iterations takes what is left of the others'.
//
// A thread resolves its four consecutive draws in PHASES, each phase running over all its reads before the next
// starts, so the dependent memory accesses of different reads overlap (pool gather -> start -> grid entry are two L2
// round trips per read; the reads of a thread being in flight together hides them):
//   1 draw  2 blocked probes, queue, drain  3 label histogram  4 target probes, queue, rounds  5 tile sum
// Warps only meet when the CTA claims tiles (one barrier, the next claim is fetched while the current one is worked
// on) or moves to the next iteration and flushes its shared label histogram.
#ifndef ESL_ITER_SLOTS
#define ESL_ITER_SLOTS 4
#endif
constexpr int kIterSlots = ESL_ITER_SLOTS; // (measured, config 3 / config 4 in G reads/s: 2 groups 140.2 / 55.0, 4 groups 140.6 / 59.3, 8 groups 141.9 / 59.4)
template <typename C, typename Src, bool kMask, bool kShared, bool kHot>
__global__ void __launch_bounds__(kBulkBlock, bulk_min_blocks<C, kShared>())
k_place_tally_bulk(const __grid_constant__ DevCtx<C> cx, const __grid_constant__ Src src, const DevOut out, const i64 *__restrict__ lo_arr,
                   const i64 lo_const, const i64 *__restrict__ hi_arr, const i64 hi_const, const int tile0,
                   const int tiles_launch, const int n_iter, u64 *__restrict__ sched)
{
    extern __shared__ u64 smem[];
    const int B = cx.n_barcodes;
    const bool quant = !Src::kHasLabel && Src::kPoolLengths && cx.qtab != nullptr; // native stream with the quantile sampler
    const BulkSmem lay = bulk_smem_layout(sizeof(C), kMask, kShared, kHot, B, kShared ? cx.sgrid_n : 0, kMask ? cx.mbits_words : 0, quant);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kItemBytes = sizeof(C) == 4 ? 20 : 24;
    const QueueRef<C> queue((char *)(smem + lay.queue) + wid * kQueueCap * kItemBytes, kQueueCap);
    const QueueRef<C> mqueue((char *)(smem + lay.mqueue) + wid * kMaskQueueCap * kItemBytes, kMaskQueueCap);
    const bool with_n = out.hits != nullptr || cx.thin; // the walks need the draw index only for hit records / thinning
    uint32_t *const mflags = (uint32_t *)(smem + lay.mflags) + wid * kBulkIpt; // bit l of word i: read i of lane l is blocked
    // private tallies of the hot targets + their rows (kHot: the host flagged some; otherwise the null pointer is a
    // compile-time constant and every hot-target branch below folds away)
    uint32_t *const s_hot = kHot ? (uint32_t *)(smem + lay.hot) : nullptr;
    uint2 *const s_lut = (uint2 *)(smem + lay.lut);             // [kCdfLut]
    i64 *const s_sched = (i64 *)(smem + lay.sched);             // [2] the claimed tile, broadcast to the CTA
    // per-warp counted bases [warps] and the CTA's reads [1] since the last flush, touched once per tile by one lane.
    // General variant: here and not in registers (at its 64-register cap the compiler kept them in local memory, and
    // the reload behind a thrashed L1 was 5.7 % of config 4's stall samples).  Small-target variant: registers (it has
    // them to spare; through shared memory config 2 loses 4 %).
    u64 *const s_acc = smem + lay.acc;
    u64 acc_bases = 0, acc_reads = 0;
    C *const s_grid = (C *)(smem + lay.sgrid);                  // [B][sgrid_n] pairs (small target sets)
    int4 *const s_desc = (int4 *)(smem + lay.desc);             // [min(B, kSmemDescs)] first-layer descriptors
    const int n_sdesc = B <= kSmemDescs ? B : 0;
    int *const s_hist = (int *)(smem + lay.hist);               // [B + 2] reads per label since the last flush
    uint32_t *const s_cdf = (uint32_t *)(smem + lay.cdf);       // [B]
    uint32_t *const s_mbits = (uint32_t *)(smem + lay.mbits);   // [mbits_words] blocked-bucket map, two bits per bucket
    uint32_t *const s_qtab = (uint32_t *)(smem + lay.qtab);     // [kQuantBins + 1] quantile table of the read lengths
    if (quant)
        for (int i = tid; i <= kQuantBins; i += kBulkBlock) s_qtab[i] = cx.qtab[i];
    const DrawTabs tabs{s_cdf, s_lut, quant ? s_qtab : nullptr};
    if (kMask)
        for (int i = tid; i < cx.mbits_words; i += kBulkBlock) s_mbits[i] = cx.mbits[i];
    const int n_sgrid = kShared ? 2 * B * cx.sgrid_n : 0;
    for (int i = tid; i < n_sgrid; i += kBulkBlock) s_grid[i] = cx.sgrid[i];
    for (int i = tid; i < n_sdesc; i += kBulkBlock) s_desc[i] = __ldg(reinterpret_cast<const int4 *>(cx.seg + i));
    for (int i = tid; i < B + 2; i += kBulkBlock) s_hist[i] = 0;
    if (kMask && lane < kBulkIpt) mflags[lane] = 0;
    if (tid <= kBulkBlock / 32) s_acc[tid] = 0;
    if (s_hot)
        for (int i = tid; i < kHotWords; i += kBulkBlock)
            s_hot[i] = i >= kHotSlots * 4 && i - kHotSlots * 4 < cx.n_hot ? (uint32_t)cx.hot_row[i - kHotSlots * 4] : 0u;
    if (!Src::kHasLabel && cx.cdf_u32 && B > 1) {
        for (int i = tid; i < B; i += kBulkBlock) s_cdf[i] = cx.cdf_u32[i];
        for (int i = tid; i < kCdfLut; i += kBulkBlock) s_lut[i] = cx.cdf_lut[i];
    }
    __syncthreads();

    const bool plain_labels = !Src::kHasLabel && B == 1 && !kMask; // every read is label 0
    int qn = 0, hot_tiles = 0, parity = 0;

    // The CTA's schedule (one loop body, so that it is compiled in line): first the iterations of its own group in
    // ascending order, then what is left of the other groups' -- their iterations are consumed in order, so walking a
    // group's list backwards until an iteration yields nothing visits everything that can still hold unclaimed tiles.
    const int n_slots = n_iter < kIterSlots ? n_iter : kIterSlots;
    const int slot = (int)(blockIdx.x % (unsigned)n_slots);
    int helping = 0; // 0: own group; k > 0: the group k places after it
    for (int it = slot;;) {
        if (helping == 0 && it >= n_iter) {
            if (++helping >= n_slots) break;
            const int other = (slot + helping) % n_slots;
            it = other + (n_iter - 1 - other) / n_slots * n_slots;
        }
        // claim and process tiles of iteration `it` until its counter runs out
        i64 l = lo_arr ? lo_arr[it] : lo_const, h = hi_arr ? hi_arr[it] : hi_const;
        if (Src::kBounded) { const i64 c = src.count(it); h = h < c ? h : c; }
        const uint32_t lo = (uint32_t)l, hi = (uint32_t)(h > l ? h : l);
        u64 *const tallies_it = out.tallies + (i64)it * cx.n_targets * 3;
        u64 *const tile_sum_it = out.tile_sum + (i64)it * out.tiles_per_iter;
        // tiles of the window that hold draws of this iteration: [t_lo, t_hi) relative to tile0
        const int t_lo = lo > (uint32_t)tile0 * kBulkTile ? (int)(lo / kBulkTile) - tile0 : 0;
        i64 t_end = ((i64)hi + kBulkTile - 1) / kBulkTile - tile0;
        const int t_hi = (int)(t_end < 0 ? 0 : (t_end > tiles_launch ? tiles_launch : t_end));
        bool any = false;
        i64 next = 0;
        if (tid == 0) next = t_lo + (i64)atomicAdd(sched + it, (u64)kBulkClaim);
        for (;;) {
            if (tid == 0) s_sched[parity] = next;
            __syncthreads();
            const i64 g0 = s_sched[parity];
            parity ^= 1;
            if (g0 >= t_hi) break;
            if (tid == 0) next = t_lo + (i64)atomicAdd(sched + it, (u64)kBulkClaim); // fetched while this claim is worked on
            any = true;
            const int g1 = (int)(g0 + kBulkClaim < t_hi ? g0 + kBulkClaim : t_hi);
            for (int t = (int)g0 + tile0; t < g1 + tile0; ++t) {
            const uint32_t base = (uint32_t)t * kBulkTile;
            if (s_hot && ++hot_tiles == kHotFlushTiles) { // (block-uniform)
                __syncthreads();
                flush_hot(s_hot, cx.n_hot, tallies_it);
                hot_tiles = 0;
                __syncthreads();
            }

            const uint32_t n0 = base + (uint32_t)tid * kBulkIpt; // this thread's draws: n0 .. n0 + 3
            uint32_t L[kBulkIpt];
            C st[kBulkIpt];
            int label[kBulkIpt];
            // phase 1: draws (Philox, pool gather, start, barcode)
            if (hi - base >= (uint32_t)kBulkTile) { // full tile (block-uniform; all but the last tile of a range)
                src.draw4(cx, tabs, it, n0, L, st, label);
            } else {
#pragma unroll
                for (int i = 0; i < kBulkIpt; ++i) {
                    // branch-free tail handling: a thread past the end of the range redraws the range's last draw and is
                    // masked out afterwards (label = invalid, length 0)
                    const uint32_t n = n0 + i;
                    src.draw(cx, tabs, it, (i64)(n < hi ? n : hi - 1), L[i], st[i], label[i]);
                    if (n >= hi) { L[i] = 0; label[i] = kLabelInvalid; }
                }
            }
            // target probes of the general path are issued here, ahead of the blocked-region phase: both only need the
            // read's start, so their L2 round trips overlap instead of queueing up behind each other
            Probe<C> pt[kShared ? 1 : kBulkIpt];
            if (!kShared) {
#pragma unroll
                for (int i = 0; i < kBulkIpt; ++i) {
                    const bool ok = label[i] >= 0 && label[i] < B;
                    const int b = ok ? label[i] : 0;
                    const int4 dv = n_sdesc ? s_desc[b] : __ldg(reinterpret_cast<const int4 *>(cx.seg + b));
                    ESL_CHK(dv.z + (int)((u64)st[i] >> (dv.w & 0xff)) + 1 < cx.n_tgrid, "target grid index");
                    const GridEntry<C> e = load_entry(cx.tgrid + dv.z + (int)((u64)st[i] >> (dv.w & 0xff)));
                    pt[kShared ? 0 : i] = Probe<C>{e.j | ((dv.w & 0x100) ? kCrowdedBit : 0), ok ? e.s : (C) ~(C)0};
                }
            }
            u64 my = 0;
            // the warp tests its queued reads against the blocked intervals, all lanes busy; verdicts go to mflags
            auto drain_mask = [&](int &mn) {
                __syncwarp();
                for (int k = lane; k < mn; k += 32) {
                    const WalkItem<C> w = mqueue.get(k, true);
                    const int slot = (int)(w.n >> 16);
                    const int owner = slot & 31; // the draw index follows from the slot: lane `owner`, read slot >> 5
                    const uint32_t nd = base + (uint32_t)((wid * 32 + owner) * kBulkIpt + (slot >> 5));
                    if (blocked_from(cx, src, it, (i64)nd, (int)(w.n & 0xffffu), w.st, (C)(w.st + w.L), w.a, (int)w.sg))
                        atomicOr(mflags + (slot >> 5), 1u << owner);
                }
                __syncwarp();
                mn = 0;
            };
            if (!Src::kHasLabel && kMask) {
                // phase 2: blocked regions -- a shared-memory bucket map clears most reads and settles most of the blocked
                // ones (buckets wholly inside certain intervals) without a global load; the others
                // probe the barcode's tree (if any barcode has one) and the "ALL" tree, all four reads of the thread in
                // flight together; reads with a candidate are queued and the warp walks them with its lanes busy
                int mn = 0;
                unsigned sure_bits = 0; // bit i: read i lies in buckets that are blocked for certain
                const int2 d_all = __ldg(reinterpret_cast<const int2 *>(cx.grp + B) + 1);
#pragma unroll
                for (int i = 0; i < kBulkIpt; ++i) {
                    const C re = (C)(st[i] + L[i]);
                    const uint32_t b0 = (uint32_t)((u64)st[i] >> cx.mbits_shift), b1 = (uint32_t)((u64)(re - 1) >> cx.mbits_shift);
                    const bool wide = b1 - b0 > 1u;
                    const uint32_t b1c = wide ? b0 : b1; // (keeps the second look-up in range whatever the length)
                    const uint32_t m0 = s_mbits[b0 >> 4] >> ((b0 & 15) * 2), m1 = s_mbits[b1c >> 4] >> ((b1c & 15) * 2);
                    const bool sure = !wide && (m0 & m1 & 2u) && label[i] >= 0;
                    const bool maybe = !sure && (wide || ((m0 | m1) & 1u));
                    sure_bits |= (unsigned)sure << i;
                    int a_bc = -1, a_all = -1;
                    if (maybe && (cx.mask_mode & 1)) {
                        const int2 d = __ldg(reinterpret_cast<const int2 *>(cx.grp + (label[i] >= 0 ? label[i] : 0)) + 1);
                        const GridEntry<C> e = load_entry(cx.mgrid + d.x + (int)((u64)st[i] >> (d.y & 0xff)));
                        if (e.s < re) a_bc = e.j | ((d.y & 0x100) ? kCrowdedBit : 0);
                    }
                    if (maybe && (cx.mask_mode & 2)) {
                        const GridEntry<C> e = load_entry(cx.mgrid + d_all.x + (int)((u64)st[i] >> (d_all.y & 0xff)));
                        if (e.s < re) a_all = e.j | ((d_all.y & 0x100) ? kCrowdedBit : 0);
                    }
                    queue_push(mqueue, mn, label[i] >= 0 && (a_bc >= 0 || a_all >= 0), st[i], L[i], a_bc, (uint32_t)a_all,
                               (uint32_t)label[i] | (uint32_t)(i * 32 + lane) << 16, true);
                    if (mn > kMaskQueueCap - 32) drain_mask(mn);
                }
                drain_mask(mn);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kBulkIpt; ++i) {
                    uint32_t counted = L[i];
                    if ((mflags[i] >> lane | sure_bits >> i) & 1) {
                        label[i] = cx.consuming ? kLabelBlockedConsuming : kLabelBlockedNotConsuming;
                        counted = cx.consuming ? L[i] : 0u;
                    }
                    my += counted;
                }
                __syncwarp();
                if (lane < kBulkIpt) mflags[lane] = 0;
            } else {
#pragma unroll
                for (int i = 0; i < kBulkIpt; ++i) my += (Src::kHasLabel && label[i] == kLabelBlockedNotConsuming) ? 0u : L[i];
            }
            if (!plain_labels) { // phase 3: label histogram
                if (B <= 2) {
                    // at most four label slots (one or two barcodes + the two blocked labels): the thread's four reads are
                    // counted in one packed word, byte s = reads of slot s, the warp adds them with ONE REDUX (a byte holds
                    // at most 4 x 32 = 128) and lanes 0..3 add a byte each to the shared histogram -- instead of a
                    // match.any + atomic per read
                    unsigned packed = 0;
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i)
                        if (label[i] != kLabelInvalid && label[i] < B) packed += 1u << (8 * label_slot(label[i], B));
                    packed = __reduce_add_sync(0xffffffffu, packed);
                    const unsigned cnt = lane < 4 ? (packed >> (8 * lane)) & 0xffu : 0u;
                    if (cnt) atomicAdd(s_hist + lane, (int)cnt);
                } else {
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i)
                        hist_add(s_hist, label[i] != kLabelInvalid && label[i] < B ? label_slot(label[i], B) : -1, 1, B + 2);
                }
            }
            if (kShared) { // phase 4, small target sets: shared-memory pre-filter, exact path for the rest
                unsigned maybe = 0; // bit i: read i may overlap a first-layer target
                if (plain_labels && hi - base >= (uint32_t)kBulkTile) { // one barcode, no mask, full tile: every read is valid
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i) {
                        const C *g = s_grid + 2 * (int)((u64)st[i] >> cx.sgrid_shift);
                        maybe |= (unsigned)(g[0] < (C)(st[i] + L[i]) && (!Src::kPoolLengths || st[i] < g[1])) << i;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i) {
                        const bool ok = label[i] >= 0 && label[i] < B;
                        ESL_CHK((int)((u64)st[i] >> cx.sgrid_shift) < cx.sgrid_n, "shared grid index");
                        const C *g = s_grid + 2 * ((ok ? label[i] : 0) * cx.sgrid_n + (int)((u64)st[i] >> cx.sgrid_shift));
                        // g[0]: first candidate's start; g[1]: largest end a read from this bucket can meet
                        maybe |= (unsigned)(ok && g[0] < (C)(st[i] + L[i]) && (!Src::kPoolLengths || st[i] < g[1])) << i;
                    }
                }
                if (maybe) {
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i)
                        if (maybe >> i & 1)
                            tally_first_layer(cx, src, out, tallies_it, it, (i64)(n0 + i), st[i], (C)(st[i] + L[i]), label[i], 1, s_hot);
                }
                if (cx.any_extra) { // the pre-filter only covers first layers
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i) {
                        if (label[i] < 0 || label[i] >= B) continue;
                        const int n_extra = (int)((uint32_t)(n_sdesc ? s_desc[label[i]].w : __ldg(&cx.seg[label[i]].flags)) >> 16);
                        if (n_extra)
                            tally_extra_layers(cx, src, out, tallies_it, it, (i64)(n0 + i), label[i], n_extra,
                                               st[i], (C)(st[i] + L[i]), 1, s_hot);
                    }
                }
            } else { // phase 4: targets -- queue the reads whose probe said "maybe", work the queue in full rounds
#pragma unroll
                for (int i = 0; i < kBulkIpt; ++i) {
                    // (a read the blocked-region phase relabelled keeps its probe but is not tallied)
                    queue_push(queue, qn, label[i] >= 0 && pt[i].s < (C)(st[i] + L[i]), st[i], L[i], pt[i].a, (uint32_t)label[i], n0 + i, with_n);
                    if (i & 1) // two reads per batch of rounds: half the copies of the round's code
                        while (qn >= 32) qn = queue_round(cx, src, out, tallies_it, it, queue, qn, with_n, s_hot);
                }
                if (cx.any_extra) { // nesting targets somewhere: layer k of every read's barcode, k ascending
                    int ne_max = 0;
#pragma unroll
                    for (int i = 0; i < kBulkIpt; ++i) {
                        const bool ok = label[i] >= 0 && label[i] < B;
                        const int ne = ok ? (int)((uint32_t)(n_sdesc ? s_desc[label[i]].w : __ldg(&cx.seg[label[i]].flags)) >> 16) : 0;
                        ne_max = ne > ne_max ? ne : ne_max;
                    }
                    ne_max = __reduce_max_sync(0xffffffffu, ne_max);
                    for (int k = 0; k < ne_max; ++k) {
#pragma unroll
                        for (int i = 0; i < kBulkIpt; ++i) { // (layer counts re-read from shared memory: keeping them live spills)
                            const bool ok = label[i] >= 0 && label[i] < B;
                            const int ne = ok ? (int)((uint32_t)(n_sdesc ? s_desc[label[i]].w : __ldg(&cx.seg[label[i]].flags)) >> 16) : 0;
                            Probe<C> pe{0, (C) ~(C)0};
                            int sg = 0;
                            if (k < ne) {
                                sg = B + __ldg(cx.extra_off + label[i]) + k;
                                const int2 d = __ldg(reinterpret_cast<const int2 *>(cx.seg + sg) + 1);
                                const GridEntry<C> e = load_entry(cx.tgrid + d.x + (int)((u64)st[i] >> (d.y & 0xff)));
                                pe = Probe<C>{e.j | ((d.y & 0x100) ? kCrowdedBit : 0), e.s};
                            }
                            queue_push(queue, qn, pe.s < (C)(st[i] + L[i]), st[i], L[i], pe.a, (uint32_t)sg, n0 + i, with_n);
                            if (i & 1)
                                while (qn >= 32) qn = queue_round(cx, src, out, tallies_it, it, queue, qn, with_n, s_hot);
                        }
                    }
                }
            }
            my = warp_sum_redux(my); // phase 5: tile sum (read only when the cut-off has to be taken back)
            if (lane == 0) {
                ESL_CHK(t >= 0 && t < out.tiles_per_iter, "tile sum index");
                if (my) atomicAdd(tile_sum_it + t, my);
                if (kShared) acc_bases += my; else s_acc[wid] += my;
            }
            if (tid == 0) {
                const uint32_t left = hi - base, got = left < (uint32_t)kBulkTile ? left : (uint32_t)kBulkTile;
                if (kShared) acc_reads += got; else s_acc[kBulkBlock / 32] += got;
            }
            } // tiles of the claim
        }
        if (any) { // (block-uniform) what the CTA gathered for this iteration: empty the walk queue, flush histogram / sums
            if (!kShared)
                while (qn > 0) qn = queue_round(cx, src, out, tallies_it, it, queue, qn, with_n, s_hot);
            __syncthreads();
            if (s_hot) { flush_hot(s_hot, cx.n_hot, tallies_it); hot_tiles = 0; }
            for (int i = tid; i < B + 2; i += kBulkBlock)
                if (s_hist[i]) { atomicAdd(out.label_counts + (i64)it * (B + 2) + i, (u64)(uint32_t)s_hist[i]); s_hist[i] = 0; }
            if (!kShared) {
                if (lane == 0) { acc_bases = s_acc[wid]; s_acc[wid] = 0; }
                if (tid == 0) { acc_reads = s_acc[kBulkBlock / 32]; s_acc[kBulkBlock / 32] = 0; }
            }
            if (lane == 0 && acc_bases) atomicAdd(out.n_bases + it, acc_bases);
            if (tid == 0 && acc_reads) {
                atomicAdd(out.n_reads + it, acc_reads);
                if (plain_labels) atomicAdd(out.label_counts + (i64)it * (B + 2), acc_reads);
            }
            acc_bases = acc_reads = 0;
            __syncthreads();
        }
        if (helping == 0) {
            it += n_slots;
        } else if (any && it >= n_slots) {
            it -= n_slots;
        } else { // this group has nothing left: on to the next one
            if (++helping >= n_slots) break;
            const int other = (slot + helping) % n_slots;
            it = other + (n_iter - 1 - other) / n_slots * n_slots;
        }
    }
}

// With blocked regions and consuming = False a blocked draw adds nothing to the covered length
// (read_operations.py:116-117), so the cut-off moves out by the unknown blocked fraction.  After the first bulk
// pass the per-iteration mean counted length per draw is known from n_bases; this kernel sizes a second bulk
// pass [n1, n2(it)) with the same z-sigma rule on the REMAINING coverage (sd bound: counted is 0 or L, so
// E[c^2] <= E[L^2]).  Any overshoot is again corrected exactly by k_cutoff_tail.
__global__ void k_plan_second_pass(const u64 *__restrict__ n_bases, const int n_iter, const i64 thr, const i64 n1,
                                   const i64 n_cap, const double pool_m2, const double z, i64 *__restrict__ lo_out,
                                   i64 *__restrict__ hi_out)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_iter) return;
    i64 n2 = n1;
    const double S = (double)n_bases[it], rem = (double)thr - S;
    if (n1 > 0 && rem > 0 && S > 0) {
        const double m = S / (double)n1;
        const double var = pool_m2 - m * m, sd = var > 0 ? sqrt(var) : 0.0;
        const double r = (-z * sd + sqrt(z * z * sd * sd + 4.0 * m * rem)) / (2.0 * m);
        i64 extra = (i64)floor(r * r) - 1;
        extra -= extra % kBulkTile;
        if (extra > 0) n2 = n1 + extra;
        if (n2 > n_cap) n2 = n_cap;
    }
    lo_out[it] = n1;
    hi_out[it] = n2;
}

// ---------------------------------------------------------------- exact cut-off kernel (K5)

// Block-wide exclusive scan of one u64 per thread; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ u64 block_excl_scan(u64 v, u64 *s_warp, u64 *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u64 incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const u64 w = s_warp[lane]; // kTailBlock / 32 == 32 warps
        const u64 wi = warp_incl_scan(w, lane);
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    const u64 ex = s_warp[wid] + incl - v;
    *total = s_warp[32];
    __syncthreads();
    return ex;
}

// Block-wide sums of two values at once (two barriers): every thread gets both totals.
__device__ __forceinline__ void block_sum2(u64 a, u64 b, u64 *s_pair, u64 &tot_a, u64 &tot_b)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { s_pair[2 * wid] = a; s_pair[2 * wid + 1] = b; }
    __syncthreads();
    if (wid == 0) {
        a = warp_sum(s_pair[2 * lane]); b = warp_sum(s_pair[2 * lane + 1]); // kTailBlock / 32 == 32 warps
        if (lane == 0) { s_pair[64] = a; s_pair[65] = b; }
    }
    __syncthreads();
    tot_a = s_pair[64]; tot_b = s_pair[65];
    __syncthreads();
}

template <typename C, typename Src>
__global__ void __launch_bounds__(kTailBlock)
k_cutoff_tail(const __grid_constant__ DevCtx<C> cx, const __grid_constant__ Src src, const DevOut out, const i64 *__restrict__ n_bulk_arr,
              const i64 n_bulk_const)
{
    extern __shared__ u64 smem[];
    u64 *s_warp = smem;                         // [33] scans
    u64 *s_pair = smem + 40;                    // [66] pair sums
    int *s_hist = (int *)(smem + 112);          // [B + 2] signed deltas to the label counts
    uint32_t *s_cdf = (uint32_t *)(s_hist + cx.n_barcodes + 2);
    uint2 *s_lut = (uint2 *)(((uintptr_t)(s_cdf + cx.n_barcodes) + 7) & ~(uintptr_t)7);
    const DrawTabs tabs{s_cdf, s_lut, cx.qtab}; // (the quantile table is read from global memory here: 0.4 % of the draws)
    __shared__ i64 s_cut_tile;
    __shared__ u64 s_cut_prefix;
    const int tid = threadIdx.x;
    const int B = cx.n_barcodes;
    const int it = blockIdx.x;
    for (int i = tid; i < B + 2; i += kTailBlock) s_hist[i] = 0;
    if (!Src::kHasLabel && cx.cdf_u32 && B > 1) {
        for (int i = tid; i < B; i += kTailBlock) s_cdf[i] = cx.cdf_u32[i];
        for (int i = tid; i < kCdfLut; i += kTailBlock) s_lut[i] = cx.cdf_lut[i];
    }
    if (tid == 0) { s_cut_tile = -1; s_cut_prefix = 0; }
    __syncthreads();

    const i64 n_total = src.count(it);
    i64 nb = n_bulk_arr ? n_bulk_arr[it] : n_bulk_const;
    if (nb > n_total) nb = n_total;
    const u64 bulk_sum = out.n_bases[it];
    const i64 thr = cx.thr;
    u64 *tallies_it = out.tallies + (i64)it * cx.n_targets * 3;

    i64 n, n_end;
    u64 covered;
    bool take_back; // true: the cut-off lies inside the bulk range; remove what bulk added beyond it
    if ((i64)bulk_sum >= thr && nb > 0) {
        take_back = true;
        const i64 tiles = (nb + kBulkTile - 1) / kBulkTile;
        const u64 *ts_it = out.tile_sum + (i64)it * out.tiles_per_iter;
        u64 run = 0;
        for (i64 b0 = 0; b0 < tiles; b0 += kTailBlock) { // first tile whose inclusive prefix reaches thr
            const i64 ti = b0 + tid;
            const u64 v = ti < tiles ? ts_it[ti] : 0;
            u64 tot;
            const u64 ex = block_excl_scan(v, s_warp, &tot);
            const bool hit = ti < tiles && (i64)(run + ex) < thr && (i64)(run + ex + v) >= thr;
            if (hit) { s_cut_tile = ti; s_cut_prefix = run + ex; }
            __syncthreads();
            if (s_cut_tile >= 0) break;
            run += tot;
        }
        n = s_cut_tile * kBulkTile;
        covered = s_cut_prefix;
        n_end = nb;
        if (tid == 0) atomicOr(out.status + it, 1);
    } else {
        take_back = false;
        n = nb;
        covered = bulk_sum;
        n_end = n_total;
    }

    i64 d_bases = 0, d_reads = 0; // deltas to the bulk totals (uniform across the block)
    bool exhausted = false;
#pragma unroll 1
    for (;;) {
        if (n >= n_end) { exhausted = !take_back; break; }
        uint32_t L[kTailIpt];
        C st[kTailIpt];
        int bc[kTailIpt], label[kTailIpt];
        uint32_t counted[kTailIpt];
        bool valid[kTailIpt];
        u64 mine = 0;
        const i64 idx0 = n + (i64)tid * kTailIpt; // thread-contiguous draw order so one scan orders the chunk
        if (idx0 + kTailIpt <= n_end) { // (n and the bulk range are multiples of 4: idx0 is aligned)
            src.draw4(cx, tabs, it, (uint32_t)idx0, L, st, bc);
#pragma unroll
            for (int i = 0; i < kTailIpt; ++i) {
                valid[i] = true;
                resolve_read(cx, src, it, idx0 + i, L[i], st[i], bc[i], label[i], counted[i]);
                mine += counted[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < kTailIpt; ++i) {
                const i64 idx = idx0 + i;
                valid[i] = idx < n_end;
                L[i] = 0; st[i] = 0; bc[i] = 0; label[i] = 0; counted[i] = 0;
                if (valid[i]) {
                    src.draw(cx, tabs, it, idx, L[i], st[i], bc[i]);
                    resolve_read(cx, src, it, idx, L[i], st[i], bc[i], label[i], counted[i]);
                    mine += counted[i];
                }
            }
        }
        u64 chunk_total;
        u64 pre = covered + block_excl_scan(mine, s_warp, &chunk_total);
        u64 inc_bases = 0; // counted bases / reads of this thread on the acted-on side of the cut
        int inc_reads = 0, excluded = 0;
#pragma unroll
        for (int i = 0; i < kTailIpt; ++i) {
            const bool included = (i64)pre < thr; // the while-condition seen by this draw (read_operations.py:100)
            pre += counted[i];
            excluded += valid[i] && !included;
            const bool act = valid[i] && (take_back ? !included : included);
            const int sign = take_back ? -1 : 1;
            hist_add(s_hist, act && label[i] < B ? label_slot(label[i], B) : -1, sign);
            if (!act) continue;
            inc_bases += counted[i];
            inc_reads += 1;
            if (label[i] >= 0 && label[i] < B)
                tally_read(cx, src, out, tallies_it, it, n + (i64)tid * kTailIpt + i, st[i], (C)(st[i] + L[i]), label[i], sign);
        }
        u64 tot_b, tot_rx; // one pass for the three totals: bases; reads and excluded draws packed (<= 4096 each)
        block_sum2(inc_bases, (u64)inc_reads | (u64)excluded << 32, s_pair, tot_b, tot_rx);
        const u64 tot_r = tot_rx & 0xffffffffull, tot_x = tot_rx >> 32;
        if (take_back) { d_bases -= (i64)tot_b; d_reads -= (i64)tot_r; }
        else { d_bases += (i64)tot_b; d_reads += (i64)tot_r; }
        covered += chunk_total;
        n += kTailChunk;
        if (!take_back && tot_x > 0) break; // the cut-off read was in this chunk
    }
    __syncthreads();
    for (int i = tid; i < B + 2; i += kTailBlock)
        if (s_hist[i]) out.label_counts[(i64)it * (B + 2) + i] += (u64)(i64)s_hist[i];
    if (tid == 0) {
        out.n_bases[it] = bulk_sum + (u64)d_bases;
        out.n_reads[it] += (u64)d_reads;
        if (exhausted && (i64)covered < thr) atomicOr(out.status + it, 2);
    }
}

// ---------------------------------------------------------------- auxiliary kernels

// Write the first n_reads draws of one native iteration in draw order (coverage track / overlap lists).
template <typename C>
__global__ void __launch_bounds__(256)
k_materialize(const __grid_constant__ DevCtx<C> cx, const __grid_constant__ PhiloxSrc src, const i64 n_reads, i64 *__restrict__ o_start,
              i64 *__restrict__ o_length, int32_t *__restrict__ o_label)
{
    extern __shared__ u64 smem[];
    uint32_t *s_cdf = (uint32_t *)smem;
    uint2 *s_lut = (uint2 *)(((uintptr_t)(s_cdf + cx.n_barcodes) + 7) & ~(uintptr_t)7);
    for (int i = threadIdx.x; i < cx.n_barcodes; i += blockDim.x) s_cdf[i] = cx.cdf_u32[i];
    for (int i = threadIdx.x; i < kCdfLut; i += blockDim.x) s_lut[i] = cx.cdf_lut[i];
    const DrawTabs tabs{s_cdf, s_lut, cx.qtab};
    __syncthreads();
    for (i64 n = (i64)blockIdx.x * blockDim.x + threadIdx.x; n < n_reads; n += (i64)gridDim.x * blockDim.x) {
        uint32_t L, counted; C st; int bc, label;
        src.draw(cx, tabs, 0, n, L, st, bc);
        resolve_read(cx, src, 0, n, L, st, bc, label, counted);
        o_start[n] = (i64)st;
        o_length[n] = (i64)L;
        o_label[n] = label;
    }
}

// Binned coverage track of materialised reads (coverage.py:52-64,76-81 without the genome-length arrays): bins
// [label slot][bin] accumulate the number of read bases falling into each bin; the host divides by the bin
// width to get the reference's per-bin mean coverage.  One read per thread, a read spans few bins.
__global__ void __launch_bounds__(256)
k_coverage_bins(const i64 *__restrict__ start, const i64 *__restrict__ length, const int32_t *__restrict__ label,
                const i64 n_reads, const i64 bin_size, const i64 n_bins, const int n_barcodes, u64 *__restrict__ bins)
{
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (i64)gridDim.x * blockDim.x) {
        const int lab = label[r];
        if (lab >= n_barcodes) continue;
        const i64 s = start[r], e = s + length[r];
        u64 *row = bins + (i64)label_slot(lab, n_barcodes) * n_bins;
        for (i64 b = s / bin_size; b < n_bins && b * bin_size < e; ++b) {
            const i64 lo = b * bin_size > s ? b * bin_size : s;
            const i64 hi = (b + 1) * bin_size < e ? (b + 1) * bin_size : e;
            if (hi > lo) atomicAdd(row + b, (u64)(hi - lo));
        }
    }
}

// Replay of the reference's pair thinning (counting.py:46,51): `hits` are the qualifying (target, read) pairs of a
// replay in the reference's visiting order (sorted by iteration, target row, draw index), u[i] the uniform the
// reference drew for pair i; a pair is kept iff u <= scaling.  `kind` comes back from the record's sign bit of the
// overlap column (set by the collector for full matches).
__global__ void __launch_bounds__(256)
k_thin_pairs(const i64 *__restrict__ hits, const double *__restrict__ u, const i64 n_hits, const double scaling,
             const i64 n_targets, u64 *__restrict__ tallies)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits || !(u[i] <= scaling)) return;
    const i64 *h = hits + i * 4;
    const i64 ov = h[3] & 0x7fffffffffffffffLL;
    const int kind = h[3] < 0 ? 0 : 1;
    u64 *t = tallies + (h[0] * n_targets + h[1]) * 3;
    atomicAdd(t + kind, 1ull);
    atomicAdd(t + 2, (u64)ov);
}

// Moments of the length pool on the device (set-up): out = {sum, sum of squares (double bits), max, min}.  The
// pool is uploaded anyway; reducing it there costs microseconds instead of a scalar pass over 10^6 host integers.
__global__ void __launch_bounds__(256)
k_pool_moments(const uint32_t *__restrict__ pool, const i64 n, u64 *__restrict__ out_sum, double *__restrict__ out_sq,
               uint32_t *__restrict__ out_max, uint32_t *__restrict__ out_min)
{
    u64 s = 0;
    double q = 0.0;
    uint32_t mx = 0, mn = 0xFFFFFFFFu;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const uint32_t v = pool[i];
        s += v; q += (double)v * (double)v;
        mx = v > mx ? v : mx; mn = v < mn ? v : mn;
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, o);
        const uint32_t a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
        mx = a > mx ? a : mx; mn = b < mn ? b : mn;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out_sum, s);
        atomicAdd(out_sq, q);
        atomicMax(out_max, mx);
        atomicMin(out_min, mn);
    }
}

// Quantile table of the sorted length pool: q[k] = sorted[floor(k * n / kQuantBins)], q[kQuantBins] = the longest read.
__global__ void __launch_bounds__(256)
k_quantile_table(const uint32_t *__restrict__ sorted, const i64 n, uint32_t *__restrict__ q)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < kQuantBins) q[k] = sorted[(i64)(((u64)k * (u64)n) / kQuantBins)];
    else if (k == kQuantBins) q[k] = sorted[n - 1];
}

// Per-target moment accumulators over the iterations of a batch (operand of the multi-GPU sum-reduce).  Thread
// (target t, slice s) folds iterations s, s + n_slices, ... and adds its partial sums with atomics: small target sets
// get their parallelism from the iterations (100 targets x 100 iterations: one thread per pair), large ones from the
// targets (n_slices = 1, coalesced 24-byte rows).  The square sum of the bases is carried as hi * 2^32 + lo; partial
// sums may leave lo above 2^32, which every consumer folds back (hi << 32) + lo.
__global__ void __launch_bounds__(256)
k_moments(const i64 *__restrict__ tallies, const int n_iter, const i64 n_targets, const int n_slices, i64 *__restrict__ mom)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (t >= n_targets) return;
    i64 sf = 0, sf2 = 0, sp = 0, sp2 = 0, sb = 0;
    unsigned __int128 sb2 = 0;
    for (int it = s; it < n_iter; it += n_slices) {
        const i64 *r = tallies + ((i64)it * n_targets + t) * 3;
        const i64 f = r[0], p = r[1], b = r[2];
        sf += f; sf2 += f * f; sp += p; sp2 += p * p; sb += b;
        sb2 += (unsigned __int128)(u64)b * (u64)b;
    }
    u64 *m = (u64 *)(mom + t * 8);
    if (s == 0) atomicAdd(m, (u64)n_iter);
    atomicAdd(m + 1, (u64)sf); atomicAdd(m + 2, (u64)sf2); atomicAdd(m + 3, (u64)sp); atomicAdd(m + 4, (u64)sp2);
    atomicAdd(m + 5, (u64)sb);
    atomicAdd(m + 6, (u64)(sb2 & 0xffffffffull));
    atomicAdd(m + 7, (u64)(sb2 >> 32));
}

} // namespace esl
