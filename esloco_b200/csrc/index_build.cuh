// Device-side construction of the TARGET index for large, non-nesting target sets (esl_prepare).
//
// The host builder (esl_api.cu: build_layers + fill_grid) handles every geometry; at 240 k targets it costs ~3 ms of
// host time per rebuild on an idle box and more when eight ranks share the cores.  The common large case -- insertion
// targets of one length, disjoint ROIs: after sorting by (barcode, start) the ends ascend too -- needs no layering
// logic at all, only a sort and two embarrassingly parallel passes, so it is built here:
//   k_idx_keys      per target: validity (known barcode, not empty, starts inside the genome), sort keys
//   radix sort      by start, then (stable) by barcode            -- the staged baseline's LSD sort kernels
//   k_idx_bounds    per barcode: record range, grid shape and offset (one CTA, block scan), layer descriptors
//   k_idx_records   16-byte records {start, end, row, next start}; flags nesting (an end below its predecessor's) and
//                   hot targets -- either sends esl_prepare back to the host builder
//   k_idx_grid      per grid bucket: first record whose end exceeds the bucket start (binary search), crowded flag
// The result is the same structure the host builder writes for a single layer per barcode (the order of equal
// starts may differ, which no tally can see).  32-bit coordinates only.
#pragma once
#include "kernels.cuh"

namespace esl {

constexpr uint32_t kIdxNested = 1u, kIdxHot = 2u;

struct IdxResult { // read back by the host
    uint32_t flags;     // kIdxNested | kIdxHot: fall back to the host builder
    int32_t n_valid;    // indexed targets
    int32_t n_grid;     // grid entries written
    int32_t n_nonempty; // barcodes with at least one target
};

__host__ __device__ inline int idx_grid_shift(u64 n_intervals, u64 G, u64 buckets_per_interval)
{
    u64 want = 64 > buckets_per_interval * n_intervals ? 64 : buckets_per_interval * n_intervals;
    if (want < (G >> 16) + 1) want = (G >> 16) + 1;
    if (want > (1ull << 21)) want = 1ull << 21;
    int shift = 0;
    while (((G + 1) >> shift) + 2 > want) ++shift;
    return shift;
}

__device__ __forceinline__ uint32_t idx_clamp(i64 v, u64 G) { return (u64)v > G + 1 ? (uint32_t)(G + 1) : (uint32_t)v; }

__global__ void __launch_bounds__(256)
k_idx_keys(const i64 *__restrict__ ts, const i64 *__restrict__ te, const int32_t *__restrict__ tb, const int n, const int B,
           const u64 G, uint32_t *__restrict__ key_start, uint32_t *__restrict__ key_bc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const i64 s = ts[i], e = te[i];
    const int32_t b = tb[i];
    const bool valid = b >= 0 && b < B && e > s && (u64)s < G;
    key_start[i] = valid ? (uint32_t)s : 0xFFFFFFFFu;
    key_bc[i] = valid ? (uint32_t)b : (uint32_t)B; // invalid targets sort behind every barcode
}

__global__ void __launch_bounds__(256)
k_idx_gather(const uint32_t *__restrict__ src, const uint32_t *__restrict__ perm, const int n, uint32_t *__restrict__ dst)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[perm[j]];
}

// One CTA of 1024 threads.  kb = barcodes of the targets in (barcode, start) order.
__global__ void __launch_bounds__(1024)
k_idx_bounds(const uint32_t *__restrict__ kb, const int n, const int B, const u64 G, const u64 buckets_per_interval,
             SegDesc *__restrict__ seg, int32_t *__restrict__ extra_off, IdxResult *__restrict__ res)
{
    __shared__ u64 s_warp[40];
    __shared__ u64 s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    int nonempty = 0;
    for (int b0 = 0; b0 < B; b0 += 1024) { // barcodes in chunks of 1024, grid offsets carried in s_run
        const int b = b0 + threadIdx.x;
        int lo = 0, hi = 0, shift = 63;
        u64 nb = 0;
        if (b < B) {
            int a = 0, z = n;
            while (a < z) { const int m = (a + z) >> 1; if (kb[m] < (uint32_t)b) a = m + 1; else z = m; }
            lo = a; z = n;
            while (a < z) { const int m = (a + z) >> 1; if (kb[m] <= (uint32_t)b) a = m + 1; else z = m; }
            hi = a;
            if (hi > lo) { shift = idx_grid_shift((u64)(hi - lo), G, buckets_per_interval); nb = ((G + 1) >> shift) + 2; ++nonempty; }
            else nb = 2; // empty run: two entries that reject every read, any position maps to bucket 0 (shift 63)
        }
        u64 tot;
        const u64 ex = block_excl_scan(nb, s_warp, &tot);
        if (b < B) {
            seg[b] = SegDesc{hi > lo ? lo : 0, hi > lo ? hi : 0, (int)(s_run + ex), shift};
            extra_off[b] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    // totals: n_valid = targets with a barcode below B; the number of non-empty barcodes via a block sum
    u64 tot;
    block_excl_scan((u64)nonempty, s_warp, &tot);
    if (threadIdx.x == 0) {
        int a = 0, z = n;
        while (a < z) { const int m = (a + z) >> 1; if (kb[m] < (uint32_t)B) a = m + 1; else z = m; }
        res->n_valid = a;
        res->n_grid = (int32_t)s_run;
        res->n_nonempty = (int32_t)tot;
    }
}

__global__ void __launch_bounds__(256)
k_idx_records(const i64 *__restrict__ ts, const i64 *__restrict__ te, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ kb,
              const SegDesc *__restrict__ seg, const int n, const int B, const u64 G, const double hot_len,
              TargetRec<uint32_t> *__restrict__ trec, uint32_t *__restrict__ tkey, IdxResult *__restrict__ res)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t b = kb[j];
    if (b >= (uint32_t)B) return; // invalid targets are not indexed
    const uint32_t row = perm[j];
    const uint32_t s = idx_clamp(ts[row], G), e = idx_clamp(te[row], G);
    const int hi = seg[b].hi, lo = seg[b].lo;
    TargetRec<uint32_t> r;
    r.s = s; r.e = e; r.row = (int)row;
    r.s_next = j + 1 < hi ? idx_clamp(ts[perm[j + 1]], G) : 0xFFFFFFFFu;
    trec[j] = r;
    tkey[j] = e; // running max of the ends = the end itself as long as the ends ascend (checked right here)
    uint32_t f = 0;
    if (j > lo && e < idx_clamp(te[perm[j - 1]], G)) f |= kIdxNested;
    if ((double)(e - s) >= hot_len) f |= kIdxHot;
    if (f) atomicOr(&res->flags, f);
}

// One thread per grid bucket over all barcodes (n_grid_cap >= the number of entries k_idx_bounds laid out).
__global__ void __launch_bounds__(256)
k_idx_grid(SegDesc *__restrict__ seg, const int B, const TargetRec<uint32_t> *__restrict__ trec, const uint32_t *__restrict__ tkey,
           const IdxResult *__restrict__ res, GridEntry<uint32_t> *__restrict__ grid)
{
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= res->n_grid) return;
    int a = 0, z = B; // the barcode whose grid holds entry g: last b with grid_off <= g
    while (z - a > 1) { const int m = (a + z) >> 1; if ((i64)seg[m].grid_off <= g) a = m; else z = m; }
    const SegDesc d = seg[a];
    const int k = (int)(g - d.grid_off), shift = d.flags & 0xff;
    auto first_above = [&](u64 edge) { // first record of the run whose end exceeds `edge`
        int lo = d.lo, hi = d.hi;
        while (lo < hi) { const int m = (lo + hi) >> 1; if ((u64)tkey[m] <= edge) lo = m + 1; else hi = m; }
        return lo;
    };
    const u64 edge = shift >= 63 ? 0 : (u64)k << shift;
    const int j = first_above(edge);
    GridEntry<uint32_t> e;
    e.j = j;
    e.s = j < d.hi ? trec[j].s : 0xFFFFFFFFu;
    grid[g] = e;
    if (d.hi > d.lo && shift < 63) {
        const int j2 = first_above((u64)(k + 1) << shift);
        if (j2 - j > kCrowdedBucket) atomicOr(&seg[a].flags, 0x100);
    }
}

} // namespace esl
