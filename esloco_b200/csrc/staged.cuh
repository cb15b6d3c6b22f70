// The STAGED pipeline the north star sketches, kept as a measured baseline next to the fused kernels:
//   S1 generate + place + block test, write (start, len, label) per draw          12 B/read
//   S2 block-scan prefix sum over the oversampled batch -> cut-off index          (K5)
//   S3 LSD radix sort of the read starts (8-bit digits, 4 passes), payload = draw index   (K6, CUB-style)
//   S4 per-target binary-search sweep over the sorted starts, one warp per target,
//      warp-shuffle reduction, one result write per target                         (K6 + K7)
// It computes exactly what the fused path computes from the same Philox draws (tests compare them bit for bit),
// but every read travels through HBM several times (96 B/read) and the sweep window must reach back by
// the longest read, so it cannot beat 96 B x reads / HBM bandwidth.  bench.py --path staged times it.
// u32 coordinates, native stream, scaling = 1 only: it is a baseline, not a product path.
#pragma once
#include "kernels.cuh"

namespace esl {

constexpr int kStageBlock = 256;
constexpr int kSortItems = 16;                        // keys per thread and radix pass
constexpr int kSortTile = kStageBlock * kSortItems;   // 4096 keys per CTA

// S1: one draw per thread; per-CTA counted-length sums for the cut-off scan.
__global__ void __launch_bounds__(kStageBlock)
k_stage_generate(const __grid_constant__ DevCtx<uint32_t> cx, const __grid_constant__ PhiloxSrc src, const int n_cap,
                 uint32_t *__restrict__ start, uint32_t *__restrict__ len, int *__restrict__ label,
                 u64 *__restrict__ block_sum)
{
    extern __shared__ u64 smem[];
    uint32_t *s_cdf = (uint32_t *)smem;
    uint2 *s_lut = (uint2 *)(((uintptr_t)(s_cdf + cx.n_barcodes) + 7) & ~(uintptr_t)7);
    __shared__ u64 s_part[kStageBlock / 32];
    if (cx.n_barcodes > 1) {
        for (int i = threadIdx.x; i < cx.n_barcodes; i += kStageBlock) s_cdf[i] = cx.cdf_u32[i];
        for (int i = threadIdx.x; i < kCdfLut; i += kStageBlock) s_lut[i] = cx.cdf_lut[i];
    }
    __syncthreads();
    const int n = blockIdx.x * kStageBlock + threadIdx.x;
    u64 counted = 0;
    if (n < n_cap) {
        uint32_t L, c; uint32_t st; int bc, lab;
        src.draw(cx, DrawTabs{s_cdf, s_lut, cx.qtab}, 0, (i64)n, L, st, bc);
        resolve_read(cx, src, 0, (i64)n, L, st, bc, lab, c);
        start[n] = st; len[n] = L; label[n] = lab;
        counted = c;
    }
    counted = warp_sum(counted);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = counted;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 s = 0;
        for (int w = 0; w < kStageBlock / 32; ++w) s += s_part[w];
        block_sum[blockIdx.x] = s;
    }
}

// S2 (K5): single CTA.  Scan the per-block sums to the block holding the cut-off, then block-scan that block's
// counted lengths for the exact draw.  out: n_reads, n_bases, status of ONE iteration (and n_reads again as an
// int the later stages read on the device).
__global__ void __launch_bounds__(1024)
k_stage_cutoff(const DevCtx<uint32_t> cx, const int n_cap, const uint32_t *__restrict__ len, const int *__restrict__ label,
               const u64 *__restrict__ block_sum, u64 *__restrict__ n_reads_out, u64 *__restrict__ n_bases_out,
               int *__restrict__ status_out, int *__restrict__ n_reads_dev)
{
    __shared__ u64 s_warp[40];
    __shared__ i64 s_blk;
    __shared__ u64 s_pref;
    __shared__ int s_cut;
    __shared__ u64 s_cov;
    const int tid = threadIdx.x;
    const int n_blocks = (n_cap + kStageBlock - 1) / kStageBlock;
    if (tid == 0) { s_blk = -1; s_pref = 0; s_cut = -1; s_cov = 0; }
    __syncthreads();
    const i64 thr = cx.thr;
    u64 run = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) { // first block whose inclusive prefix reaches thr
        const int b = b0 + tid;
        const u64 v = b < n_blocks ? block_sum[b] : 0;
        u64 tot;
        const u64 ex = block_excl_scan(v, s_warp, &tot);
        if (b < n_blocks && (i64)(run + ex) < thr && (i64)(run + ex + v) >= thr) { s_blk = b; s_pref = run + ex; }
        __syncthreads();
        if (s_blk >= 0) break;
        run += tot;
    }
    int n_reads;
    u64 covered;
    if (thr <= 0) { n_reads = 0; covered = 0; }
    else if (s_blk < 0) { n_reads = n_cap; covered = run; if (tid == 0) atomicOr(status_out, 2); } // not reached
    else {
        const int n = (int)s_blk * kStageBlock + tid; // kStageBlock draws, first 256 threads
        const bool in = tid < kStageBlock && n < n_cap;
        u64 c = 0;
        if (in) { const int lab = label[n]; c = lab == kLabelBlockedNotConsuming ? 0 : len[n]; }
        u64 tot;
        const u64 ex = block_excl_scan(c, s_warp, &tot);
        // the cut-off draw is the first whose inclusive prefix reaches thr; it is still placed
        if (in && (i64)(s_pref + ex) < thr && (i64)(s_pref + ex + c) >= thr) { s_cut = n; s_cov = s_pref + ex + c; }
        __syncthreads();
        n_reads = s_cut + 1;
        covered = s_cov;
    }
    if (tid == 0) { *n_reads_out = (u64)n_reads; *n_bases_out = covered; *n_reads_dev = n_reads; }
}

// label histogram of the placed reads (counting.py:63-71): shared 32-bit bins per CTA, one flush per CTA
__global__ void __launch_bounds__(kStageBlock)
k_stage_labels(const int n_barcodes, const int *__restrict__ label, const int *__restrict__ n_dev, u64 *__restrict__ label_counts)
{
    extern __shared__ u64 smem[];
    int *s_hist = (int *)smem;
    const int B = n_barcodes, n = *n_dev;
    for (int i = threadIdx.x; i < B + 2; i += kStageBlock) s_hist[i] = 0;
    __syncthreads();
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    for (int j0 = lo; j0 < hi; j0 += kStageBlock) { // whole warps enter hist_add together
        const int j = j0 + threadIdx.x;
        hist_add(s_hist, j < hi ? label_slot(label[j], B) : -1, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B + 2; i += kStageBlock)
        if (s_hist[i]) atomicAdd(label_counts + i, (u64)(uint32_t)s_hist[i]);
}

// ---- S3: LSD radix sort, 8 bits per pass -----------------------------------------------------------------------

// digit histogram of each CTA's tile: hist[digit * n_blocks + block]
__global__ void __launch_bounds__(kStageBlock)
k_rs_hist(const uint32_t *__restrict__ keys, const int *__restrict__ n_dev, const int shift, const int n_blocks,
          uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_h[256];
    const int n = *n_dev;
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int j = base + i * kStageBlock + threadIdx.x;
        if (j < n) atomicAdd(&s_h[(keys[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * n_blocks + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan of the (digit-major) histogram: CTA d scans the row of digit d in place (offsets of the CTAs'
// tiles inside the digit) and leaves the digit's total; a second tiny kernel turns the 256 totals into digit bases
__global__ void __launch_bounds__(1024)
k_rs_scan_rows(uint32_t *__restrict__ hist, const int n_blocks, uint32_t *__restrict__ digit_tot)
{
    __shared__ u64 s_warp[40];
    uint32_t *row = hist + (size_t)blockIdx.x * n_blocks;
    u64 run = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const u64 v = i < n_blocks ? row[i] : 0;
        u64 tot;
        const u64 ex = block_excl_scan(v, s_warp, &tot);
        if (i < n_blocks) row[i] = (uint32_t)(run + ex);
        run += tot;
    }
    if (threadIdx.x == 0) digit_tot[blockIdx.x] = (uint32_t)run;
}
__global__ void __launch_bounds__(1024)
k_rs_scan_digits(uint32_t *__restrict__ digit_tot)
{
    __shared__ u64 s_warp[40];
    const u64 v = threadIdx.x < 256 ? digit_tot[threadIdx.x] : 0;
    u64 tot;
    const u64 ex = block_excl_scan(v, s_warp, &tot);
    if (threadIdx.x < 256) digit_tot[threadIdx.x] = (uint32_t)ex;
}

// stable scatter: rank of a key = global base of (digit, CTA) + keys of the same digit earlier in the CTA's tile
__global__ void __launch_bounds__(kStageBlock)
k_rs_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ keys_out,
             uint32_t *__restrict__ vals_out, const int *__restrict__ n_dev, const int shift, const int n_blocks,
             const uint32_t *__restrict__ hist, const uint32_t *__restrict__ digit_base, const int first_pass)
{
    __shared__ uint32_t s_base[256];              // running output position of each digit for this CTA
    __shared__ uint32_t s_cnt[kStageBlock / 32][256]; // per-warp digit counts of the current slice
    const int n = *n_dev;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    s_base[tid] = digit_base[tid] + hist[tid * n_blocks + blockIdx.x];
    const int base = blockIdx.x * kSortTile;
    for (int i = 0; i < kSortItems; ++i) { // slices of 256 consecutive keys keep the pass stable
        for (int k = tid; k < (kStageBlock / 32) * 256; k += kStageBlock) (&s_cnt[0][0])[k] = 0;
        __syncthreads();
        const int j = base + i * kStageBlock + tid;
        const bool in = j < n;
        const uint32_t key = in ? keys_in[j] : 0xFFFFFFFFu;
        const uint32_t val = in ? (first_pass ? (uint32_t)j : vals_in[j]) : 0u;
        const uint32_t d = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, in ? d : 256u + lane); // out-of-range lanes stay alone
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1));
        if (in && rank_in_warp == 0) s_cnt[wid][d] = __popc(peers);
        __syncthreads();
        uint32_t before = 0; // same digit in earlier warps of this slice
        if (in)
            for (int w = 0; w < wid; ++w) before += s_cnt[w][d];
        if (in) {
            const uint32_t pos = s_base[d] + before + rank_in_warp;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        { // advance the running bases by the slice's digit totals (thread tid owns digit tid)
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < kStageBlock / 32; ++w) tot += s_cnt[w][tid];
            s_base[tid] += tot;
        }
        __syncthreads();
    }
}

// ---- S4 (K6 + K7): one warp per target sweeps the reads whose start lies in (ts - Lmax, te) -------------------
__global__ void __launch_bounds__(kStageBlock)
k_stage_sweep(const i64 *__restrict__ t_start, const i64 *__restrict__ t_end, const int *__restrict__ t_bc, const i64 n_targets,
              const uint32_t *__restrict__ sorted_start, const uint32_t *__restrict__ sorted_idx, const int *__restrict__ n_dev,
              const uint32_t *__restrict__ len, const int *__restrict__ label, const uint32_t max_len, const i64 min_overlap,
              const int n_barcodes, u64 *__restrict__ tallies)
{
    const i64 t = ((i64)blockIdx.x * kStageBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_targets) return;
    const int n = *n_dev;
    const i64 s = t_start[t], e = t_end[t];
    const int bc = t_bc[t];
    u64 full = 0, part = 0, otb = 0;
    if (bc >= 0 && bc < n_barcodes && e > s) {
        // window of candidate reads: start < e and start > s - max_len (a read ends at most max_len after it starts)
        const i64 lo_key = s - (i64)max_len;
        int a = 0, z = n;
        if (lo_key >= 0) {
            while (a < z) { const int mid = (a + z) >> 1; if ((i64)sorted_start[mid] <= lo_key) a = mid + 1; else z = mid; }
        }
        const int w0 = a;
        a = w0; z = n;
        while (a < z) { const int mid = (a + z) >> 1; if ((i64)sorted_start[mid] < e) a = mid + 1; else z = mid; }
        const int w1 = a;
        for (int j = w0 + lane; j < w1; j += 32) {
            const uint32_t idx = sorted_idx[j];
            if (label[idx] != bc) continue;
            const i64 rs = sorted_start[j], re = rs + len[idx];
            const i64 lo = s > rs ? s : rs, hi = e < re ? e : re;
            if (lo < hi && hi - lo >= min_overlap) {
                if (rs <= s && e <= re) ++full; else ++part;
                otb += (u64)(hi - lo);
            }
        }
    }
    full = warp_sum(full); part = warp_sum(part); otb = warp_sum(otb);
    if (lane == 0) { u64 *o = tallies + t * 3; o[0] = full; o[1] = part; o[2] = otb; }
}

} // namespace esl
