// C ABI of the engine (include/esloco_b200.h): context, host-side index construction, kernel launches.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/esloco_b200.h"
#include "kernels.cuh"
#include "staged.cuh"
#include "index_build.cuh"

using esl::i64;
using esl::u64;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    // Set-up-time upload, ordered on `st`.  The source is pageable host memory, so the call returns once the data sit
    // in the driver's staging buffer: the caller may free `src` right away.  Growing the buffer frees the old one, which
    // waits for the device; set-up calls are allowed to block (include/esloco_b200.h).
    cudaError_t upload(const void *src, size_t n, cudaStream_t st)
    {
        if (n > bytes) {
            if (p) cudaFree(p);
            p = nullptr; bytes = 0;
            cudaError_t e = cudaMalloc(&p, n ? n : 1);
            if (e != cudaSuccess) return e;
            bytes = n;
        }
        return n ? cudaMemcpyAsync(p, src, n, cudaMemcpyHostToDevice, st) : cudaSuccess;
    }
    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <typename T> T *as() const { return (T *)p; }
};

// Host array in page-locked memory (the targets: set-up uploads them straight from here).  Just enough of std::vector.
template <typename T>
struct PinnedVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    ~PinnedVec() { if (p) cudaFreeHost(p); }
    bool assign(const T *first, const T *last)
    {
        const size_t m = (size_t)(last - first);
        if (m > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            if (cudaMallocHost((void **)&p, (m + m / 4 + 64) * sizeof(T)) != cudaSuccess) { p = nullptr; n = 0; return false; }
            cap = m + m / 4 + 64;
        }
        if (m) memcpy(p, first, m * sizeof(T));
        n = m;
        return true;
    }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    const T *data() const { return p; }
    const T &operator[](size_t i) const { return p[i]; }
};

struct SortScratch { DevBuf keys[2], vals[2], hist; };

} // namespace

struct esl_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    i64 launches = 0;

    u64 G = 0;
    int B = 0;
    std::vector<double> cdf;
    PinnedVec<i64> t_start, t_end;
    PinnedVec<int32_t> t_bc;
    int index_builder = ESL_INDEX_AUTO; // esl_set_index_builder
    bool index_on_device = false;       // how the current target index was built
    std::vector<i64> m_start, m_end;
    std::vector<double> m_w;
    std::vector<int32_t> m_bc;
    bool have_targets = false, have_pool = false;
    int n_seg = 0, n_tgrid = 0, n_trec = 0, n_mgrid = 0, n_mrec = 0;
    bool index_dirty = true;
    bool raw_dirty = true;
    bool wide = false; // 64-bit coordinates
    int n_layers = 0, n_first_layers = 0, cdf_walk = 0, n_hot = 0;
    i64 n_indexed = 0, n_mask = 0;

    // pool statistics
    uint32_t n_pool = 0, pool_max = 0;
    double pool_mean = 0, pool_sd = 0;

    DevBuf d_pool, d_cdf, d_cdf_u32, d_cdf_lut, d_trec, d_te, d_seg, d_tgrid, d_extra_off, d_mrec, d_mpmax, d_grp, d_mgrid, d_hot_row;
    DevBuf d_raw_ts, d_raw_te, d_raw_tb; // targets in caller order (staged baseline)
    DevBuf d_pool_stats;
    DevBuf d_sgrid;
    struct esl_host_scratch *host = nullptr; // persistent host-side scratch of esl_prepare (defined below)
    int sgrid_n = 0, sgrid_shift = 0; // shared-memory copy of the first-layer grids (0 = does not fit)
    // quantile length sampler (esl_set_length_sampler): sorted pool + quantile table, built by esl_prepare
    int sampler = ESL_SAMPLER_POOL;
    bool qtab_dirty = true;
    DevBuf d_pool_sorted, d_qtab;
    SortScratch sort;
    DevBuf d_idx_key_s, d_idx_key_b, d_idx_res; // device index builder
    DevBuf d_mbits;
    int mbits_shift = 0, mbits_words = 0; // blocked-bucket map, two bits per bucket (kernels.cuh DevCtx::mbits)
    u64 sgrid_lmax = 0;               // longest read its upper bounds were built for

    // optional hit-record buffer (esl_set_hit_buffer)
    int64_t *d_hits = nullptr;
    int64_t *d_hit_count = nullptr;
    int64_t hit_cap = 0;

    // optional per-kernel timing (esl_set_profiling): events around the bulk and the cut-off kernel
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;

    int fail(int code, const char *fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return ctx->fail(ESL_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

namespace {

// ---- host-side index construction ------------------------------------------------------------------

struct Rec { i64 s, e; int32_t row; };

// Split one barcode's targets (sorted by start, end) into layers the device can scan front to back.  A read
// [rs, re) walks a layer from the first record whose running max of ends exceeds rs until a start reaches re, so
// every record in between that ended before rs is a wasted ("dead") load.  A layer whose ends never decrease has
// none; strict patience sorting gives that, but random ROIs of mixed lengths nest all the time (50 k ROIs of
// 1-100 kb: 9 strict layers, each costing every read a probe).  So nesting is tolerated while it stays cheap:
// a target that ends before the layer's running max is nested in an earlier record; it is accepted (first-fit over
// the layers) if (a) it lies at most kDeadScan records behind the OLDEST record that outlives it -- a walk starts at
// such a record, so no read ever passes more than kDeadScan dead records -- and (b) the layer's summed dead spans
// (running max - end) stay below a quarter of the genome (< 0.25 dead loads per read on average).  Targets whose
// ends keep ascending are always accepted, however densely they overlap.  Whole-chromosome ROIs over many small
// ones still split into two layers; 50 k mixed ROIs become one.  Layers past the 64th are strictly monotone
// (binary search), so pathological nesting stays O(n log n).
constexpr int kDeadScan = 16;
// Small target sets whose targets meet more than this share of the reads skip the shared-memory pre-filter (24
// whole-chromosome ROIs + 76 small ones: 63.7 -> 71.1 G reads/s without it).
constexpr double kPrefilterMaxCover = 0.5;
#ifndef ESL_MBITS_MAX_BUCKETS
#define ESL_MBITS_MAX_BUCKETS 65536
#endif
constexpr u64 MBITS_MAX_BUCKETS = ESL_MBITS_MAX_BUCKETS; // blocked-bucket map: two bits per bucket in shared memory
constexpr size_t kRelaxedLayers = 64;
// Sort targets by (start, end, row).  Large sets: stable LSD radix sort on the start (11-bit digits), then the rare runs
// of equal starts are ordered by (end, row) -- 10 k targets take ~0.1 ms instead of 0.7 ms with a comparison sort, and
// the index of a 240 k-target set is rebuilt on every esl_prepare.
void sort_recs(Rec *first_rec, Rec *last_rec)
{
    auto less = [](const Rec &a, const Rec &b) {
        if (a.s != b.s) return a.s < b.s;
        if (a.e != b.e) return a.e < b.e;
        return a.row < b.row;
    };
    const size_t n = (size_t)(last_rec - first_rec);
    if (n < 512) { std::sort(first_rec, last_rec, less); return; }
    i64 mx = 0;
    for (const Rec *r = first_rec; r != last_rec; ++r) mx = std::max(mx, r->s);
    std::vector<Rec> tmp(n);
    Rec *src = first_rec, *dst = tmp.data();
    for (int shift = 0; shift < 63 && (mx >> shift) != 0; shift += 11) {
        size_t cnt[2049] = {0};
        for (size_t i = 0; i < n; ++i) ++cnt[((u64)src[i].s >> shift & 2047u) + 1];
        for (int d = 0; d < 2048; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; ++i) dst[cnt[(u64)src[i].s >> shift & 2047u]++] = src[i];
        std::swap(src, dst);
    }
    if (src != first_rec) std::copy(src, src + n, first_rec);
    for (size_t i = 0; i < n;) { // equal starts: order by (end, row)
        size_t j = i + 1;
        while (j < n && first_rec[j].s == first_rec[i].s) ++j;
        if (j - i > 1) std::sort(first_rec + i, first_rec + j, less);
        i = j;
    }
}

void build_layers(Rec *first_rec, Rec *last_rec, std::vector<std::vector<Rec>> &layers, u64 G)
{
    sort_recs(first_rec, last_rec);
    struct Relaxed {
        std::vector<i64> pm_end;    // the records that raised the running max of ends: their ends (ascending) ...
        std::vector<size_t> pm_idx; // ... and their positions in the layer
        long double dead = 0;
    };
    std::vector<Relaxed> relaxed;
    std::vector<i64> last; // last end of each strict layer (index kRelaxedLayers + k); decreasing under first-fit
    const long double budget = 0.25L * (long double)G;
    for (const Rec *it = first_rec; it != last_rec; ++it) {
        const Rec &r = *it;
        bool placed = false;
        for (size_t l = 0; l < relaxed.size() && !placed; ++l) {
            Relaxed &o = relaxed[l];
            std::vector<Rec> &lay = layers[l];
            if (r.e >= o.pm_end.back()) { // not nested: no read will ever see it dead
                if (r.e > o.pm_end.back()) { o.pm_end.push_back(r.e); o.pm_idx.push_back(lay.size()); }
                lay.push_back(r);
                placed = true;
                break;
            }
            // oldest record that ends at or after this one = first running-max record with end >= r.e
            const size_t k = (size_t)(std::lower_bound(o.pm_end.begin(), o.pm_end.end(), r.e) - o.pm_end.begin());
            const long double dead = (long double)(o.pm_end.back() - r.e);
            if (lay.size() - o.pm_idx[k] <= (size_t)kDeadScan && o.dead + dead <= budget) {
                lay.push_back(r);
                o.dead += dead;
                placed = true;
            }
        }
        if (placed) continue;
        if (relaxed.size() < kRelaxedLayers) {
            relaxed.emplace_back();
            relaxed.back().pm_end.push_back(r.e);
            relaxed.back().pm_idx.push_back(0);
            layers.emplace_back(1, r);
            continue;
        }
        size_t lo = 0, hi = last.size();
        while (lo < hi) { // first strict layer with last <= r.e
            size_t mid = (lo + hi) / 2;
            if (last[mid] <= r.e) hi = mid; else lo = mid + 1;
        }
        if (lo == last.size()) { last.push_back(r.e); layers.emplace_back(); }
        else last[lo] = r.e;
        layers[kRelaxedLayers + lo].push_back(r);
    }
}

template <typename C> C clampc(i64 v, u64 G) { return (C)((u64)v > G + 1 ? G + 1 : (u64)v); }

// Bucket grid over one sorted run: keys[lo..hi) ascending (the running max of the ends of a target layer or of a
// group of blocked intervals), starts[lo..hi) the interval starts.  At least 32 buckets per interval, so a bucket
// rarely holds more than one interval end; runs where some bucket holds more than kCrowdedBucket ends are flagged
// so the device bisects inside the bucket.
// Bucket width ~ min(G / (32 * intervals), 64 kb): fine enough that "first candidate starts before the read ends" is
// rarely a false alarm even for sparse runs, yet at most 2^21 entries (16 MB) per run.  (Config 3, 10 k targets per
// barcode: 8 buckets per target queue 17 % of the reads for 6.5 % true overlaps, 32 buckets 9 % -- 128 -> 136 G reads/s
// for 36 MB of L2-resident grids instead of 9 MB.)
constexpr u64 kBucketsPerInterval = 32;
inline int grid_shift(u64 n_intervals, u64 G)
{
    const u64 want = std::min<u64>(std::max<u64>(std::max<u64>(64, kBucketsPerInterval * n_intervals), (G >> 16) + 1), 1ull << 21);
    int shift = 0;
    while (((G + 1) >> shift) + 2 > want) ++shift;
    return shift;
}
// reads start below G: bucket index <= G >> shift; one more entry for the upper bound of the last bucket
inline u64 grid_buckets(int shift, u64 G) { return ((G + 1) >> shift) + 2; }

// Writes the nb entries of one run's grid to out[]; returns 1 when some bucket is crowded.
template <typename C>
int fill_grid(const C *keys, const C *starts, int lo, int hi, int shift, u64 nb, esl::GridEntry<C> *out)
{
    int j = lo, crowded = 0, prev = lo;
    for (u64 b = 0; b < nb; ++b) {
        const u64 edge = b << shift;
        while (j < hi && (u64)keys[j] <= edge) ++j; // first key > bucket start
        if (j - prev > esl::kCrowdedBucket) crowded = 1;
        prev = j;
        esl::GridEntry<C> e{};
        e.j = j;
        e.s = j < hi ? starts[j] : (C) ~(C)0;
        out[b] = e;
    }
    return crowded;
}

// Vector form: appends the run's grid to `grid`.
template <typename C>
esl::SegDesc make_grid(const C *keys, const C *starts, int lo, int hi, u64 G, std::vector<esl::GridEntry<C>> &grid,
                       int forced_shift = -1)
{
    esl::SegDesc d;
    d.lo = lo; d.hi = hi; d.grid_off = (int)grid.size();
    const int shift = forced_shift >= 0 ? forced_shift : grid_shift((u64)(hi - lo), G);
    const u64 nb = grid_buckets(shift, G);
    grid.resize(grid.size() + nb);
    const int crowded = fill_grid<C>(keys, starts, lo, hi, shift, nb, grid.data() + d.grid_off);
    d.flags = shift | (crowded << 8);
    return d;
}

// Grid of an empty run: two entries that reject every read; any position maps to bucket 0 (shift 63).
template <typename C>
esl::SegDesc empty_run(std::vector<esl::GridEntry<C>> &grid)
{
    esl::SegDesc d{0, 0, (int)grid.size(), 63};
    esl::GridEntry<C> e{};
    e.j = 0; e.s = (C) ~(C)0;
    grid.push_back(e); grid.push_back(e);
    return d;
}

// ESL_TIMING=1 prints where esl_prepare's host time goes (stderr); off by default.
struct PhaseTimer {
    bool on = getenv("ESL_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char *what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[esl_prepare] %-40s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// Page-locked host staging buffer kept between esl_prepare calls: the index arrays are written straight into it by
// the builder threads and copied from there (fresh pageable vectors cost more in first-touch page faults and staged
// copies than the index build itself: 30 MB per call at 240 k targets).
struct PinnedBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        const size_t want = n + n / 4 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    template <typename T> T *as() const { return (T *)p; }
};

} // namespace

struct esl_host_scratch {
    std::vector<Rec> recs;               // targets bucketed by barcode
    PinnedBuf trec, tp, tgrid, ts, te;   // index arrays in their final order (ts / te: host-only helpers)
    cudaEvent_t uploaded = nullptr;      // the last esl_prepare's copies out of the pinned buffers
    ~esl_host_scratch() { if (uploaded) cudaEventDestroy(uploaded); }
};

namespace {

template <typename F>
void parallel_for(int n, int n_thr, F f)
{
    std::atomic<int> next{0};
    auto work = [&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
}

struct LayerPlan { int barcode, lo, hi, grid_off, shift; u64 nb; const std::vector<Rec> *recs; };

// The target index built on the host: any geometry (nesting layers, hot targets, the shared pre-filter of small sets).
template <typename C>
int build_targets_host(esl_ctx *ctx, cudaStream_t st, PhaseTimer &tm)
{
    const int B = ctx->B;
    const u64 G = ctx->G;
    esl_host_scratch &hs = *ctx->host;
    if (!hs.uploaded) CU(cudaEventCreateWithFlags(&hs.uploaded, cudaEventDisableTiming));
    CU(cudaEventSynchronize(hs.uploaded)); // an earlier prepare may still be copying out of the pinned buffers
    // targets: seg[b] = first layer of barcode b, extra layers appended behind the B first-layer descriptors.
    // (1) one counting pass buckets the targets by barcode; (2) the barcodes are sorted and split into layers
    // independently on a few host threads (config 3: 24 barcodes x 10 k targets); (3) the layout of the final arrays
    // follows from the layer sizes; (4) the threads write records and grids straight to their final places.
    const size_t n_all = ctx->t_start.size();
    // host threads of the build: at most 8, and a fair share of the cores when several ranks of one launch (torchrun's
    // LOCAL_WORLD_SIZE) rebuild their indexes at the same time
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char *lws = getenv("LOCAL_WORLD_SIZE")) hw = std::max(1u, hw / (unsigned)std::max(1, atoi(lws)));
    const int n_thr = (int)std::min<size_t>({(size_t)hw, (size_t)8, n_all / 4096 + 1});
    std::vector<size_t> off(B + 1, 0);
    {
        // slice s of the targets counts its valid targets per barcode, the counts become write positions (barcode-major,
        // slices in order, so the caller's order survives inside a barcode), then every slice scatters its own targets
        const int n_sl = n_thr;
        const size_t per = (n_all + n_sl - 1) / std::max(n_sl, 1);
        std::vector<size_t> cnt((size_t)n_sl * B, 0);
        // indexed: targets some read can overlap -- a known barcode, not empty, and starting inside the genome (a read ends
        // before G; insertion coordinates pushed past G by later insertions, create_insertion_genome.py:60-64, keep their
        // all-zero rows but would crowd the last grid bucket)
        auto valid = [&](size_t i) {
            const int32_t b = ctx->t_bc[i];
            return b >= 0 && b < B && ctx->t_end[i] > ctx->t_start[i] && (u64)ctx->t_start[i] < G;
        };
        parallel_for(n_sl, n_thr, [&](int sl) {
            size_t *c = cnt.data() + (size_t)sl * B;
            for (size_t i = sl * per, e = std::min(n_all, (sl + 1) * per); i < e; ++i)
                if (valid(i)) ++c[ctx->t_bc[i]];
        });
        size_t run = 0;
        for (int b = 0; b < B; ++b) {
            off[b] = run;
            for (int sl = 0; sl < n_sl; ++sl) { const size_t c = cnt[(size_t)sl * B + b]; cnt[(size_t)sl * B + b] = run; run += c; }
        }
        off[B] = run;
        hs.recs.resize(run);
        parallel_for(n_sl, n_thr, [&](int sl) {
            size_t *at = cnt.data() + (size_t)sl * B;
            for (size_t i = sl * per, e = std::min(n_all, (sl + 1) * per); i < e; ++i)
                if (valid(i)) hs.recs[at[ctx->t_bc[i]]++] = Rec{ctx->t_start[i], ctx->t_end[i], (int32_t)i};
        });
    }
    tm.mark("bucket by barcode");
    std::vector<std::vector<std::vector<Rec>>> layers(B);
    parallel_for(B, n_thr, [&](int b) { build_layers(hs.recs.data() + off[b], hs.recs.data() + off[b + 1], layers[b], G); });
    tm.mark("sort + layers per barcode (threads)");
    std::vector<LayerPlan> plan;
    std::vector<int32_t> extra_off(B, 0);
    std::vector<int> first_plan(B, -1);
    size_t n_rec = 0, n_grid = 0, n_extra = 0;
    ctx->n_layers = 0;
    ctx->n_first_layers = 0;
    for (int b = 0; b < B; ++b) {
        if (layers[b].size() > 0xFFFF) return ctx->fail(ESL_ERR_UNSUPPORTED, "barcode %d needs %zu nesting layers (max 65535)", b, layers[b].size());
        extra_off[b] = (int32_t)n_extra;
        for (size_t l = 0; l < layers[b].size(); ++l) {
            const size_t n = layers[b][l].size();
            const int shift = grid_shift(n, G);
            const u64 nb = grid_buckets(shift, G);
            if (l == 0) first_plan[b] = (int)plan.size(); else ++n_extra;
            plan.push_back(LayerPlan{b, (int)n_rec, (int)(n_rec + n), (int)n_grid, shift, nb, &layers[b][l]});
            n_rec += n; n_grid += nb;
        }
        ctx->n_layers += (int)layers[b].size();
        ctx->n_first_layers += layers[b].empty() ? 0 : 1;
    }
    // barcodes without targets: two entries each that reject every read, placed BEHIND all real grids (below) -- so they
    // are added only now, after every real grid has its offset
    n_grid += 2 * (size_t)(B - ctx->n_first_layers);
    if (n_rec > 0x7fffffffull || n_grid > 0x7fffffffull) return ctx->fail(ESL_ERR_UNSUPPORTED, "target index too large");
    if ((size_t)B + n_extra >= ((size_t)1 << esl::kQueueSegBits)) return ctx->fail(ESL_ERR_UNSUPPORTED, "too many target layers (barcodes + nesting levels must stay below 2^22)");
    CU(hs.trec.reserve(std::max<size_t>(n_rec, 1) * sizeof(esl::TargetRec<C>)));
    CU(hs.tp.reserve(std::max<size_t>(n_rec, 1) * sizeof(C)));
    CU(hs.ts.reserve(std::max<size_t>(n_rec, 1) * sizeof(C)));
    CU(hs.te.reserve(std::max<size_t>(n_rec, 1) * sizeof(C)));
    CU(hs.tgrid.reserve(std::max<size_t>(n_grid, 1) * sizeof(esl::GridEntry<C>)));
    esl::TargetRec<C> *trec = hs.trec.as<esl::TargetRec<C>>();
    C *tp = hs.tp.as<C>(), *ts = hs.ts.as<C>(), *te = hs.te.as<C>(); // tp: running max of te inside each layer = grid / bisection keys
    esl::GridEntry<C> *tgrid = hs.tgrid.as<esl::GridEntry<C>>();
    std::vector<int> crowded(plan.size(), 0);
    parallel_for((int)plan.size(), n_thr, [&](int i) {
        const LayerPlan &pl = plan[i];
        C run = 0;
        int k = pl.lo;
        for (const Rec &r : *pl.recs) {
            ts[k] = clampc<C>(r.s, G); te[k] = clampc<C>(r.e, G);
            run = std::max(run, te[k]);
            tp[k] = run;
            ++k;
        }
        k = pl.lo;
        for (const Rec &r : *pl.recs) {
            esl::TargetRec<C> t{};
            t.s = ts[k]; t.e = te[k]; t.row = r.row;
            t.s_next = k + 1 < pl.hi ? ts[k + 1] : (C) ~(C)0;
            trec[k++] = t;
        }
        crowded[i] = fill_grid<C>(tp, ts, pl.lo, pl.hi, pl.shift, pl.nb, tgrid + pl.grid_off);
    });
    std::vector<esl::SegDesc> first(B), extra;
    {
        size_t empty_at = n_grid; // the empty runs' entries sit behind all real grids
        for (int b = 0; b < B; ++b)
            if (layers[b].empty()) {
                empty_at -= 2;
                esl::GridEntry<C> e{};
                e.j = 0; e.s = (C) ~(C)0;
                tgrid[empty_at] = e; tgrid[empty_at + 1] = e;
                first[b] = esl::SegDesc{0, 0, (int)empty_at, 63}; // any position maps to bucket 0 (shift 63)
            }
        for (size_t i = 0; i < plan.size(); ++i) {
            const LayerPlan &pl = plan[i];
            esl::SegDesc d{pl.lo, pl.hi, pl.grid_off, pl.shift | (crowded[i] << 8)};
            if ((int)i == first_plan[pl.barcode]) { d.flags |= (int)((layers[pl.barcode].size() - 1) << 16); first[pl.barcode] = d; }
            else extra.push_back(d);
        }
    }
    tm.mark("records + grids in place (threads)");
    // Shared-memory pre-filter for small target sets (typical esloco runs: a handful of ROIs / insertions per
    // barcode): per first layer a coarser grid, all barcodes at one common bucket width, holding per bucket the
    // pair {start of the bucket's first candidate, largest end among the targets a read STARTING in the bucket can
    // reach}.  A read [rs, re) can only overlap something if first < re and rs < that end; the second test needs
    // the longest read, so it only serves draws from the length pool (rounded up to a power of two: a longer pool
    // makes the index stale, index_stale).  Used if it fits ~36 KB with at least four buckets per target.
    ctx->sgrid_n = 0;
    ctx->sgrid_lmax = 1;
    while (ctx->sgrid_lmax < (u64)ctx->pool_max) ctx->sgrid_lmax <<= 1;
    {
        const size_t budget = 36 * 1024 / sizeof(C) / 2;
        const size_t per_bc = B > 0 ? budget / (size_t)B : 0;
        int max_t = 0;
        for (int b = 0; b < B; ++b) max_t = std::max(max_t, first[b].hi - first[b].lo);
        int sshift = 0;
        while (((G + 1) >> sshift) + 2 > per_bc && sshift < 63) ++sshift;
        const size_t n_sm = (size_t)((G + 1) >> sshift) + 2;
        // share of the reads that meet a target: when most do (whole-chromosome ROIs), the pre-filter answers "maybe" for
        // nearly every read and the candidates are better served by the general path's full-lane queue rounds
        double cover = 0;
        for (size_t k = 0; k < n_rec; ++k) cover += (double)(te[k] - ts[k]) + ctx->pool_mean;
        cover /= (double)G * (double)std::max(B, 1);
        const bool prefilter_pays = cover < kPrefilterMaxCover;
        if (per_bc >= 64 && n_sm <= per_bc && n_sm >= 4 * (size_t)max_t && ctx->n_first_layers > 0 && prefilter_pays) {
            std::vector<C> sg;
            sg.reserve(2 * n_sm * (size_t)B);
            for (int b = 0; b < B; ++b) {
                std::vector<esl::GridEntry<C>> tmp;
                make_grid<C>(tp, ts, first[b].lo, first[b].hi, G, tmp, sshift);
                for (size_t k = 0; k < n_sm; ++k) {
                    const esl::GridEntry<C> &e = tmp[std::min(k, tmp.size() - 1)];
                    // targets a read starting in bucket k can reach: from the first candidate on, starting before the
                    // end of the bucket plus the longest read
                    const long double reach = (long double)(((u64)k + 1) << sshift) + (long double)ctx->sgrid_lmax;
                    C hi_end = 0;
                    for (int j = e.j; j < first[b].hi && (long double)ts[j] < reach; ++j) hi_end = std::max(hi_end, te[j]);
                    sg.push_back(e.s);
                    sg.push_back(hi_end);
                }
            }
            CU(ctx->d_sgrid.upload(sg.data(), sg.size() * sizeof(C), st));
            ctx->sgrid_n = (int)n_sm;
            ctx->sgrid_shift = sshift;
        }
    }
    tm.mark("shared pre-filter grid");
    std::vector<esl::SegDesc> seg(first);
    seg.insert(seg.end(), extra.begin(), extra.end());
    ctx->n_indexed = (i64)n_rec;
    // hot targets: a read overlaps them with probability >= 1/512 (whole-chromosome ROIs and the like); the
    // kHotSlots hottest get a private shared-memory tally per CTA (their record carries the slot, hot_row[] the row)
    std::vector<int32_t> hot_row;
    {
        // p_hit = (target length + mean read length) / G >= 1/512, compared without a division per target
        std::vector<std::pair<double, int>> hot; // (-(length + mean read length), record index): hottest first, ties by position
        const double hot_len = (double)G / 512.0 - ctx->pool_mean;
        for (size_t k = 0; k < n_rec; ++k) {
            const double len = (double)(te[k] - ts[k]);
            if (len >= hot_len) hot.emplace_back(-(len + ctx->pool_mean), (int)k);
        }
        std::sort(hot.begin(), hot.end());
        if (hot.size() > (size_t)esl::kHotSlots) hot.resize(esl::kHotSlots);
        for (size_t h = 0; h < hot.size(); ++h) {
            hot_row.push_back(trec[hot[h].second].row);
            trec[hot[h].second].row = (int)(0x80000000u | (uint32_t)h);
        }
    }
    tm.mark("hot targets");
    auto copy_pinned = [&](DevBuf &d, const void *src, size_t n) -> cudaError_t {
        cudaError_t e = d.reserve(std::max<size_t>(n, 1));
        if (e != cudaSuccess || n == 0) return e;
        return cudaMemcpyAsync(d.p, src, n, cudaMemcpyHostToDevice, st);
    };
    CU(copy_pinned(ctx->d_trec, trec, n_rec * sizeof(esl::TargetRec<C>)));
    CU(copy_pinned(ctx->d_te, tp, n_rec * sizeof(C))); // bisection keys of crowded buckets
    CU(copy_pinned(ctx->d_tgrid, tgrid, n_grid * sizeof(esl::GridEntry<C>)));
    CU(cudaEventRecord(hs.uploaded, st));
    ctx->n_hot = (int)hot_row.size();
    hot_row.resize(esl::kHotSlots, 0);
    CU(ctx->d_hot_row.upload(hot_row.data(), hot_row.size() * sizeof(int32_t), st));
    ctx->n_seg = (int)seg.size(); ctx->n_tgrid = (int)n_grid; ctx->n_trec = (int)n_rec;
    ctx->raw_dirty = true; // targets in caller order: uploaded by the staged baseline only, on its first call
    CU(ctx->d_seg.upload(seg.data(), seg.size() * sizeof(esl::SegDesc), st));
    CU(ctx->d_extra_off.upload(extra_off.data(), extra_off.size() * sizeof(int32_t), st));
    tm.mark("target uploads");
    ctx->index_on_device = false;
    return ESL_OK;
}

// LSD radix sort of n 32-bit keys (+ 32-bit payload; vals_in == NULL: the payload is the element's index) on the
// device with the staged baseline's sort kernels, 8-bit digits, as many passes as max_key needs.  Asynchronous;
// *keys_out / *vals_out point into the scratch buffers.
int device_radix_sort(esl_ctx *ctx, SortScratch &sc, const uint32_t *keys_in, const uint32_t *vals_in, i64 n, uint32_t max_key,
                      const uint32_t **keys_out, const uint32_t **vals_out, cudaStream_t st)
{
    const int sort_blocks = (int)((n + esl::kSortTile - 1) / esl::kSortTile);
    for (int i = 0; i < 2; ++i) {
        CU(sc.keys[i].reserve((size_t)n * 4));
        CU(sc.vals[i].reserve((size_t)n * 4));
    }
    CU(sc.hist.reserve((size_t)sort_blocks * 256 * 4 + 256 * 4 + 16));
    uint32_t *hist = sc.hist.as<uint32_t>(), *digit_tot = hist + (size_t)sort_blocks * 256;
    int *n_dev = (int *)(digit_tot + 256);
    const int n_host = (int)n;
    CU(cudaMemcpyAsync(n_dev, &n_host, sizeof(int), cudaMemcpyHostToDevice, st));
    int passes = 1;
    while (passes < 4 && (max_key >> (8 * passes)) != 0) ++passes;
    const uint32_t *kin = keys_in, *vin = vals_in;
    for (int pass = 0; pass < passes; ++pass) {
        uint32_t *ko = sc.keys[pass & 1].as<uint32_t>(), *vo = sc.vals[pass & 1].as<uint32_t>();
        esl::k_rs_hist<<<sort_blocks, esl::kStageBlock, 0, st>>>(kin, n_dev, pass * 8, sort_blocks, hist);
        esl::k_rs_scan_rows<<<256, 1024, 0, st>>>(hist, sort_blocks, digit_tot);
        esl::k_rs_scan_digits<<<1, 1024, 0, st>>>(digit_tot);
        esl::k_rs_scatter<<<sort_blocks, esl::kStageBlock, 0, st>>>(kin, vin, ko, vo, n_dev, pass * 8, sort_blocks, hist, digit_tot,
                                                                  pass == 0 && vals_in == nullptr);
        kin = ko; vin = vo;
        ctx->launches += 4;
    }
    CU(cudaGetLastError());
    *keys_out = kin;
    if (vals_out) *vals_out = vin;
    return ESL_OK;
}

// The target index built on the device (index_build.cuh): large target sets with 32-bit coordinates whose ends ascend
// with their starts inside every barcode and that hold no hot target.  *built = false: the geometry needs the host builder.
int build_targets_device(esl_ctx *ctx, cudaStream_t st, PhaseTimer &tm, bool *built)
{
    *built = false;
    const int B = ctx->B;
    const u64 G = ctx->G;
    const int n = (int)ctx->t_start.size();
    // raw targets in caller order, straight from the page-locked copies esl_set_targets made
    CU(ctx->d_raw_ts.reserve((size_t)n * 8)); CU(ctx->d_raw_te.reserve((size_t)n * 8)); CU(ctx->d_raw_tb.reserve((size_t)n * 4));
    CU(cudaMemcpyAsync(ctx->d_raw_ts.p, ctx->t_start.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_raw_te.p, ctx->t_end.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_raw_tb.p, ctx->t_bc.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    ctx->raw_dirty = false;
    CU(ctx->d_idx_key_s.reserve((size_t)n * 4)); CU(ctx->d_idx_key_b.reserve((size_t)n * 4));
    CU(ctx->d_idx_res.reserve(sizeof(esl::IdxResult)));
    CU(ctx->d_trec.reserve((size_t)n * sizeof(esl::TargetRec<uint32_t>)));
    CU(ctx->d_te.reserve((size_t)n * 4));
    CU(ctx->d_seg.reserve((size_t)B * sizeof(esl::SegDesc)));
    CU(ctx->d_extra_off.reserve((size_t)B * 4));
    const size_t grid_cap = (size_t)kBucketsPerInterval * (size_t)n + (size_t)B * (size_t)((G >> 16) + 67);
    CU(ctx->d_tgrid.reserve(grid_cap * sizeof(esl::GridEntry<uint32_t>)));
    CU(ctx->d_hot_row.reserve(esl::kHotSlots * sizeof(int32_t)));
    CU(cudaMemsetAsync(ctx->d_idx_res.p, 0, sizeof(esl::IdxResult), st));
    const i64 *ts = ctx->d_raw_ts.as<i64>(), *te = ctx->d_raw_te.as<i64>();
    uint32_t *key_s = ctx->d_idx_key_s.as<uint32_t>(), *key_b = ctx->d_idx_key_b.as<uint32_t>();
    const unsigned blocks = (unsigned)((n + 255) / 256);
    esl::k_idx_keys<<<blocks, 256, 0, st>>>(ts, te, ctx->d_raw_tb.as<int32_t>(), n, B, G, key_s, key_b);
    const uint32_t *sorted = nullptr, *perm = nullptr;
    int rc = device_radix_sort(ctx, ctx->sort, key_s, nullptr, n, 0xFFFFFFFFu, &sorted, &perm, st); // by start; payload = row
    if (rc) return rc;
    CU(cudaMemcpyAsync(key_s, perm, (size_t)n * 4, cudaMemcpyDeviceToDevice, st)); // the rows in start order, out of the sort's way
    uint32_t *bc_by_start = ctx->d_te.as<uint32_t>();                              // (scratch until k_idx_records fills it)
    esl::k_idx_gather<<<blocks, 256, 0, st>>>(key_b, key_s, n, bc_by_start);
    const uint32_t *kb = nullptr;
    rc = device_radix_sort(ctx, ctx->sort, bc_by_start, key_s, n, (uint32_t)B, &kb, &perm, st); // stable: by barcode, starts stay ordered
    if (rc) return rc;
    esl::IdxResult *res = ctx->d_idx_res.as<esl::IdxResult>();
    esl::k_idx_bounds<<<1, 1024, 0, st>>>(kb, n, B, G, kBucketsPerInterval, ctx->d_seg.as<esl::SegDesc>(), ctx->d_extra_off.as<int32_t>(), res);
    const double hot_len = (double)G / 512.0 - ctx->pool_mean;
    esl::k_idx_records<<<blocks, 256, 0, st>>>(ts, te, perm, kb, ctx->d_seg.as<esl::SegDesc>(), n, B, G, hot_len,
                                                ctx->d_trec.as<esl::TargetRec<uint32_t>>(), ctx->d_te.as<uint32_t>(), res);
    esl::k_idx_grid<<<(unsigned)((grid_cap + 255) / 256), 256, 0, st>>>(ctx->d_seg.as<esl::SegDesc>(), B, ctx->d_trec.as<esl::TargetRec<uint32_t>>(),
                                                                        ctx->d_te.as<uint32_t>(), res, ctx->d_tgrid.as<esl::GridEntry<uint32_t>>());
    CU(cudaGetLastError());
    ctx->launches += 5;
    esl::IdxResult h;
    CU(cudaMemcpyAsync(&h, res, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st)); // one small read-back decides whether the geometry was simple enough
    tm.mark("device target index");
    if (h.flags) return ESL_OK; // nested or hot targets: the host builder takes over
    ctx->n_layers = ctx->n_first_layers = h.n_nonempty;
    ctx->n_indexed = h.n_valid;
    ctx->n_hot = 0;
    ctx->sgrid_n = 0;
    ctx->sgrid_lmax = 1;
    while (ctx->sgrid_lmax < (u64)ctx->pool_max) ctx->sgrid_lmax <<= 1;
    ctx->n_seg = B; ctx->n_tgrid = h.n_grid; ctx->n_trec = h.n_valid;
    ctx->index_on_device = true;
    *built = true;
    return ESL_OK;
}

template <typename C>
int upload_index(esl_ctx *ctx, cudaStream_t st)
{
    const int B = ctx->B;
    const u64 G = ctx->G;
    PhaseTimer tm;
    // who builds the target index: the device for large sets with 32-bit coordinates (it hands back to the host when it
    // meets nesting or hot targets), the host otherwise; esl_set_index_builder overrides the size rule
    bool on_device = false;
    const bool try_device = sizeof(C) == 4 && ctx->index_builder != ESL_INDEX_HOST && ctx->t_start.size() > 0 &&
                            (ctx->index_builder == ESL_INDEX_DEVICE || ctx->t_start.size() >= 16384);
    if (try_device) {
        int rc = build_targets_device(ctx, st, tm, &on_device);
        if (rc) return rc;
    }
    if (!on_device) {
        int rc = build_targets_host<C>(ctx, st, tm);
        if (rc) return rc;
    }
    // blocked intervals: group g < B = barcode g, group B = "ALL"; (start, end) order, running max of ends
    std::vector<C> ms, me, mp;
    std::vector<double> mw;
    std::vector<esl::SegDesc> grp;
    std::vector<esl::GridEntry<C>> mgrid;
    i64 n_real = 0;
    for (int g = 0; g <= B; ++g) {
        const int want = g < B ? g : -1;
        std::vector<size_t> idx;
        for (size_t i = 0; i < ctx->m_start.size(); ++i)
            if (ctx->m_bc[i] == want) idx.push_back(i);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b2) {
            if (ctx->m_start[a] != ctx->m_start[b2]) return ctx->m_start[a] < ctx->m_start[b2];
            return ctx->m_end[a] < ctx->m_end[b2];
        });
        C run = 0;
        const int glo = (int)ms.size();
        for (size_t i : idx) {
            const C s = clampc<C>(std::max<i64>(ctx->m_start[i], 0), G), e = clampc<C>(std::max<i64>(ctx->m_end[i], 0), G);
            run = std::max(run, e);
            ms.push_back(s); me.push_back(e); mp.push_back(run); mw.push_back(ctx->m_w[i]);
        }
        if ((int)ms.size() == glo) grp.push_back(empty_run<C>(mgrid));
        else grp.push_back(make_grid<C>(mp.data(), ms.data(), glo, (int)ms.size(), G, mgrid));
        n_real += (i64)ms.size() - glo;
        // sentinel behind the group (outside [lo, hi)): a start of all ones ends any walk without a bound check
        ms.push_back((C) ~(C)0); me.push_back((C) ~(C)0); mp.push_back((C) ~(C)0); mw.push_back(0.0);
    }
    ctx->n_mask = n_real;
    {   // blocked-bucket map, two bits per bucket (kernels.cuh DevCtx::mbits): at most 64 Ki buckets (16 KB of shared
        // memory; hg38: 47 k buckets of 64 kb, 12 KB -- measured on config 4: 58.7 G reads/s against 57.5 with 128-kb
        // buckets in 6 KB and 56.5 without the second bit), at least 64 kb wide.  Bit 0: some blocked interval of any tree touches the bucket.  Bit 1: the bucket
        // lies entirely inside the union of the "ALL" intervals of weight >= 1 -- a read inside such buckets is blocked
        // whatever its barcode and whatever uniform would be drawn.
        int sh = 16;
        while (((G + 1) >> sh) + 1 > MBITS_MAX_BUCKETS) ++sh;
        const size_t nb = (size_t)((G + 1) >> sh) + 2;
        std::vector<uint32_t> bits((nb + 15) / 16, 0u);
        std::vector<std::pair<u64, u64>> sure;
        for (size_t i = 0; i < ctx->m_start.size(); ++i) {
            const u64 a = (u64)std::max<i64>(ctx->m_start[i], 0), e = (u64)std::max<i64>(ctx->m_end[i], 0);
            if (e <= a || a > G) continue;
            for (u64 b = a >> sh, b1 = (std::min<u64>(e, G + 1) - 1) >> sh; b <= b1; ++b) bits[b >> 4] |= 1u << ((b & 15) * 2);
            if (ctx->m_bc[i] < 0 && ctx->m_w[i] >= 1.0) sure.emplace_back(a, e >= G ? ~0ull : e); // (reads end at or before G)
        }
        std::sort(sure.begin(), sure.end());
        for (size_t i = 0; i < sure.size();) { // merged runs of certain intervals -> the buckets they cover completely
            u64 a = sure[i].first, e = sure[i].second;
            size_t j = i + 1;
            while (j < sure.size() && sure[j].first <= e) { e = std::max(e, sure[j].second); ++j; }
            const u64 first = (a + ((1ull << sh) - 1)) >> sh;
            const u64 last = e == ~0ull ? nb : std::min<u64>(e >> sh, nb); // buckets [first, last) are inside [a, e)
            for (u64 b = first; b < last; ++b) bits[b >> 4] |= 2u << ((b & 15) * 2);
            i = j;
        }
        ctx->mbits_shift = sh;
        ctx->mbits_words = (int)bits.size();
        CU(ctx->d_mbits.upload(bits.data(), bits.size() * sizeof(uint32_t), st));
    }
    std::vector<esl::MaskRec<C>> mrec(ms.size());
    for (size_t k = 0; k < ms.size(); ++k) { mrec[k] = esl::MaskRec<C>{}; mrec[k].s = ms[k]; mrec[k].e = me[k]; mrec[k].w = mw[k]; }
    ctx->n_mgrid = (int)mgrid.size(); ctx->n_mrec = (int)mrec.size();
    CU(ctx->d_mrec.upload(mrec.data(), mrec.size() * sizeof(esl::MaskRec<C>), st));
    CU(ctx->d_mpmax.upload(mp.data(), mp.size() * sizeof(C), st));
    CU(ctx->d_grp.upload(grp.data(), grp.size() * sizeof(esl::SegDesc), st));
    CU(ctx->d_mgrid.upload(mgrid.data(), mgrid.size() * sizeof(esl::GridEntry<C>), st));
    // barcode cdf: doubles (replay), integer thresholds and a 1024-slice start table (native)
    std::vector<uint32_t> cu(B);
    for (int k = 0; k < B; ++k) {
        const double v = std::ceil(ctx->cdf[k] * 4294967296.0);
        cu[k] = v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
    }
    std::vector<uint2> lut(esl::kCdfLut);
    int k = 0, walk = 0;
    for (int q = 0; q < esl::kCdfLut; ++q) {
        const uint32_t w = (uint32_t)q << 22; // lowest draw of the slice
        while (k < B - 1 && cu[k] <= w) ++k;
        lut[q] = make_uint2((uint32_t)k | (k < B - 1 ? 0x10000u : 0u), k < B - 1 ? cu[k] : 0xFFFFFFFFu);
        const uint32_t w_hi = w | 0x3FFFFFu; // highest draw of the slice
        int k2 = k;
        while (k2 < B - 1 && cu[k2] <= w_hi) ++k2;
        walk = std::max(walk, k2 - k);
    }
    ctx->cdf_walk = walk;
    CU(ctx->d_cdf.upload(ctx->cdf.data(), B * sizeof(double), st));
    CU(ctx->d_cdf_u32.upload(cu.data(), B * sizeof(uint32_t), st));
    CU(ctx->d_cdf_lut.upload(lut.data(), lut.size() * sizeof(uint2), st));
    tm.mark("blocked groups + cdf");
    return ESL_OK;
}

// Set-up state checks shared by esl_prepare and the run calls.
int check_inputs(esl_ctx *ctx, bool need_pool)
{
    if (ctx->G == 0) return ctx->fail(ESL_ERR_STATE, "esl_set_genome has not been called");
    if (ctx->B <= 0) return ctx->fail(ESL_ERR_STATE, "esl_set_barcode_cdf has not been called");
    if (!ctx->have_targets) return ctx->fail(ESL_ERR_STATE, "esl_set_targets has not been called");
    if (need_pool && !ctx->have_pool) return ctx->fail(ESL_ERR_STATE, "esl_set_length_table has not been called");
    return ESL_OK;
}

// The shared pre-filter's upper bounds assume no read is longer than sgrid_lmax: a longer pool needs a rebuild.
bool index_stale(const esl_ctx *ctx)
{
    return ctx->index_dirty || (ctx->have_pool && ctx->sgrid_n && (u64)ctx->pool_max > ctx->sgrid_lmax) ||
           (ctx->have_pool && ctx->sampler == ESL_SAMPLER_QUANTILE && ctx->qtab_dirty);
}

// Quantile sampler set-up: the resident pool sorted on the device, and its quantile table.  Asynchronous.
int build_quantile_table(esl_ctx *ctx, cudaStream_t st)
{
    const i64 n = ctx->n_pool;
    CU(ctx->d_pool_sorted.reserve((size_t)n * 4));
    CU(ctx->d_qtab.reserve((size_t)(esl::kQuantBins + 1) * 4));
    const uint32_t *sorted = nullptr;
    int rc = device_radix_sort(ctx, ctx->sort, ctx->d_pool.as<uint32_t>(), nullptr, n, ctx->pool_max, &sorted, nullptr, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->d_pool_sorted.p, sorted, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    esl::k_quantile_table<<<(esl::kQuantBins + 256) / 256, 256, 0, st>>>(ctx->d_pool_sorted.as<uint32_t>(), n, ctx->d_qtab.as<uint32_t>());
    CU(cudaGetLastError());
    ctx->launches++;
    ctx->qtab_dirty = false;
    return ESL_OK;
}

// Run calls neither allocate nor block: they require the index esl_prepare built for the current inputs.
int require_index(esl_ctx *ctx, bool need_pool)
{
    int rc = check_inputs(ctx, need_pool);
    if (rc) return rc;
    if (index_stale(ctx)) return ctx->fail(ESL_ERR_STATE, "inputs changed since the last esl_prepare: call esl_prepare before running");
    return ESL_OK;
}

template <typename C>
esl::DevCtx<C> make_devctx(const esl_ctx *ctx, double coverage, int consuming, i64 min_overlap, double scaling)
{
    esl::DevCtx<C> d;
    d.G = ctx->G;
    d.thr = (i64)std::ceil(coverage * (double)ctx->G);
    d.min_overlap = min_overlap;
    d.scaling = scaling;
    d.n_barcodes = ctx->B;
    d.consuming = consuming ? 1 : 0;
    d.has_mask = ctx->n_mask > 0;
    d.mask_mode = 0;
    for (int32_t b : ctx->m_bc) d.mask_mode |= b >= 0 ? 1 : 2;
    d.mbits = ctx->d_mbits.as<uint32_t>();
    d.mbits_shift = ctx->mbits_shift;
    d.mbits_words = ctx->mbits_words;
    d.n_pool = ctx->n_pool;
    d.pool = ctx->d_pool.as<uint32_t>();
    const bool quant = ctx->sampler == ESL_SAMPLER_QUANTILE && ctx->have_pool;
    d.qtab = quant ? ctx->d_qtab.as<uint32_t>() : nullptr;
    d.pool_sorted = ctx->d_pool_sorted.as<uint32_t>();
    d.q_top_lo = (uint32_t)(((u64)(esl::kQuantBins - 1) * (u64)ctx->n_pool) / esl::kQuantBins);
    d.q_top_n = ctx->n_pool - d.q_top_lo;
    d.cdf = ctx->d_cdf.as<double>();
    d.cdf_u32 = ctx->d_cdf_u32.as<uint32_t>();
    d.cdf_lut = ctx->d_cdf_lut.as<uint2>();
    d.cdf_walk = ctx->cdf_walk;
    d.min_ov32 = min_overlap <= 0 ? 0u : (min_overlap > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)min_overlap);
    d.thin = scaling < 1.0;
    d.trec = ctx->d_trec.as<esl::TargetRec<C>>();
    d.te = ctx->d_te.as<C>();
    d.seg = ctx->d_seg.as<esl::SegDesc>();
    d.tgrid = ctx->d_tgrid.as<esl::GridEntry<C>>();
    d.extra_off = ctx->d_extra_off.as<int32_t>();
    d.any_extra = ctx->n_layers > ctx->n_first_layers;
    d.hot_row = ctx->d_hot_row.as<int32_t>();
    d.n_hot = ctx->n_hot;
    d.n_seg = ctx->n_seg; d.n_tgrid = ctx->n_tgrid; d.n_trec = ctx->n_trec; d.n_mgrid = ctx->n_mgrid; d.n_mrec = ctx->n_mrec;
    d.sgrid = ctx->sgrid_n ? ctx->d_sgrid.as<C>() : nullptr;
    d.sgrid_n = ctx->sgrid_n;
    d.sgrid_shift = ctx->sgrid_shift;
    d.n_targets = (i64)ctx->t_start.size();
    d.mrec = ctx->d_mrec.as<esl::MaskRec<C>>();
    d.mpmax = ctx->d_mpmax.as<C>();
    d.grp = ctx->d_grp.as<esl::SegDesc>();
    d.mgrid = ctx->d_mgrid.as<esl::GridEntry<C>>();
    return d;
}

// Draws per iteration handed to the bulk kernel: the largest N with N*mean + z*sd*sqrt(N) <= coverage*G,
// i.e. z standard deviations of the length sum below the expected cut-off (z = 8; Cantelli bounds the
// overshoot probability by 1/(1+z^2) for ANY length distribution, and the tail kernel corrects an
// overshoot exactly, so this only ever costs time).  Blocking with consuming=False only moves the
// cut-off later, so the bound stays valid.
i64 bulk_reads(const esl_ctx *ctx, double coverage)
{
    if (!ctx->have_pool || ctx->pool_mean <= 0) return 0;
    const double thr = coverage * (double)ctx->G, z = 8.0, mu = ctx->pool_mean, sd = ctx->pool_sd;
    if (!(thr > 0)) return 0;
    const double r = (-z * sd + std::sqrt(z * z * sd * sd + 4.0 * mu * thr)) / (2.0 * mu);
    i64 n = (i64)std::floor(r * r) - 1;
    n -= n % esl::kBulkTile; // whole tiles
    return n < 2 * esl::kBulkTile ? 0 : n;
}

size_t bulk_smem(const esl_ctx *ctx, bool quant)
{
    return esl::bulk_smem_layout(ctx->wide ? 8 : 4, ctx->n_mask > 0, ctx->sgrid_n > 0, ctx->n_hot > 0, ctx->B, ctx->sgrid_n, ctx->mbits_words, quant).total;
}
size_t tail_smem(int B) { return (size_t)112 * 8 + (size_t)(2 * B + 2) * 4 + (size_t)esl::kCdfLut * 8 + 24; }

// Upper bound on the draws per iteration any bulk pass may cover.  Without a mask (or with consuming = True)
// every draw counts and the first-pass bound n1 is already tight.  With a mask a blocked draw may count nothing,
// so a second pass can extend up to this cap; the blocked fraction is bounded from above by the union bound
// sum_k w_k (len_k + mean read length) / G.  The cap only sizes scratch and grids -- the per-iteration extent of
// the second pass is decided on the device from the measured counted mean, and k_cutoff_tail finishes exactly.
i64 bulk_cap(const esl_ctx *ctx, double coverage)
{
    const i64 n1 = bulk_reads(ctx, coverage);
    if (ctx->m_start.empty() || n1 == 0) return n1;
    long double blocked = 0;
    for (size_t k = 0; k < ctx->m_start.size(); ++k) {
        const long double len = (long double)std::max<i64>(0, ctx->m_end[k] - ctx->m_start[k]);
        blocked += std::min<long double>(1.0L, std::max<long double>(0.0L, ctx->m_w[k])) * (len + ctx->pool_mean);
    }
    const double f = (double)std::min<long double>(0.9L, blocked / (long double)ctx->G);
    const double expect = coverage * (double)ctx->G / ctx->pool_mean / (1.0 - f);
    i64 cap = (i64)(expect * 1.02) + 4 * esl::kBulkTile;
    cap = std::max(cap, n1);
    cap = std::min<i64>(cap, 0xFFFF0000ll);
    cap += (esl::kBulkTile - cap % esl::kBulkTile) % esl::kBulkTile;
    return cap;
}

struct BulkRange {
    const i64 *lo_arr; i64 lo_const;
    const i64 *hi_arr; i64 hi_const;
    i64 tile0, tiles; // absolute tile window of the launch (per iteration)
};

template <typename C, typename Src>
int launch_bulk(esl_ctx *ctx, const esl::DevCtx<C> &dc, const Src &src, const esl::DevOut &out, const BulkRange &r,
                int n_iter, u64 *d_sched, cudaStream_t st)
{
    const i64 total_tiles = r.tiles * n_iter;
    if (total_tiles <= 0) return ESL_OK;
    // one instantiation per (blocked regions?, small target set with a shared pre-filter?, hot targets?)
    using Kern = void (*)(esl::DevCtx<C>, Src, esl::DevOut, const i64 *, i64, const i64 *, i64, int, int, int, u64 *);
    static const Kern table[8] = {
        esl::k_place_tally_bulk<C, Src, false, false, false>, esl::k_place_tally_bulk<C, Src, false, false, true>,
        esl::k_place_tally_bulk<C, Src, false, true, false>,  esl::k_place_tally_bulk<C, Src, false, true, true>,
        esl::k_place_tally_bulk<C, Src, true, false, false>,  esl::k_place_tally_bulk<C, Src, true, false, true>,
        esl::k_place_tally_bulk<C, Src, true, true, false>,   esl::k_place_tally_bulk<C, Src, true, true, true>};
    const Kern kern = table[(dc.has_mask ? 4 : 0) | (dc.sgrid ? 2 : 0) | (dc.n_hot ? 1 : 0)];
    int occ = 0;
    if (bulk_smem(ctx, dc.qtab != nullptr && Src::kPoolLengths) > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem(ctx, dc.qtab != nullptr && Src::kPoolLengths)));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, esl::kBulkBlock, bulk_smem(ctx, dc.qtab != nullptr && Src::kPoolLengths)));
    if (occ < 1) occ = 1;
    const i64 grid = std::min<i64>((total_tiles + esl::kBulkClaim - 1) / esl::kBulkClaim, (i64)ctx->sm_count * occ);
    kern<<<(unsigned)grid, esl::kBulkBlock, bulk_smem(ctx, dc.qtab != nullptr && Src::kPoolLengths), st>>>(dc, src, out, r.lo_arr, r.lo_const, r.hi_arr, r.hi_const,
                                                                      (int)r.tile0, (int)r.tiles, n_iter, d_sched);
    CU(cudaGetLastError());
    ctx->launches++;
    return ESL_OK;
}

// Workspace layout: [tile sums: tiles * n_iter u64][second-pass ranges: 2 * n_iter i64][per-iteration tile counters of
// the two bulk launches: 2 * n_iter u64]
struct Workspace {
    u64 *tile_sum;
    i64 *range;
    u64 *sched;
    size_t tile_bytes;
};
size_t ws_tile_bytes(i64 cap, int n_iter) { return (size_t)((cap + esl::kBulkTile - 1) / esl::kBulkTile) * (size_t)n_iter * sizeof(u64); }
size_t ws_total_bytes(i64 cap, int n_iter) { return ((ws_tile_bytes(cap, n_iter) + 15) & ~(size_t)15) + (size_t)(4 * n_iter + 2) * sizeof(i64) + 64; }
Workspace ws_carve(void *d_ws, i64 cap, int n_iter)
{
    Workspace w;
    w.tile_bytes = ws_tile_bytes(cap, n_iter);
    w.tile_sum = (u64 *)d_ws;
    w.range = (i64 *)((char *)d_ws + ((w.tile_bytes + 15) & ~(size_t)15));
    w.sched = (u64 *)(w.range + 2 * n_iter);
    return w;
}

// Bulk pass(es) + exact cut-off kernel.  n1 = draws per iteration of the first pass; n_cap >= n1 = bound of the
// optional second pass; ws.tile_sum holds ceil(max(n1, n_cap) / tile) entries per iteration and is zeroed by the
// caller when the cut-off kernel will read it.
template <typename C, typename Src>
int launch_pipeline(esl_ctx *ctx, const esl::DevCtx<C> &dc, const Src &src, esl::DevOut out, i64 n1, i64 n_cap,
                    const Workspace &ws, bool second_pass, int n_iter, bool run_tail, cudaStream_t st)
{
    const i64 T = esl::kBulkTile;
    out.tile_sum = ws.tile_sum;
    out.tiles_per_iter = (int)((std::max(n1, n_cap) + T - 1) / T);
    const bool prof = ctx->profiling && ctx->ev[0];
    ctx->ev_valid = false;
    CU(cudaMemsetAsync(ws.sched, 0, 2 * (size_t)n_iter * sizeof(u64), st));
    if (prof) CU(cudaEventRecord(ctx->ev[0], st));
    int rc = launch_bulk<C, Src>(ctx, dc, src, out, BulkRange{nullptr, 0, nullptr, n1, 0, (n1 + T - 1) / T}, n_iter, ws.sched, st);
    if (rc) return rc;
    const i64 *d_hi = nullptr;
    if (second_pass && n_cap > n1 && n1 > 0) {
        i64 *d_lo = ws.range, *d_hi2 = ws.range + n_iter;
        const double m2 = ctx->pool_sd * ctx->pool_sd + ctx->pool_mean * ctx->pool_mean;
        esl::k_plan_second_pass<<<(n_iter + 127) / 128, 128, 0, st>>>(out.n_bases, n_iter, dc.thr, n1, n_cap, m2, 8.0, d_lo, d_hi2);
        CU(cudaGetLastError());
        ctx->launches++;
        rc = launch_bulk<C, Src>(ctx, dc, src, out, BulkRange{d_lo, 0, d_hi2, 0, n1 / T, (n_cap - n1) / T}, n_iter, ws.sched + n_iter, st);
        if (rc) return rc;
        d_hi = d_hi2;
    }
    if (prof) { CU(cudaEventRecord(ctx->ev[1], st)); CU(cudaEventRecord(ctx->ev[2], st)); }
    if (run_tail) {
        auto kern = esl::k_cutoff_tail<C, Src>;
        kern<<<(unsigned)n_iter, esl::kTailBlock, tail_smem(ctx->B), st>>>(dc, src, out, d_hi, n1);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    if (prof) { CU(cudaEventRecord(ctx->ev[3], st)); ctx->ev_valid = true; }
    return ESL_OK;
}

int zero_outputs(esl_ctx *ctx, int n_iter, int64_t *d_tallies, int64_t *d_label_counts, int64_t *d_n_bases,
                 int64_t *d_n_reads, int32_t *d_status, cudaStream_t st)
{
    const size_t T = ctx->t_start.size();
    CU(cudaMemsetAsync(d_tallies, 0, (size_t)n_iter * T * 3 * sizeof(int64_t), st));
    CU(cudaMemsetAsync(d_label_counts, 0, (size_t)n_iter * (ctx->B + 2) * sizeof(int64_t), st));
    CU(cudaMemsetAsync(d_n_bases, 0, (size_t)n_iter * sizeof(int64_t), st));
    CU(cudaMemsetAsync(d_n_reads, 0, (size_t)n_iter * sizeof(int64_t), st));
    CU(cudaMemsetAsync(d_status, 0, (size_t)n_iter * sizeof(int32_t), st));
    return ESL_OK;
}

int check_outputs(esl_ctx *ctx, const void *a, const void *b, const void *c, const void *d, const void *e)
{
    if ((!a && !ctx->t_start.empty()) || !b || !c || !d || !e) return ctx->fail(ESL_ERR_INVALID, "an output pointer is NULL");
    return ESL_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------------

extern "C" {

int esl_abi_version(void) { return 2; }

const char *esl_last_error(const esl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int esl_create(esl_ctx **out, int device)
{
    if (!out) { g_create_error = "ctx out-pointer is NULL"; return ESL_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ESL_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return ESL_ERR_INVALID; }
    cudaDeviceProp p;
    if ((e = cudaGetDeviceProperties(&p, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return ESL_ERR_CUDA; }
    if (p.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library carries sm_100a code only", device, p.major, p.minor);
        g_create_error = buf;
        return ESL_ERR_NO_DEVICE;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return ESL_ERR_CUDA; }
    esl_ctx *c = new esl_ctx();
    c->host = new esl_host_scratch();
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    *out = c;
    return ESL_OK;
}

int esl_destroy(esl_ctx *ctx)
{
    if (!ctx) return ESL_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (cudaEvent_t e : ctx->ev) if (e) cudaEventDestroy(e);
    delete ctx->host;
    delete ctx;
    return ESL_OK;
}

int esl_set_genome(esl_ctx *ctx, uint64_t genome_size)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (genome_size == 0) return ctx->fail(ESL_ERR_INVALID, "genome_size must be > 0");
    if (genome_size > (1ull << 62)) return ctx->fail(ESL_ERR_INVALID, "genome_size too large");
    ctx->G = genome_size;
    ctx->index_dirty = true;
    return ESL_OK;
}

int esl_set_barcode_cdf(esl_ctx *ctx, const double *cdf, int32_t n_barcodes)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (!cdf || n_barcodes < 1 || n_barcodes > esl::kMaxBarcodes)
        return ctx->fail(ESL_ERR_INVALID, "n_barcodes must be in [1, %d]", esl::kMaxBarcodes);
    for (int k = 0; k < n_barcodes; ++k)
        if (!(cdf[k] >= 0.0) || (k && cdf[k] < cdf[k - 1])) return ctx->fail(ESL_ERR_INVALID, "cdf must be non-decreasing and >= 0");
    ctx->cdf.assign(cdf, cdf + n_barcodes);
    ctx->B = n_barcodes;
    ctx->index_dirty = true;
    return ESL_OK;
}

int esl_set_targets(esl_ctx *ctx, const int64_t *start, const int64_t *end, const int32_t *barcode, int64_t n)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n < 0 || n > 0x7fffffff / 3 || (n && (!start || !end || !barcode))) return ctx->fail(ESL_ERR_INVALID, "bad target arrays");
    for (int64_t i = 0; i < n; ++i)
        if (start[i] < 0 || end[i] < 0) return ctx->fail(ESL_ERR_INVALID, "target %lld has a negative coordinate", (long long)i);
    CU(cudaSetDevice(ctx->device));
    if (!ctx->t_start.assign(start, start + n) || !ctx->t_end.assign(end, end + n) || !ctx->t_bc.assign(barcode, barcode + n))
        return ctx->fail(ESL_ERR_CUDA, "cannot allocate page-locked memory for %lld targets", (long long)n);
    ctx->have_targets = true;
    ctx->index_dirty = true;
    return ESL_OK;
}

int32_t esl_targets_equal(const esl_ctx *ctx, const int64_t *start, const int64_t *end, const int32_t *barcode, int64_t n)
{
    if (!ctx || !ctx->have_targets || n < 0 || (size_t)n != ctx->t_start.size()) return 0;
    if (n == 0) return 1;
    if (!start || !end || !barcode) return 0;
    return memcmp(start, ctx->t_start.data(), (size_t)n * 8) == 0 && memcmp(end, ctx->t_end.data(), (size_t)n * 8) == 0 &&
           memcmp(barcode, ctx->t_bc.data(), (size_t)n * 4) == 0;
}

int esl_set_blocked(esl_ctx *ctx, const int64_t *start, const int64_t *end, const double *weight,
                    const int32_t *barcode, int64_t n)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n < 0 || n > 0x7fffffff || (n && (!start || !end || !weight || !barcode))) return ctx->fail(ESL_ERR_INVALID, "bad blocked arrays");
    for (int64_t i = 0; i < n; ++i)
        if (barcode[i] < -1) return ctx->fail(ESL_ERR_INVALID, "blocked interval %lld: barcode must be >= -1", (long long)i);
    ctx->m_start.assign(start, start + n);
    ctx->m_end.assign(end, end + n);
    ctx->m_w.assign(weight, weight + n);
    ctx->m_bc.assign(barcode, barcode + n);
    ctx->index_dirty = true;
    return ESL_OK;
}

int esl_set_length_table(esl_ctx *ctx, const uint32_t *lengths, int64_t n, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (!lengths || n < 1 || n > 0xFFFFFFFFll) return ctx->fail(ESL_ERR_INVALID, "length table must hold 1 .. 2^32-1 entries");
    // upload, then reduce on the device: sum (exact), sum of squares (double), max and min -- all on the caller's stream
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(ctx->d_pool.upload(lengths, (size_t)n * sizeof(uint32_t), st));
    CU(ctx->d_pool_stats.reserve(32));
    struct { u64 sum; double sq; uint32_t mx, mn; } h = {0, 0.0, 0u, 0xFFFFFFFFu};
    CU(cudaMemcpyAsync(ctx->d_pool_stats.p, &h, sizeof h, cudaMemcpyHostToDevice, st));
    {
        char *b8 = (char *)ctx->d_pool_stats.p;
        const unsigned grid = (unsigned)std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 4);
        esl::k_pool_moments<<<grid, 256, 0, st>>>(ctx->d_pool.as<uint32_t>(), n, (u64 *)b8, (double *)(b8 + 8), (uint32_t *)(b8 + 16), (uint32_t *)(b8 + 20));
        CU(cudaGetLastError());
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(&h, ctx->d_pool_stats.p, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st)); // the bulk range is planned on the host from these statistics
    if (h.mn == 0) return ctx->fail(ESL_ERR_INVALID, "length table holds a zero-length read");
    ctx->n_pool = (uint32_t)n;
    ctx->pool_max = h.mx;
    const long double mean = (long double)h.sum / (long double)n;
    const long double var = (long double)h.sq / (long double)n - mean * mean;
    ctx->pool_mean = (double)mean;
    ctx->pool_sd = var > 0 ? (double)sqrtl(var) : 0.0;
    ctx->have_pool = true;
    ctx->qtab_dirty = true;
    return ESL_OK;
}

int esl_set_index_builder(esl_ctx *ctx, int32_t mode)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (mode != ESL_INDEX_AUTO && mode != ESL_INDEX_HOST && mode != ESL_INDEX_DEVICE) return ctx->fail(ESL_ERR_INVALID, "unknown index builder %d", mode);
    if (mode != ctx->index_builder) ctx->index_dirty = true;
    ctx->index_builder = mode;
    return ESL_OK;
}

int32_t esl_index_built_on_device(const esl_ctx *ctx) { return ctx && ctx->index_on_device ? 1 : 0; }

int esl_set_length_sampler(esl_ctx *ctx, int32_t mode)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (mode != ESL_SAMPLER_POOL && mode != ESL_SAMPLER_QUANTILE) return ctx->fail(ESL_ERR_INVALID, "unknown length sampler %d", mode);
    if (mode != ctx->sampler) ctx->qtab_dirty = true;
    ctx->sampler = mode;
    return ESL_OK;
}

int esl_prepare(esl_ctx *ctx, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    int rc = check_inputs(ctx, false);
    if (rc) return rc;
    for (int32_t b : ctx->m_bc)
        if (b >= ctx->B) return ctx->fail(ESL_ERR_INVALID, "blocked interval names barcode %d but n_barcodes is %d", b, ctx->B);
    if (!index_stale(ctx)) return ESL_OK;
    CU(cudaSetDevice(ctx->device));
    if (ctx->have_pool && ctx->sampler == ESL_SAMPLER_QUANTILE && ctx->qtab_dirty && (rc = build_quantile_table(ctx, (cudaStream_t)stream))) return rc;
    if (ctx->index_dirty || (ctx->have_pool && ctx->sgrid_n && (u64)ctx->pool_max > ctx->sgrid_lmax)) {
        ctx->wide = ctx->G > 0xFFFFFFF0ull;
        rc = ctx->wide ? upload_index<uint64_t>(ctx, (cudaStream_t)stream) : upload_index<uint32_t>(ctx, (cudaStream_t)stream);
        if (rc) return rc;
        ctx->index_dirty = false;
    }
    return ESL_OK;
}

int esl_pool_stats(const esl_ctx *ctx, double *mean, double *sd, uint32_t *longest)
{
    if (!ctx || !ctx->have_pool) return ESL_ERR_STATE;
    if (mean) *mean = ctx->pool_mean;
    if (sd) *sd = ctx->pool_sd;
    if (longest) *longest = ctx->pool_max;
    return ESL_OK;
}

int64_t esl_bulk_reads(const esl_ctx *ctx, double coverage) { return ctx ? bulk_reads(ctx, coverage) : 0; }

int64_t esl_workspace_bytes(const esl_ctx *ctx, int32_t n_iter, double coverage, int64_t max_draws_per_iter)
{
    if (!ctx || n_iter < 0) return ESL_ERR_INVALID;
    const i64 cap = max_draws_per_iter > 0 ? max_draws_per_iter : bulk_cap(ctx, coverage);
    return (int64_t)ws_total_bytes(cap, n_iter) + 192;
}

int esl_set_hit_buffer(esl_ctx *ctx, int64_t *d_hits, int64_t capacity, int64_t *d_count)
{
    if (!ctx) return ESL_ERR_INVALID;
    if ((d_hits == nullptr) != (d_count == nullptr) || capacity < 0 || (d_hits && capacity == 0))
        return ctx->fail(ESL_ERR_INVALID, "hit buffer needs both pointers and a positive capacity (or all NULL / 0)");
    ctx->d_hits = d_hits;
    ctx->d_hit_count = d_count;
    ctx->hit_cap = d_hits ? capacity : 0;
    return ESL_OK;
}

int esl_set_profiling(esl_ctx *ctx, int32_t on)
{
    if (!ctx) return ESL_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (on && !ctx->ev[0])
        for (cudaEvent_t &e : ctx->ev) CU(cudaEventCreate(&e));
    ctx->profiling = on != 0;
    ctx->ev_valid = false;
    return ESL_OK;
}

int esl_get_timing(esl_ctx *ctx, float *bulk_ms, float *cutoff_ms)
{
    if (!ctx || !bulk_ms || !cutoff_ms) return ESL_ERR_INVALID;
    if (!ctx->ev_valid) return ctx->fail(ESL_ERR_STATE, "no profiled launch: call esl_set_profiling(ctx, 1) before the run");
    CU(cudaEventSynchronize(ctx->ev[3]));
    CU(cudaEventElapsedTime(bulk_ms, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(cutoff_ms, ctx->ev[2], ctx->ev[3]));
    return ESL_OK;
}

int64_t esl_launch_count(const esl_ctx *ctx) { return ctx ? ctx->launches : 0; }
int32_t esl_target_layers(const esl_ctx *ctx) { return ctx ? ctx->n_layers : 0; }

int32_t esl_plan_target_layers(uint64_t genome_size, const int64_t *start, const int64_t *end, int64_t n, int32_t *layer_out)
{
    if (n < 0 || (n > 0 && (!start || !end || !layer_out)) || genome_size == 0) return ESL_ERR_INVALID;
    if (n > 0x7fffffffLL) return ESL_ERR_UNSUPPORTED; // rows are 31-bit, as in esl_set_targets
    std::vector<Rec> v;
    for (int64_t i = 0; i < n; ++i) {
        layer_out[i] = -1;
        if (end[i] > start[i]) v.push_back(Rec{start[i], end[i], (int32_t)i});
    }
    std::vector<std::vector<Rec>> layers;
    build_layers(v.data(), v.data() + v.size(), layers, genome_size);
    for (size_t l = 0; l < layers.size(); ++l)
        for (const Rec &r : layers[l]) layer_out[r.row] = (int32_t)l;
    return (int32_t)layers.size();
}

int esl_run_batch(esl_ctx *ctx, uint64_t seed, uint32_t combo, uint32_t iter_begin, int32_t n_iter, double coverage,
                  int32_t consuming, int64_t min_overlap, double scaling, int64_t *d_tallies, int64_t *d_label_counts,
                  int64_t *d_n_bases, int64_t *d_n_reads, int32_t *d_status, void *d_ws, size_t ws_bytes, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_iter < 1) return ctx->fail(ESL_ERR_INVALID, "n_iter must be >= 1");
    if (!(coverage >= 0) || !std::isfinite(coverage)) return ctx->fail(ESL_ERR_INVALID, "coverage must be finite and >= 0");
    if (combo >= (1u << 28)) return ctx->fail(ESL_ERR_INVALID, "combo index too large");
    int rc = require_index(ctx, true);
    if (rc) return rc;
    if ((rc = check_outputs(ctx, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status))) return rc;
    if ((u64)ctx->pool_max >= ctx->G) // numpy's randint(0, G - L) raises "low >= high" in the reference (read_operations.py:103)
        return ctx->fail(ESL_ERR_INVALID, "longest read (%u) is not shorter than the genome (%llu)", ctx->pool_max, (unsigned long long)ctx->G);
    const double expect = coverage * (double)ctx->G / ctx->pool_mean;
    if (expect > 3.5e9) return ctx->fail(ESL_ERR_UNSUPPORTED, "more than 2^32 draws per iteration");
    const i64 need = esl_workspace_bytes(ctx, n_iter, coverage, 0);
    if (!d_ws || (i64)ws_bytes < need) return ctx->fail(ESL_ERR_WORKSPACE, "workspace %zu bytes < required %lld", ws_bytes, (long long)need);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = zero_outputs(ctx, n_iter, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status, st))) return rc;
    esl::DevOut out{(u64 *)d_tallies, (u64 *)d_label_counts, (u64 *)d_n_bases, (u64 *)d_n_reads, d_status, (u64 *)d_ws, 0, ctx->d_hits, (u64 *)ctx->d_hit_count, ctx->hit_cap};
    esl::PhiloxSrc src{(uint32_t)seed, (uint32_t)(seed >> 32), combo, iter_begin, philox_keys((uint32_t)seed, (uint32_t)(seed >> 32))};
    const i64 n1 = bulk_reads(ctx, coverage);
    const i64 cap = (ctx->n_mask > 0 && !consuming) ? bulk_cap(ctx, coverage) : n1;
    const Workspace ws = ws_carve(d_ws, std::max(cap, n1), n_iter);
    CU(cudaMemsetAsync(d_ws, 0, ws.tile_bytes, st));
    if (ctx->wide)
        return launch_pipeline<uint64_t>(ctx, make_devctx<uint64_t>(ctx, coverage, consuming, min_overlap, scaling), src, out, n1, cap, ws, true, n_iter, true, st);
    return launch_pipeline<uint32_t>(ctx, make_devctx<uint32_t>(ctx, coverage, consuming, min_overlap, scaling), src, out, n1, cap, ws, true, n_iter, true, st);
}

// ---- staged baseline ---------------------------------------------------------------------------------------------

static i64 staged_cap(const esl_ctx *ctx, double coverage)
{
    const double thr = coverage * (double)ctx->G, mu = ctx->pool_mean, sd = ctx->pool_sd;
    const double r0 = thr / mu;
    i64 cap = (i64)(r0 + 16.0 * sd * std::sqrt(std::max(r0, 1.0)) / mu) + 8192;
    cap = std::max(cap, bulk_cap(ctx, coverage));
    return std::min<i64>(cap, 0x7FFF0000ll);
}

int64_t esl_staged_workspace_bytes(const esl_ctx *ctx, double coverage)
{
    if (!ctx || !ctx->have_pool) return ESL_ERR_STATE;
    const i64 cap = staged_cap(ctx, coverage);
    const i64 sort_blocks = (cap + esl::kSortTile - 1) / esl::kSortTile;
    return (int64_t)(cap * 28 + ((cap + esl::kStageBlock - 1) / esl::kStageBlock) * 8 + sort_blocks * 256 * 4 + 8192);
}

int esl_run_batch_staged(esl_ctx *ctx, uint64_t seed, uint32_t combo, uint32_t iter_begin, int32_t n_iter, double coverage,
                         int32_t consuming, int64_t min_overlap, int64_t *d_tallies, int64_t *d_label_counts,
                         int64_t *d_n_bases, int64_t *d_n_reads, int32_t *d_status, void *d_ws, size_t ws_bytes, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_iter < 1) return ctx->fail(ESL_ERR_INVALID, "n_iter must be >= 1");
    int rc = require_index(ctx, true);
    if (rc) return rc;
    if ((rc = check_outputs(ctx, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status))) return rc;
    if (ctx->wide) return ctx->fail(ESL_ERR_UNSUPPORTED, "the staged baseline supports 32-bit coordinates only");
    if ((u64)ctx->pool_max >= ctx->G) return ctx->fail(ESL_ERR_INVALID, "longest read is not shorter than the genome");
    const i64 need = esl_staged_workspace_bytes(ctx, coverage);
    if (!d_ws || (i64)ws_bytes < need) return ctx->fail(ESL_ERR_WORKSPACE, "workspace %zu bytes < required %lld", ws_bytes, (long long)need);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = zero_outputs(ctx, n_iter, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status, st))) return rc;
    if (ctx->raw_dirty) { // baseline only: the sweep reads the targets in caller order (a set-up step on its first call)
        CU(ctx->d_raw_ts.upload(ctx->t_start.data(), ctx->t_start.size() * sizeof(i64), st));
        CU(ctx->d_raw_te.upload(ctx->t_end.data(), ctx->t_end.size() * sizeof(i64), st));
        CU(ctx->d_raw_tb.upload(ctx->t_bc.data(), ctx->t_bc.size() * sizeof(int32_t), st));
        ctx->raw_dirty = false;
    }
    const int cap = (int)staged_cap(ctx, coverage);
    const int gen_blocks = (cap + esl::kStageBlock - 1) / esl::kStageBlock;
    const int sort_blocks = (cap + esl::kSortTile - 1) / esl::kSortTile;
    uint32_t *w = (uint32_t *)d_ws;
    uint32_t *start = w, *len = w + cap, *kA = w + 3 * (size_t)cap, *vA = w + 4 * (size_t)cap, *kB = w + 5 * (size_t)cap, *vB = w + 6 * (size_t)cap;
    int *label = (int *)(w + 2 * (size_t)cap);
    u64 *block_sum = (u64 *)(w + 7 * (size_t)cap + (((size_t)cap * 7) & 1));
    uint32_t *hist = (uint32_t *)(block_sum + gen_blocks);
    uint32_t *digit_tot = hist + (size_t)sort_blocks * 256;
    int *n_dev = (int *)(digit_tot + 256);
    const esl::DevCtx<uint32_t> dc = make_devctx<uint32_t>(ctx, coverage, consuming, min_overlap, 1.0);
    const i64 T = (i64)ctx->t_start.size();
    const size_t gen_smem = (size_t)ctx->B * 4 + (size_t)esl::kCdfLut * 8 + 24;
    const bool prof = ctx->profiling && ctx->ev[0];
    ctx->ev_valid = false;
    if (prof) CU(cudaEventRecord(ctx->ev[0], st));
    for (int i = 0; i < n_iter; ++i) {
        esl::PhiloxSrc src{(uint32_t)seed, (uint32_t)(seed >> 32), combo, iter_begin + (uint32_t)i, philox_keys((uint32_t)seed, (uint32_t)(seed >> 32))};
        esl::k_stage_generate<<<gen_blocks, esl::kStageBlock, gen_smem, st>>>(dc, src, cap, start, len, label, block_sum);
        esl::k_stage_cutoff<<<1, 1024, 0, st>>>(dc, cap, len, label, block_sum, (u64 *)d_n_reads + i, (u64 *)d_n_bases + i, d_status + i, n_dev);
        esl::k_stage_labels<<<ctx->sm_count * 2, esl::kStageBlock, (size_t)(ctx->B + 2) * 4 + 16, st>>>(ctx->B, label, n_dev, (u64 *)d_label_counts + (i64)i * (ctx->B + 2));
        const uint32_t *kin = start, *vin = nullptr;
        uint32_t *kout = kA, *vout = vA;
        for (int pass = 0; pass < 4; ++pass) {
            esl::k_rs_hist<<<sort_blocks, esl::kStageBlock, 0, st>>>(kin, n_dev, pass * 8, sort_blocks, hist);
            esl::k_rs_scan_rows<<<256, 1024, 0, st>>>(hist, sort_blocks, digit_tot);
            esl::k_rs_scan_digits<<<1, 1024, 0, st>>>(digit_tot);
            esl::k_rs_scatter<<<sort_blocks, esl::kStageBlock, 0, st>>>(kin, vin, kout, vout, n_dev, pass * 8, sort_blocks, hist, digit_tot, pass == 0);
            kin = kout; vin = vout;
            if (kout == kA) { kout = kB; vout = vB; } else { kout = kA; vout = vA; }
        }
        if (T > 0)
            esl::k_stage_sweep<<<(unsigned)((T * 32 + esl::kStageBlock - 1) / esl::kStageBlock), esl::kStageBlock, 0, st>>>(
                ctx->d_raw_ts.as<i64>(), ctx->d_raw_te.as<i64>(), ctx->d_raw_tb.as<int32_t>(), T, kin, vin, n_dev, len, label,
                ctx->pool_max, min_overlap, ctx->B, (u64 *)d_tallies + (i64)i * T * 3);
        ctx->launches += 20;
    }
    CU(cudaGetLastError());
    if (prof) { CU(cudaEventRecord(ctx->ev[1], st)); CU(cudaEventRecord(ctx->ev[2], st)); CU(cudaEventRecord(ctx->ev[3], st)); ctx->ev_valid = true; }
    return ESL_OK;
}

int esl_replay_reads(esl_ctx *ctx, const int64_t *d_start, const int64_t *d_length, const int32_t *d_label,
                     const int64_t *d_seg, int64_t total_reads, int64_t max_reads_per_iter, int32_t n_iter,
                     int64_t min_overlap, int64_t *d_tallies,
                     int64_t *d_label_counts, int64_t *d_n_bases, int64_t *d_n_reads, int32_t *d_status, void *d_ws,
                     size_t ws_bytes, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_iter < 1 || total_reads < 0) return ctx->fail(ESL_ERR_INVALID, "bad n_iter / total_reads");
    if (total_reads && (!d_start || !d_length || !d_label)) return ctx->fail(ESL_ERR_INVALID, "a read array is NULL");
    if (!d_seg) return ctx->fail(ESL_ERR_INVALID, "seg_offsets is NULL");
    int rc = require_index(ctx, false);
    if (rc) return rc;
    if ((rc = check_outputs(ctx, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status))) return rc;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = zero_outputs(ctx, n_iter, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status, st))) return rc;
    // the bulk kernel also records per-tile sums; replay level 1 has no cut-off and never reads them (no zeroing needed)
    if (total_reads >= 0xFFFF0000ll) return ctx->fail(ESL_ERR_UNSUPPORTED, "more than 2^32 reads in one replay call");
    if (max_reads_per_iter <= 0 || max_reads_per_iter > total_reads) max_reads_per_iter = total_reads;
    const i64 need = esl_workspace_bytes(ctx, n_iter, 0.0, std::max<i64>(max_reads_per_iter, 1));
    if (!d_ws || (i64)ws_bytes < need) return ctx->fail(ESL_ERR_WORKSPACE, "workspace %zu bytes < required %lld", ws_bytes, (long long)need);
    esl::DevOut out{(u64 *)d_tallies, (u64 *)d_label_counts, (u64 *)d_n_bases, (u64 *)d_n_reads, d_status, (u64 *)d_ws, 0, ctx->d_hits, (u64 *)ctx->d_hit_count, ctx->hit_cap};
    esl::ReadSrc src{d_start, d_length, d_label, d_seg};
    const i64 big = max_reads_per_iter; // every iteration's reads are all "bulk" (clipped to its own count on device)
    const Workspace ws = ws_carve(d_ws, std::max<i64>(big, 1), n_iter);
    if (ctx->wide)
        return launch_pipeline<uint64_t>(ctx, make_devctx<uint64_t>(ctx, 0.0, 0, min_overlap, 1.0), src, out, big, big, ws, false, n_iter, false, st);
    return launch_pipeline<uint32_t>(ctx, make_devctx<uint32_t>(ctx, 0.0, 0, min_overlap, 1.0), src, out, big, big, ws, false, n_iter, false, st);
}

int esl_replay_draws(esl_ctx *ctx, const double *d_ubc, const int64_t *d_length, const int64_t *d_start,
                     const int64_t *d_uoff, const double *d_ublock, const int64_t *d_seg, int64_t max_draws,
                     int64_t bulk_draws, int32_t n_iter, double coverage, int32_t consuming, int64_t min_overlap,
                     int64_t *d_tallies, int64_t *d_label_counts, int64_t *d_n_bases, int64_t *d_n_reads,
                     int32_t *d_status, void *d_ws, size_t ws_bytes, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_iter < 1 || max_draws < 0) return ctx->fail(ESL_ERR_INVALID, "bad n_iter / max_draws_per_iter");
    if (!d_ubc || !d_length || !d_start || !d_seg) return ctx->fail(ESL_ERR_INVALID, "a draw array is NULL");
    int rc = require_index(ctx, false);
    if (rc) return rc;
    if (ctx->n_mask > 0 && (!d_uoff || !d_ublock)) return ctx->fail(ESL_ERR_INVALID, "a mask is set but the blocked-draw arrays are NULL");
    if ((rc = check_outputs(ctx, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status))) return rc;
    if (bulk_draws < 0) bulk_draws = 0;
    if (bulk_draws > max_draws) bulk_draws = max_draws;
    bulk_draws -= bulk_draws % esl::kBulkTile;
    const i64 need = esl_workspace_bytes(ctx, n_iter, coverage, std::max<i64>(bulk_draws, 1));
    if (!d_ws || (i64)ws_bytes < need) return ctx->fail(ESL_ERR_WORKSPACE, "workspace %zu bytes < required %lld", ws_bytes, (long long)need);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = zero_outputs(ctx, n_iter, d_tallies, d_label_counts, d_n_bases, d_n_reads, d_status, st))) return rc;
    if (max_draws >= 0xFFFF0000ll) return ctx->fail(ESL_ERR_UNSUPPORTED, "more than 2^32 draws per iteration");
    const Workspace ws = ws_carve(d_ws, std::max<i64>(bulk_draws, 1), n_iter);
    CU(cudaMemsetAsync(d_ws, 0, ws.tile_bytes, st));
    esl::DevOut out{(u64 *)d_tallies, (u64 *)d_label_counts, (u64 *)d_n_bases, (u64 *)d_n_reads, d_status, (u64 *)d_ws, 0, ctx->d_hits, (u64 *)ctx->d_hit_count, ctx->hit_cap};
    esl::DrawSrc src{d_ubc, d_length, d_start, d_uoff, d_ublock, d_seg};
    if (ctx->wide)
        return launch_pipeline<uint64_t>(ctx, make_devctx<uint64_t>(ctx, coverage, consuming, min_overlap, 1.0), src, out, bulk_draws, bulk_draws, ws, false, n_iter, true, st);
    return launch_pipeline<uint32_t>(ctx, make_devctx<uint32_t>(ctx, coverage, consuming, min_overlap, 1.0), src, out, bulk_draws, bulk_draws, ws, false, n_iter, true, st);
}

int esl_materialize_reads(esl_ctx *ctx, uint64_t seed, uint32_t combo, uint32_t iteration, int32_t consuming,
                          int64_t n_reads, int64_t *d_start, int64_t *d_length, int32_t *d_label, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_reads < 0 || (n_reads && (!d_start || !d_length || !d_label))) return ctx->fail(ESL_ERR_INVALID, "bad output arrays");
    int rc = require_index(ctx, true);
    if (rc) return rc;
    if (n_reads == 0) return ESL_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    esl::PhiloxSrc src{(uint32_t)seed, (uint32_t)(seed >> 32), combo, iteration, philox_keys((uint32_t)seed, (uint32_t)(seed >> 32))};
    const unsigned grid = (unsigned)std::min<i64>((n_reads + 255) / 256, (i64)ctx->sm_count * 8);
    const size_t sm = (size_t)ctx->B * 4 + (size_t)esl::kCdfLut * 8 + 24;
    if (ctx->wide)
        esl::k_materialize<uint64_t><<<grid, 256, sm, st>>>(make_devctx<uint64_t>(ctx, 0.0, consuming, 1, 1.0), src, n_reads, d_start, d_length, d_label);
    else
        esl::k_materialize<uint32_t><<<grid, 256, sm, st>>>(make_devctx<uint32_t>(ctx, 0.0, consuming, 1, 1.0), src, n_reads, d_start, d_length, d_label);
    CU(cudaGetLastError());
    ctx->launches++;
    return ESL_OK;
}

int esl_coverage_bins(esl_ctx *ctx, const int64_t *d_start, const int64_t *d_length, const int32_t *d_label,
                      int64_t n_reads, int64_t bin_size, int64_t n_bins, int64_t *d_bins, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_reads < 0 || bin_size < 1 || n_bins < 1 || !d_bins || (n_reads && (!d_start || !d_length || !d_label)))
        return ctx->fail(ESL_ERR_INVALID, "bad arguments");
    if (ctx->B <= 0) return ctx->fail(ESL_ERR_STATE, "esl_set_barcode_cdf has not been called");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemsetAsync(d_bins, 0, (size_t)(ctx->B + 2) * (size_t)n_bins * sizeof(int64_t), st));
    if (n_reads == 0) return ESL_OK;
    const unsigned grid = (unsigned)std::min<i64>((n_reads + 255) / 256, (i64)ctx->sm_count * 8);
    esl::k_coverage_bins<<<grid, 256, 0, st>>>(d_start, d_length, d_label, n_reads, bin_size, n_bins, ctx->B, (u64 *)d_bins);
    CU(cudaGetLastError());
    ctx->launches++;
    return ESL_OK;
}

int esl_replay_thin(esl_ctx *ctx, const int64_t *d_hits, const double *d_u, int64_t n_hits, double scaling,
                    int32_t n_iter, int64_t *d_tallies, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (n_hits < 0 || n_iter < 1 || !d_tallies || (n_hits && (!d_hits || !d_u))) return ctx->fail(ESL_ERR_INVALID, "bad arguments");
    if (!ctx->have_targets) return ctx->fail(ESL_ERR_STATE, "esl_set_targets has not been called");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const i64 T = (i64)ctx->t_start.size();
    CU(cudaMemsetAsync(d_tallies, 0, (size_t)n_iter * T * 3 * sizeof(int64_t), st));
    if (n_hits == 0) return ESL_OK;
    esl::k_thin_pairs<<<(unsigned)((n_hits + 255) / 256), 256, 0, st>>>(d_hits, d_u, n_hits, scaling, T, (u64 *)d_tallies);
    CU(cudaGetLastError());
    ctx->launches++;
    return ESL_OK;
}

int esl_moments(esl_ctx *ctx, const int64_t *d_tallies, int32_t n_iter, int64_t *d_moments, void *stream)
{
    if (!ctx) return ESL_ERR_INVALID;
    if (!d_tallies || !d_moments || n_iter < 1) return ctx->fail(ESL_ERR_INVALID, "bad arguments");
    if (!ctx->have_targets) return ctx->fail(ESL_ERR_STATE, "esl_set_targets has not been called");
    const i64 T = (i64)ctx->t_start.size();
    if (T == 0) return ESL_OK;
    CU(cudaSetDevice(ctx->device));
    // enough threads to fill the GPU: iterations are split over grid.y when there are few targets
    const int n_slices = (int)std::max<i64>(1, std::min<i64>({(i64)n_iter, (i64)65536 / std::max<i64>(T, 1), (i64)65535}));
    esl::k_moments<<<dim3((unsigned)((T + 255) / 256), (unsigned)n_slices), 256, 0, (cudaStream_t)stream>>>(d_tallies, n_iter, T, n_slices, d_moments);
    CU(cudaGetLastError());
    ctx->launches++;
    return ESL_OK;
}

} // extern "C"
