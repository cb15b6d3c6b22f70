/*
 * esloco_oracle.c -- CPU restatement of esloco's read-placement / target-tally path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the timed CPU baseline.
 * The product path (esloco_b200/) never imports it and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference (aweich/esloco v1.0.6, pure Python) has
 * no tests or golden vectors of its own, so tests/golden/make_golden.py runs the
 * unmodified reference in the build container and commits its outputs; this
 * restatement reproduces every fixture bit for bit FROM THE SEED ALONE
 * (tests/test_oracle_golden.py), i.e. including the numpy legacy RandomState
 * stream the reference draws from.
 *
 * The arithmetic lives partly in a third-party dependency that is not vendored
 * in /root/reference: numpy's legacy global RandomState (pyproject.toml:19,
 * unpinned; container has numpy 2.3.5).  Its published algorithm is restated
 * here (MT19937, 53-bit doubles, masked-rejection bounded integers, polar
 * Box-Muller gauss) and checked against numpy itself in
 * tests/test_oracle_rng.py.  `intervaltree` (pyproject.toml:29, absent) is
 * restated as a scan with the package's half-open overlap rule.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/src/esloco/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MT_N 624
#define MT_M 397

typedef struct {
    uint32_t key[MT_N];
    int pos;
    int has_gauss;
    double gauss;
} eslo_rng;

size_t eslo_rng_sizeof(void) { return sizeof(eslo_rng); }

/* np.random.seed(int) -> numpy legacy mt19937_seed (init_genrand).  Call site: esloco.py:100. */
void eslo_seed(eslo_rng *st, uint32_t seed)
{
    for (int pos = 0; pos < MT_N; pos++) {
        st->key[pos] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
    }
    st->pos = MT_N;
    st->has_gauss = 0;
    st->gauss = 0.0;
}

static void mt_refill(eslo_rng *st)
{
    uint32_t *k = st->key;
    int i;
    uint32_t y;
    for (i = 0; i < MT_N - MT_M; i++) {
        y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
        k[i] = k[i + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; i < MT_N - 1; i++) {
        y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu);
        k[i] = k[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (k[MT_N - 1] & 0x80000000u) | (k[0] & 0x7fffffffu);
    k[MT_N - 1] = k[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    st->pos = 0;
}

uint32_t eslo_u32(eslo_rng *st)
{
    if (st->pos == MT_N) mt_refill(st);
    uint32_t y = st->key[st->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* np.random.rand() / random_sample(): 53-bit double, high part drawn first. */
double eslo_double(eslo_rng *st)
{
    uint32_t a = eslo_u32(st) >> 5, b = eslo_u32(st) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

/* numpy legacy bounded integer in [0, rng] by masked rejection; 64-bit words are (hi<<32)|lo.
 * Backs np.random.randint(0, hi) (read_operations.py:103) and np.random.choice(int_array)
 * (read_operations.py:102, combined_calculations.py:57). */
uint64_t eslo_bounded(eslo_rng *st, uint64_t rng)
{
    if (rng == 0) return 0;
    if (rng <= 0xFFFFFFFFull) {
        if (rng == 0xFFFFFFFFull) return eslo_u32(st);
        uint32_t mask = (uint32_t)rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = eslo_u32(st) & mask; } while (v > (uint32_t)rng);
        return v;
    }
    uint64_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    if (rng == 0xFFFFFFFFFFFFFFFFull) {
        uint64_t hi = eslo_u32(st);
        return (hi << 32) | eslo_u32(st);
    }
    uint64_t v;
    do {
        uint64_t hi = eslo_u32(st);
        v = ((hi << 32) | eslo_u32(st)) & mask;
    } while (v > rng);
    return v;
}

/* numpy legacy_gauss: polar Box-Muller with one cached value. */
double eslo_gauss(eslo_rng *st)
{
    if (st->has_gauss) {
        double t = st->gauss;
        st->has_gauss = 0;
        st->gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * eslo_double(st) - 1.0;
        x2 = 2.0 * eslo_double(st) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    st->gauss = f * x1;
    st->has_gauss = 1;
    return f * x2;
}

/* generate_read_length_distribution(..., 'lognormal'), read_operations.py:8-27:
 * round(lognormal(ln(mean) - 0.5, sigma, n)).astype(int), keep values > min_read_length.
 * Returns how many were kept (written to out[0..kept)). */
int64_t eslo_lognormal_pool(eslo_rng *st, int64_t n, double mean_read_length, double sigma,
                            int64_t min_read_length, int64_t *out)
{
    const double mu = log(mean_read_length) - 0.5;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; i++) {
        double x = exp(mu + sigma * eslo_gauss(st));
        int64_t v = (int64_t)rint(x);
        if (v > min_read_length) out[kept++] = v;
    }
    return kept;
}

/* np.random.choice(src, size=n) (combined_calculations.py:57): n successive bounded(len-1). */
void eslo_choice(eslo_rng *st, const int64_t *src, int64_t nsrc, int64_t n, int64_t *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = src[eslo_bounded(st, (uint64_t)(nsrc - 1))];
}

/* check_if_blocked_region (read_operations.py:59-67) over build_mask_tree's trees (utils.py:89-101).
 * Intervals are flat arrays with barcode >= 0 (that barcode's tree) or -1 (the "ALL" tree); the tree
 * of the read's barcode is visited first, then "ALL".  Within a tree hits are visited in array order
 * (callers pass them sorted by (start, end): the canonical order, the package's set order being
 * unspecified).  One rand() per visited hit; u < weight blocks.  Draws are appended to ublock. */
static int blocked_test(eslo_rng *st, int barcode, int64_t rs, int64_t re, int64_t K, const int64_t *ms,
                        const int64_t *me, const double *mw, const int32_t *mb, double *ublock,
                        int64_t *n_ublock, int64_t ublock_cap)
{
    for (int pass = 0; pass < 2; pass++) {
        int want = pass == 0 ? barcode : -1;
        for (int64_t k = 0; k < K; k++) {
            if (mb[k] != want) continue;
            if (ms[k] < re && me[k] > rs) {
                double u = eslo_double(st);
                if (ublock && *n_ublock < ublock_cap) ublock[*n_ublock] = u;
                (*n_ublock)++;
                if (u < mw[k]) return 1;
            }
        }
    }
    return 0;
}

/* generate_reads_based_on_coverage (read_operations.py:86-123).
 *   cdf      : cumsum(p)/cumsum(p)[-1] exactly as numpy's choice(p=) builds it (read_operations.py:96-97,101)
 *   pool     : precomputed_lengths (combined_calculations.py:57)
 *   labels   : barcode k >= 0, -1 = "Blocked/Consuming", -2 = "Blocked/NotConsuming"
 * Records every draw up to and including the cut-off read (n_placed), then keeps drawing n_extra more
 * reads from the same stream (an oversampled tail for replay tests; NOT part of the reference's result
 * and not added to covered; it stops early if a picked length is not shorter than the genome).
 * Optional outputs may be NULL.  ublock_off has cap+1 entries.
 * Returns 0, -1 if cap is too small, -2 if a read is not shorter than the genome (numpy raises there). */
int eslo_place(eslo_rng *st, int64_t G, double coverage, const int64_t *pool, int64_t npool, int n_barcodes,
               const double *cdf, int consuming, int64_t K, const int64_t *ms, const int64_t *me,
               const double *mw, const int32_t *mb, int64_t cap, int64_t n_extra, int64_t *start,
               int64_t *length, int32_t *label, double *u_barcode, int64_t *ublock_off, double *ublock,
               int64_t ublock_cap, int64_t *covered_out, int64_t *n_placed_out, int64_t *n_ublock_out,
               int64_t *n_total_out)
{
    const double thr = coverage * (double)G; /* `covered_length < coverage * genome_size`, :100 */
    int64_t covered = 0, n = 0, n_ublock = 0, placed = -1;
    for (;;) {
        if (placed < 0 && !((double)covered < thr)) placed = n;
        if (placed >= 0 && n >= placed + n_extra) break;
        if (n >= cap) return -1;
        double u = eslo_double(st); /* :101 */
        int k = 0;
        while (k < n_barcodes && cdf[k] <= u) k++; /* searchsorted(cdf, u, side='right') */
        if (k >= n_barcodes) k = n_barcodes - 1;
        int64_t L = pool[eslo_bounded(st, (uint64_t)(npool - 1))]; /* :102 */
        if (G - L <= 0) {
            if (placed >= 0) break; /* oversampled tail only: stop early instead of raising */
            return -2;
        }
        int64_t s = (int64_t)eslo_bounded(st, (uint64_t)(G - L - 1)); /* :103 */
        if (ublock_off) ublock_off[n] = n_ublock;
        int blocked = K > 0 ? blocked_test(st, k, s, s + L, K, ms, me, mw, mb, ublock, &n_ublock, ublock_cap) : 0;
        int32_t lab;
        int64_t counted;
        if (blocked) { /* :108-117 */
            lab = consuming ? -1 : -2;
            counted = consuming ? L : 0;
        } else { /* :119-121 */
            lab = k;
            counted = L;
        }
        if (start) start[n] = s;
        if (length) length[n] = L;
        if (label) label[n] = lab;
        if (u_barcode) u_barcode[n] = u;
        if (placed < 0) covered += counted;
        n++;
    }
    if (ublock_off) ublock_off[n] = n_ublock;
    *covered_out = covered;
    *n_placed_out = placed;
    if (n_ublock_out) *n_ublock_out = n_ublock;
    if (n_total_out) *n_total_out = n; /* placed + extras actually drawn */
    return 0;
}

/* count_matches (counting.py:7-61): target-major, reads in draw order.  A read matches a target iff the
 * read's label is the target's barcode (:37; blocked labels never match), the half-open intervals
 * intersect (:40) and the overlap is >= min_overlap (:43).  ONE rand() is drawn per qualifying pair, even
 * when scaling == 1 (:46,51); `u <= scaling` keeps it.  full iff read_start <= start and end <= read_end.
 * ov_off (T+1) / ov (cap ov_cap) receive the per-target overlap lists in read order when non-NULL; upair
 * (cap upair_cap) the uniform of every qualifying pair in visiting order, kept or not (*n_pairs_out = how many).
 * st == NULL skips the draw (valid only for scaling >= 1: the outcome does not depend on u). */
int64_t eslo_count_matches(eslo_rng *st, int64_t T, const int64_t *ts, const int64_t *te, const int32_t *tb,
                           int64_t R, const int64_t *rs, const int64_t *re, const int32_t *rl, double scaling,
                           int64_t min_overlap, int64_t *full, int64_t *partial, int64_t *otb, int64_t *ov_off,
                           int64_t *ov, int64_t ov_cap, double *upair, int64_t upair_cap, int64_t *n_pairs_out)
{
    int64_t nov = 0, npairs = 0;
    for (int64_t t = 0; t < T; t++) {
        int64_t f = 0, p = 0, sum = 0;
        const int64_t s = ts[t], e = te[t];
        const int32_t b = tb[t];
        if (ov_off) ov_off[t] = nov;
        if (b >= 0) {
            for (int64_t r = 0; r < R; r++) {
                if (rl[r] != b) continue;
                int64_t lo = s > rs[r] ? s : rs[r], hi = e < re[r] ? e : re[r];
                if (lo < hi) {
                    int64_t o = hi - lo;
                    if (o >= min_overlap) {
                        double u = st ? eslo_double(st) : 0.0;
                        if (upair && npairs < upair_cap) upair[npairs] = u; /* the draw of this qualifying pair */
                        npairs++;
                        if (u <= scaling) {
                            if (rs[r] <= s && e <= re[r]) f++; else p++;
                            sum += o;
                            if (ov && nov < ov_cap) ov[nov] = o;
                            nov++;
                        }
                    }
                }
            }
        }
        full[t] = f;
        partial[t] = p;
        otb[t] = sum;
    }
    if (ov_off) ov_off[T] = nov;
    if (n_pairs_out) *n_pairs_out = npairs;
    return nov;
}

/* Indexed tally for sizes where the reference's O(T*R) loop cannot run (configs 2-4).  Same outcome as
 * eslo_count_matches with scaling >= 1 (checked against it in tests/test_oracle_golden.py): targets are
 * grouped per barcode and sorted by start; a read probes the window of targets whose start lies in
 * (read_start - maxlen_b, read_end) and applies the identical pair test.  Checker accelerator only. */
typedef struct { int64_t s, e; int64_t id; int32_t b; } tgt_rec;
static int tgt_cmp(const void *a, const void *b)
{
    const tgt_rec *x = a, *y = b;
    if (x->b != y->b) return x->b < y->b ? -1 : 1;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}
int eslo_count_matches_indexed(int64_t T, const int64_t *ts, const int64_t *te, const int32_t *tb, int n_barcodes,
                               int64_t R, const int64_t *rs, const int64_t *re, const int32_t *rl,
                               int64_t min_overlap, int64_t *full, int64_t *partial, int64_t *otb)
{
    tgt_rec *v = malloc(sizeof(tgt_rec) * (size_t)(T > 0 ? T : 1));
    int64_t *boff = calloc((size_t)n_barcodes + 1, sizeof(int64_t));
    int64_t *maxlen = calloc((size_t)n_barcodes + 1, sizeof(int64_t));
    if (!v || !boff || !maxlen) { free(v); free(boff); free(maxlen); return -1; }
    int64_t m = 0;
    for (int64_t t = 0; t < T; t++) {
        full[t] = partial[t] = otb[t] = 0;
        if (tb[t] < 0 || tb[t] >= n_barcodes) continue;
        v[m].s = ts[t]; v[m].e = te[t]; v[m].id = t; v[m].b = tb[t]; m++;
        if (te[t] - ts[t] > maxlen[tb[t]]) maxlen[tb[t]] = te[t] - ts[t];
    }
    qsort(v, (size_t)m, sizeof(tgt_rec), tgt_cmp);
    for (int64_t i = 0; i < m; i++) boff[v[i].b + 1]++;
    for (int b = 0; b < n_barcodes; b++) boff[b + 1] += boff[b];
    for (int64_t r = 0; r < R; r++) {
        int32_t b = rl[r];
        if (b < 0 || b >= n_barcodes) continue;
        int64_t lo = boff[b], hi = boff[b + 1];
        const int64_t key = rs[r] - maxlen[b]; /* candidates have start > key */
        int64_t a = lo, z = hi;
        while (a < z) { int64_t mid = a + (z - a) / 2; if (v[mid].s <= key) a = mid + 1; else z = mid; }
        for (int64_t i = a; i < hi && v[i].s < re[r]; i++) {
            int64_t s = v[i].s, e = v[i].e;
            int64_t l2 = s > rs[r] ? s : rs[r], h2 = e < re[r] ? e : re[r];
            if (l2 < h2 && h2 - l2 >= min_overlap) {
                if (rs[r] <= s && e <= re[r]) full[v[i].id]++; else partial[v[i].id]++;
                otb[v[i].id] += h2 - l2;
            }
        }
    }
    free(v); free(boff); free(maxlen);
    return 0;
}

/* count_barcode_occurrences (counting.py:63-71): histogram of read labels; slots 0..B-1 are barcodes,
 * B = "Blocked/Consuming", B+1 = "Blocked/NotConsuming".  first_seen[slot] = index of the first read
 * carrying that label (-1 if none) so the caller can rebuild the dict's first-appearance key order. */
void eslo_label_hist(int64_t R, const int32_t *rl, int n_barcodes, int64_t *counts, int64_t *first_seen)
{
    for (int i = 0; i < n_barcodes + 2; i++) { counts[i] = 0; first_seen[i] = -1; }
    for (int64_t r = 0; r < R; r++) {
        int slot = rl[r] >= 0 ? rl[r] : (rl[r] == -1 ? n_barcodes : n_barcodes + 1);
        if (counts[slot]++ == 0) first_seen[slot] = r;
    }
}
