// Direct writer for the rows of `{exp}_matches_table.csv` (host side) -- SURVEY.md section 8(f) row 1.
// The reference builds one dict per (iteration, target), splits strings per row with pandas and calls to_csv
// (esloco.py:155-216): config 3 has 2.4e8 such rows.  Here the per-target text that does not change between
// iterations (key prefix; mean_read_length / coverage / target / barcode / insertion columns) is prepared once by the
// Python host and this function only formats the three integers of each row next to it, straight from the tally
// array, into a large buffer.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>

#include "../../include/esloco_b200.h"

namespace {

inline char *put_u64(char *p, uint64_t v)
{
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char *put_i64(char *p, int64_t v)
{
    if (v < 0) { *p++ = '-'; return put_u64(p, (uint64_t)0 - (uint64_t)v); }
    return put_u64(p, (uint64_t)v);
}

// Buffered row sink: one FILE*, one uninitialised buffer no larger than the text it will hold.
struct RowSink {
    FILE *f = nullptr;
    std::unique_ptr<char[]> buf;
    size_t cap = 0, used = 0;
    int64_t written = 0;
    bool open(const char *path, int append, size_t want)
    {
        f = fopen(path, append ? "ab" : "wb");
        if (!f) return false;
        cap = std::min<size_t>((size_t)8 << 20, std::max<size_t>(want, 4096));
        buf.reset(new char[cap]);
        return true;
    }
    bool flush()
    {
        if (used && fwrite(buf.get(), 1, used, f) != used) return false;
        written += (int64_t)used;
        used = 0;
        return true;
    }
    int64_t close(bool ok)
    {
        if (f) { ok = flush() && ok; ok = (fclose(f) == 0) && ok; f = nullptr; }
        return ok ? written : (int64_t)ESL_ERR_INVALID;
    }
};

// Rows of one (iteration, combination) block; returns 0 or a negative esl_status.
int write_block(RowSink &o, const char *pre_blob, const int64_t *pre_off, const char *mid_blob, const int64_t *mid_off,
                const char *overlap_field, size_t ov_len, const int64_t *tallies, int64_t n_targets, const char *iteration,
                size_t it_len)
{
    for (int64_t k = 0; k < n_targets; ++k) {
        const size_t pre_len = (size_t)(pre_off[k + 1] - pre_off[k]), mid_len = (size_t)(mid_off[k + 1] - mid_off[k]);
        const size_t need = pre_len + mid_len + 2 * it_len + ov_len + 3 * 21 + 8;
        if (need > o.cap) return ESL_ERR_UNSUPPORTED; // a single row longer than the buffer
        if (o.used + need > o.cap && !o.flush()) return ESL_ERR_INVALID;
        char *p = o.buf.get() + o.used;
        // target_region = "<key>_<iteration>"
        memcpy(p, pre_blob + pre_off[k], pre_len); p += pre_len;
        memcpy(p, iteration, it_len); p += it_len;
        *p++ = '\t'; p = put_i64(p, tallies[3 * k]);         // full_matches
        *p++ = '\t'; p = put_i64(p, tallies[3 * k + 1]);     // partial_matches
        *p++ = '\t'; memcpy(p, overlap_field, ov_len); p += ov_len; // overlap (the empty list unless lists were asked for)
        *p++ = '\t'; p = put_i64(p, tallies[3 * k + 2]);     // on_target_bases
        memcpy(p, mid_blob + mid_off[k], mid_len); p += mid_len; // "\t<mrl>\t<cov>\t<target>\t<barcode>[\t<insertion>]\t"
        memcpy(p, iteration, it_len); p += it_len;
        *p++ = '\n';
        o.used = (size_t)(p - o.buf.get());
    }
    return ESL_OK;
}

// Upper bound of the text of one block (sizes the buffer for small tables: 100 targets need a few KB, not 8 MB).
size_t block_bytes(const int64_t *pre_off, const int64_t *mid_off, int64_t n_targets, size_t ov_len)
{
    return (size_t)(pre_off[n_targets] - pre_off[0]) + (size_t)(mid_off[n_targets] - mid_off[0]) +
           (size_t)n_targets * (2 * 20 + ov_len + 3 * 21 + 8);
}

} // namespace

extern "C" int64_t esl_write_match_rows(const char *path, int32_t append, const char *pre_blob, const int64_t *pre_off,
                                        const char *mid_blob, const int64_t *mid_off, const char *overlap_field,
                                        const int64_t *tallies, int64_t n_targets, const char *iteration)
{
    if (!path || !pre_off || !mid_off || !overlap_field || !iteration || n_targets < 0) return ESL_ERR_INVALID;
    if (n_targets > 0 && (!pre_blob || !mid_blob || !tallies)) return ESL_ERR_INVALID;
    const size_t ov_len = strlen(overlap_field), it_len = strlen(iteration);
    RowSink o;
    if (!o.open(path, append, block_bytes(pre_off, mid_off, n_targets, ov_len) + 2 * it_len * (size_t)n_targets)) return ESL_ERR_INVALID;
    const int rc = write_block(o, pre_blob, pre_off, mid_blob, mid_off, overlap_field, ov_len, tallies, n_targets, iteration, it_len);
    const int64_t w = o.close(rc == ESL_OK);
    return rc != ESL_OK ? rc : w;
}

extern "C" int64_t esl_write_match_rows_batch(const char *path, int32_t append, const char *pre_blob, const int64_t *pre_off,
                                              int32_t n_combos, const char *const *mid_blobs, const int64_t *const *mid_offs,
                                              const char *overlap_field, const int64_t *const *tallies, int64_t n_targets,
                                              int64_t n_iter, const int64_t *iteration_ids)
{
    if (!path || !pre_off || !overlap_field || n_targets < 0 || n_iter < 0 || n_combos < 0) return ESL_ERR_INVALID;
    if (n_combos > 0 && (!mid_blobs || !mid_offs || !tallies)) return ESL_ERR_INVALID;
    if (n_iter > 0 && !iteration_ids) return ESL_ERR_INVALID;
    for (int32_t c = 0; c < n_combos; ++c)
        if (!mid_offs[c] || (n_targets > 0 && (!pre_blob || !mid_blobs[c] || !tallies[c]))) return ESL_ERR_INVALID;
    const size_t ov_len = strlen(overlap_field);
    size_t want = 0;
    for (int32_t c = 0; c < n_combos; ++c) want += block_bytes(pre_off, mid_offs[c], n_targets, ov_len);
    RowSink o;
    if (!o.open(path, append, want * (size_t)std::min<int64_t>(std::max<int64_t>(n_iter, 1), 4096))) return ESL_ERR_INVALID;
    int rc = ESL_OK;
    char it_text[24];
    for (int64_t i = 0; i < n_iter && rc == ESL_OK; ++i) { // rows: iteration -> combination -> target (esloco.py:155-160)
        char *e = put_i64(it_text, iteration_ids[i]);
        *e = 0;
        for (int32_t c = 0; c < n_combos && rc == ESL_OK; ++c)
            rc = write_block(o, pre_blob, pre_off, mid_blobs[c], mid_offs[c], overlap_field, ov_len,
                             tallies[c] + (size_t)i * (size_t)n_targets * 3, n_targets, it_text, (size_t)(e - it_text));
    }
    const int64_t w = o.close(rc == ESL_OK);
    return rc != ESL_OK ? rc : w;
}
