// Direct writer for the rows of `{exp}_matches_table.csv` (host side) -- SURVEY.md section 8(f) row 1.
// The reference builds one dict per (iteration, target), splits strings per row with pandas and calls to_csv
// (esloco.py:155-216): config 3 has 2.4e8 such rows.  Here the per-target text that does not change between
// iterations (key prefix; mean_read_length / coverage / target / barcode / insertion columns) is prepared once by the
// Python host and this function only formats the three integers of each row next to it, straight from the tally
// array, into a large buffer.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

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

} // namespace

extern "C" int64_t esl_write_match_rows(const char *path, int32_t append, const char *pre_blob, const int64_t *pre_off,
                                        const char *mid_blob, const int64_t *mid_off, const char *overlap_field,
                                        const int64_t *tallies, int64_t n_targets, const char *iteration)
{
    if (!path || !pre_off || !mid_off || !overlap_field || !iteration || n_targets < 0) return ESL_ERR_INVALID;
    if (n_targets > 0 && (!pre_blob || !mid_blob || !tallies)) return ESL_ERR_INVALID;
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return ESL_ERR_INVALID;
    const size_t it_len = strlen(iteration), ov_len = strlen(overlap_field);
    std::vector<char> buf;
    buf.resize((size_t)8 << 20);
    size_t used = 0;
    int64_t written = 0;
    for (int64_t k = 0; k < n_targets; ++k) {
        const size_t pre_len = (size_t)(pre_off[k + 1] - pre_off[k]), mid_len = (size_t)(mid_off[k + 1] - mid_off[k]);
        const size_t need = pre_len + mid_len + 2 * it_len + ov_len + 3 * 21 + 8;
        if (need > buf.size()) { fclose(f); return ESL_ERR_UNSUPPORTED; } // a single row longer than the buffer
        if (used + need > buf.size()) {
            if (fwrite(buf.data(), 1, used, f) != used) { fclose(f); return ESL_ERR_INVALID; }
            written += (int64_t)used;
            used = 0;
        }
        char *p = buf.data() + used;
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
        used = (size_t)(p - buf.data());
    }
    if (used && fwrite(buf.data(), 1, used, f) != used) { fclose(f); return ESL_ERR_INVALID; }
    written += (int64_t)used;
    if (fclose(f) != 0) return ESL_ERR_INVALID;
    return written;
}
