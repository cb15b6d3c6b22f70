// Streaming FASTA / FASTQ read-length scanner (host side of the length table) -- SURVEY.md section 8(f) row 3.
// Replaces the per-iteration Bio.SeqIO parse of the reference (config_handler.py:16-51, re-run from
// combined_calculations.py:37-39): one pass over the file with a 4 MB buffer, plain or gzip, only lengths kept.
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/esloco_b200.h"

namespace {

struct Reader { // buffered line reader over gzFile (zlib reads plain files transparently)
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    bool open(const char *path)
    {
        f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(4 << 20);
        return true;
    }
    ~Reader() { if (f) gzclose(f); }
    bool fill()
    {
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        pos = 0;
        len = n > 0 ? (size_t)n : 0;
        return len > 0;
    }
    // Reads one line; returns false at EOF with nothing read.  *first = first character (0 for an empty line),
    // *n = number of characters excluding "\r\n"; trailing/leading blanks of sequence lines are not stripped
    // (sequence files do not carry them).
    bool line(char *first, uint64_t *n)
    {
        uint64_t count = 0;
        char f0 = 0;
        bool any = false;
        for (;;) {
            if (pos == len && !fill()) break;
            any = true;
            const char *p = buf.data() + pos;
            const char *nl = (const char *)memchr(p, '\n', len - pos);
            const size_t take = nl ? (size_t)(nl - p) : len - pos;
            if (count == 0 && take > 0) f0 = p[0];
            count += take;
            pos += take + (nl ? 1 : 0);
            if (nl) {
                if (take > 0 && p[take - 1] == '\r') --count;
                break;
            }
        }
        *first = f0;
        *n = count;
        return any;
    }
};

} // namespace

extern "C" int64_t esl_scan_read_lengths(const char *path, int32_t is_fastq, int64_t min_read_length, uint32_t *out,
                                         int64_t capacity, double *mean_out)
{
    if (!path) return ESL_ERR_INVALID;
    Reader r;
    if (!r.open(path)) return ESL_ERR_INVALID;
    int64_t kept = 0;
    long double sum = 0;
    char c;
    uint64_t n;
    auto emit = [&](uint64_t len) {
        if ((int64_t)len > min_read_length) {
            if (out && kept < capacity) out[kept] = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)len;
            ++kept;
            sum += (long double)len;
        }
    };
    if (is_fastq) {
        for (;;) {
            if (!r.line(&c, &n)) break;          // @header
            if (n == 0) continue;                // tolerate blank lines between records
            if (c != '@') return ESL_ERR_UNSUPPORTED;
            uint64_t seq;
            if (!r.line(&c, &seq)) return ESL_ERR_UNSUPPORTED;
            if (!r.line(&c, &n) || c != '+') return ESL_ERR_UNSUPPORTED;
            uint64_t qual;
            if (!r.line(&c, &qual)) return ESL_ERR_UNSUPPORTED;
            if (seq > 0 && qual == 0) return ESL_ERR_UNSUPPORTED; // no quality string: not FASTQ (config_handler.py:40-41)
            emit(seq);
        }
    } else {
        bool seen = false;
        uint64_t cur = 0;
        while (r.line(&c, &n)) {
            if (c == '>') {
                if (seen) emit(cur);
                seen = true;
                cur = 0;
            } else if (seen) {
                cur += n;
            }
        }
        if (seen) emit(cur);
    }
    if (mean_out) *mean_out = kept ? (double)(sum / kept) : 0.0;
    return kept;
}

// Host helper: the reference's sequential coordinate shift of insertion sites (create_insertion_genome.py:60-64,76-80):
// insertion i lands at positions[i] and pushes every EARLIER insertion that currently starts at or after it right by
// insertion_length.  The reference walks a dict per insertion; this is the same O(n^2) recurrence as two flat loops
// the compiler vectorises (10 k insertions: ~10 ms; the Python walk takes minutes at 24 barcodes).
extern "C" int esl_shift_insertions(const int64_t *positions, int64_t n, int64_t insertion_length, int64_t *starts)
{
    if (n < 0 || (n > 0 && (!positions || !starts))) return ESL_ERR_INVALID;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t p = positions[i];
        for (int64_t j = 0; j < i; ++j) starts[j] += starts[j] >= p ? insertion_length : 0;
        starts[i] = p;
    }
    return ESL_OK;
}
