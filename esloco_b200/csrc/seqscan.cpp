// Streaming FASTA / FASTQ read-length scanner (host side of the length table) -- SURVEY.md section 8(f) row 3.
// Replaces the per-iteration Bio.SeqIO parse of the reference (config_handler.py:16-51, re-run from
// combined_calculations.py:37-39): one pass over the file with a 4 MB buffer, plain or gzip, only lengths kept.
#include <zlib.h>
#ifdef __SSE2__
#include <emmintrin.h>
#endif

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/esloco_b200.h"

namespace {

struct Reader { // buffered line reader: gzip files through zlib, plain files straight from read() (no second copy)
    gzFile f = nullptr;
    FILE *plain = nullptr;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    bool open(const char *path)
    {
        plain = fopen(path, "rb");
        if (!plain) return false;
        unsigned char magic[2] = {0, 0};
        const size_t got = fread(magic, 1, 2, plain);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) { // gzip
            fclose(plain);
            plain = nullptr;
            f = gzopen(path, "rb");
            if (!f) return false;
            gzbuffer(f, 1 << 20);
        } else {
            rewind(plain);
            setvbuf(plain, nullptr, _IONBF, 0); // fread of 4 MB blocks goes straight into buf
        }
        buf.resize(4 << 20);
        return true;
    }
    ~Reader()
    {
        if (f) gzclose(f);
        if (plain) fclose(plain);
    }
    bool fill()
    {
        size_t n;
        if (plain) n = fread(buf.data(), 1, buf.size(), plain);
        else { const int g = gzread(f, buf.data(), (unsigned)buf.size()); n = g > 0 ? (size_t)g : 0; }
        pos = 0;
        len = n;
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

// Number of line breaks, blanks and tabs in [p, e): 16 bytes per step where SSE2 is there (every x86-64), else a loop.
inline uint64_t count_blanks(const char *p, const char *e)
{
    uint64_t total = 0;
#ifdef __SSE2__
    const __m128i nl = _mm_set1_epi8('\n'), cr = _mm_set1_epi8('\r'), sp = _mm_set1_epi8(' '), tb = _mm_set1_epi8('\t');
    while (e - p >= 16) {
        __m128i acc = _mm_setzero_si128(); // per-byte counts, at most 255 steps before they are summed
        for (int k = 0; k < 255 && e - p >= 16; ++k, p += 16) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p));
            const __m128i m = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, nl), _mm_cmpeq_epi8(v, cr)),
                                           _mm_or_si128(_mm_cmpeq_epi8(v, sp), _mm_cmpeq_epi8(v, tb)));
            acc = _mm_sub_epi8(acc, m); // a match is 0xff = -1
        }
        const __m128i sums = _mm_sad_epu8(acc, _mm_setzero_si128());
        total += (uint64_t)_mm_cvtsi128_si64(sums) + (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(sums, sums));
    }
#endif
    for (; p < e; ++p) total += (uint64_t)((*p == '\n') | (*p == '\r') | (*p == ' ') | (*p == '\t'));
    return total;
}

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

// Contig ids and lengths of a FASTA in one streaming pass (fasta_operations.py:3-22 loads every sequence into memory
// and joins them; only the ids and lengths are ever used).  A record starts at a line whose first character is '>'
// (as in Bio.SeqIO); its id is the first whitespace-delimited token of that line; its length counts the characters of
// the following lines except line breaks, blanks and tabs.  Lines before the first header are ignored.
extern "C" int64_t esl_scan_fasta_contigs(const char *path, char *names, int64_t names_cap, int64_t *lengths, int64_t max_records,
                                          int64_t *names_bytes)
{
    if (!path) return ESL_ERR_INVALID;
    Reader r;
    if (!r.open(path)) return ESL_ERR_INVALID;
    int64_t n_rec = 0, name_at = 0;
    uint64_t cur = 0;
    bool seen = false, in_header = false, line_start = true;
    std::string header;
    auto close_record = [&]() {
        if (seen && lengths && n_rec - 1 < max_records) lengths[n_rec - 1] = (int64_t)cur;
    };
    auto open_record = [&]() { // header holds the line behind '>' (without the line break)
        size_t a = 0;
        while (a < header.size() && (header[a] == ' ' || header[a] == '\t' || header[a] == '\r' || header[a] == '\v' || header[a] == '\f')) ++a;
        size_t b = a;
        while (b < header.size() && !(header[b] == ' ' || header[b] == '\t' || header[b] == '\r' || header[b] == '\v' || header[b] == '\f')) ++b;
        const int64_t need = (int64_t)(b - a) + 1;
        if (names && n_rec < max_records && name_at + need <= names_cap) {
            memcpy(names + name_at, header.data() + a, b - a);
            names[name_at + need - 1] = 0;
        }
        name_at += need;
        ++n_rec;
        seen = true;
        cur = 0;
    };
    for (;;) {
        if (r.pos == r.len && !r.fill()) break;
        const char *p = r.buf.data() + r.pos, *end = r.buf.data() + r.len;
        while (p < end) {
            if (in_header) {
                const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
                header.append(p, nl ? nl : end);
                p = nl ? nl + 1 : end;
                if (nl) { open_record(); in_header = false; line_start = true; }
            } else if (line_start && *p == '>') {
                close_record();
                header.clear();
                in_header = true;
                line_start = false;
                ++p;
            } else { // sequence lines up to the next line that starts a record: counted as one block
                const char *stop = end; // the next '>' that follows a line break ('>' is rare inside sequence lines: memchr)
                for (const char *g = p + 1; g < end && (g = (const char *)memchr(g, '>', (size_t)(end - g))) != nullptr; ++g)
                    if (g[-1] == '\n') { stop = g; break; }
                if (seen) cur += (uint64_t)(stop - p) - count_blanks(p, stop);
                line_start = stop[-1] == '\n';
                p = stop;
            }
        }
        r.pos = r.len;
    }
    if (in_header) open_record(); // the file ends inside a header line
    close_record();
    if (names_bytes) *names_bytes = name_at;
    return n_rec;
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
