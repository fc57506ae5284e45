// Host packer at speed: ancestral_sequences.fasta -> node-major bit planes, by column tile,
// so that packing of tile t+1 overlaps the upload (pb_put_planes) and K1 of tile t.
//
// Record semantics follow the reference's reader (compiled/Fasta_Manager.class next(), used by build_seg_sites,
// HC.java:1241-1306): a header line starts with '>', node name = header.substring(1).trim(); the sequence is the
// concatenation of the following lines up to the next header or EOF, lines split the way BufferedReader.readLine
// does ("\n", "\r\n" or "\r"), nothing else trimmed.  Characters are classified as in pb_pack_chars:
// upper-case A/C/G/T are alleles, everything else is missing (HC.java:1600, 1845).
//
// The file is mmap'ed; records are indexed once (parallel-friendly, single pass); each record remembers its
// first-line width so that site -> file offset is O(1) for regular FASTA (fixed line width or one line per
// record); irregular records fall back to a per-record line table.
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/poutine_b200.h"
#include "pack_simd.h"

struct FastaRecord {
    int64_t name_off, name_len;   // trimmed header text
    int64_t seq_off;              // file offset of the first sequence char
    int64_t seq_len;              // chars after concatenating lines
    int64_t line_w;               // chars per full line (regular layout), 0 = irregular -> `lines`
    int64_t line_stride;          // bytes from one line start to the next (line_w + eol bytes)
    std::vector<int64_t> lines;   // irregular only: (file offset, cumulative chars before) pairs
};

struct pb_fasta {
    int fd = -1;
    const char* data = nullptr;
    int64_t size = 0;
    std::vector<FastaRecord> recs;
    std::string err;
};

static thread_local std::string g_fasta_err;

static inline bool is_eol(char c) { return c == '\n' || c == '\r'; }

// false: the file does not start with a header line — where Fasta_Manager.next() prints "First line of Fasta file muust be a
// header line." (or, for an empty first line, "the header line should never be an empty string") and calls System.exit(-1).
// Every later record can only begin at a '>' line, so the first line is the one place this can happen.
static bool index_records(pb_fasta* f) {
    const char* d = f->data;
    const int64_t n = f->size;
    int64_t pos = 0, next_nl = -1;
    if (n > 0 && d[0] != '>') {
        f->err = is_eol(d[0]) ? "pb_fasta_open: the first line is empty, the header line should never be an empty string"
                              : "pb_fasta_open: first line of the FASTA file must be a header line ('>')";
        return false;
    }
    while (pos < n) {
        FastaRecord r{};
        // header line
        int64_t h0 = pos + 1, h1 = h0;
        while (h1 < n && !is_eol(d[h1])) h1++;
        int64_t a = h0, b = h1;
        while (a < b && (unsigned char)d[a] <= ' ') a++;          // String.trim(): chars <= U+0020
        while (b > a && (unsigned char)d[b - 1] <= ' ') b--;
        r.name_off = a; r.name_len = b - a;
        pos = h1;
        if (pos < n && d[pos] == '\r') pos++;
        if (pos < n && d[pos] == '\n') pos++;
        r.seq_off = pos;
        // sequence lines: (file offset, chars before) per readLine() line
        int64_t total = 0;
        std::vector<int64_t> lines;
        bool header_after_cr = false;
        while (pos < n && d[pos] != '>') {
            const int64_t l0 = pos;
            if (next_nl < pos) {                       // cached: a file with bare-'\r' line ends has few (or no) '\n'
                const char* q = (const char*)memchr(d + pos, '\n', (size_t)(n - pos));
                next_nl = q ? (q - d) : n;
            }
            const bool nl = next_nl < n;
            const int64_t l1 = next_nl;               // physical line end (exclusive of '\n')
            int64_t e = l1;
            if (e > l0 && d[e - 1] == '\r') e--;      // "\r\n"
            int64_t s = l0;
            for (;;) {                                 // a bare '\r' also ends a readLine() line
                const char* cr = (const char*)memchr(d + s, '\r', (size_t)(e - s));
                const int64_t se = cr ? (cr - d) : e;
                lines.push_back(s); lines.push_back(total);
                total += se - s;
                if (!cr) break;
                s = se + 1;
                if (s < e && d[s] == '>') { header_after_cr = true; break; }   // the bare '\r' ended the line before a header line
            }
            if (header_after_cr) { pos = s; break; }
            pos = nl ? l1 + 1 : n;
        }
        r.seq_len = total;
        const int64_t n_lines = (int64_t)lines.size() / 2;
        bool regular = n_lines > 0;
        if (n_lines == 1) {
            r.line_w = total > 0 ? total : 1; r.line_stride = r.line_w + 1;
        } else if (n_lines > 1) {
            const int64_t w0 = lines[3] - lines[1], st0 = lines[2] - lines[0];
            for (int64_t i = 1; i < n_lines && regular; i++) {
                const int64_t wi = ((i + 1 < n_lines) ? lines[2 * i + 3] : total) - lines[2 * i + 1];
                if (lines[2 * i] - lines[2 * i - 2] != st0) regular = false;
                if (i + 1 < n_lines ? wi != w0 : wi > w0) regular = false;
            }
            if (regular && w0 > 0) { r.line_w = w0; r.line_stride = st0; }
            else { r.line_w = 0; r.lines = std::move(lines); }
        } else {
            r.line_w = 1; r.line_stride = 2;           // empty record
        }
        f->recs.push_back(std::move(r));
    }
    return true;
}

extern "C" const char* pb_fasta_last_error(void) { return g_fasta_err.c_str(); }

extern "C" int pb_fasta_open(const char* path, pb_fasta** out) {
    if (!path || !out) { g_fasta_err = "pb_fasta_open: null argument"; return PB_ERR_INVALID; }
    *out = nullptr;
    pb_fasta* f = new pb_fasta();
    f->fd = open(path, O_RDONLY);
    if (f->fd < 0) { g_fasta_err = std::string("pb_fasta_open: cannot open ") + path; delete f; return PB_ERR_INVALID; }
    struct stat st;
    if (fstat(f->fd, &st) != 0) { g_fasta_err = "pb_fasta_open: fstat failed"; close(f->fd); delete f; return PB_ERR_INVALID; }
    f->size = (int64_t)st.st_size;
    if (f->size > 0) {
        void* p = mmap(nullptr, (size_t)f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
        if (p == MAP_FAILED) { g_fasta_err = "pb_fasta_open: mmap failed"; close(f->fd); delete f; return PB_ERR_OOM; }
        f->data = (const char*)p;
        madvise(p, (size_t)f->size, MADV_WILLNEED);
    }
    if (!index_records(f)) {
        g_fasta_err = f->err;
        pb_fasta_close(f);
        return PB_ERR_INVALID;
    }
    *out = f;
    return PB_OK;
}

extern "C" void pb_fasta_close(pb_fasta* f) {
    if (!f) return;
    if (f->data) munmap((void*)f->data, (size_t)f->size);
    if (f->fd >= 0) close(f->fd);
    delete f;
}

extern "C" int64_t pb_fasta_num_records(pb_fasta* f) { return f ? (int64_t)f->recs.size() : -1; }

extern "C" int pb_fasta_record(pb_fasta* f, int64_t i, const char** name, int64_t* name_len, int64_t* seq_len) {
    if (!f || i < 0 || i >= (int64_t)f->recs.size()) return PB_ERR_INVALID;
    const FastaRecord& r = f->recs[(size_t)i];
    if (name) *name = f->data + r.name_off;
    if (name_len) *name_len = r.name_len;
    if (seq_len) *seq_len = r.seq_len;
    return PB_OK;
}

// copy sites [s0, s1) of record r into buf (concatenated-line coordinates)
static void gather_chars(const pb_fasta* f, const FastaRecord& r, int64_t s0, int64_t s1, unsigned char* buf) {
    const char* d = f->data;
    if (r.line_w > 0) {
        int64_t s = s0;
        while (s < s1) {
            const int64_t line = s / r.line_w, col = s % r.line_w;
            const int64_t take = std::min(r.line_w - col, s1 - s);
            memcpy(buf + (s - s0), d + r.seq_off + line * r.line_stride + col, (size_t)take);
            s += take;
        }
    } else {
        // irregular: binary search the line table (pairs: offset, cumulative chars)
        const int64_t nl = (int64_t)r.lines.size() / 2;
        int64_t lo = 0, hi = nl - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (r.lines[2 * mid + 1] <= s0) lo = mid; else hi = mid - 1;
        }
        int64_t s = s0, li = lo;
        while (s < s1 && li < nl) {
            const int64_t cum = r.lines[2 * li + 1];
            const int64_t next_cum = (li + 1 < nl) ? r.lines[2 * li + 3] : r.seq_len;
            const int64_t col = s - cum, avail = next_cum - s;
            const int64_t take = std::min(avail, s1 - s);
            if (take > 0) memcpy(buf + (s - s0), d + r.lines[2 * li] + col, (size_t)take);
            s += take;
            li++;
        }
    }
}

static void pack_records(const pb_fasta* f, const int32_t* node_of_record, int64_t r0, int64_t r1, int64_t site_begin, int64_t n_sites_tile,
                         uint64_t* B0, uint64_t* B1, uint64_t* V, int64_t pitch_words) {
    const pb_pack64_fn pack64 = pb_pack64_select();
    const int64_t W = (n_sites_tile + 63) / 64;
    std::vector<unsigned char> buf((size_t)W * 64);
    for (int64_t i = r0; i < r1; i++) {
        const int32_t v = node_of_record[i];
        if (v < 0) continue;
        const FastaRecord& r = f->recs[(size_t)i];
        const int64_t s0 = std::min(site_begin, r.seq_len), s1 = std::min(site_begin + n_sites_tile, r.seq_len);
        memset(buf.data(), 0, buf.size());           // short records: the rest is "missing"
        if (s1 > s0) gather_chars(f, r, s0, s1, buf.data());
        uint64_t* b0 = B0 + (int64_t)v * pitch_words;
        uint64_t* b1 = B1 + (int64_t)v * pitch_words;
        uint64_t* vv = V + (int64_t)v * pitch_words;
        for (int64_t w = 0; w < W; w++) pack64(buf.data() + w * 64, b0 + w, b1 + w, vv + w);
    }
}

// Pack the column tile [site_begin, site_begin + n_sites_tile) (site_begin % 64 == 0) of every record into row
// node_of_record[i] of the tile-local planes (row pitch pitch_words >= ceil(n_sites_tile / 64)).  Records mapped
// to -1 are skipped; rows of nodes without a record are left untouched (the caller zero-fills).
extern "C" int pb_fasta_pack_tile(pb_fasta* f, const int32_t* node_of_record, int64_t site_begin, int64_t n_sites_tile, uint64_t* B0,
                                  uint64_t* B1, uint64_t* V, int64_t pitch_words, int32_t n_threads) {
    if (!f || !node_of_record || !B0 || !B1 || !V || site_begin < 0 || (site_begin & 63) || n_sites_tile < 1 ||
        pitch_words < (n_sites_tile + 63) / 64) {
        g_fasta_err = "pb_fasta_pack_tile: invalid arguments";
        return PB_ERR_INVALID;
    }
    const int64_t nrec = (int64_t)f->recs.size();
    int64_t nt = n_threads > 0 ? n_threads : (int64_t)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > nrec) nt = nrec > 0 ? nrec : 1;
    if (nt == 1) { pack_records(f, node_of_record, 0, nrec, site_begin, n_sites_tile, B0, B1, V, pitch_words); return PB_OK; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; t++)
        th.emplace_back(pack_records, f, node_of_record, nrec * t / nt, nrec * (t + 1) / nt, site_begin, n_sites_tile, B0, B1, V, pitch_words);
    for (auto& t : th) t.join();
    return PB_OK;
}

// =====================================================================================================================
// One pass from genotype characters to the COMPACT upload format of pb_put_planes_biallelic (no three-plane detour, no
// compaction pass): the host side of build_seg_sites (HC.java:1241-1306) for the usual alignment — every genotype an
// upper-case A/C/G/T, at most two bases per site.
//   reference row  = the first mapped record's characters (ref[s]: the site's "first base", colmask A1:A0)
//   X bit          = this node's character differs from ref[s]                (one byte compare + movemask per 32 sites)
//   second[s]      = THE other character of the site: every differing character must equal it (checked while packing,
//                    per worker thread, merged at the end); colmask D1:D0 = code(ref) ^ code(second)
// A tile qualifies (return 1) when ref[s] and second[s] are upper-case ACGT for every site, every record covers the tile
// and every node got a row; otherwise 0 (outputs undefined): pack that tile with pb_fasta_pack_tile (+ pb_compact_planes).
// Validity needs no per-character test: a character equals ref[s] or second[s], and those two are checked once per site.
namespace {

struct XPackCtx {
    const unsigned char* ref;      // tile chars of the reference row
    int64_t n;                     // sites in the tile
};

// one row: X words out, second[] updated; returns false when a third character shows up at some site
static bool xpack_row_scalar(const unsigned char* x, const unsigned char* ref, unsigned char* sec, int64_t n, uint64_t* out) {
    bool bad = false;
    for (int64_t w = 0; w * 64 < n; w++) {
        uint64_t bits = 0;
        const int64_t lim = std::min<int64_t>(64, n - w * 64);
        for (int64_t j = 0; j < lim; j++) {
            const int64_t i = w * 64 + j;
            if (x[i] != ref[i]) {
                bits |= 1ull << j;
                if (sec[i] == 0) sec[i] = x[i];
                else if (sec[i] != x[i]) bad = true;
            }
        }
        out[w] = bits;
    }
    return !bad;
}

#ifdef PB_HAVE_X86
__attribute__((target("avx2"))) static bool xpack_row_avx2(const unsigned char* x, const unsigned char* ref, unsigned char* sec, int64_t n,
                                                           uint64_t* out) {
    __m256i badv = _mm256_setzero_si256();
    const __m256i zero = _mm256_setzero_si256();
    const int64_t full = n / 64;
    for (int64_t w = 0; w < full; w++) {
        uint64_t bits = 0;
        for (int h = 0; h < 2; h++) {
            const int64_t i = w * 64 + 32 * h;
            const __m256i xv = _mm256_loadu_si256((const __m256i*)(x + i)), rv = _mm256_loadu_si256((const __m256i*)(ref + i));
            const __m256i eq = _mm256_cmpeq_epi8(xv, rv);
            const uint32_t eqm = (uint32_t)_mm256_movemask_epi8(eq);
            bits |= (uint64_t)(~eqm) << (32 * h);
            if (eqm != 0xFFFFFFFFu) {                      // rows mostly equal the reference: the second-base check is the rare path
                const __m256i sv = _mm256_loadu_si256((const __m256i*)(sec + i));
                const __m256i unset = _mm256_cmpeq_epi8(sv, zero), eqs = _mm256_cmpeq_epi8(xv, sv);
                const __m256i take = _mm256_andnot_si256(eq, unset);                       // differs, second not known yet
                _mm256_storeu_si256((__m256i*)(sec + i), _mm256_blendv_epi8(sv, xv, take));
                // differs, second known, and it is not the second: a third character
                badv = _mm256_or_si256(badv, _mm256_andnot_si256(_mm256_or_si256(_mm256_or_si256(eq, unset), eqs), _mm256_set1_epi8((char)0xFF)));
            }
        }
        out[w] = bits;
    }
    bool ok = _mm256_testz_si256(badv, badv) != 0;
    if (full * 64 < n) ok = xpack_row_scalar(x + full * 64, ref + full * 64, sec + full * 64, n - full * 64, out + full) && ok;
    return ok;
}
#endif

typedef bool (*xpack_row_fn)(const unsigned char*, const unsigned char*, unsigned char*, int64_t, uint64_t*);
static xpack_row_fn xpack_select() {
#ifdef PB_HAVE_X86
    if (__builtin_cpu_supports("avx2")) return xpack_row_avx2;
#endif
    return xpack_row_scalar;
}

static inline int base_code(unsigned char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// merge the workers' second[] arrays, validate, write the colmask words {A0, A1, D0, D1}; false = the tile does not qualify
static bool xpack_finish(const unsigned char* ref, std::vector<std::vector<unsigned char>>& secs, int64_t n, uint64_t* colmask) {
    const int64_t W = (n + 63) / 64;
    for (int64_t w = 0; w < W; w++) {
        uint64_t a0 = 0, a1 = 0, d0 = 0, d1 = 0;
        const int64_t lim = std::min<int64_t>(64, n - w * 64);
        for (int64_t j = 0; j < lim; j++) {
            const int64_t i = w * 64 + j;
            unsigned char s2 = 0;
            for (auto& sv : secs) {
                const unsigned char c = sv[(size_t)i];
                if (!c) continue;
                if (s2 == 0) s2 = c; else if (s2 != c) return false;          // two workers saw different second characters
            }
            const int ca = base_code(ref[i]);
            if (ca < 0) return false;
            int cd = 0;
            if (s2) { const int cs = base_code(s2); if (cs < 0) return false; cd = ca ^ cs; }
            a0 |= (uint64_t)(ca & 1) << j; a1 |= (uint64_t)(ca >> 1) << j;
            d0 |= (uint64_t)(cd & 1) << j; d1 |= (uint64_t)(cd >> 1) << j;
        }
        colmask[4 * w] = a0; colmask[4 * w + 1] = a1; colmask[4 * w + 2] = d0; colmask[4 * w + 3] = d1;
    }
    return true;
}

}  // namespace

extern "C" int pb_fasta_pack_tile_biallelic(pb_fasta* f, const int32_t* node_of_record, int64_t n_nodes, int64_t site_begin,
                                            int64_t n_sites_tile, uint64_t* X, int64_t pitch_words, uint64_t* colmask, int32_t n_threads) {
    if (!f || !node_of_record || !X || !colmask || n_nodes < 1 || site_begin < 0 || (site_begin & 63) || n_sites_tile < 1 ||
        pitch_words < (n_sites_tile + 63) / 64) {
        g_fasta_err = "pb_fasta_pack_tile_biallelic: invalid arguments";
        return PB_ERR_INVALID;
    }
    const int64_t nrec = (int64_t)f->recs.size(), n = n_sites_tile, W = (n + 63) / 64;
    int64_t first = -1, mapped = 0;
    for (int64_t i = 0; i < nrec; i++) {
        const int32_t v = node_of_record[i];
        if (v < 0) continue;
        if (v >= n_nodes) { g_fasta_err = "pb_fasta_pack_tile_biallelic: node index out of range"; return PB_ERR_INVALID; }
        if (f->recs[(size_t)i].seq_len < site_begin + n) return 0;            // a short record: missing genotypes
        if (first < 0) first = i;
        mapped++;
    }
    if (mapped != n_nodes) return 0;                                           // some node has no row (or two): not all-valid
    std::vector<unsigned char> ref((size_t)W * 64, 0);
    gather_chars(f, f->recs[(size_t)first], site_begin, site_begin + n, ref.data());
    int64_t nt = n_threads > 0 ? n_threads : (int64_t)std::thread::hardware_concurrency();
    nt = std::max<int64_t>(1, std::min(nt, nrec));
    std::vector<std::vector<unsigned char>> secs((size_t)nt);
    std::vector<char> ok((size_t)nt, 1);
    const xpack_row_fn xrow = xpack_select();
    auto work = [&](int64_t t) {
        std::vector<unsigned char>& sec = secs[(size_t)t];
        sec.assign((size_t)W * 64, 0);
        std::vector<unsigned char> buf((size_t)W * 64, 0);
        for (int64_t i = nrec * t / nt; i < nrec * (t + 1) / nt; i++) {
            const int32_t v = node_of_record[i];
            if (v < 0) continue;
            gather_chars(f, f->recs[(size_t)i], site_begin, site_begin + n, buf.data());
            if (!xrow(buf.data(), ref.data(), sec.data(), n, X + (int64_t)v * pitch_words)) { ok[(size_t)t] = 0; return; }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int64_t t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    for (char o : ok) if (!o) return 0;
    return xpack_finish(ref.data(), secs, n, colmask) ? 1 : 0;
}

// the same from a character matrix (node-major, row v at seq + v*seq_pitch)
extern "C" int pb_pack_chars_biallelic(const char* seq, int64_t n_nodes, int64_t n_sites, int64_t seq_pitch, uint64_t* X,
                                       int64_t x_pitch_words, uint64_t* colmask, int32_t n_threads) {
    const int64_t W = (n_sites + 63) / 64;
    if (!seq || !X || !colmask || n_nodes < 1 || n_sites < 1 || seq_pitch < n_sites || x_pitch_words < W) return PB_ERR_INVALID;
    const unsigned char* ref = (const unsigned char*)seq;      // row 0
    int64_t nt = n_threads > 0 ? n_threads : (int64_t)std::thread::hardware_concurrency();
    nt = std::max<int64_t>(1, std::min(nt, n_nodes));
    if (n_nodes * n_sites < (1 << 20)) nt = 1;
    std::vector<std::vector<unsigned char>> secs((size_t)nt);
    std::vector<char> ok((size_t)nt, 1);
    const xpack_row_fn xrow = xpack_select();
    auto work = [&](int64_t t) {
        secs[(size_t)t].assign((size_t)W * 64, 0);
        for (int64_t v = n_nodes * t / nt; v < n_nodes * (t + 1) / nt; v++) {
            uint64_t* out = X + v * x_pitch_words;
            if (!xrow((const unsigned char*)seq + v * seq_pitch, ref, secs[(size_t)t].data(), n_sites, out)) { ok[(size_t)t] = 0; return; }
            for (int64_t w = W; w < x_pitch_words; w++) out[w] = 0;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int64_t t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    for (char o : ok) if (!o) return 0;
    return xpack_finish(ref, secs, n_sites, colmask) ? 1 : 0;
}
