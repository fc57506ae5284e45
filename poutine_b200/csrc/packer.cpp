// Host packer: `seg_sites` characters -> node-major bit planes (the Fasta_Record -> plane step the
// Java host performs after build_seg_sites, HC.java:1241-1306).
//
// 2-bit base code A0 C1 G2 T3 in (B1,B0); V = genotype is an UPPER-CASE A/C/G/T.  Everything else
// ('N', '-', lower-case, IUPAC codes) is "missing" exactly as in the reference, whose tests are
// `geno == 'A' || geno == 'C' || geno == 'G' || geno == 'T'` (HC.java:1600, 1845).
// Bit j of word w = site 64*w + j.  Rows are independent: one std::thread per slice of nodes.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/poutine_b200.h"
#include "pack_simd.h"

static void pack_rows(const char* seq, int64_t v0, int64_t v1, int64_t n_sites, int64_t seq_pitch, uint64_t* B0, uint64_t* B1,
                      uint64_t* V, int64_t pitch_words) {
    const pb_pack64_fn pack64 = pb_pack64_select();
    const int64_t W = (n_sites + 63) / 64;
    for (int64_t v = v0; v < v1; v++) {
        const unsigned char* row = (const unsigned char*)seq + v * seq_pitch;
        uint64_t* b0 = B0 + v * pitch_words;
        uint64_t* b1 = B1 + v * pitch_words;
        uint64_t* vv = V + v * pitch_words;
        for (int64_t w = 0; w < W; w++) {
            const int64_t s0 = w * 64;
            if (n_sites - s0 >= 64) { pack64(row + s0, b0 + w, b1 + w, vv + w); continue; }
            unsigned char tail[64];
            memset(tail, 0, sizeof(tail));                 // padding sites are "missing": V = 0
            memcpy(tail, row + s0, (size_t)(n_sites - s0));
            pack64(tail, b0 + w, b1 + w, vv + w);
        }
        for (int64_t w = W; w < pitch_words; w++) { b0[w] = 0; b1[w] = 0; vv[w] = 0; }
    }
}

extern "C" int pb_pack_chars(const char* seq, int64_t n_nodes, int64_t n_sites, int64_t seq_pitch, uint64_t* B0, uint64_t* B1,
                             uint64_t* V, int64_t pitch_words) {
    if (!seq || !B0 || !B1 || !V || n_nodes < 1 || n_sites < 1 || seq_pitch < n_sites || pitch_words < (n_sites + 63) / 64)
        return PB_ERR_INVALID;
    unsigned hw = std::thread::hardware_concurrency();
    int64_t nt = hw ? hw : 1;
    if (nt > n_nodes) nt = n_nodes;
    if (n_nodes * n_sites < (1 << 20)) nt = 1;
    if (nt == 1) { pack_rows(seq, 0, n_nodes, n_sites, seq_pitch, B0, B1, V, pitch_words); return PB_OK; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; t++)
        th.emplace_back(pack_rows, seq, n_nodes * t / nt, n_nodes * (t + 1) / nt, n_sites, seq_pitch, B0, B1, V, pitch_words);
    for (auto& t : th) t.join();
    return PB_OK;
}

// ---- biallelic compaction of a column tile ---------------------------------------------------------------
// A site that carries at most two distinct bases over ALL nodes (what the reference's biallelic SNP alignments look
// like: build_seg_sites keeps every node's genotype, HC.java:1241-1306) needs one bit per node, not three:
//     X bit  = the node's valid genotype differs from the site's first valid base (A1:A0)
//     B0 = A0 ^ (X & D0),  B1 = A1 ^ (X & D1)   with (D1:D0) = first base XOR second base, per site
// pb_put_planes_biallelic uploads X (and V only when some genotype is not an upper-case ACGT) and expands on the
// device: a third of the PCIe bytes.  Under V = 0 the expanded B bits are the first base instead of the caller's
// (don't-care: every consumer masks with V).  One thread per slice of word columns, column blocks of 64 words so that
// a row visit reads 512 contiguous bytes per plane.
namespace {
struct CompactJob {
    const uint64_t *B0, *B1, *V;
    int64_t n_nodes, w0, w1, pitch, n_sites;
    uint64_t* X; int64_t x_pitch;
    uint64_t* colmask;
    bool ok = true, all_valid = true;
};

void compact_cols(CompactJob* j) {
    constexpr int CB = 64;
    const int64_t W = (j->n_sites + 63) / 64;
    for (int64_t c0 = j->w0; c0 < j->w1 && j->ok; c0 += CB) {
        const int nc = (int)std::min<int64_t>(CB, j->w1 - c0);
        uint64_t A0[CB], A1[CB], seen[CB], D0[CB], D1[CB], dset[CB], bad = 0, allv[CB];
        for (int c = 0; c < nc; c++) { A0[c] = A1[c] = seen[c] = D0[c] = D1[c] = dset[c] = 0; allv[c] = ~0ull; }
        for (int64_t v = 0; v < j->n_nodes; v++) {      // pass 1: the site's first valid base
            const uint64_t *b0 = j->B0 + v * j->pitch + c0, *b1 = j->B1 + v * j->pitch + c0, *vv = j->V + v * j->pitch + c0;
            for (int c = 0; c < nc; c++) {
                const uint64_t nw = vv[c] & ~seen[c];
                A0[c] |= b0[c] & nw; A1[c] |= b1[c] & nw; seen[c] |= nw; allv[c] &= vv[c];
            }
        }
        for (int64_t v = 0; v < j->n_nodes; v++) {      // pass 2: X, the second base, and the at-most-two-bases check
            const uint64_t *b0 = j->B0 + v * j->pitch + c0, *b1 = j->B1 + v * j->pitch + c0, *vv = j->V + v * j->pitch + c0;
            uint64_t* x = j->X + v * j->x_pitch + c0;
            for (int c = 0; c < nc; c++) {
                const uint64_t d0 = (b0[c] ^ A0[c]) & vv[c], d1 = (b1[c] ^ A1[c]) & vv[c], xx = d0 | d1;
                const uint64_t nw = xx & ~dset[c];
                D0[c] |= d0 & nw; D1[c] |= d1 & nw; dset[c] |= nw;
                bad |= xx & ((d0 ^ D0[c]) | (d1 ^ D1[c]));   // on a differing site the difference must be THE second base
                x[c] = xx;
            }
            if (bad) { j->ok = false; return; }
        }
        for (int c = 0; c < nc; c++) {
            uint64_t* m = j->colmask + 4 * (c0 + c);
            m[0] = A0[c]; m[1] = A1[c]; m[2] = D0[c]; m[3] = D1[c];
            const int64_t w = c0 + c;
            uint64_t care = ~0ull;                        // n_sites is the TILE's: the padding bits of its last word do not count
            if (w == W - 1 && (j->n_sites & 63)) care = (1ull << (j->n_sites & 63)) - 1;
            if ((allv[c] & care) != care) j->all_valid = false;
        }
    }
}
}  // namespace

extern "C" int pb_compact_planes(const uint64_t* B0, const uint64_t* B1, const uint64_t* V, int64_t n_nodes, int64_t n_sites,
                                 int64_t pitch_words, uint64_t* X, int64_t x_pitch_words, uint64_t* colmask, int32_t* all_valid,
                                 int32_t n_threads) {
    const int64_t W = (n_sites + 63) / 64;
    if (!B0 || !B1 || !V || !X || !colmask || n_nodes < 1 || n_sites < 1 || pitch_words < W || x_pitch_words < W) return PB_ERR_INVALID;
    int64_t nt = n_threads > 0 ? n_threads : (int64_t)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    nt = std::min<int64_t>(nt, (W + 63) / 64);
    if (n_nodes * W < (1 << 16)) nt = 1;
    std::vector<CompactJob> jobs((size_t)nt);
    std::vector<std::thread> th;
    const int64_t blocks = (W + 63) / 64;
    for (int64_t t = 0; t < nt; t++) {
        CompactJob& j = jobs[(size_t)t];
        j.B0 = B0; j.B1 = B1; j.V = V; j.n_nodes = n_nodes; j.pitch = pitch_words; j.n_sites = n_sites;
        j.w0 = std::min(W, 64 * (blocks * t / nt)); j.w1 = std::min(W, 64 * (blocks * (t + 1) / nt));
        j.X = X; j.x_pitch = x_pitch_words; j.colmask = colmask;
        if (nt > 1) th.emplace_back(compact_cols, &j); else compact_cols(&j);
    }
    for (auto& t : th) t.join();
    bool ok = true, av = true;
    for (auto& j : jobs) { ok = ok && j.ok; av = av && j.all_valid; }
    if (all_valid) *all_valid = av ? 1 : 0;
    return ok ? 1 : 0;
}
