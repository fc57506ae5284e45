// Host packer: `seg_sites` characters -> node-major bit planes (the Fasta_Record -> plane step the
// Java host performs after build_seg_sites, HC.java:1241-1306).
//
// 2-bit base code A0 C1 G2 T3 in (B1,B0); V = genotype is an UPPER-CASE A/C/G/T.  Everything else
// ('N', '-', lower-case, IUPAC codes) is "missing" exactly as in the reference, whose tests are
// `geno == 'A' || geno == 'C' || geno == 'G' || geno == 'T'` (HC.java:1600, 1845).
// Bit j of word w = site 64*w + j.  Rows are independent: one std::thread per slice of nodes.
#include <stdint.h>
#include <string.h>

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
