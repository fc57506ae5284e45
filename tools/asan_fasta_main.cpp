#include <cstdio>
#include <cstdlib>
#include <vector>
#include "poutine_b200.h"
int main(int argc, char** argv) {
    pb_fasta* f = nullptr;
    if (pb_fasta_open(argv[1], &f) != 0) { std::printf("refused: %s\n", pb_fasta_last_error()); return 0; }
    int64_t n = pb_fasta_num_records(f), S = 0;
    for (int64_t i = 0; i < n; i++) { const char* nm; int64_t nl, sl; pb_fasta_record(f, i, &nm, &nl, &sl); if (sl > S) S = sl; }
    std::vector<int32_t> nor(n); for (int64_t i = 0; i < n; i++) nor[i] = (int32_t)i;
    if (n && S) {
        int64_t W = (S + 63) / 64;
        for (int64_t tw : {W, (int64_t)1}) {
            std::vector<uint64_t> b0(n * tw), b1(n * tw), v(n * tw), x(n * tw), cm(4 * tw);
            for (int64_t w0 = 0; w0 < W; w0 += tw) {
                int64_t ns = std::min<int64_t>(S, 64 * (w0 + tw)) - 64 * w0;
                if (pb_fasta_pack_tile(f, nor.data(), 64 * w0, ns, b0.data(), b1.data(), v.data(), tw, argc > 2 ? atoi(argv[2]) : 2) != 0) { std::printf("pack error\n"); return 1; }
                int32_t av = 0;
                pb_compact_planes(b0.data(), b1.data(), v.data(), n, ns, tw, x.data(), tw, cm.data(), &av, 2);
            }
        }
    }
    std::printf("%lld records, S=%lld\n", (long long)n, (long long)S);
    pb_fasta_close(f);
    return 0;
}
