// 64 genotype characters -> one word of each bit plane.  'A' 'C' 'G' 'T' (upper case) are alleles, everything else is
// missing (HC.java:1600, 1845).  2-bit base code A0 C1 G2 T3: B0 = C|T, B1 = G|T, V = A|C|G|T.
// AVX2 path (4 byte-compares + 4 movemasks per 32 chars) selected at run time; portable scalar fallback.
#pragma once
#include <stdint.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define PB_HAVE_X86 1
#endif

static inline void pb_pack64_scalar(const unsigned char* c, uint64_t* b0, uint64_t* b1, uint64_t* v) {
    uint64_t a = 0, cc = 0, g = 0, t = 0;
    for (int j = 0; j < 64; j++) {
        const unsigned char ch = c[j];
        a |= (uint64_t)(ch == 'A') << j;
        cc |= (uint64_t)(ch == 'C') << j;
        g |= (uint64_t)(ch == 'G') << j;
        t |= (uint64_t)(ch == 'T') << j;
    }
    *b0 = cc | t; *b1 = g | t; *v = a | cc | g | t;
}

#ifdef PB_HAVE_X86
__attribute__((target("avx2"))) static inline void pb_pack64_avx2(const unsigned char* c, uint64_t* b0, uint64_t* b1, uint64_t* v) {
    const __m256i A = _mm256_set1_epi8('A'), C = _mm256_set1_epi8('C'), G = _mm256_set1_epi8('G'), T = _mm256_set1_epi8('T');
    uint64_t m[4] = {0, 0, 0, 0};
    for (int h = 0; h < 2; h++) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(c + 32 * h));
        m[0] |= (uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, A)) << (32 * h);
        m[1] |= (uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, C)) << (32 * h);
        m[2] |= (uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, G)) << (32 * h);
        m[3] |= (uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, T)) << (32 * h);
    }
    *b0 = m[1] | m[3]; *b1 = m[2] | m[3]; *v = m[0] | m[1] | m[2] | m[3];
}
#endif

typedef void (*pb_pack64_fn)(const unsigned char*, uint64_t*, uint64_t*, uint64_t*);
static inline pb_pack64_fn pb_pack64_select() {
#ifdef PB_HAVE_X86
    if (__builtin_cpu_supports("avx2")) return pb_pack64_avx2;
#endif
    return pb_pack64_scalar;
}
