// K2 + K3 — observed tallies and the phenotype-permutation null (sm_100a).
//
// Replaces count_homoplasies/counts_by_allele (HC.java:3695-3761), Replicate.run
// (HC.java:3229-3272), permute_phenos (HC.java:4890-4923) and
// calc_binomial_test_stats_memoization_concurrent (HC.java:4196-4280).
//
// Reference semantics per replicate: permute the P pheno-file labels uniformly over the P samples,
// recount (n_r, k_r) = (#labelled, #case) over each in-use allele's extant MRCA leaves, look up
// p_r = f(n_r, k_r) (one-sided binomial, commons-math), r_a += (p_r <= p_obs), maxT[rep] = min p_r.
//
// B200 design: replicates are bit-parallel.
//   K3a  builds the label bit-matrix X[sample][replicate] (64 replicates per word), either from
//        counter-based Philox4x32-10 streams (one thread = one replicate, sequential draw without
//        replacement, warp ballot -> words) or from explicit replay permutations.  In Philox mode the
//        first generation tile is launched by pb_run_events on a second stream (overlaps the K1 tail).
//   K3b  alleles are classified once (k4_classify_kernel) into work lists:
//          0  nothing to sweep: r follows from per-sample column popcounts (n <= 1) or is 0 / n_reps;
//          1  short lists (n < 8) on the tight path: one warp sweeps one allele over all replicate words of the
//             tile, rows streamed coalesced from the (L2-resident) matrix, "k_r >= k*" evaluated bit-sliced,
//             r accumulated in registers and written once — no atomics;
//          2  long lists (n >= 8): a whole block per allele, general bit-sliced path;
//          3  short lists that need the general path (possible maxT candidates, non-monotone table rows).
//        "p_r <= p_obs" as a threshold k_r >= k* is valid because f(n,.) is checked per allele to be an
//        up-set w.r.t. p_obs (k4_observed_kernel); with missing labels in the pheno file the check covers every
//        n_r = n - j a replicate can see, with one threshold per j (AlleleMeta.kpack).  Otherwise, and for
//        maxT, replicate bits are extracted to scalars and looked up in f exactly as the reference does.
//        maxT uses a candidate threshold tau: only p-values <= tau are min-ed atomically; replicates left
//        without a certified minimum are recomputed exactly by k3_fallback_kernel.
#include <stdlib.h>

#include <algorithm>

#include "pb_internal.h"

// ---------------------------------------------------------------------------------------------------
// K2: observed tallies
// ---------------------------------------------------------------------------------------------------
__global__ void k2_tally_kernel(AlleleMeta* __restrict__ alleles, int64_t n_slots, const int32_t* __restrict__ csr,
                                const uint8_t* __restrict__ label) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_slots) return;
    AlleleMeta m = alleles[a];
    int cs = 0, ct = 0;
    if (m.flags & PB_AF_IN_USE)
        for (int e = 0; e < m.n; e++) {
            const uint8_t l = label[csr[m.off + e]];
            cs += l == PB_LABEL_CASE;    // pheno.equals("1") -> case   (HC.java:3747)
            ct += l == PB_LABEL_CONTROL; // pheno.equals("0") -> control (HC.java:3749)
        }
    alleles[a].obs_case = cs;
    alleles[a].obs_ctrl = ct;
}

// ---------------------------------------------------------------------------------------------------
// K3a: label bit-matrix
// ---------------------------------------------------------------------------------------------------
#ifndef PB_PHILOX_ROUNDS
#define PB_PHILOX_ROUNDS 10   // experiment knob only (7 rounds: generator 0.52 -> 0.43 ms at C2, not adopted; the oracle restates 10)
#endif
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < PB_PHILOX_ROUNDS; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#ifndef GEN_THREADS
#define GEN_THREADS 256       // replicates per block (a multiple of 32; Rw32 assumes tiles padded to 256 replicates)
#endif
static_assert(256 % GEN_THREADS == 0 && GEN_THREADS % 32 == 0, "generator blocks tile the 256-replicate padding unit");
#define GEN_WORDS (GEN_THREADS / 32)
#define GEN_ROWS 128
#ifndef GEN_ILP
#define GEN_ILP 4          // Philox blocks computed together (GEN_ROWS must be a multiple of 4 * GEN_ILP)
#endif

// MODE 0: batched draws (P < 65536; sampling spec v3 in DESIGN.md section 3, restated by oracle/oracle.cpp or_philox_labels)
// MODE 1: replay (sample j gets label[perms[rep*P + j]])
// MODE 2: Philox, one 32-bit word per draw (P >= 65536: four ranges would overflow a 64-bit product; spec v1)
// One thread = one replicate; a warp ballot turns 32 replicates' labels of one sample into a matrix word.
//
// Spec v3.  Replicate `rep` labels samples 0..P-1 sequentially without replacement, four at a time from one 64-bit word:
// group g = samples 4g..4g+c-1 (c = min(4, P - 4g)), x = (o[2g+1] : o[2g]) with o[i] = the i-th output of the replicate's
// xoshiro128++ stream, whose state starts as Philox4x32-10(ctr = (0, rep_lo, rep_hi, 2), key = seed) (all zero -> s0 = 1).
// For j = 0..c-1 with R_j = P - 4g - j:  x * R_j = u_j * 2^64 + x'  (u_j = the draw in [0, R_j), x' carries on as x) —
// Lemire's multiply-shift applied to the product range R_0..R_{c-1} at once: floor(x * prod / 2^64) in mixed radix.  The
// group is exact-uniform iff the final x' >= 2^64 mod prod(R_j); otherwise x is replaced by the next word of the retry
// stream (word q of Philox(ctr = (q>>1, rep_lo, rep_hi, 1)), q = 0, 1, ...) and the group is drawn again (probability
// < prod / 2^64 < 2^-11 per group for P = 5000); the main stream's position does not depend on the retries.  u_j <
// cases_left -> case, < cases_left + controls_left -> control, else missing.
// Spec v2 (superseded) took x from Philox(ctr = (g>>1, rep_lo, rep_hi, 0)): 2.5 of its 4.5 wide multiplies per label were
// the ten Philox rounds, on the pipe that bounds the kernel (profiles/r2_k3a_gen_v2_C2_ncu.txt).  The xoshiro step is nine
// shift / xor / add instructions per 32-bit output and no multiply; Philox still seeds every replicate's stream, so a
// replicate's labels remain a pure function of (seed, rep).
__device__ __forceinline__ uint32_t xoshiro128pp_next(uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
    const uint32_t a = s0 + s3;
    const uint32_t result = __funnelshift_l(a, a, 7) + s0;
    const uint32_t t = s1 << 9;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3;
    s2 ^= t;
    s3 = __funnelshift_l(s3, s3, 11);
    return result;
}
// (xh:xl) * R = u * 2^64 + (xh':xl').  Two 32x32->64 multiplies; the cross carry goes through the integer ALU (add.cc / addc),
// not through a 64-bit multiply-add: the generator is bound by the multiply pipe, and a mad.wide with a 64-bit addend also
// costs a register move on that pipe to build the addend's zero high word.
__device__ __forceinline__ void mul64x32(uint32_t& xl, uint32_t& xh, uint32_t R, uint32_t& u) {
    uint32_t a_lo, a_hi, b_lo, b_hi;
    asm("{\n\t.reg .u64 a, b;\n\t"
        "mul.wide.u32 a, %4, %6;\n\t"
        "mul.wide.u32 b, %5, %6;\n\t"
        "mov.b64 {%0, %1}, a;\n\t"
        "mov.b64 {%2, %3}, b;\n\t}"
        : "=r"(a_lo), "=r"(a_hi), "=r"(b_lo), "=r"(b_hi) : "r"(xl), "r"(xh), "r"(R));
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(b_lo), "+r"(b_hi) : "r"(a_hi));
    u = b_hi; xl = a_lo; xh = b_lo;
}

#ifndef GEN_MIN_BLOCKS
#define GEN_MIN_BLOCKS 3   // register cap of the 256-thread generator through __launch_bounds__ (3: 85, 5: 48, 6: 40 registers)
#endif
#define GEN_PASS (8 * GEN_ILP)   // samples per pass of the batched generator: 4 * GEN_ILP stream outputs = 2 * GEN_ILP words x 4 draws
static_assert(GEN_ROWS % GEN_PASS == 0 && GEN_ROWS % (4 * GEN_ILP) == 0, "a staging tile is a whole number of passes");

// THREADS: replicates per block.  256 normally; 64 when a sharded run leaves this rank so few replicates that 256-thread
// blocks would occupy a third of the SMs with two warps per scheduler each — the multiply pipe is per scheduler, so the
// same warps spread one per scheduler finish in half the time.
template <int MODE, bool MISS, int THREADS = GEN_THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 256 ? GEN_MIN_BLOCKS : 1)) k3_gen_kernel(uint64_t seed, int64_t rep_base, int64_t n_reps, int32_t P, int32_t n_case,
                                                             int32_t n_ctrl, const uint8_t* __restrict__ label,
                                                             const uint32_t* __restrict__ perms, uint32_t* __restrict__ Xc32,
                                                             uint32_t* __restrict__ Xm32, int64_t Rw32, int force_replay = 0,
                                                             unsigned long long* __restrict__ replay_err = nullptr, int64_t blk0 = 0) {
    __shared__ uint32_t sc[GEN_ROWS][THREADS / 32], sm[MISS ? GEN_ROWS : 1][THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t gblk = (int64_t)blockIdx.x + blk0;      // blk0: a rank of a sharded run generates only its own slice of the tile's replicates
    const int64_t rep_local = gblk * THREADS + tid;
    const bool active = rep_local < n_reps;
    const uint64_t rep = (uint64_t)(rep_base + rep_local);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), r0 = (uint32_t)rep, r1 = (uint32_t)(rep >> 32);
    uint32_t rem_case = (uint32_t)n_case, rem_ctrl = (uint32_t)n_ctrl, q = 0;
    if (MODE != 1 && !MISS && !active) rem_case = 0;     // padding replicates of the last block never draw a case: bits stay 0
    const uint32_t* prow = (MODE == 1 && active) ? perms + rep_local * (int64_t)P : nullptr;
    uint32_t range = (uint32_t)P;                         // samples not yet labelled (Philox modes)
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;              // spec v3: the replicate's xoshiro128++ state
    if (MODE == 0) {
        uint32_t sw[4];
        philox4x32_10(0u, r0, r1, 2u, k0, k1, sw);
        s0 = sw[0]; s1 = sw[1]; s2 = sw[2]; s3 = sw[3];
        if ((s0 | s1 | s2 | s3) == 0) s0 = 1;
    }
    // replay mode: the caller's rows are checked while they are consumed — an index >= P is clamped and flagged (it would be an
    // out-of-bounds read of label[]), and a row whose index sum / sum of squares is not a permutation's is flagged as well (a
    // repeated index changes the replicate's case / control totals).  pb_run_assoc returns PB_ERR_INVALID on either.
    uint64_t psum = 0, psq = 0;
    bool pbad = false;

    // label of one draw u (Philox modes) -> one 32-replicate word per warp and matrix
    auto emit = [&](uint32_t u, int row) {
        uint32_t lab;
        if (!MISS) {                      // no missing labels: whatever is not a case is a control
            lab = u < rem_case ? 1u : 0u;
            rem_case -= lab;
        } else if (u < rem_case) { lab = 1; rem_case--; }
        else if (u < rem_case + rem_ctrl) { lab = 0; rem_ctrl--; }
        else lab = 2;
        const uint32_t bc = !MISS ? __ballot_sync(0xffffffffu, lab != 0) : __ballot_sync(0xffffffffu, active && lab == 1);
        if (lane == 0) sc[row][warp] = bc;
        if (MISS) {
            const uint32_t bm = __ballot_sync(0xffffffffu, active && lab == 2);
            if (lane == 0) sm[row][warp] = bm;
        }
    };

    // spec v1 / replay: one sample
    auto draw = [&](uint32_t x, int i, int row) {
        if (MODE == 2) {
            uint64_t prod = (uint64_t)x * range;
            uint32_t lo = (uint32_t)prod;
            if (lo < range) {  // taken with probability < range / 2^32
                const uint32_t t = (0u - range) % range;
                while (lo < t) {
                    uint32_t rw[4];
                    philox4x32_10(q >> 2, r0, r1, 1u, k0, k1, rw);
                    const uint32_t y = (q & 2) ? ((q & 1) ? rw[3] : rw[2]) : ((q & 1) ? rw[1] : rw[0]);
                    q++;
                    prod = (uint64_t)y * range;
                    lo = (uint32_t)prod;
                }
            }
            range--;
            emit((uint32_t)(prod >> 32), row);
        } else {
            uint32_t idx = active ? prow[i] : 0u;
            if (idx >= (uint32_t)P) { pbad = true; idx = 0; }
            psum += idx; psq += (uint64_t)idx * idx;
            const uint32_t lab = active ? label[idx] : 0;
            const uint32_t bc = __ballot_sync(0xffffffffu, active && lab == 1);
            if (lane == 0) sc[row][warp] = bc;
            if (MISS) {
                const uint32_t bm = __ballot_sync(0xffffffffu, active && lab == 2);
                if (lane == 0) sm[row][warp] = bm;
            }
        }
    };

    // spec v3: one 64-bit word -> `count` (1..4, warp-uniform) samples, exact rejection
    auto draw_word = [&](uint32_t xl, uint32_t xh, int row, int count) {
        uint32_t us[4] = {0, 0, 0, 0};
        for (;;) {
            uint32_t l = xl, h = xh, R = range;
            uint64_t prod = 1;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (j < count) { mul64x32(l, h, R, us[j]); prod *= R; R--; }
            const uint64_t lo = ((uint64_t)h << 32) | l;
            if (lo < prod) {                               // probability < prod / 2^64
                const uint64_t t = (0ull - prod) % prod;   // 2^64 mod prod
                if (lo < t) {
                    uint32_t rw[4];
                    philox4x32_10(q >> 1, r0, r1, 1u, k0, k1, rw);
                    xl = (q & 1) ? rw[2] : rw[0]; xh = (q & 1) ? rw[3] : rw[1];
                    q++;
                    continue;
                }
            }
            break;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < count) { range--; emit(us[j], row + j); }
    };

    for (int i0 = 0; i0 < P; i0 += GEN_ROWS) {
        const int rows = min(GEN_ROWS, P - i0);
        if (MODE == 0) {
            for (int ii = 0; ii < rows; ii += GEN_PASS) {
                if (!MISS && ii + GEN_PASS <= rows) {
                    // Full pass, binary labels: the draws run without their per-word rejection test.  A retry is needed only if
                    // some word's final low part is below its range product (probability < prod / 2^64 per word): the pass keeps
                    // the minimum high half of the low parts, one vote checks it against the high half of the pass's largest
                    // product (a superset of the true condition), and the rare hit replays the whole pass through the exact path
                    // from the saved state (same words are overwritten, so how often the vote fires cannot change a result).
                    const uint32_t rem0 = rem_case, range0 = range;
                    const uint32_t t0 = s0, t1 = s1, t2 = s2, t3 = s3;
                    uint32_t w[GEN_ILP][4];
#pragma unroll
                    for (int q4 = 0; q4 < GEN_ILP; q4++) {
#pragma unroll
                        for (int k = 0; k < 4; k++) w[q4][k] = xoshiro128pp_next(s0, s1, s2, s3);
                    }
                    // Two phases, so that a lone warp (a sharded run leaves one or two per scheduler) is not one long dependent
                    // chain: first every draw value — the 2*GEN_ILP words are independent multiply chains, advanced in step —
                    // then the labels, whose only serial dependency is the case counter.
                    uint32_t us[GEN_PASS];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
#pragma unroll
                        for (int wi = 0; wi < 2 * GEN_ILP; wi++)
                            mul64x32(w[wi >> 1][2 * (wi & 1)], w[wi >> 1][2 * (wi & 1) + 1], range0 - (uint32_t)(4 * wi + j), us[4 * wi + j]);
                    }
                    uint32_t mn = ~0u;              // smallest HIGH half of the words' final low parts: a conservative test is enough
#pragma unroll
                    for (int wi = 0; wi < 2 * GEN_ILP; wi++) mn = min(mn, w[wi >> 1][2 * (wi & 1) + 1]);
                    range = range0 - GEN_PASS;
#pragma unroll
                    for (int k = 0; k < GEN_PASS; k++) {
                        uint32_t bc;
                        asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %0;\n\t@q add.u32 %0, %0, -1;\n\tvote.sync.ballot.b32 %1, q, 0xffffffff;\n\t}"
                                     : "+r"(rem_case), "=r"(bc) : "r"(us[k]));
                        sc[ii + k][warp] = bc;   // every lane stores the same word: no lane test, one broadcast store
                    }
                    const uint64_t prodmax = (uint64_t)range0 * (range0 - 1) * (uint64_t)(range0 - 2) * (range0 - 3);   // range0 >= GEN_PASS; < 2^64 for P < 65536
                    if (__any_sync(0xffffffffu, mn <= (uint32_t)(prodmax >> 32)) || force_replay) {   // force_replay: test hook (PB_GEN_FORCE_REPLAY)
                        rem_case = rem0; range = range0;
                        s0 = t0; s1 = t1; s2 = t2; s3 = t3;       // cold path: the stream is run again from the pass's first output
#pragma unroll 1
                        for (int wi = 0; wi < 2 * GEN_ILP; wi++) {
                            const uint32_t xl = xoshiro128pp_next(s0, s1, s2, s3), xh = xoshiro128pp_next(s0, s1, s2, s3);
                            draw_word(xl, xh, ii + 4 * wi, 4);
                        }
                    }
                } else {
#pragma unroll 1
                    for (int wi = 0; wi < 2 * GEN_ILP; wi++) {
                        const int row = ii + 4 * wi;
                        const int count = min(4, rows - row);     // block-uniform
                        if (count > 0) {
                            const uint32_t xl = xoshiro128pp_next(s0, s1, s2, s3), xh = xoshiro128pp_next(s0, s1, s2, s3);
                            draw_word(xl, xh, row, count);
                        }
                    }
                }
            }
        } else {
            for (int ii = 0; ii < rows; ii += 4 * GEN_ILP) {
                uint32_t w[GEN_ILP][4];
#pragma unroll
                for (int q4 = 0; q4 < GEN_ILP; q4++) {
                    w[q4][0] = w[q4][1] = w[q4][2] = w[q4][3] = 0;
                    if (MODE == 2) philox4x32_10((uint32_t)((i0 + ii) >> 2) + q4, r0, r1, 0u, k0, k1, w[q4]);
                }
                if (ii + 4 * GEN_ILP <= rows) {        // full group: no per-sample bounds checks (block-uniform)
#pragma unroll
                    for (int jj = 0; jj < 4 * GEN_ILP; jj++) draw(w[jj >> 2][jj & 3], i0 + ii + jj, ii + jj);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4 * GEN_ILP; jj++)
                        if (ii + jj < rows) draw(w[jj >> 2][jj & 3], i0 + ii + jj, ii + jj);
                }
            }
        }
        __syncthreads();
        constexpr int WORDS = THREADS / 32;
        for (int t = tid; t < rows * WORDS; t += THREADS) {
            const int row = t / WORDS, wd = t % WORDS;
            const int64_t o = (int64_t)(i0 + row) * Rw32 + gblk * WORDS + wd;
            Xc32[o] = sc[row][wd];
            if (MISS) Xm32[o] = sm[row][wd];
        }
        __syncthreads();
    }
    if (MODE == 1 && active && replay_err) {
        const uint64_t n = (uint64_t)P;
        const uint64_t want_sum = n * (n - 1) / 2;
        // (n-1) n (2n-1) / 6 without overflow for n < 2^21: one of n, n-1 is even, one of n-1, n, 2n-1 is a multiple of 3
        uint64_t a = n - 1, b = n, c = 2 * n - 1;
        if (a % 2 == 0) a /= 2; else b /= 2;
        if (a % 3 == 0) a /= 3; else if (b % 3 == 0) b /= 3; else c /= 3;
        if (pbad) atomicOr(replay_err, 1ull);
        else if (psum != want_sum || psq != a * b * c) atomicOr(replay_err, 2ull);
    }
}

// ---------------------------------------------------------------------------------------------------
// K3b: replicate counting
// ---------------------------------------------------------------------------------------------------
struct K3Params {
    const AlleleMeta* __restrict__ alleles;
    int64_t n_slots;
    const int32_t* __restrict__ csr;
    const double* __restrict__ pobs;      // [n_slots]
    const double* __restrict__ ptable;    // triangular
    const int32_t* __restrict__ ktau;     // [n_max+1]
    const double* __restrict__ tau;
    const uint64_t* __restrict__ Xc;
    const uint64_t* __restrict__ Xm;      // nullptr when the pheno file has no missing labels
    int64_t Rw;                           // pitch in 64-bit words
    int64_t rep_base, n_reps;             // this tile
    uint32_t* __restrict__ r;
    unsigned long long* __restrict__ maxT;  // [m] double bits, global replicate index
};

template <int NB>
__device__ __forceinline__ uint64_t ge_const(const uint64_t (&c)[NB], int T) {
    // bit-sliced (k >= T) for a warp-uniform constant 0 < T < 2^NB
    uint64_t gt = 0, eq = ~0ull;
#pragma unroll
    for (int j = NB - 1; j >= 0; j--) {
        if ((T >> j) & 1) eq &= c[j];
        else { gt |= eq & c[j]; eq &= ~c[j]; }
    }
    return gt | eq;
}

template <int NB>
__device__ __forceinline__ int extract(const uint64_t (&c)[NB], int b) {
    int v = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) v |= (int)((c[j] >> b) & 1ull) << j;
    return v;
}

// One warp sweeps ONE allele over every replicate word of the tile: lane l owns words l, l+32, ...
// The allele's sample rows are streamed (coalesced 256 B per row per step, L2-resident matrix),
// r is accumulated in registers and written once — no atomics on r.
template <int NB>
__device__ __forceinline__ void sweep_allele(const K3Params& p, const AlleleMeta& m, int64_t slot, int lane, double tau, int w0 = -1,
                                             int wstride = 32, int64_t wend = -1) {
    // default: one warp owns the allele (lane = first word, stride 32, plain add); otherwise several warps share it
    // (a thread block striding over all words, or warps owning word segments [w0 - lane, wend)) and r is added atomically
    constexpr int NREG = (NB <= 3) ? ((1 << NB) - 1) : 1;   // small lists live in registers
    const int n = m.n;
    const int32_t* __restrict__ ev = p.csr + m.off;
    int32_t evr[NREG];
    if (NB <= 3) {
#pragma unroll
        for (int e = 0; e < NREG; e++) evr[e] = (e < n) ? __ldg(ev + e) : 0;
    }
    const bool has_miss = p.Xm != nullptr;
    const bool fast = m.flags & PB_AF_FAST;
    const double pobs = p.pobs[slot];
    const int kstar = m.kstar;
    const int kt = p.ktau[n];
    const double* __restrict__ row_n = p.ptable + (int64_t)n * (n + 1) / 2;
    const int64_t words = (p.n_reps + 63) >> 6;
    const int64_t wstop = (wend < 0 || wend > words) ? words : wend;
    const int tail = (int)(p.n_reps & 63);
    unsigned int rcount = 0;
#pragma unroll 2
    for (int64_t w = (w0 < 0 ? lane : w0); w < wstop; w += wstride) {
        const uint64_t valid = (w == words - 1 && tail) ? ((1ull << tail) - 1) : ~0ull;
        uint64_t c[NB], u[NB];
#pragma unroll
        for (int j = 0; j < NB; j++) { c[j] = 0; u[j] = 0; }
        if (NB <= 3) {
#pragma unroll
            for (int e = 0; e < NREG; e++) {
                if (e < n) {
                    const int64_t o = (int64_t)evr[e] * p.Rw + w;
                    uint64_t x = __ldg(p.Xc + o);
#pragma unroll
                    for (int j = 0; j < NB; j++) { const uint64_t t = c[j] & x; c[j] ^= x; x = t; }
                    if (has_miss) {
                        uint64_t y = __ldg(p.Xm + o);
#pragma unroll
                        for (int j = 0; j < NB; j++) { const uint64_t t = u[j] & y; u[j] ^= y; y = t; }
                    }
                }
            }
        } else {
            for (int e = 0; e < n; e++) {
                const int64_t o = (int64_t)__ldg(ev + e) * p.Rw + w;
                uint64_t x = __ldg(p.Xc + o);
#pragma unroll
                for (int j = 0; j < NB; j++) { const uint64_t t = c[j] & x; c[j] ^= x; x = t; }
                if (has_miss) {
                    uint64_t y = __ldg(p.Xm + o);
#pragma unroll
                    for (int j = 0; j < NB; j++) { const uint64_t t = u[j] & y; u[j] ^= y; y = t; }
                }
            }
        }
        uint64_t exc = 0;
        if (has_miss) {
#pragma unroll
            for (int j = 0; j < NB; j++) exc |= u[j];
        }
        if (!fast) exc = ~0ull;
        exc &= valid;
        const uint64_t fastmask = valid & ~exc;
        const uint64_t ge = (kstar <= 0) ? ~0ull : (kstar > n ? 0ull : ge_const<NB>(c, kstar));
        rcount += (unsigned int)__popcll(ge & fastmask);
        const uint64_t cand = (kt <= 0) ? ~0ull : (kt > n ? 0ull : ge_const<NB>(c, kt));
        uint64_t todo = (cand & fastmask) | exc;
        while (todo) {
            const int b = __ffsll((long long)todo) - 1;
            todo &= todo - 1;
            const int k = extract<NB>(c, b);
            double pv;
            if ((exc >> b) & 1ull) {
                const int nr = n - (has_miss ? extract<NB>(u, b) : 0);
                pv = p.ptable[(int64_t)nr * (nr + 1) / 2 + k];
                rcount += (pv <= pobs) ? 1u : 0u;   // HC.java:4239 / 4273
            } else {
                pv = row_n[k];
            }
            if (pv <= tau) atomicMin(p.maxT + (p.rep_base + w * 64 + b), (unsigned long long)__double_as_longlong(pv));
        }
    }
    rcount = __reduce_add_sync(0xffffffffu, rcount);
    if (lane == 0 && rcount) {
        if (w0 < 0) p.r[slot] += rcount;   // this warp is the slot's only writer in this launch
        else atomicAdd(p.r + slot, rcount);
    }
}

// ---- tight sweeps for the common case ---------------------------------------------------------------
// Preconditions (checked by k4_classify_kernel): PB_AF_FAST, no missing labels in the pheno file, no maxT
// candidate possible (ktau[n] > n), 1 <= kstar <= N == m.n.  Padding replicate bits of the matrix are zero,
// and k* >= 1, so they never count.  r += #{replicates : k_r >= k*}.
template <int N, int K>
__device__ __forceinline__ uint64_t at_least(const uint64_t (&x)[N]) {
    if (N == 1) return x[0];
    if (N == 2) return K == 1 ? (x[0] | x[1]) : (x[0] & x[1]);
    if (N == 3) return K == 1 ? (x[0] | x[1] | x[2]) : K == 2 ? ((x[0] & x[1]) | (x[2] & (x[0] | x[1]))) : (x[0] & x[1] & x[2]);
    return 0;
}

// The sweeps read the label matrix with 16-byte loads: a lane owns the word PAIR (2q, 2q+1), q = lane, lane + 32, ...
// Measured on this B200 (tools/microbench/l2_peak, profiles/r2_l2_peak.json): an L2-resident stream moves 15.0 TB/s with
// 8 bytes per lane and 26.6 TB/s with 16 — the request rate, not the sector rate, is what an 8-byte sweep runs into
// (round 1: 16.7 TB/s at C2).  Rows are 16-byte aligned (pitch = a multiple of 4 words, count tiles start at multiples of
// 32 words) and the word behind an odd word count is zero padding written by the generator, which never counts (k* >= 1).
#ifndef K3_SWEEP_UNROLL
#define K3_SWEEP_UNROLL 4   // 16-byte loads per row in flight (round 1: 8 x 8 bytes)
#endif
#ifndef K3_TIGHT_BLOCKS
#define K3_TIGHT_BLOCKS 4
#endif
__device__ __forceinline__ ulonglong2 ldg_pair(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}

template <int N, int K>
__device__ __forceinline__ unsigned int sweep_fixed(const K3Params& p, const AlleleMeta& m, int lane, int words) {
    const ulonglong2* __restrict__ rp[N];
#pragma unroll
    for (int e = 0; e < N; e++) rp[e] = reinterpret_cast<const ulonglong2*>(p.Xc + (int64_t)__ldg(p.csr + m.off + e) * p.Rw) + lane;
    unsigned int rc = 0;
    const int pairs = (words + 1) >> 1;
    constexpr int UNR = K3_SWEEP_UNROLL;
#pragma unroll UNR
    for (int q = lane, off = 0; q < pairs; q += 32, off += 32) {
        uint64_t a[N], b[N];
#pragma unroll
        for (int e = 0; e < N; e++) { const ulonglong2 v = ldg_pair(rp[e] + off); a[e] = v.x; b[e] = v.y; }
        rc += (unsigned int)__popcll(at_least<N, K>(a)) + (unsigned int)__popcll(at_least<N, K>(b));
    }
    return rc;
}

// 4 <= n <= 7: bit-sliced 3-bit count, k >= k* with a warp-uniform k*
__device__ __forceinline__ unsigned int sweep_upto7(const K3Params& p, const AlleleMeta& m, int lane, int words) {
    const int n = m.n, kstar = m.kstar;
    const ulonglong2* __restrict__ rp[7];
#pragma unroll
    for (int e = 0; e < 7; e++) rp[e] = reinterpret_cast<const ulonglong2*>(p.Xc + (int64_t)((e < n) ? __ldg(p.csr + m.off + e) : 0) * p.Rw) + lane;
    unsigned int rc = 0;
    const int pairs = (words + 1) >> 1;
#pragma unroll 1
    for (int q = lane, off = 0; q < pairs; q += 32, off += 32) {
        uint64_t c[3] = {0, 0, 0}, d[3] = {0, 0, 0};
#pragma unroll
        for (int e = 0; e < 7; e++) {
            if (e < n) {
                const ulonglong2 v = ldg_pair(rp[e] + off);
                uint64_t x = v.x, y = v.y;
#pragma unroll
                for (int j = 0; j < 3; j++) { const uint64_t t = c[j] & x; c[j] ^= x; x = t; }
#pragma unroll
                for (int j = 0; j < 3; j++) { const uint64_t t = d[j] & y; d[j] ^= y; y = t; }
            }
        }
        rc += (unsigned int)__popcll(ge_const<3>(c, kstar)) + (unsigned int)__popcll(ge_const<3>(d, kstar));
    }
    return rc;
}

// missing labels in the pheno file, n <= 7: a replicate sees n_r = n - j labelled leaves (j = its missing draws
// among the allele's samples) and counts when k_r >= k*_j.  Bit-sliced k (cases) and j (missing), one pass per j.
template <int NB>
__device__ __forceinline__ uint64_t miss_word(const uint64_t (&c)[NB], const uint64_t (&u)[NB], int n, uint32_t kpack) {
    uint64_t acc = 0;
    for (int j = 0; j <= n; j++) {                       // warp-uniform
        const int kj = (int)((kpack >> (4 * j)) & 15u);  // k* for n_r = n - j
        uint64_t eq = ~0ull;                             // (missing count == j)
#pragma unroll
        for (int b = 0; b < NB; b++) eq &= ((j >> b) & 1) ? u[b] : ~u[b];
        const uint64_t ge = (kj <= 0) ? ~0ull : (kj > n - j ? 0ull : ge_const<NB>(c, kj));
        acc |= eq & ge;
    }
    return acc;
}

template <int NB>
__device__ __forceinline__ unsigned int sweep_miss(const K3Params& p, const AlleleMeta& m, int lane, int words) {
    constexpr int NMAX = (1 << NB) - 1;
    const int n = m.n;
    const uint32_t kpack = m.kpack;
    const ulonglong2* __restrict__ rc_[NMAX];
    const ulonglong2* __restrict__ rm_[NMAX];
#pragma unroll
    for (int e = 0; e < NMAX; e++) {
        const int64_t o = (int64_t)((e < n) ? __ldg(p.csr + m.off + e) : 0) * p.Rw;
        rc_[e] = reinterpret_cast<const ulonglong2*>(p.Xc + o) + lane; rm_[e] = reinterpret_cast<const ulonglong2*>(p.Xm + o) + lane;
    }
    const int tail = (int)(p.n_reps & 63);
    const int pairs = (words + 1) >> 1;
    unsigned int rc = 0;
#pragma unroll 1
    for (int q = lane, off = 0; q < pairs; q += 32, off += 32) {
        uint64_t c0[NB], u0[NB], c1[NB], u1[NB];
#pragma unroll
        for (int j = 0; j < NB; j++) { c0[j] = 0; u0[j] = 0; c1[j] = 0; u1[j] = 0; }
#pragma unroll
        for (int e = 0; e < NMAX; e++) {
            if (e < n) {
                const ulonglong2 xv = ldg_pair(rc_[e] + off), yv = ldg_pair(rm_[e] + off);
                uint64_t x = xv.x, y = yv.x;
#pragma unroll
                for (int j = 0; j < NB; j++) { const uint64_t t = c0[j] & x; c0[j] ^= x; x = t; }
#pragma unroll
                for (int j = 0; j < NB; j++) { const uint64_t t = u0[j] & y; u0[j] ^= y; y = t; }
                x = xv.y; y = yv.y;
#pragma unroll
                for (int j = 0; j < NB; j++) { const uint64_t t = c1[j] & x; c1[j] ^= x; x = t; }
#pragma unroll
                for (int j = 0; j < NB; j++) { const uint64_t t = u1[j] & y; u1[j] ^= y; y = t; }
            }
        }
        // padding replicates (k = 0, no missing draw) would count wherever k*_0 <= 0: mask the last word's tail and the word behind it
        const int w0 = 2 * q, w1 = 2 * q + 1;
        uint64_t a0 = miss_word<NB>(c0, u0, n, kpack), a1 = miss_word<NB>(c1, u1, n, kpack);
        if (w0 == words - 1 && tail) a0 &= (1ull << tail) - 1;
        if (w1 >= words) a1 = 0; else if (w1 == words - 1 && tail) a1 &= (1ull << tail) - 1;
        rc += (unsigned int)__popcll(a0) + (unsigned int)__popcll(a1);
    }
    return rc;
}

// n <= 7, every label present, maxT candidates possible for k >= kt (1 <= kt <= n): the tight sweep plus a scalar
// look-up + atomic min for the (rare) candidate replicates.  k* may be anything (<= 0: all count, > n: none).
__device__ __forceinline__ unsigned int sweep_upto7_cand(const K3Params& p, const AlleleMeta& m, int lane, int words, int kt, double tau,
                                                         int wbeg, int wend) {
    const int n = m.n, kstar = m.kstar;
    const uint64_t* __restrict__ rp[7];
#pragma unroll
    for (int e = 0; e < 7; e++) rp[e] = p.Xc + (int64_t)((e < n) ? __ldg(p.csr + m.off + e) : 0) * p.Rw + lane;
    const double* __restrict__ row_n = p.ptable + (int64_t)n * (n + 1) / 2;
    const int tail = (int)(p.n_reps & 63);
    unsigned int rc = 0;
    for (int w = wbeg + lane, off = wbeg; w < wend; w += 32, off += 32) {   // wbeg is a multiple of 32, wend <= words
        uint64_t c[3] = {0, 0, 0};
#pragma unroll
        for (int e = 0; e < 7; e++) {
            if (e < n) {
                uint64_t x = __ldg(rp[e] + off);
#pragma unroll
                for (int j = 0; j < 3; j++) { const uint64_t t = c[j] & x; c[j] ^= x; x = t; }
            }
        }
        uint64_t ge = (kstar <= 0) ? ~0ull : (kstar > n ? 0ull : ge_const<3>(c, kstar));
        if (w == words - 1 && tail) ge &= (1ull << tail) - 1;        // k* <= 0 would count the padding replicates
        rc += (unsigned int)__popcll(ge);
        uint64_t cand = ge_const<3>(c, kt);                           // kt >= 1: padding bits (k = 0) are never candidates
        while (cand) {
            const int b = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            const double pv = row_n[extract<3>(c, b)];
            if (pv <= tau) atomicMin(p.maxT + (p.rep_base + (int64_t)w * 64 + b), (unsigned long long)__double_as_longlong(pv));
        }
    }
    return rc;
}

#define K3_THREADS 256
// work list 1: short lists (n < 8) on the tight path.  Few registers -> many resident warps: the sweep is bound by the
// latency of the label-matrix loads.  MISS = the pheno file has missing labels.
template <bool MISS>
__global__ void __launch_bounds__(K3_THREADS, MISS ? 4 : K3_TIGHT_BLOCKS) k3_count_tight_kernel(const K3Params p, const int32_t* __restrict__ list, const unsigned int* __restrict__ n_list_dev) {
    const int64_t n_list = *n_list_dev;   // list lengths stay on the device (k4_classify_kernel): no host round trip
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    const int words = (int)((p.n_reps + 63) >> 6);
    for (int64_t i = gwarp; i < n_list; i += nwarps) {
        const int64_t slot = list[i];
        const AlleleMeta m = p.alleles[slot];
        const int n = m.n, ks = m.kstar;
        unsigned int rc;
        if (MISS) {
            rc = n < 2 ? sweep_miss<1>(p, m, lane, words) : n < 4 ? sweep_miss<2>(p, m, lane, words) : sweep_miss<3>(p, m, lane, words);
        } else {
            if (ks <= 0) rc = (lane == 0) ? (unsigned int)p.n_reps : 0u;   // every replicate counts
            else if (ks > n) rc = 0;
            else if (n == 1) rc = sweep_fixed<1, 1>(p, m, lane, words);
            else if (n == 2) rc = ks == 1 ? sweep_fixed<2, 1>(p, m, lane, words) : sweep_fixed<2, 2>(p, m, lane, words);
            else if (n == 3) rc = ks == 1 ? sweep_fixed<3, 1>(p, m, lane, words) : ks == 2 ? sweep_fixed<3, 2>(p, m, lane, words)
                                                                                           : sweep_fixed<3, 3>(p, m, lane, words);
            else rc = sweep_upto7(p, m, lane, words);
        }
        rc = __reduce_add_sync(0xffffffffu, rc);
        if (lane == 0 && rc) p.r[slot] += rc;
    }
}

// work list 3: short lists with possible maxT candidates (tight sweep + scalar candidate look-ups) or that need the
// general path (non-monotone table rows, missing labels with candidates, test hooks)
// These alleles are few (thousands at C2) and each sweep is long, so one task = one allele x one segment of the
// replicate words: K3_SMALL_SEGS warps share an allele and add their partial r atomically.
#define K3_SMALL_SEGS 8
// (a) candidate short lists on the tight path.  Kept apart from the general path below so that it stays at 64 registers and
// its blocks fit on an SM next to the tight sweep's (it runs on the side stream, concurrently with k3_count_tight_kernel).
__global__ void __launch_bounds__(K3_THREADS, 4) k3_count_cand_kernel(const K3Params p, const int32_t* __restrict__ list, const unsigned int* __restrict__ n_list_dev) {
    const int64_t n_tasks = (int64_t)*n_list_dev * K3_SMALL_SEGS;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    const double tau = *p.tau;
    const int words = (int)((p.n_reps + 63) >> 6);
    const int seg = ((words + K3_SMALL_SEGS - 1) / K3_SMALL_SEGS + 31) & ~31;   // whole warp steps
    for (int64_t t = gwarp; t < n_tasks; t += nwarps) {
        const int wbeg = (int)(t % K3_SMALL_SEGS) * seg, wend = min(words, wbeg + seg);
        if (wbeg >= words) continue;
        const int64_t slot = list[t / K3_SMALL_SEGS];
        const AlleleMeta m = p.alleles[slot];
        if (!(m.flags & PB_AF_TIGHT_CAND)) continue;
        unsigned int rc = sweep_upto7_cand(p, m, lane, words, p.ktau[m.n], tau, wbeg, wend);
        rc = __reduce_add_sync(0xffffffffu, rc);
        if (lane == 0 && rc) atomicAdd(p.r + slot, rc);
    }
}

// (b) whatever needs the general path
__global__ void __launch_bounds__(K3_THREADS, 2) k3_count_small_kernel(const K3Params p, const int32_t* __restrict__ list, const unsigned int* __restrict__ n_list_dev) {
    const int64_t n_tasks = (int64_t)*n_list_dev * K3_SMALL_SEGS;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    const double tau = *p.tau;
    const int words = (int)((p.n_reps + 63) >> 6);
    const int seg = ((words + K3_SMALL_SEGS - 1) / K3_SMALL_SEGS + 31) & ~31;
    for (int64_t t = gwarp; t < n_tasks; t += nwarps) {
        const int wbeg = (int)(t % K3_SMALL_SEGS) * seg, wend = min(words, wbeg + seg);
        if (wbeg >= words) continue;
        const int64_t slot = list[t / K3_SMALL_SEGS];
        const AlleleMeta m = p.alleles[slot];
        const int n = m.n;
        if (m.flags & PB_AF_TIGHT_CAND) continue;
        if (n < 2) sweep_allele<1>(p, m, slot, lane, tau, wbeg + lane, 32, wend);
        else if (n < 4) sweep_allele<2>(p, m, slot, lane, tau, wbeg + lane, 32, wend);
        else sweep_allele<3>(p, m, slot, lane, tau, wbeg + lane, 32, wend);
    }
}

// work list C: the long-list tail (n >= 8) and whatever needs the general path there.  These are few and long:
// a whole thread block shares one allele (thread = word, stride 256) so that no single warp becomes the long pole.
__global__ void __launch_bounds__(K3_THREADS, 1) k3_count_large_kernel(const K3Params p, const int32_t* __restrict__ list, const unsigned int* __restrict__ n_list_dev) {
    const int64_t n_list = *n_list_dev;
    const int lane = threadIdx.x & 31;
    const double tau = *p.tau;
    for (int64_t i = blockIdx.x; i < n_list; i += gridDim.x) {
        const int64_t slot = list[i];
        const AlleleMeta m = p.alleles[slot];
        const int n = m.n;
        if (n < 16) sweep_allele<4>(p, m, slot, lane, tau, threadIdx.x, K3_THREADS);
        else if (n < 64) sweep_allele<6>(p, m, slot, lane, tau, threadIdx.x, K3_THREADS);
        else if (n < 256) sweep_allele<8>(p, m, slot, lane, tau, threadIdx.x, K3_THREADS);
        else if (n < 4096) sweep_allele<12>(p, m, slot, lane, tau, threadIdx.x, K3_THREADS);
        else sweep_allele<17>(p, m, slot, lane, tau, threadIdx.x, K3_THREADS);
    }
}

// work list A: n <= 1 alleles of the tight class need no sweep at all — the per-sample column popcount is shared
__global__ void k3_colpop_kernel(const uint64_t* __restrict__ Xc, const uint64_t* __restrict__ Xm, int64_t Rw, int64_t words, int32_t P,
                                 uint32_t* __restrict__ colpop /* [2][P]: cases, missing */) {
    __shared__ unsigned int sh[2][4];
    const uint64_t* rc = Xc + (int64_t)blockIdx.x * Rw;
    const uint64_t* rm = Xm ? Xm + (int64_t)blockIdx.x * Rw : nullptr;
    unsigned int c = 0, u = 0;
    for (int64_t w = threadIdx.x; w < words; w += blockDim.x) {
        c += (unsigned int)__popcll(rc[w]);
        if (rm) u += (unsigned int)__popcll(rm[w]);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    u = __reduce_add_sync(0xffffffffu, u);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = c; sh[1][threadIdx.x >> 5] = u; }
    __syncthreads();
    if (threadIdx.x == 0) {
        colpop[blockIdx.x] = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
        colpop[P + blockIdx.x] = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
    }
}
__global__ void k3_trivial_kernel(const K3Params p, const int32_t* __restrict__ list, const unsigned int* __restrict__ n_list_dev,
                                  const uint32_t* __restrict__ colpop, int32_t P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)*n_list_dev) return;
    const int64_t slot = list[i];
    const AlleleMeta m = p.alleles[slot];
    unsigned int rc = 0;
    if (p.Xm == nullptr) {
        if (m.kstar <= 0) rc = (unsigned int)p.n_reps;
        else if (m.n == 1 && m.kstar == 1) rc = colpop[p.csr[m.off]];
    } else {
        // missing labels (n <= 1): k*_0 for a labelled draw (n_r = n), k*_1 for a missing one (n_r = 0)
        const int k0 = (int)(m.kpack & 15u), k1 = (int)((m.kpack >> 4) & 15u);
        if (m.n == 0) rc = k0 <= 0 ? (unsigned int)p.n_reps : 0u;
        else {
            const int32_t s = p.csr[m.off];
            const unsigned int pc = colpop[s], pm = colpop[P + s], pt = (unsigned int)p.n_reps - pc - pm;
            rc = (k1 <= 0 ? pm : 0u) + (k0 <= 1 ? pc : 0u) + (k0 <= 0 ? pt : 0u);
        }
    }
    if (rc) p.r[slot] += rc;
}

// ---------------------------------------------------------------------------------------------------
// maxT fallback: replicates whose min could not be certified (> tau) are recomputed exactly
// ---------------------------------------------------------------------------------------------------
__global__ void k3_collect_fallback_kernel(const unsigned long long* __restrict__ maxT, int64_t rep_base, int64_t n_reps,
                                           const double* __restrict__ tau, int32_t* __restrict__ list, unsigned int* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reps) return;
    const double v = __longlong_as_double((long long)maxT[rep_base + i]);
    if (!(v <= *tau)) list[atomicAdd(count, 1u)] = (int32_t)i;
}

__global__ void __launch_bounds__(256) k3_fallback_kernel(const K3Params p, const int32_t* __restrict__ list, const unsigned int* __restrict__ count,
                                                          unsigned long long* __restrict__ total) {
    __shared__ double smin[256];
    const unsigned int n = *count;                      // device-side list length: the kernel is always launched
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(total, (unsigned long long)n);
    for (unsigned int li = blockIdx.x; li < n; li += gridDim.x) {
    const int64_t rl = list[li];
    const int64_t word = rl >> 6;
    const int b = (int)(rl & 63);
    double best = __longlong_as_double(0x7FF0000000000000ll);
    for (int64_t slot = threadIdx.x; slot < p.n_slots; slot += blockDim.x) {
        const AlleleMeta m = p.alleles[slot];
        if (!(m.flags & PB_AF_IN_USE)) continue;
        int k = 0, u = 0;
        for (int e = 0; e < m.n; e++) {
            const int64_t o = (int64_t)p.csr[m.off + e] * p.Rw + word;
            k += (int)((p.Xc[o] >> b) & 1ull);
            if (p.Xm) u += (int)((p.Xm[o] >> b) & 1ull);
        }
        const int nr = m.n - u;
        const double pv = p.ptable[(int64_t)nr * (nr + 1) / 2 + k];
        best = fmin(best, pv);
    }
    smin[threadIdx.x] = best;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) p.maxT[p.rep_base + rl] = (unsigned long long)__double_as_longlong(smin[0]);
    __syncthreads();
    }
}

__global__ void fill_u64_kernel(unsigned long long* p, int64_t n, unsigned long long v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ===================================================================================================
static int ensure(pb_ctx* ctx, void** ptr, int64_t* cap, int64_t need_bytes) {
    if (need_bytes <= *cap) return PB_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    const int64_t bytes = need_bytes + need_bytes / 8 + 256;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        *cap = 0;
        return PB_ERR_OOM;
    }
    *cap = bytes;
    return PB_OK;
}

__global__ void k3_headstart_kernel(unsigned int ns) {
    const unsigned long long t0 = clock64();
    // ~2 cycles per ns at 1.9 GHz; __nanosleep keeps the SM slot idle instead of spinning hot
    while (clock64() - t0 < 2ull * ns) __nanosleep(1000);
}

// replicate tiling of the label matrices + their buffers
struct K3Plan { int64_t tile, ctile, Rw32, Rw; bool has_miss; int nbuf; int64_t buf_words; };  // tile = generation tile, ctile = count tile

static int k3_plan(pb_ctx* ctx, K3Plan* pl) {
    const int64_t m = ctx->cfg.n_replicates;
    pl->has_miss = ctx->n_miss > 0;
    // tile size: K3b streams label-matrix rows from L2 (13 TB/s measured at C2), so a tile's matrices should stay
    // L2-resident (126 MB L2): everything in one tile when that is <= 100 MB, else tiles of ~64 MB
    const int64_t bytes_per_rep = ((int64_t)ctx->P * (pl->has_miss ? 2 : 1) + 7) / 8;
    int64_t budget = (m * bytes_per_rep <= (100ll << 20)) ? (100ll << 20) : (64ll << 20);
    if (const char* e = getenv("PB_K3_TILE_MB")) budget = atoll(e) << 20;
    if (const char* e = getenv("PB_K3_TILE_KB")) budget = atoll(e) << 10;   // test hook: forces many small tiles
    int64_t ctile = budget / bytes_per_rep;
    ctile = (ctile / 2048) * 2048;
    if (ctile < 2048) ctile = 2048;
    const int64_t m_pad = ((m + 255) / 256) * 256;
    if (ctile > m_pad) ctile = m_pad;
    pl->ctile = ctile;
    // generation tile: K3a is one thread per replicate with P sequential draws, so it wants as many replicates per
    // launch as memory allows (<= 2 GB per matrix by default), a whole number of count tiles
    int64_t gen_budget = 2048ll << 20;
    if (const char* e = getenv("PB_K3_GEN_MB")) gen_budget = atoll(e) << 20;
    if (const char* e = getenv("PB_K3_GEN_KB")) gen_budget = atoll(e) << 10;   // test hook
    int64_t tile = (gen_budget / bytes_per_rep / ctile) * ctile;
    if (tile < ctile) tile = ctile;
    if (tile > m_pad) tile = m_pad;
    pl->tile = tile;
    pl->Rw32 = ((tile + 255) / 256) * 8;  // uint32 pitch
    pl->Rw = pl->Rw32 / 2;
    int rc;
    // two matrix buffers when the null is tiled: tile t+1 is generated on the second stream while tile t is counted
    pl->nbuf = (m > tile && ctx->cfg.mode == PB_MODE_PHILOX && !getenv("PB_NO_PREGEN")) ? 2 : 1;
    pl->buf_words = (int64_t)ctx->P * pl->Rw;
    const int64_t need = pl->buf_words * 8 * pl->nbuf;
    void* pc = ctx->d_Xcase;
    if ((rc = ensure(ctx, &pc, &ctx->X_cap, need))) { ctx->d_Xcase = nullptr; return rc; }
    ctx->d_Xcase = (uint64_t*)pc;
    if (pl->has_miss) {
        void* pmx = ctx->d_Xmiss;
        if ((rc = ensure(ctx, &pmx, &ctx->Xm_cap, need))) { ctx->d_Xmiss = nullptr; return rc; }
        ctx->d_Xmiss = (uint64_t*)pmx;
    }
    return PB_OK;
}

static int k3_launch_philox(pb_ctx* ctx, const K3Plan& pl, int64_t base, int64_t nr, cudaStream_t st, int buf) {
    // the grid covers the tile rounded up to 256 replicates: the padding replicates write zero bits, which the 16-byte
    // sweeps of K3b rely on (the word behind an odd word count)
    // slices are cut in units of 256 replicates (4 matrix words), whatever the block size
    const int64_t units = (nr + 255) / 256;
    int64_t u0 = 0, u1 = units;
    // Sharded run with mapped label matrices (pb_comm.cu, "label sharing"): every rank needs the SAME matrix, so each generates
    // 1/world of the tile's replicate columns and pulls the rest from its peers over NVLink instead of drawing all of it.
    const bool shared = pbx_labels_shared(ctx, pl.buf_words * 8 * pl.nbuf);
    int rc;
    if (shared) {
        if ((rc = pbx_labels_begin(ctx, buf, st))) return rc;          // peers have pulled what this buffer held before
        u0 = units * ctx->rank / ctx->world;
        u1 = units * (ctx->rank + 1) / ctx->world;
    }
    const bool timed = base == 0;        // ms_gen_kernel: the generator's own time for the first tile (events on its stream)
    if (timed) PB_CUDA(cudaEventRecord(ctx->ev_gen_t0, st));
    const int force_replay = getenv("PB_GEN_FORCE_REPLAY") ? 1 : 0;   // every group also goes through the exact Lemire path
    uint32_t* xc = (uint32_t*)(ctx->d_Xcase + buf * pl.buf_words);
    uint32_t* xm = pl.has_miss ? (uint32_t*)(ctx->d_Xmiss + buf * pl.buf_words) : nullptr;
    const bool batched = ctx->P < 65536;      // spec v3 (four draws per 64-bit word); spec v1 above that
    // fewer than two warps per scheduler: 64-thread blocks spread them over all SMs (see k3_gen_kernel)
    const bool small = batched && ((u1 - u0) * 8 < (int64_t)ctx->sm_count * 8 || getenv("PB_GEN_SMALL_BLOCKS")) && !getenv("PB_GEN_BIG_BLOCKS");   // (env: test hooks)
    const int threads = small ? 64 : GEN_THREADS;
    const unsigned g = (unsigned)((u1 - u0) * (256 / threads));
    const int64_t blk0 = u0 * (256 / threads);
    if (g == 0) {
    } else if (small && pl.has_miss)
        k3_gen_kernel<0, true, 64><<<g, 64, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label, nullptr, xc, xm, pl.Rw32, force_replay, nullptr, blk0);
    else if (small)
        k3_gen_kernel<0, false, 64><<<g, 64, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label, nullptr, xc, nullptr, pl.Rw32, force_replay, nullptr, blk0);
    else if (pl.has_miss && batched)
        k3_gen_kernel<0, true><<<g, GEN_THREADS, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label, nullptr, xc, xm, pl.Rw32, force_replay, nullptr, blk0);
    else if (batched)
        k3_gen_kernel<0, false><<<g, GEN_THREADS, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label, nullptr, xc, nullptr, pl.Rw32, force_replay, nullptr, blk0);
    else if (pl.has_miss)
        k3_gen_kernel<2, true><<<g, GEN_THREADS, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label, nullptr, xc, xm, pl.Rw32, 0, nullptr, blk0);
    else
        k3_gen_kernel<2, false><<<g, GEN_THREADS, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label, nullptr, xc, nullptr, pl.Rw32, 0, nullptr, blk0);
    if (timed) { PB_CUDA(cudaEventRecord(ctx->ev_gen_t1, st)); ctx->gen_timed = true; }
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    ctx->tm.gen_shard_frac = shared ? 1.0f / (float)ctx->world : 1.0f;
    if (shared && (rc = pbx_labels_exchange(ctx, buf, (int64_t)buf * pl.buf_words, ctx->P, pl.Rw, units, 4, pl.has_miss, st))) return rc;
    return PB_OK;
}

// Called by pb_run_events right after the K1 sweep is enqueued: the label matrix of the first replicate tile
// depends only on the phenotype table and the seed, so it is generated on a second stream while the (small,
// latency-bound) K1 tail, K2 and the p-table kernels run; pb_run_assoc picks it up through an event.
// immediately: called while the alignment is still uploading (first tile of an in-order one-plane upload): nothing to wait
// for and nothing to race against, the generator simply starts.
int pbk_pregen_labels(pb_ctx* ctx, bool immediately) {
    ctx->pregen_valid = false;
    if (ctx->cfg.mode != PB_MODE_PHILOX || ctx->P == 0 || ctx->n_case + ctx->n_ctrl == 0) return PB_OK;
    if (getenv("PB_NO_PREGEN")) return PB_OK;
    K3Plan pl;
    int rc;
    if ((rc = k3_plan(ctx, &pl))) return rc;
    const int64_t m = ctx->cfg.n_replicates;
    const int64_t nr = m < pl.tile ? m : pl.tile;
    // one wave of generator blocks -> the high-priority stream (resident before the tail); many waves -> the low-priority one
    const bool many_waves = (nr + 31) / 32 > (int64_t)ctx->sm_count * 40 && !getenv("PB_GEN_ALWAYS_HIGH");   // > 40 generator warps per SM
    cudaStream_t gs = many_waves ? ctx->stream2_lo : ctx->stream2;
    // default: the generator starts when K1 has finished.  Experiments: PB_PREGEN_EARLY — it is ordered behind the previous
    // step only (ev[0] is recorded just before the K1 launch), so the low-priority stream may pick up SM slots while
    // K1's last wave drains; PB_PREGEN_FIRST — pb_run_events calls this BEFORE it enqueues the sweep (ordered behind
    // ev[12], the start of the call), so the generator's blocks are resident first and the sweep fills the rest of each SM.
    const bool early = getenv("PB_PREGEN_EARLY") != nullptr, first = getenv("PB_PREGEN_FIRST") != nullptr;
    if (immediately) {
        PB_CUDA(cudaEventRecord(ctx->ev_gen_started, gs));
    } else if (first) {
        PB_CUDA(cudaStreamWaitEvent(gs, ctx->ev[12], 0));
        PB_CUDA(cudaEventRecord(ctx->ev_gen_started, gs));
    } else if (early) {
        PB_CUDA(cudaStreamWaitEvent(gs, ctx->ev[0], 0));
    } else {
        PB_CUDA(cudaEventRecord(ctx->ev_k1_done, ctx->stream));
        PB_CUDA(cudaStreamWaitEvent(gs, ctx->ev_k1_done, 0));
        PB_CUDA(cudaEventRecord(ctx->ev_gen_started, gs));
    }
    if ((rc = k3_launch_philox(ctx, pl, 0, nr, gs, 0))) return rc;
    PB_CUDA(cudaEventRecord(ctx->ev_gen_done, gs));
    // The generator is ONE wave of long-running blocks: it finishes ~0.65 ms after its LAST block starts, and the main
    // stream's tail kernels (queued right behind K1) would otherwise win the race for the SM slots K1 frees.  Hold the
    // main stream for a few microseconds behind the K1 dependency so that the generator's blocks become resident first;
    // the short tail kernels then fill in around them.  (Measured at C2: 2.76 ms per step without the head start, 2.44 ms with 3-12 us; PB_GEN_HEADSTART_US tunes it.)
    {
        int us = 8;
        if (const char* e = getenv("PB_GEN_HEADSTART_US")) us = atoi(e);
        if (us > 0 && !early && !immediately && !many_waves) {
            PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_gen_started, 0));
            k3_headstart_kernel<<<1, 1, 0, ctx->stream>>>((unsigned)us * 1000u);
            ctx->tm.n_launches += 1;
        }
    }
    ctx->pregen_valid = true;
    ctx->pregen_tile = pl.tile;
    return PB_OK;
}

// Every allocation pb_run_assoc needs (sizes follow from S_inf, n_max, P, m: all known after pb_run_events), so that the
// pass itself issues no cudaMalloc / cudaFree.  In a sharded run a shard's last kernel waits for its peers inside the
// exchange; an allocation on a device that is in that state may wait for the device — pb_group therefore prepares ALL its
// shards before any of them starts the pass.
static int k3_alloc(pb_ctx* ctx, K3Plan* pl) {
    const int64_t m = ctx->cfg.n_replicates;
    int rc;
    if (!ctx->xchg.active) {
        int64_t mcap_bytes = ctx->maxT_cap * 8;
        void* pm = ctx->d_maxT_own;
        if ((rc = ensure(ctx, &pm, &mcap_bytes, m * 8))) { ctx->d_maxT_own = nullptr; return rc; }
        ctx->d_maxT_own = (unsigned long long*)pm; ctx->maxT_cap = mcap_bytes / 8;
    }
    if ((rc = k3_plan(ctx, pl))) return rc;
    if (!ctx->d_colpop || ctx->colpop_cap < ctx->P) {
        if (ctx->d_colpop) cudaFree(ctx->d_colpop);
        ctx->d_colpop = nullptr;
        PB_CUDA(cudaMalloc(&ctx->d_colpop, (size_t)ctx->P * 2 * 4));
        ctx->colpop_cap = ctx->P;
    }
    void* pf = ctx->d_fb_reps;
    if ((rc = ensure(ctx, &pf, &ctx->fb_cap, (pl->tile + 256) * 4))) { ctx->d_fb_reps = nullptr; return rc; }
    ctx->d_fb_reps = (int32_t*)pf;
    if (ctx->cfg.mode == PB_MODE_REPLAY) {       // one tile of caller permutations at a time
        const int64_t nr = m < pl->tile ? m : pl->tile;
        int64_t cap = ctx->perms_cap;
        void* pp = ctx->d_perms;
        if ((rc = ensure(ctx, &pp, &cap, nr * (int64_t)ctx->P * 4))) { ctx->d_perms = nullptr; ctx->perms_cap = 0; return rc; }
        ctx->d_perms = (uint32_t*)pp; ctx->perms_cap = cap;
    }
    return PB_OK;
}

int pbk_labels_alloc(pb_ctx* ctx) {
    if (ctx->cfg.mode != PB_MODE_PHILOX || ctx->P == 0 || ctx->n_case + ctx->n_ctrl == 0) return PB_OK;
    K3Plan pl;
    return k3_plan(ctx, &pl);
}

int pbk_assoc_prepare(pb_ctx* ctx) {
    int rc;
    K3Plan pl;
    if ((rc = pbk_ptable(ctx, true))) return rc;
    if ((rc = pbk_thresholds(ctx, true))) return rc;
    if ((rc = k3_alloc(ctx, &pl))) return rc;
    return pbk_pvalues(ctx, true);
}

int pbk_assoc(pb_ctx* ctx, const uint32_t* replay_perms) {
    const int64_t m = ctx->cfg.n_replicates;
    const int64_t n_slots = 2 * ctx->S_inf;
    const int T = 256;
    cudaStream_t st = ctx->stream;
    int rc;

    // ---- K2 observed tallies
    PB_CUDA(cudaEventRecord(ctx->ev[2], st));
    pb_trace_mark(ctx, "start");
    k2_tally_kernel<<<(unsigned)std::max<int64_t>(1, (n_slots + T - 1) / T), T, 0, st>>>(ctx->d_alleles, n_slots, ctx->d_csr, ctx->d_label);
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaEventRecord(ctx->ev[3], st));
    pb_trace_mark(ctx, "tally");

    // ---- p-table, observed p-values, thresholds
    if ((rc = pbk_ptable(ctx))) return rc;
    if ((rc = pbk_thresholds(ctx))) return rc;
    PB_CUDA(cudaEventRecord(ctx->ev[4], st));
    pb_trace_mark(ctx, "ptable+thresholds");

    // ---- null distribution, tiled over replicates
    K3Plan pl;
    if ((rc = k3_alloc(ctx, &pl))) return rc;
    if (ctx->xchg.active) {
        if ((rc = pbx_begin_epoch(ctx))) return rc;      // the epoch's buffer of the exchange block (peers read it)
    } else {
        ctx->d_maxT = ctx->d_maxT_own;
    }
    fill_u64_kernel<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(ctx->d_maxT, m, 0x7FF0000000000000ull);
    PB_CUDA(cudaMemsetAsync(ctx->d_r, 0, (size_t)n_slots * 4, st));
    PB_CUDA(cudaMemsetAsync(ctx->d_class_counts + 12, 0, 8, st));   // replay-permutation check flag (k3_gen_kernel<1, .>)
    ctx->tm.n_launches += 1;

    const bool has_miss = pl.has_miss;
    const int64_t tile = pl.tile, Rw32 = pl.Rw32, Rw = pl.Rw;
    // No host round trip inside the loop: list lengths, the fallback list and its length stay on the device; buffer
    // reuse between the two streams is ordered with events.  (Replay mode uploads each tile's permutations.)
    ctx->tm.n_maxt_fallback_reps = 0;
    const unsigned int* lc = ctx->d_lcount;
    int64_t bps = 32;
    if (const char* e = getenv("PB_K3_BLOCKS_PER_SM")) bps = atoll(e);
    const int64_t warps_per_block = K3_THREADS / 32;
    const unsigned tight_blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * bps, (n_slots + warps_per_block - 1) / warps_per_block));
    const unsigned small_blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 8, (n_slots + warps_per_block - 1) / warps_per_block));
    const unsigned large_blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 4, n_slots));
    PB_CUDA(cudaEventRecord(ctx->ev[5], st));
    bool first_gen_timed = false;
    for (int64_t base = 0; base < m; base += tile) {
        const int64_t nr = (m - base < tile) ? (m - base) : tile;
        const unsigned gblocks = (unsigned)(((nr + 255) / 256) * (256 / GEN_THREADS));
        const int buf = (pl.nbuf == 2) ? (int)((base / tile) & 1) : 0;
        if (ctx->cfg.mode == PB_MODE_REPLAY) {
            const int64_t bytes = nr * (int64_t)ctx->P * 4;       // (buffer sized for a whole tile by k3_alloc)
            PB_CUDA(cudaMemcpyAsync(ctx->d_perms, replay_perms + base * (int64_t)ctx->P, (size_t)bytes, cudaMemcpyHostToDevice, st));
            if (has_miss)
                k3_gen_kernel<1, true><<<gblocks, GEN_THREADS, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label,
                                                                       ctx->d_perms, (uint32_t*)ctx->d_Xcase, (uint32_t*)ctx->d_Xmiss, Rw32, 0,
                                                                       ctx->d_class_counts + 12);
            else
                k3_gen_kernel<1, false><<<gblocks, GEN_THREADS, 0, st>>>(ctx->cfg.seed, base, nr, ctx->P, ctx->n_case, ctx->n_ctrl, ctx->d_label,
                                                                        ctx->d_perms, (uint32_t*)ctx->d_Xcase, nullptr, Rw32, 0,
                                                                        ctx->d_class_counts + 12);
            ctx->tm.n_launches += 1;
            // replay_perms is caller-owned pageable memory: the copy above is staged synchronously by the runtime
        } else if (base == 0 && ctx->pregen_valid && ctx->pregen_tile == tile) {
            PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_gen_done, 0));   // tile 0 was generated alongside the K1 tail
        } else if (base > 0 && pl.nbuf == 2) {
            PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_gen_b[buf], 0)); // generated on the second stream during the previous tile's count
        } else {
            if ((rc = k3_launch_philox(ctx, pl, base, nr, st, buf))) return rc;
        }
        if (pl.nbuf == 2 && base + tile < m) {
            // next tile into the other buffer, on the second stream, concurrently with this tile's count; the buffer's
            // last reader was the count of the previous tile
            const int64_t nb = base + tile, nnr = (m - nb < tile) ? (m - nb) : tile;
            if (base > 0) PB_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev_cnt_b[buf ^ 1], 0));
            if ((rc = k3_launch_philox(ctx, pl, nb, nnr, ctx->stream2, buf ^ 1))) return rc;
            PB_CUDA(cudaEventRecord(ctx->ev_gen_b[buf ^ 1], ctx->stream2));
        }
        ctx->pregen_valid = false;
        if (!first_gen_timed) { PB_CUDA(cudaEventRecord(ctx->ev[6], st)); first_gen_timed = true; }
        pb_trace_mark(ctx, "fill+wait_gen");

        // count in L2-sized column tiles of the generated matrix
        for (int64_t cb = 0; cb < nr; cb += pl.ctile) {
            const int64_t cnr = (nr - cb < pl.ctile) ? (nr - cb) : pl.ctile;
            K3Params p;
            p.alleles = ctx->d_alleles; p.n_slots = n_slots; p.csr = ctx->d_csr; p.pobs = ctx->d_pobs; p.ptable = ctx->d_ptable;
            p.ktau = ctx->d_ktau; p.tau = ctx->d_tau;
            p.Xc = ctx->d_Xcase + buf * pl.buf_words + cb / 64;
            p.Xm = has_miss ? ctx->d_Xmiss + buf * pl.buf_words + cb / 64 : nullptr;
            p.Rw = Rw;
            p.rep_base = base + cb; p.n_reps = cnr; p.r = ctx->d_r; p.maxT = ctx->d_maxT;
            const int64_t words = (cnr + 63) / 64;
            // the few heavy sweeps (candidate short lists, long lists) go to a side stream: they are instruction-bound and
            // fill the issue slots the load-latency-bound tight sweep leaves idle; the lists are disjoint, maxT is atomicMin
            const bool side = !getenv("PB_K3_NO_SIDE_STREAM");
            cudaStream_t ss = side ? ctx->stream3 : st;
            if (side) {
                PB_CUDA(cudaEventRecord(ctx->ev_side_fork, st));
                PB_CUDA(cudaStreamWaitEvent(ss, ctx->ev_side_fork, 0));
            }
            k3_count_cand_kernel<<<small_blocks, K3_THREADS, 0, ss>>>(p, ctx->d_work + 3 * n_slots, lc + 3);
            k3_count_small_kernel<<<small_blocks, K3_THREADS, 0, ss>>>(p, ctx->d_work + 3 * n_slots, lc + 3);
            k3_count_large_kernel<<<large_blocks, K3_THREADS, 0, ss>>>(p, ctx->d_work + 2 * n_slots, lc + 2);
            k3_colpop_kernel<<<(unsigned)ctx->P, 128, 0, ss>>>(p.Xc, p.Xm, Rw, words, ctx->P, ctx->d_colpop);
            k3_trivial_kernel<<<(unsigned)std::max<int64_t>(1, (n_slots + T - 1) / T), T, 0, ss>>>(p, ctx->d_work, lc + 0, ctx->d_colpop, ctx->P);
            if (side) PB_CUDA(cudaEventRecord(ctx->ev_side_join, ss));
            if (has_miss) k3_count_tight_kernel<true><<<tight_blocks, K3_THREADS, 0, st>>>(p, ctx->d_work + n_slots, lc + 1);
            else k3_count_tight_kernel<false><<<tight_blocks, K3_THREADS, 0, st>>>(p, ctx->d_work + n_slots, lc + 1);
            pb_trace_mark(ctx, "tight");
            if (side) PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_side_join, 0));
            // replicates whose minimum was not certified by a candidate are recomputed exactly
            PB_CUDA(cudaMemsetAsync(ctx->d_fb_count, 0, 4, st));
            k3_collect_fallback_kernel<<<(unsigned)((cnr + T - 1) / T), T, 0, st>>>(ctx->d_maxT, p.rep_base, cnr, ctx->d_tau, ctx->d_fb_reps,
                                                                                 ctx->d_fb_count);
            k3_fallback_kernel<<<(unsigned)std::min<int64_t>(cnr, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(p, ctx->d_fb_reps, ctx->d_fb_count,
                                                                                                           ctx->d_fb_total);
            pb_trace_mark(ctx, "join+fallback");
            ctx->tm.n_launches += 8;
        }
        if (pl.nbuf == 2) PB_CUDA(cudaEventRecord(ctx->ev_cnt_b[buf], st));
        PB_CUDA(cudaGetLastError());
    }
    PB_CUDA(cudaEventRecord(ctx->ev[7], st));

    // ---- the one exchange: min of maxT over the site shards.  Peer-memory mode (pb_comm_connect / pb_group): only the
    // ready signal is issued here, the reduction itself happens inside K4's sort kernel (ms_allreduce ~ 0, the wait for
    // the slowest rank shows up in ms_pvalues).  NCCL mode (pb_comm_init): an all-reduce on the stream.
    PB_CUDA(cudaEventRecord(ctx->ev[9], st));
    if (ctx->xchg.active) {
        if ((rc = pbx_signal(ctx))) return rc;
    } else if (ctx->world > 1) {
        if ((rc = pbn_allreduce_min_f64(ctx, (double*)ctx->d_maxT, m))) return rc;
    }
    PB_CUDA(cudaEventRecord(ctx->ev[10], st));

    // ---- K4
    if ((rc = pbk_pvalues(ctx))) return rc;
    PB_CUDA(cudaEventRecord(ctx->ev[11], st));
    pb_trace_mark(ctx, "allreduce+pvalues");
    struct { unsigned long long nf; unsigned int lc[4]; unsigned long long fb, swept; } hb;   // layout of the scratch block in pbk_thresholds
    PB_CUDA(cudaMemcpyAsync(&hb, ctx->d_nfast, sizeof(hb), cudaMemcpyDeviceToHost, st));
    unsigned long long replay_err = 0;
    PB_CUDA(cudaMemcpyAsync(&replay_err, ctx->d_class_counts + 12, 8, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    PB_CUDA(cudaGetLastError());
    if ((rc = pbx_check(ctx))) return rc;
    if (replay_err) {
        ctx->err = (replay_err & 1ull) ? "pb_run_assoc: replay_perms holds an index >= n_samples"
                                       : "pb_run_assoc: a row of replay_perms is not a permutation of 0 .. n_samples-1";
        return PB_ERR_INVALID;
    }
    float t;
    cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3]); ctx->tm.ms_tally = t;
    cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[4]); ctx->tm.ms_ptable = t;
    cudaEventElapsedTime(&t, ctx->ev[5], ctx->ev[6]); ctx->tm.ms_null_gen = t;      // first tile, as seen by the main stream
    cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]); ctx->tm.ms_null_count = t;    // everything after it (counts, fallback, later tiles)
    ctx->tm.ms_null_fallback = 0;
    cudaEventElapsedTime(&t, ctx->ev[9], ctx->ev[10]); ctx->tm.ms_allreduce = t;
    cudaEventElapsedTime(&t, ctx->ev[10], ctx->ev[11]); ctx->tm.ms_pvalues = t;
    cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[11]); ctx->tm.ms_assoc_total = t;
    pb_trace_dump(ctx, "assoc");
    ctx->tm.n_fast_alleles = (int64_t)hb.nf;
    for (int c = 0; c < 4; c++) ctx->n_work[c] = (int64_t)hb.lc[c];
    ctx->tm.n_maxt_fallback_reps = (int64_t)hb.fb;
    ctx->tm.n_swept_rows = (int64_t)hb.swept;
    if (ctx->gen_timed) { cudaEventElapsedTime(&t, ctx->ev_gen_t0, ctx->ev_gen_t1); ctx->tm.ms_gen_kernel = t; }
    if (getenv("PB_VERBOSE"))
        fprintf(stderr, "[poutine_b200] work lists: trivial %u, tight %u, long %u, general %u; fallback replicates %llu\n", hb.lc[0], hb.lc[1], hb.lc[2],
                hb.lc[3], hb.fb);
    return PB_OK;
}
