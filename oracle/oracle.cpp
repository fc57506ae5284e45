// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of POUTINE's live homoplasy-counting + association path, following
// /root/reference/src/Homoplasy_Counter.java ("HC.java") line by line.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
// library; the product (poutine_b200/) never does.
//
// PARITY PIN STATUS
//   * binomial p-values: PINNED — checked bit-for-bit against the commons-math3 3.6.1 bytecode
//     executed by oracle/tools/minijvm.py (tests/golden/binom_golden.json).
//   * event detection (rows a1-a5 of SURVEY.md section 8): PINNED — the reference's own compiled
//     count_all_homoplasy_events (compiled/Homoplasy_Counter.class + coevolution.jar) is executed by
//     oracle/tools/refjvm.py (bytecode interpreter + JDK-collection emulation, HashMap iteration order
//     included); tests/golden/events_refbytecode.json holds its outputs for 5 alignments on nasty
//     topologies (every Homoplasy_Events field, both MRCA sets, the logged class counters) and this
//     file reproduces all of them (tests/test_oracle_events.py).
//   * observed tallies, binomial statistics, replicate counting, maxT, point-wise and family-wise
//     estimates (rows a6-a11): PINNED the same way — subset_by_min_hcount and the live resampling
//     driver run from the reference's bytecode, BinomialTest from commons-math's jar, with
//     Collections.shuffle seeded and the replicates serial; tests/golden/assoc_refbytecode.json records
//     the shuffles and the results, and this file reproduces every integer and every double bit for
//     bit in replay mode.
//
// Deviations from a literal transcription (none change results):
//   * nodes are integers, not name strings (HashMap<String,..> keyed by unique node names
//     == arrays / hash sets keyed by node index);
//   * `map_seg_site_to_tree` (HC.java:1383-1394) builds a per-site name->char HashMap; here the
//     column is read in place from the node-major char matrix;
//   * the racy `r_a1++` (HC.java:4240, 4274) is summed serially — the reference's intended
//     semantics.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "commons_math_port.h"

extern "C" {

// ---------------------------------------------------------------------------------------------
// 1. commons-math chain (exported for the pinning tests)
// ---------------------------------------------------------------------------------------------
double or_fm_exp(double x) { return cmport::fm_exp(x); }
double or_fm_log(double x) { return cmport::fm_log(x); }
double or_fm_log1p(double x) { return cmport::fm_log1p(x); }
double or_log_gamma(double x) { return cmport::gamma_logGamma(x); }
double or_log_beta(double a, double b) { return cmport::beta_logBeta(a, b); }
double or_regularized_beta(double x, double a, double b) { return cmport::beta_regularizedBeta(x, a, b); }
// HC.java:3050 — binom_test.binomialTest(n, k, p, AlternativeHypothesis.GREATER_THAN)
double or_binomial_test(int n, int k, double p) { return cmport::binomial_test_greater(n, k, p); }

// Triangular table f[n][k], 0 <= k <= n <= n_max, laid out at index n*(n+1)/2 + k.
void or_binomial_table(int n_max, double p, double* out) {
    for (int n = 0; n <= n_max; n++)
        for (int k = 0; k <= n; k++) out[(int64_t)n * (n + 1) / 2 + k] = cmport::binomial_test_greater(n, k, p);
}

// ---------------------------------------------------------------------------------------------
// 2. java.util.Random + Collections.shuffle (JDK-standard) — defines REPLAY perms
// ---------------------------------------------------------------------------------------------
struct JavaRandom {
    uint64_t seed;
    explicit JavaRandom(int64_t s) : seed(((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
    int32_t next(int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)((int64_t)seed >> (48 - bits));
    }
    int32_t nextInt(int32_t bound) {
        int32_t r = next(31);
        int32_t m = bound - 1;
        if ((bound & m) == 0) {
            r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        } else {
            for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = next(31)) {}
        }
        return r;
    }
};

// out[i] = nextInt(bound) (bound > 0) or nextInt() (bound == 0) of `new Random(seed)` — known-answer tests
void or_java_random_ints(int64_t seed, int32_t n, int32_t bound, int32_t* out) {
    JavaRandom rnd(seed);
    for (int32_t i = 0; i < n; i++) out[i] = bound > 0 ? rnd.nextInt(bound) : rnd.next(32);
}

// perms[rep*P + i] = index (into the pheno-file order) of the label that sample i receives in
// replicate rep.  Replicate rep uses `new Random(base_seed + rep)` and Collections.shuffle:
//   for (i = size; i > 1; i--) swap(list, i-1, rnd.nextInt(i));
void or_java_shuffle_perms(int64_t base_seed, int64_t m, int32_t P, uint32_t* perms) {
    for (int64_t rep = 0; rep < m; rep++) {
        uint32_t* p = perms + rep * P;
        for (int32_t i = 0; i < P; i++) p[i] = (uint32_t)i;
        JavaRandom rnd(base_seed + rep);
        for (int32_t i = P; i > 1; i--) std::swap(p[i - 1], p[rnd.nextInt(i)]);
    }
}

// ---------------------------------------------------------------------------------------------
// 3. Philox4x32-10 label sampler — restates the PRODUCT's own sampling spec so
//    the GPU's Philox mode can be checked bit-exactly too.  Not reference behaviour: the
//    reference draws from an unseeded shared java.util.Random (HC.java:4911).
// ---------------------------------------------------------------------------------------------
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Labels of the P pheno-file samples in replicate `rep`: a uniformly random arrangement of
// n_case 1s, n_ctrl 0s and n_miss 2s, drawn sequentially (sample 0 first) without replacement.
//
// Spec v1 (P >= 65536).  Draw for sample i: u uniform in [0, P-i) by Lemire's multiply-shift with exact rejection;
// main stream word i = philox(ctr=(i>>2, rep_lo, rep_hi, 0), key=seed)[i&3]; rejected draws are
// replaced from the retry stream word q = philox(ctr=(q>>2, rep_lo, rep_hi, 1))[q&3], q = 0,1,..
static void philox_labels_v1(uint64_t seed, uint64_t rep, int32_t P, int32_t n_case, int32_t n_ctrl, uint8_t* out) {
    uint32_t rem_case = (uint32_t)n_case, rem_ctrl = (uint32_t)n_ctrl;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t r0 = (uint32_t)rep, r1 = (uint32_t)(rep >> 32);
    uint32_t w[4] = {0, 0, 0, 0}, rw[4];
    uint32_t q = 0;
    for (int32_t i = 0; i < P; i++) {
        if ((i & 3) == 0) philox4x32_10((uint32_t)(i >> 2), r0, r1, 0u, k0, k1, w);
        const uint32_t range = (uint32_t)(P - i);
        uint32_t x = w[i & 3];
        uint64_t prod = (uint64_t)x * range;
        uint32_t lo = (uint32_t)prod;
        if (lo < range) {
            const uint32_t t = (0u - range) % range;
            while (lo < t) {
                philox4x32_10(q >> 2, r0, r1, 1u, k0, k1, rw);
                x = rw[q & 3];
                q++;
                prod = (uint64_t)x * range;
                lo = (uint32_t)prod;
            }
        }
        const uint32_t u = (uint32_t)(prod >> 32);
        uint8_t lab;
        if (u < rem_case) { lab = 1; rem_case--; }
        else if (u < rem_case + rem_ctrl) { lab = 0; rem_ctrl--; }
        else lab = 2;
        out[i] = lab;
    }
}

// Spec v3 (P < 65536; DESIGN.md "Philox sampling spec").  Four samples per 64-bit word: group g = samples 4g..4g+c-1,
// c = min(4, P-4g); x = (o[2g+1] : o[2g]), o[i] = the i-th output of the replicate's xoshiro128++ stream (Blackman & Vigna:
// result = rotl(s0 + s3, 7) + s0; t = s1 << 9; s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3, 11)) whose
// 128-bit state starts as philox(ctr=(0, rep_lo, rep_hi, 2), key=seed) (an all-zero block becomes s0 = 1).  For j < c with
// R_j = P-4g-j: x * R_j = u_j * 2^64 + x' (u_j = the draw in [0, R_j); x' goes on as x) — Lemire's multiply-shift for the
// product range at once, in mixed radix.  Exact iff the final x' >= 2^64 mod prod(R_j); otherwise the group is redrawn from
// retry word q = (rw[2(q&1)+1] : rw[2(q&1)]) of philox(ctr=(q>>1, rep_lo, rep_hi, 1)), q = 0,1,.. (the main stream's position
// does not depend on the retries).  Spec v2 (round 2, superseded) took x from philox(ctr=(g>>1, rep_lo, rep_hi, 0)) instead.
static inline uint32_t rotl32(uint32_t v, int k) { return (v << k) | (v >> (32 - k)); }
static inline uint32_t xoshiro128pp_next(uint32_t s[4]) {
    const uint32_t result = rotl32(s[0] + s[3], 7) + s[0];
    const uint32_t t = s[1] << 9;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl32(s[3], 11);
    return result;
}
static void philox_labels_v3(uint64_t seed, uint64_t rep, int32_t P, int32_t n_case, int32_t n_ctrl, uint8_t* out) {
    uint32_t rem_case = (uint32_t)n_case, rem_ctrl = (uint32_t)n_ctrl;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t r0 = (uint32_t)rep, r1 = (uint32_t)(rep >> 32);
    uint32_t s[4], rw[4];
    philox4x32_10(0u, r0, r1, 2u, k0, k1, s);
    if ((s[0] | s[1] | s[2] | s[3]) == 0) s[0] = 1;
    uint32_t q = 0;
    for (int32_t g = 0; 4 * g < P; g++) {
        const uint32_t xl = xoshiro128pp_next(s), xh = xoshiro128pp_next(s);
        uint64_t x = ((uint64_t)xh << 32) | xl;
        const int c = std::min<int32_t>(4, P - 4 * g);
        const uint32_t range = (uint32_t)(P - 4 * g);
        uint32_t us[4] = {0, 0, 0, 0};
        for (;;) {
            uint64_t l = x, prod = 1;
            uint32_t R = range;
            for (int j = 0; j < c; j++) {
                const unsigned __int128 p = (unsigned __int128)l * R;
                us[j] = (uint32_t)(uint64_t)(p >> 64);
                l = (uint64_t)p;
                prod *= R; R--;
            }
            if (l < prod) {
                const uint64_t t = (0ull - prod) % prod;
                if (l < t) {
                    philox4x32_10(q >> 1, r0, r1, 1u, k0, k1, rw);
                    x = ((uint64_t)rw[2 * (q & 1) + 1] << 32) | rw[2 * (q & 1)];
                    q++;
                    continue;
                }
            }
            break;
        }
        for (int j = 0; j < c; j++) {
            const uint32_t u = us[j];
            uint8_t lab;
            if (u < rem_case) { lab = 1; rem_case--; }
            else if (u < rem_case + rem_ctrl) { lab = 0; rem_ctrl--; }
            else lab = 2;
            out[4 * g + j] = lab;
        }
    }
}

void or_philox_labels(uint64_t seed, uint64_t rep, int32_t P, int32_t n_case, int32_t n_ctrl, uint8_t* out) {
    if (P < 65536) philox_labels_v3(seed, rep, P, n_case, n_ctrl, out);
    else philox_labels_v1(seed, rep, P, n_case, n_ctrl, out);
}

// ---------------------------------------------------------------------------------------------
// 4. L4 — count_all_homoplasy_events (HC.java:1402-1511)
// ---------------------------------------------------------------------------------------------
struct or_site_t {  // mirrors Homoplasy_Events (HC.java:5531-5550), 32 bytes
    int32_t segsite_id;  // 1-based (HC.java:1419)
    int32_t physical_pos;
    int32_t allele1, allele2;  // chars
    int32_t a1_count, a2_count, a1_extant, a2_extant;
};

struct or_stat_t {  // mirrors Binomial_Test_Stat (HC.java:5876-5887) + obs_counts, 96 bytes
    int32_t event_index;  // index into the biallelic-site list (all_events)
    int32_t segsite_id;
    int32_t a1_in_use, a2_in_use;
    int32_t obs_counts[4];  // a1case, a1ctrl, a2case, a2ctrl (HC.java:3714-3717)
    int32_t r_a1, r_a2, r_maxT_a1, r_maxT_a2;
    double obs_p_a1, obs_p_a2, pointwise_a1, pointwise_a2, familywise_a1, familywise_a2;
};

struct Oracle {
    int32_t N = 0, L = 0, root = -1;
    std::vector<int32_t> parent, leaves;
    std::vector<uint8_t> is_leaf;
    // results of L4
    std::vector<or_site_t> events;                // all_events (biallelic sites only, site order)
    std::vector<std::vector<int32_t>> mrca[2];    // per event: allele1_MRCAs / allele2_MRCAs keys
    int64_t n_mono = 0, n_bi = 0, n_tri = 0, n_quad = 0, n_other = 0;
    // results of L5
    std::vector<or_stat_t> stats;
    std::vector<double> maxT_sorted;
    std::vector<double> maxT_unsorted;
    double p_success = 0;
};

static inline bool is_acgt(char g) { return g == 'A' || g == 'C' || g == 'G' || g == 'T'; }

void* or_create(int32_t N, const int32_t* parent, int32_t L, const int32_t* leaf_nodes) {
    Oracle* o = new Oracle();
    o->N = N; o->L = L;
    o->parent.assign(parent, parent + N);
    o->leaves.assign(leaf_nodes, leaf_nodes + L);
    o->is_leaf.assign(N, 0);
    for (int32_t i = 0; i < L; i++) o->is_leaf[leaf_nodes[i]] = 1;
    for (int32_t v = 0; v < N; v++) if (parent[v] < 0) o->root = v;
    return o;
}
void or_destroy(void* h) { delete (Oracle*)h; }

// seq is node-major: genotype of node v at site i = seq[v*S + i]  (HashMap<String,char[]> seg_sites)
struct L4Part {
    std::vector<or_site_t> events;
    std::vector<std::vector<int32_t>> mrca[2];
    int64_t n_mono = 0, n_bi = 0, n_tri = 0, n_quad = 0, n_other = 0;
};

static void l4_range(const Oracle& o, int64_t S, const char* seq, const int32_t* physical_poss, int64_t i0, int64_t i1, L4Part& out) {
    const int32_t root = o.root;
    for (int64_t i = i0; i < i1; i++) {                     // HC.java:1418
        auto geno = [&](int32_t v) -> char { return seq[(int64_t)v * S + i]; };  // get_node_genotype, HC.java:1823-1830
        // get_alleles (HC.java:1593-1603): leaf genotypes in [ACGT] grouped by allele.
        // HashMap<Character,...> iteration order = bucket order (c & 15): A(1) C(3) T(4) G(7).
        int64_t cnt[4] = {0, 0, 0, 0};                      // order A, C, T, G
        for (int32_t leaf : o.leaves) {
            const char g = geno(leaf);
            if (g == 'A') cnt[0]++; else if (g == 'C') cnt[1]++; else if (g == 'T') cnt[2]++; else if (g == 'G') cnt[3]++;
        }
        static const char ORDER[4] = {'A', 'C', 'T', 'G'};
        int n_alleles = 0;
        for (int c = 0; c < 4; c++) n_alleles += cnt[c] > 0;
        switch (n_alleles) {                                 // HC.java:1443-1473
            case 2: out.n_bi++; break;
            case 1: out.n_mono++; continue;
            case 3: out.n_tri++; continue;
            case 4: out.n_quad++; continue;
            default: out.n_other++; continue;
        }
        // identify_major_minor_alleles (HC.java:1712-1743)
        int first = -1, second = -1;
        for (int c = 0; c < 4; c++) if (cnt[c] > 0) { if (first < 0) first = c; else second = c; }
        char allele1, allele2;
        if (cnt[first] >= cnt[second]) { allele1 = ORDER[first]; allele2 = ORDER[second]; }
        else { allele1 = ORDER[second]; allele2 = ORDER[first]; }
        // identify_homoplasy_events_for_seg_site (HC.java:1647-1698)
        std::unordered_set<int32_t> a1_MRCAs, a2_MRCAs;
        std::vector<int32_t> a1_order, a2_order;
        auto add_MRCA = [&](int32_t node) {                  // HC.java:1856-1888
            if (geno(node) == allele1) { if (a1_MRCAs.insert(node).second) a1_order.push_back(node); }
            else { if (a2_MRCAs.insert(node).second) a2_order.push_back(node); }
        };
        for (int32_t leaf : o.leaves) {                      // HC.java:1666
            int32_t prev_node = leaf;
            for (int32_t curr = o.parent[leaf]; curr != root; curr = o.parent[curr]) {  // HC.java:1674
                // is_MRCA_found (HC.java:1837-1852)
                const char curr_geno = geno(curr), prev_geno = geno(prev_node);
                bool found = false;
                if (!is_acgt(prev_geno)) found = false;
                else if (curr_geno != prev_geno) found = true;
                if (found) add_MRCA(prev_node);
                prev_node = curr;
            }
            add_MRCA(root);                                  // HC.java:1689
        }
        if (a1_MRCAs.size() < 2) { a1_MRCAs.clear(); a1_order.clear(); }   // remove_non_homoplasic_mutations, HC.java:1954-1962
        if (a2_MRCAs.size() < 2) { a2_MRCAs.clear(); a2_order.clear(); }
        or_site_t e;
        e.segsite_id = (int32_t)(i + 1);
        e.physical_pos = physical_poss ? physical_poss[i] : 0;
        e.allele1 = allele1; e.allele2 = allele2;
        e.a1_count = (int32_t)a1_MRCAs.size(); e.a2_count = (int32_t)a2_MRCAs.size();
        e.a1_extant = e.a2_extant = 0;                       // calc_num_extant_homoplasic_mutations, HC.java:5586-5600
        for (int32_t v : a1_order) e.a1_extant += o.is_leaf[v];
        for (int32_t v : a2_order) e.a2_extant += o.is_leaf[v];
        std::sort(a1_order.begin(), a1_order.end());
        std::sort(a2_order.begin(), a2_order.end());
        out.events.push_back(e);
        out.mrca[0].push_back(std::move(a1_order));
        out.mrca[1].push_back(std::move(a2_order));
    }
}

// n_threads > 1 splits the site loop into contiguous ranges (sites are independent, HC.java:1418 carries
// only the class counters) and concatenates in site order.  The reference itself runs this loop on ONE
// thread (README.md:127); the threaded form exists so the CPU baseline can use every host core.
int64_t or_count_all_homoplasy_events_mt(void* h, int64_t S, const char* seq, const int32_t* physical_poss, int n_threads) {
    Oracle& o = *(Oracle*)h;
    o.events.clear(); o.mrca[0].clear(); o.mrca[1].clear();
    o.n_mono = o.n_bi = o.n_tri = o.n_quad = o.n_other = 0;
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > S) n_threads = (int)(S > 0 ? S : 1);
    std::vector<L4Part> parts(n_threads);
    if (n_threads == 1) l4_range(o, S, seq, physical_poss, 0, S, parts[0]);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t] { l4_range(o, S, seq, physical_poss, S * t / n_threads, S * (t + 1) / n_threads, parts[t]); });
        for (auto& t : th) t.join();
    }
    for (L4Part& p : parts) {
        o.n_mono += p.n_mono; o.n_bi += p.n_bi; o.n_tri += p.n_tri; o.n_quad += p.n_quad; o.n_other += p.n_other;
        o.events.insert(o.events.end(), p.events.begin(), p.events.end());
        for (int a = 0; a < 2; a++)
            for (auto& v : p.mrca[a]) o.mrca[a].push_back(std::move(v));
    }
    return (int64_t)o.events.size();
}
int64_t or_count_all_homoplasy_events(void* h, int64_t S, const char* seq, const int32_t* physical_poss) {
    return or_count_all_homoplasy_events_mt(h, S, seq, physical_poss, 1);
}

void or_get_events(void* h, or_site_t* out) {
    Oracle& o = *(Oracle*)h;
    if (!o.events.empty()) std::memcpy(out, o.events.data(), o.events.size() * sizeof(or_site_t));
}
void or_get_class_counts(void* h, int64_t* out5) {  // mono, bi, tri, quad, other  (HC.java:1477-1480)
    Oracle& o = *(Oracle*)h;
    out5[0] = o.n_mono; out5[1] = o.n_bi; out5[2] = o.n_tri; out5[3] = o.n_quad; out5[4] = o.n_other;
}
int64_t or_get_mrca(void* h, int64_t event_index, int allele /*0=a1,1=a2*/, int32_t* out, int64_t cap) {
    Oracle& o = *(Oracle*)h;
    const auto& v = o.mrca[allele][event_index];
    if (out) for (size_t j = 0; j < v.size() && (int64_t)j < cap; j++) out[j] = v[j];
    return (int64_t)v.size();
}

// ---------------------------------------------------------------------------------------------
// 5. L5 — assoc -> resample_all_mutations_binom_test_combined_nulldists_memoization_concurrent
//    (HC.java:2029-2036, 2096-2134, 3013-3186)
// ---------------------------------------------------------------------------------------------
// pheno file = P rows: sample_node[j] (a leaf), label[j] in {0 ctrl "0", 1 case "1", 2 anything else}.
// mode 0 (REPLAY): perms = m*P indices, sample j gets label[perms[rep*P+j]] in replicate rep.
// mode 1 (PHILOX): labels of replicate rep from or_philox_labels(seed, rep, ...).
// Returns #informative sites, or -1 when min_hcount filtered out everything (HC.java:3866-3869).
int64_t or_assoc(void* h, int32_t P, const int32_t* sample_node, const uint8_t* label, int64_t m, int32_t min_hcount,
                 int mode, const uint32_t* perms, uint64_t seed, int n_threads) {
    Oracle& o = *(Oracle*)h;
    // read_phenos totals (HC.java:1999-2007)
    int32_t tot_cases = 0, tot_controls = 0;
    for (int32_t j = 0; j < P; j++) { tot_cases += label[j] == 1; tot_controls += label[j] == 0; }
    std::vector<int32_t> sample_of_node(o.N, -1);  // phenos.containsKey(node_name)
    for (int32_t j = 0; j < P; j++) sample_of_node[sample_node[j]] = j;

    // subset_by_min_hcount (HC.java:3783-3905), EXTANT_NODES_ONLY = true (HC.java:173)
    struct Inf { int32_t ev; bool a1_in_use, a2_in_use; };
    std::vector<Inf> inf;
    for (size_t e = 0; e < o.events.size(); e++) {
        const or_site_t& s = o.events[e];
        const bool u1 = s.a1_extant >= min_hcount, u2 = s.a2_extant >= min_hcount;
        if (u1 || u2) inf.push_back({(int32_t)e, u1, u2});
    }
    o.stats.clear(); o.maxT_sorted.clear(); o.maxT_unsorted.clear();
    if (inf.empty()) return -1;
    const size_t S_inf = inf.size();

    // calc_background_p_success (HC.java:3334-3353)
    const double p_success = (double)tot_cases / (double)(tot_cases + tot_controls);
    o.p_success = p_success;

    // MRCA nodes of in-use alleles that are pheno-file samples, as sample indices
    // (counts_by_allele only looks at MRCA names present in phenos, HC.java:3741-3743).
    std::vector<std::vector<int32_t>> ev_samples[2];
    ev_samples[0].resize(S_inf); ev_samples[1].resize(S_inf);
    for (size_t s = 0; s < S_inf; s++)
        for (int a = 0; a < 2; a++) {
            if (!(a == 0 ? inf[s].a1_in_use : inf[s].a2_in_use)) continue;
            for (int32_t v : o.mrca[a][inf[s].ev]) if (sample_of_node[v] >= 0) ev_samples[a][s].push_back(sample_of_node[v]);
        }
    // count_homoplasies / counts_by_allele (HC.java:3695-3761)
    auto count_homoplasies = [&](const uint8_t* lab, size_t s, int32_t out4[4]) {
        out4[0] = out4[1] = out4[2] = out4[3] = 0;
        for (int a = 0; a < 2; a++)
            for (int32_t j : ev_samples[a][s]) {
                if (lab[j] == 1) out4[2 * a]++;
                else if (lab[j] == 0) out4[2 * a + 1]++;
            }
    };

    // observed statistics (HC.java:3035-3058)
    o.stats.resize(S_inf);
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    for (size_t s = 0; s < S_inf; s++) {
        or_stat_t& t = o.stats[s];
        std::memset(&t, 0, sizeof(t));
        t.event_index = inf[s].ev;
        t.segsite_id = o.events[inf[s].ev].segsite_id;
        t.a1_in_use = inf[s].a1_in_use; t.a2_in_use = inf[s].a2_in_use;
        count_homoplasies(label, s, t.obs_counts);
        // Binomial_Test_Stat ctor sentinels (HC.java:5891-5906)
        if (!t.a1_in_use) { t.r_a1 = -1; t.r_maxT_a1 = -1; t.pointwise_a1 = t.obs_p_a1 = t.familywise_a1 = NaN; }
        if (!t.a2_in_use) { t.r_a2 = -1; t.r_maxT_a2 = -1; t.pointwise_a2 = t.obs_p_a2 = t.familywise_a2 = NaN; }
        if (t.a1_in_use) t.obs_p_a1 = cmport::binomial_test_greater(t.obs_counts[0] + t.obs_counts[1], t.obs_counts[0], p_success);
        if (t.a2_in_use) t.obs_p_a2 = cmport::binomial_test_greater(t.obs_counts[2] + t.obs_counts[3], t.obs_counts[2], p_success);
    }

    // replicates (HC.java:3098-3130; Replicate.run 3229-3272)
    std::vector<double> maxT((size_t)m);
    if (n_threads < 1) n_threads = 1;
    std::vector<std::vector<int32_t>> r_local(n_threads, std::vector<int32_t>(2 * S_inf, 0));
    auto worker = [&](int tid) {
        std::vector<uint8_t> perm_lab(P);
        std::unordered_map<uint64_t, double> cache;  // resampled_pvalue_cache keyed "n_k" (HC.java:3096, 4226)
        std::vector<int32_t>& r = r_local[tid];
        const int64_t lo = m * tid / n_threads, hi = m * (tid + 1) / n_threads;
        for (int64_t rep = lo; rep < hi; rep++) {
            // permute_phenos (HC.java:4890-4923)
            if (mode == 0) { const uint32_t* pr = perms + rep * P; for (int32_t j = 0; j < P; j++) perm_lab[j] = label[pr[j]]; }
            else or_philox_labels(seed, (uint64_t)rep, P, tot_cases, tot_controls, perm_lab.data());
            double rep_min = std::numeric_limits<double>::infinity();
            for (size_t s = 0; s < S_inf; s++) {
                int32_t c[4];
                count_homoplasies(perm_lab.data(), s, c);
                for (int a = 0; a < 2; a++) {               // calc_binomial_test_stats_memoization_concurrent (HC.java:4196-4280)
                    if (!(a == 0 ? inf[s].a1_in_use : inf[s].a2_in_use)) continue;
                    const int32_t n = c[2 * a] + c[2 * a + 1], k = c[2 * a];
                    const uint64_t key = ((uint64_t)(uint32_t)n << 32) | (uint32_t)k;
                    double pv;
                    auto it = cache.find(key);
                    if (it != cache.end()) pv = it->second;
                    else { pv = cmport::binomial_test_greater(n, k, p_success); cache.emplace(key, pv); }
                    const double obs = a == 0 ? o.stats[s].obs_p_a1 : o.stats[s].obs_p_a2;
                    if (pv <= obs) r[2 * s + a]++;
                    if (pv < rep_min) rep_min = pv;         // HC.java:3239-3243 (min over a1 ∪ a2)
                }
            }
            maxT[(size_t)rep] = rep_min;
        }
    };
    if (n_threads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    for (size_t s = 0; s < S_inf; s++) {
        or_stat_t& t = o.stats[s];
        for (int tid = 0; tid < n_threads; tid++) {
            if (t.a1_in_use) t.r_a1 += r_local[tid][2 * s];
            if (t.a2_in_use) t.r_a2 += r_local[tid][2 * s + 1];
        }
    }
    o.maxT_unsorted = maxT;
    // calc_pointwise_estimate_binom (HC.java:4812-4838)
    for (or_stat_t& t : o.stats) {
        if (t.a1_in_use) t.pointwise_a1 = ((double)(t.r_a1 + 1)) / ((double)(m + 1));
        if (t.a2_in_use) t.pointwise_a2 = ((double)(t.r_a2 + 1)) / ((double)(m + 1));
    }
    // calc_familywise_estimates_binom_combined_nulldists (HC.java:4558-4654)
    std::sort(maxT.begin(), maxT.end());
    for (or_stat_t& t : o.stats) {
        if (t.a1_in_use) {
            const int64_t c = std::upper_bound(maxT.begin(), maxT.end(), t.obs_p_a1) - maxT.begin();  // #{maxT <= obs}
            t.r_maxT_a1 = (int32_t)c;
            t.familywise_a1 = ((double)c + 1) / ((double)m + 1);
        }
        if (t.a2_in_use) {
            const int64_t c = std::upper_bound(maxT.begin(), maxT.end(), t.obs_p_a2) - maxT.begin();
            t.r_maxT_a2 = (int32_t)c;
            t.familywise_a2 = ((double)c + 1) / ((double)m + 1);
        }
    }
    o.maxT_sorted = std::move(maxT);
    return (int64_t)S_inf;
}

void or_get_stats(void* h, or_stat_t* out) {
    Oracle& o = *(Oracle*)h;
    if (!o.stats.empty()) std::memcpy(out, o.stats.data(), o.stats.size() * sizeof(or_stat_t));
}
void or_get_maxT(void* h, double* sorted_out, double* unsorted_out) {
    Oracle& o = *(Oracle*)h;
    if (sorted_out && !o.maxT_sorted.empty()) std::memcpy(sorted_out, o.maxT_sorted.data(), o.maxT_sorted.size() * 8);
    if (unsorted_out && !o.maxT_unsorted.empty()) std::memcpy(unsorted_out, o.maxT_unsorted.data(), o.maxT_unsorted.size() * 8);
}
double or_get_p_success(void* h) { return ((Oracle*)h)->p_success; }

}  // extern "C"
