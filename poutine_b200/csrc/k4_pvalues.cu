// K4 — binomial p-value table, observed statistics, thresholds, and the point-wise / family-wise
// empirical p-values (sm_100a; FP64, compiled with --fmad=false).
//
// Replaces: observed binomialTest calls (HC.java:3047-3057), the "n_k" p-value cache
// (HC.java:4226-4236), calc_pointwise_estimate_binom (HC.java:4812-4838) and
// calc_familywise_estimates_binom_combined_nulldists (HC.java:4558-4654).
#include <cooperative_groups.h>

#include <algorithm>

#include "binom_device.cuh"
#include "pb_internal.h"

namespace cg = cooperative_groups;

// ---- f[n][k] = binomialTest(n, k, p, GREATER_THAN), triangular layout n*(n+1)/2 + k -------------
__global__ void k4_ptable_kernel(double* __restrict__ tab, int32_t n_max, double p) {
    const int64_t total = (int64_t)(n_max + 1) * (n_max + 2) / 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // invert idx = n(n+1)/2 + k
    int64_t n = (int64_t)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while (n * (n + 1) / 2 > idx) n--;
    while ((n + 1) * (n + 2) / 2 <= idx) n++;
    const int k = (int)(idx - n * (n + 1) / 2);
    tab[idx] = pbm::binomial_test_greater((int)n, k, p);
}

// ---- observed p-value, k* and the n-histogram -------------------------------------------------------
#define OBS_SH 256
__global__ void k4_observed_kernel(AlleleMeta* __restrict__ alleles, int64_t n_slots, const double* __restrict__ tab,
                                   double* __restrict__ pobs, unsigned int* __restrict__ n_hist, int force_general, int has_miss,
                                   unsigned long long* __restrict__ n_fast) {
    // block-local histogram of the common small n (global atomics only for the tail and once per block)
    __shared__ unsigned int s_hist[OBS_SH], s_fast;
    for (int i = threadIdx.x; i < OBS_SH; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_fast = 0;
    __syncthreads();
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_slots) {
        AlleleMeta m = alleles[a];
        if (!(m.flags & PB_AF_IN_USE)) {
            pobs[a] = __longlong_as_double(0x7FF8000000000000ll);  // Double.NaN sentinel (HC.java:5891-5906)
        } else {
            const int no = m.obs_case + m.obs_ctrl;
            const double po = tab[(int64_t)no * (no + 1) / 2 + m.obs_case];  // HC.java:3050 / 3056
            pobs[a] = po;
            if (m.n < OBS_SH) atomicAdd(&s_hist[m.n], 1u); else atomicAdd(n_hist + m.n, 1u);
            // k* = min{k : f(n,k) <= p_obs}; fast path only if {k : f(n,k) <= p_obs} = [k*, n]
            const double* row = tab + (int64_t)m.n * (m.n + 1) / 2;
            int kstar = m.n + 1;
            bool upset = true;
            for (int k = 0; k <= m.n; k++) {
                const bool le = row[k] <= po;
                if (le && kstar > m.n) kstar = k;
                if (!le && kstar <= m.n) upset = false;
            }
            // missing labels: a replicate sees n_r = n - j labelled leaves; k*_j per j, packed (short lists only)
            uint32_t kpack = 0;
            bool upset_all = has_miss && m.n <= 7;
            if (upset_all) {
                    for (int j = 0; j <= m.n; j++) {
                        const int nr = m.n - j;
                        const double* rj = tab + (int64_t)nr * (nr + 1) / 2;
                        int kj = nr + 1;
                        for (int k = 0; k <= nr; k++) {
                            const bool le = rj[k] <= po;
                            if (le && kj > nr) kj = k;
                            if (!le && kj <= nr) upset_all = false;
                        }
                        kpack |= (uint32_t)kj << (4 * j);
                    }
            }
            alleles[a].kpack = kpack;
            uint32_t fl = m.flags & ~(PB_AF_FAST | PB_AF_MISS_OK);
            if (upset && !force_general) { fl |= PB_AF_FAST; atomicAdd(&s_fast, 1u); }
            if (upset_all && !force_general) fl |= PB_AF_MISS_OK;
            alleles[a].flags = fl;
            alleles[a].kstar = kstar;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OBS_SH; i += blockDim.x)
        if (s_hist[i]) atomicAdd(n_hist + i, s_hist[i]);
    if (threadIdx.x == 0 && s_fast) atomicAdd(n_fast, (unsigned long long)s_fast);
}

// ---- candidate threshold tau for the maxT pass ----------------------------------------------------
// expected #candidate alleles per replicate at tau_j = 2^-j:  sum_n hist[n] * f(n, ktau_n(tau_j))
#define TAU_LEVELS 64
__global__ void k4_tau_expect_kernel(const double* __restrict__ tab, const unsigned int* __restrict__ n_hist, int32_t n_max,
                                     double* __restrict__ expect) {
    __shared__ double sh[256];
    const int j = blockIdx.x;
    const double tau = ldexp(1.0, -j);
    double acc = 0;
    for (int n = threadIdx.x; n <= n_max; n += blockDim.x) {
        const unsigned int h = n_hist[n];
        if (!h) continue;
        const double* row = tab + (int64_t)n * (n + 1) / 2;
        for (int k = 0; k <= n; k++)
            if (row[k] <= tau) { acc += (double)h * row[k]; break; }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) expect[j] = sh[0];
}
__global__ void k4_tau_pick_kernel(const double* __restrict__ expect, double target, double user_tau, double* __restrict__ tau) {
    if (user_tau > 0) { *tau = user_tau; return; }
    int best = 0;
    for (int j = 0; j < TAU_LEVELS; j++)
        if (expect[j] >= target) best = j;
    *tau = ldexp(1.0, -best);
}
// ktau[n] = min{k : f(n,k) <= tau} (n+1 if none): {k >= ktau[n]} is a superset of {k : f(n,k) <= tau}
__global__ void k4_ktau_kernel(const double* __restrict__ tab, int32_t n_max, const double* __restrict__ tau, int32_t* __restrict__ ktau) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n > n_max) return;
    const double t = *tau;
    const double* row = tab + (int64_t)n * (n + 1) / 2;
    int kt = n + 1;
    for (int k = 0; k <= n; k++)
        if (row[k] <= t) { kt = k; break; }
    ktau[n] = kt;
}

// ---- work lists for K3b: 0 = no sweep needed, 1 = short lists (n < 8) on the tight path, 2 = long lists (n >= 8),
//      3 = short lists that need the general path -----------------------------------------------------------
__global__ void k4_classify_kernel(AlleleMeta* __restrict__ alleles, int64_t n_slots, const int32_t* __restrict__ ktau, int has_miss,
                                   int32_t* __restrict__ lists, unsigned int* __restrict__ counts, unsigned long long* __restrict__ swept_rows) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1;
    unsigned int my_rows = 0;       // label-matrix rows this allele's sweep reads per replicate word (K3b's algorithmic traffic)
    if (a < n_slots) {
        const AlleleMeta m = alleles[a];
        if (m.flags & PB_AF_IN_USE) {
            // no maxT candidate possible: f(n',k) > tau for every k, for every n' a replicate can see
            bool cand_free = ktau[m.n] > m.n;
            if (has_miss) {
                cand_free = m.n <= 7;
                for (int n2 = 0; n2 <= m.n && cand_free; n2++) cand_free = ktau[n2] > n2;
            }
            const bool tight = ((m.flags & (has_miss ? PB_AF_MISS_OK : PB_AF_FAST)) != 0) && cand_free;
            // short list, every label present, candidates possible but only for k >= ktau[n] >= 1 (so never for padding bits)
            const bool tight_cand = !tight && !has_miss && (m.flags & PB_AF_FAST) && m.n < 8 && ktau[m.n] >= 1;
            alleles[a].flags = (m.flags & ~(PB_AF_TIGHT | PB_AF_TIGHT_CAND)) | (tight ? PB_AF_TIGHT : 0u) | (tight_cand ? PB_AF_TIGHT_CAND : 0u);
            const bool no_sweep = tight && (has_miss ? m.n <= 1 : (m.n <= 1 || m.kstar <= 0 || m.kstar > m.n));
            cls = no_sweep ? 0 : (m.n >= 8 ? 2 : (tight ? 1 : 3));
            if (cls > 0) my_rows = (unsigned int)m.n;
        }
    }
    my_rows = __reduce_add_sync(0xffffffffu, my_rows);
    // list append: one global atomic per class per BLOCK (per-warp ballots -> block prefix in shared memory)
    __shared__ unsigned int s_wcount[8][4], s_base[4], s_rows[8];
    const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) s_rows[warp] = my_rows;
    unsigned int my_rank = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned int mask = __ballot_sync(0xffffffffu, cls == c);
        if (lane == 0) s_wcount[warp][c] = (unsigned int)__popc(mask);
        if (cls == c) my_rank = (unsigned int)__popc(mask & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned int tot = 0;
        for (int w = 0; w < 8; w++) { const unsigned int t = s_wcount[w][threadIdx.x]; s_wcount[w][threadIdx.x] = tot; tot += t; }
        s_base[threadIdx.x] = tot ? atomicAdd(counts + threadIdx.x, tot) : 0u;
    }
    if (threadIdx.x == 32) {      // one atomic per block (all blocks hit the same word)
        unsigned long long rows = 0;
        for (int w = 0; w < 8; w++) rows += s_rows[w];
        if (rows) atomicAdd(swept_rows, rows);
    }
    __syncthreads();
    if (cls >= 0) lists[(int64_t)cls * n_slots + s_base[cls] + s_wcount[warp][cls] + my_rank] = (int32_t)a;
}

// ---- the exchange, fused into the sort's load (protocol: pb_comm.cu) ---------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// every thread of the block calls it: threads 0..world-1 each wait for one peer's flag (in this rank's own block)
__device__ __forceinline__ void xchg_wait(const XchgView& x) {
    if (x.peers == nullptr) return;
    if ((int)threadIdx.x < x.world) {
        const unsigned long long* own = x.peers[x.rank];
        const unsigned long long* f = own + threadIdx.x;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(f) < x.epoch) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ull) { atomicExch((unsigned long long*)own + PB_XCHG_MAX_WORLD, 1ull); break; }   // a peer died
        }
    }
    __syncthreads();
}
// pooled minimum of replicate i over all ranks' buffers of the epoch (p >= 0: the bit patterns order like the doubles)
__device__ __forceinline__ unsigned long long xchg_load_min(const XchgView& x, const unsigned long long* __restrict__ own, int64_t i) {
    if (x.peers == nullptr) return own[i];
    unsigned long long v = own[i];
    for (int q = 0; q < x.world; q++)
        if (q != x.rank) v = min(v, ld_relaxed_sys(x.peers[q] + x.buf_off + i));
    x.pooled[i] = v;
    return v;
}

// ---- bitonic sort of maxT (m doubles in [0,1]; padded with +inf to a power of two) -----------------
#define BSORT_TILE 2048
// whole bitonic network in ONE cooperative launch: tile-local stages in shared memory, tile-crossing stages
// in global memory with a grid barrier in between (replaces ~40 dependent launches at m = 1e5)
// Sharded runs: the tile load IS the all-reduce(min) of maxT over the site shards — the peers' buffers are read over
// NVLink while the tile is filled (xchg_load_min), after the block has seen every rank's ready flag (xchg_wait).
__global__ void __launch_bounds__(1024) k4_sort_coop_kernel(const unsigned long long* __restrict__ maxT, int64_t m, int64_t n_pad,
                                                            double* __restrict__ a, const XchgView xv,
                                                            const unsigned int* __restrict__ run_if) {
    // run_if != nullptr: this launch is the fall-back behind the counting path (below) and runs only when that path met a
    // value outside its table (*run_if != 0); the decision stays on the device, the launch is always enqueued
    if (run_if != nullptr && *run_if == 0) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[BSORT_TILE];
    const int t = threadIdx.x;
    const int64_t n_tiles = n_pad / BSORT_TILE;
    const double INF = __longlong_as_double(0x7FF0000000000000ll);
    xchg_wait(xv);
    // local sort of every tile (all k <= BSORT_TILE), fused with the load / +inf padding
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * BSORT_TILE;
        const int64_t i0 = base + t, i1 = base + t + 1024;
        sh[t] = i0 < m ? __longlong_as_double((long long)xchg_load_min(xv, maxT, i0)) : INF;
        sh[t + 1024] = i1 < m ? __longlong_as_double((long long)xchg_load_min(xv, maxT, i1)) : INF;
        __syncthreads();
        for (int k = 2; k <= BSORT_TILE; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
                const double x = sh[lo], y = sh[hi];
                const bool up = ((base + lo) & k) == 0;
                if ((x > y) == up) { sh[lo] = y; sh[hi] = x; }
                __syncthreads();
            }
        a[i0] = sh[t];
        a[i1] = sh[t + 1024];
        __syncthreads();
    }
    grid.sync();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + t, gthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = 2 * BSORT_TILE; k <= n_pad; k <<= 1) {
        for (int64_t j = k >> 1; j >= BSORT_TILE; j >>= 1) {
            for (int64_t idx = gtid; idx < n_pad / 2; idx += gthreads) {
                const int64_t i = ((idx / j) * 2 * j) + (idx % j), l = i + j;
                const double x = a[i], y = a[l];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[l] = x; }
            }
            grid.sync();
        }
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t base = tile * BSORT_TILE;
            sh[t] = a[base + t];
            sh[t + 1024] = a[base + t + 1024];
            __syncthreads();
            const bool up = (base & k) == 0;   // k > BSORT_TILE: the direction is constant inside a tile
            for (int j = BSORT_TILE / 2; j > 0; j >>= 1) {
                const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
                const double x = sh[lo], y = sh[hi];
                if ((x > y) == up) { sh[lo] = y; sh[hi] = x; }
                __syncthreads();
            }
            a[base + t] = sh[t];
            a[base + t + 1024] = sh[t + 1024];
            __syncthreads();
        }
        grid.sync();
    }
}

// ---- counting path: sort maxT by counting it against the sorted p-value table ---------------------------------------------
// Every maxT[rep] is a table value f(n, k) (a minimum over table look-ups; +inf when nothing was in use), and the table is small
// (n_max <= 126 -> <= 8 128 entries) where m is 1e5 .. 1e7.  So: sort the TABLE once (it is kept across passes like the table
// itself), bin every replicate's pooled minimum by binary search (shared-memory histogram), prefix-sum the bins, and write
// maxT_sorted by expanding the runs.  ~25 us instead of the ~90 us bitonic network at m = 1e5, 0.4 instead of 5.9 ms at 1e7.
// The bin pass is also where a sharded run pools its shards (xchg_wait / xchg_load_min, as in the sort kernel).  A pooled
// value that is NOT in this rank's table (a peer saw a larger n_max) raises a flag and the bitonic network runs after all,
// on the pooled values.
#define TAB_MAX 8192          // table entries (+ one +inf bin) the counting path handles: n_max <= 126
__global__ void __launch_bounds__(1024) k4_tabsort_kernel(const double* __restrict__ tab, int total, unsigned long long* __restrict__ out, int n_pad) {
    extern __shared__ unsigned long long sk[];        // n_pad keys (p >= 0: bit patterns order like the values)
    const unsigned long long INF = 0x7FF0000000000000ull;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) sk[i] = i < total ? (unsigned long long)__double_as_longlong(tab[i]) : (i == total ? INF : ~0ull);
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n_pad / 2; t += blockDim.x) {
                const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
                const unsigned long long x = sk[lo], y = sk[hi];
                const bool up = (lo & k) == 0;
                if ((x > y) == up) { sk[lo] = y; sk[hi] = x; }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) out[i] = sk[i];
}

// bins: index of the FIRST table entry equal to the value (lower bound); entries [0, n_bins) are valid keys
__global__ void __launch_bounds__(512) k4_bin_kernel(const unsigned long long* __restrict__ maxT, int64_t m, const unsigned long long* __restrict__ keys,
                                                     int n_bins, unsigned int* __restrict__ hist, unsigned int* __restrict__ miss, const XchgView xv,
                                                     unsigned long long* __restrict__ pooled_out) {
    extern __shared__ unsigned int sh_hist[];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) sh_hist[i] = 0;
    xchg_wait(xv);                                     // (ends in a __syncthreads)
    __syncthreads();
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = xchg_load_min(xv, maxT, i);
        if (xv.peers == nullptr && pooled_out != nullptr) pooled_out[i] = v;
        int lo = 0, hi = n_bins;                       // first key >= v
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid] < v) lo = mid + 1; else hi = mid;
        }
        if (lo < n_bins && keys[lo] == v) atomicAdd(&sh_hist[lo], 1u);
        else bad = true;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(hist + i, sh_hist[i]);
    if (bad) atomicExch(miss, 1u);
}

// one block: inclusive prefix of the bins -> cum[i] = #replicates with value <= keys[i]
__global__ void __launch_bounds__(1024) k4_binscan_kernel(const unsigned int* __restrict__ hist, int n_bins, unsigned long long* __restrict__ cum) {
    __shared__ unsigned long long s_part[1024];
    const int per = (n_bins + 1023) / 1024;
    unsigned long long acc = 0;
    for (int k = 0; k < per; k++) { const int i = threadIdx.x * per + k; if (i < n_bins) acc += hist[i]; }
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - acc;
    for (int k = 0; k < per; k++) { const int i = threadIdx.x * per + k; if (i < n_bins) { run += hist[i]; cum[i] = run; } }
}

// maxT_sorted[i] = the key of the first bin whose cumulative count exceeds i
__global__ void k4_expand_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ cum, int n_bins, int64_t m,
                                 const unsigned int* __restrict__ miss, double* __restrict__ out) {
    if (*miss) return;                                  // the fall-back sort writes the array
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int lo = 0, hi = n_bins;                            // first bin with cum > i
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] <= (unsigned long long)i) lo = mid + 1; else hi = mid;
    }
    out[i] = __longlong_as_double((long long)keys[lo < n_bins ? lo : n_bins - 1]);
}

// ---- final per-site statistics -----------------------------------------------------------------------
__global__ void k4_stats_kernel(const int32_t* __restrict__ inf_site, int64_t S_inf, const AlleleMeta* __restrict__ alleles,
                                const double* __restrict__ pobs, const uint32_t* __restrict__ r, const double* __restrict__ maxT_sorted,
                                int64_t m, int64_t site_offset, int64_t site_index_base, pb_stat_out* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_inf) return;
    const double NaN = __longlong_as_double(0x7FF8000000000000ll);
    pb_stat_out o;
    o.site_index = (int32_t)(site_index_base + inf_site[i]);
    o.segsite_id = (int32_t)(site_offset + inf_site[i] + 1);
    int32_t rr[2], rm[2];
    double pw[2], fw[2], po[2];
    for (int a = 0; a < 2; a++) {
        const AlleleMeta mt = alleles[2 * i + a];
        const bool use = mt.flags & PB_AF_IN_USE;
        o.obs_counts[2 * a] = use ? mt.obs_case : 0;
        o.obs_counts[2 * a + 1] = use ? mt.obs_ctrl : 0;
        if (!use) { rr[a] = -1; rm[a] = -1; pw[a] = fw[a] = po[a] = NaN; continue; }
        po[a] = pobs[2 * i + a];
        rr[a] = (int32_t)r[2 * i + a];
        pw[a] = ((double)(rr[a] + 1)) / ((double)(m + 1));  // HC.java:4829-4833
        // r_maxT = #{rep : maxT[rep] <= p_obs}  (HC.java:4630): upper bound in the sorted array
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (maxT_sorted[mid] <= po[a]) lo = mid + 1; else hi = mid;
        }
        rm[a] = (int32_t)lo;
        fw[a] = ((double)lo + 1) / ((double)m + 1);        // HC.java:4632
    }
    o.a1_in_use = (alleles[2 * i].flags & PB_AF_IN_USE) ? 1 : 0;
    o.a2_in_use = (alleles[2 * i + 1].flags & PB_AF_IN_USE) ? 1 : 0;
    o.r_a1 = rr[0]; o.r_a2 = rr[1]; o.r_maxT_a1 = rm[0]; o.r_maxT_a2 = rm[1];
    o.obs_p_a1 = po[0]; o.obs_p_a2 = po[1]; o.pointwise_a1 = pw[0]; o.pointwise_a2 = pw[1];
    o.familywise_a1 = fw[0]; o.familywise_a2 = fw[1];
    out[i] = o;
}

// =====================================================================================================
static int ensure_bytes(pb_ctx* ctx, void** ptr, int64_t* cap, int64_t need) {
    if (need <= *cap) return PB_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    const int64_t bytes = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); *cap = 0; return PB_ERR_OOM; }
    *cap = bytes;
    return PB_OK;
}

// alloc_only (here and below): do the allocations of this step and return — pbk_assoc_prepare runs every allocation of
// pb_run_assoc up front, so that no cudaMalloc / cudaFree is issued once a sharded pass has kernels in flight that wait
// for a peer (an allocation may synchronise the device).
int pbk_ptable(pb_ctx* ctx, bool alloc_only) {
    // table rows 0..n_max for this context's p (the table is a pure function of (n_max, p): cached)
    const int32_t n_max = ctx->n_max;
    const double p = (double)ctx->n_case / (double)(ctx->n_case + ctx->n_ctrl);  // calc_background_p_success (HC.java:3335)
    if (ctx->ptable_nmax >= n_max && ctx->p_success == p && ctx->d_ptable) return PB_OK;
    ctx->p_success = p;
    const int64_t total = (int64_t)(n_max + 1) * (n_max + 2) / 2;
    int rc;
    int64_t cap = ctx->ptable_cap;
    void* pt = ctx->d_ptable;
    if ((rc = ensure_bytes(ctx, &pt, &cap, total * 8))) { ctx->d_ptable = nullptr; ctx->ptable_cap = 0; return rc; }
    ctx->d_ptable = (double*)pt; ctx->ptable_cap = cap;
    if (alloc_only) { ctx->ptable_nmax = -1; ctx->tabkeys_valid = false; return PB_OK; }      // (the buffer may have moved: the table is rebuilt by the real call)
    ctx->tabkeys_valid = false;
    k4_ptable_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_ptable, n_max, p);
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    ctx->ptable_nmax = n_max;
    return PB_OK;
}

int pbk_thresholds(pb_ctx* ctx, bool alloc_only) {
    const int64_t n_slots = 2 * ctx->S_inf;
    const int32_t n_max = ctx->n_max;
    cudaStream_t st = ctx->stream;
    int rc;
    // one scratch allocation: [expect 64 f64][tau f64][n_fast u64][list counts 4 x u32][fallback total u64][swept rows u64]
    //                         [ktau i32 x (n_max+1)][hist u32 x (n_max+1)]
    const int64_t n1 = (int64_t)n_max + 1;
    const int64_t need = TAU_LEVELS * 8 + 8 + 8 + 16 + 8 + 8 + n1 * 4 + n1 * 4;
    if ((rc = ensure_bytes(ctx, &ctx->d_thr, &ctx->thr_cap, need))) return rc;
    if (n_slots > ctx->work_cap) {
        if (ctx->d_work) cudaFree(ctx->d_work);
        ctx->d_work = nullptr;
        ctx->work_cap = n_slots + n_slots / 8 + 1024;
        PB_CUDA(cudaMalloc(&ctx->d_work, (size_t)ctx->work_cap * 4 * 4));
    }
    if (alloc_only) return PB_OK;
    char* base = (char*)ctx->d_thr;
    double* d_expect = (double*)base;
    ctx->d_tau = (double*)(base + TAU_LEVELS * 8);
    unsigned long long* d_nfast = (unsigned long long*)(base + TAU_LEVELS * 8 + 8);
    unsigned int* d_lcount = (unsigned int*)(base + TAU_LEVELS * 8 + 16);
    ctx->d_nfast = d_nfast;                 // {n_fast u64, list counts 4 x u32, fallback total u64}: read back once, at the end of pbk_assoc
    ctx->d_lcount = d_lcount;
    ctx->d_fb_total = (unsigned long long*)(base + TAU_LEVELS * 8 + 32);
    unsigned long long* d_swept = (unsigned long long*)(base + TAU_LEVELS * 8 + 40);
    ctx->d_ktau = (int32_t*)(base + TAU_LEVELS * 8 + 48);
    unsigned int* d_hist = (unsigned int*)(base + TAU_LEVELS * 8 + 48 + n1 * 4);
    PB_CUDA(cudaMemsetAsync(ctx->d_thr, 0, (size_t)need, st));
    const int T = 256;
    const int force_general = (ctx->cfg.flags & PB_FLAG_FORCE_GENERAL_NULL) ? 1 : 0;
    k4_observed_kernel<<<(unsigned)std::max<int64_t>(1, (n_slots + T - 1) / T), T, 0, st>>>(ctx->d_alleles, n_slots, ctx->d_ptable, ctx->d_pobs, d_hist,
                                                                      force_general, ctx->n_miss > 0 ? 1 : 0, d_nfast);
    k4_tau_expect_kernel<<<TAU_LEVELS, 256, 0, st>>>(ctx->d_ptable, d_hist, n_max, d_expect);
    k4_tau_pick_kernel<<<1, 1, 0, st>>>(d_expect, 16.0, ctx->cfg.maxt_tau, ctx->d_tau);
    k4_ktau_kernel<<<(unsigned)((n1 + T - 1) / T), T, 0, st>>>(ctx->d_ptable, n_max, ctx->d_tau, ctx->d_ktau);
    k4_classify_kernel<<<(unsigned)std::max<int64_t>(1, (n_slots + T - 1) / T), T, 0, st>>>(ctx->d_alleles, n_slots, ctx->d_ktau, ctx->n_miss > 0 ? 1 : 0,
                                                                          ctx->d_work, d_lcount, d_swept);
    ctx->tm.n_launches += 5;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pbk_pvalues(pb_ctx* ctx, bool alloc_only) {
    const int64_t m = ctx->cfg.n_replicates;
    cudaStream_t st = ctx->stream;
    int rc;
    int64_t n_pad = BSORT_TILE;
    while (n_pad < m) n_pad <<= 1;
    {
        int64_t cap = ctx->sorted_cap * 8;
        void* ps = ctx->d_maxT_sorted;
        if ((rc = ensure_bytes(ctx, &ps, &cap, n_pad * 8))) { ctx->d_maxT_sorted = nullptr; ctx->sorted_cap = 0; return rc; }
        ctx->d_maxT_sorted = (double*)ps; ctx->sorted_cap = cap / 8;
    }
    {
        int64_t cap = ctx->stats_cap * (int64_t)sizeof(pb_stat_out);
        void* ps = ctx->d_stats;
        if ((rc = ensure_bytes(ctx, &ps, &cap, ctx->S_inf * (int64_t)sizeof(pb_stat_out)))) { ctx->d_stats = nullptr; ctx->stats_cap = 0; return rc; }
        ctx->d_stats = (pb_stat_out*)ps; ctx->stats_cap = cap / (int64_t)sizeof(pb_stat_out);
    }
    if (ctx->sort_blocks_per_sm < 0)   // occupancy belongs to the context's device (contexts of one process may sit on different GPUs / threads)
        PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->sort_blocks_per_sm, k4_sort_coop_kernel, 1024, 0));
    if (!ctx->d_tabkeys) {             // counting path scratch (fixed size)
        PB_CUDA(cudaMalloc(&ctx->d_tabkeys, (size_t)TAB_MAX * 8));
        PB_CUDA(cudaMalloc(&ctx->d_binhist, (size_t)(TAB_MAX + 2) * 4));
        PB_CUDA(cudaMalloc(&ctx->d_bincum, (size_t)(TAB_MAX + 1) * 8));
        PB_CUDA(cudaFuncSetAttribute(k4_tabsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAB_MAX * 8));
    }
    if (!ctx->xchg.active && ctx->pooled_own_cap < m) {     // single context: a copy of maxT for the fall-back sort to read
        if (ctx->d_pooled_own) cudaFree(ctx->d_pooled_own);
        ctx->d_pooled_own = nullptr;
        PB_CUDA(cudaMalloc(&ctx->d_pooled_own, (size_t)m * 8));
        ctx->pooled_own_cap = m;
    }
    if (alloc_only) return PB_OK;
    const int T = 256;
    const int64_t grid = std::min<int64_t>(n_pad / BSORT_TILE, (int64_t)std::max(ctx->sort_blocks_per_sm, 1) * ctx->sm_count);
    const int64_t tab_total = (int64_t)(ctx->ptable_nmax + 1) * (ctx->ptable_nmax + 2) / 2;
    const bool counting = tab_total + 1 <= TAB_MAX && m >= 4096 && !getenv("PB_K4_BITONIC");
    if (counting) {
        const int n_bins = (int)tab_total + 1;          // + the +inf bin
        int tp = 2;
        while (tp < n_bins) tp <<= 1;
        if (!ctx->tabkeys_valid) {                      // sorted table keys: a function of the table, kept with it
            k4_tabsort_kernel<<<1, 1024, (size_t)tp * 8, st>>>(ctx->d_ptable, (int)tab_total, ctx->d_tabkeys, tp);
            ctx->tabkeys_valid = true;
            ctx->tm.n_launches += 1;
        }
        PB_CUDA(cudaMemsetAsync(ctx->d_binhist, 0, (size_t)(TAB_MAX + 2) * 4, st));       // bins + the miss flag behind them
        unsigned int* d_miss = ctx->d_binhist + TAB_MAX + 1;
        const XchgView xv = pbx_view(ctx);
        const unsigned bblocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((m + 2047) / 2048, (int64_t)ctx->sm_count * 2));
        k4_bin_kernel<<<bblocks, 512, (size_t)n_bins * 4, st>>>(ctx->d_maxT, m, ctx->d_tabkeys, n_bins, ctx->d_binhist, d_miss, xv,
                                                                 ctx->xchg.active ? nullptr : ctx->d_pooled_own);
        k4_binscan_kernel<<<1, 1024, 0, st>>>(ctx->d_binhist, n_bins, ctx->d_bincum);
        k4_expand_kernel<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(ctx->d_tabkeys, ctx->d_bincum, n_bins, m, d_miss, ctx->d_maxT_sorted);
        // fall-back, decided on the device: the bitonic network over the POOLED values when some value was not in the table
        const unsigned long long* a_maxT = ctx->xchg.active ? ctx->xchg.d_pooled : ctx->d_pooled_own;
        int64_t a_m = m, a_npad = n_pad;
        double* a_out = ctx->d_maxT_sorted;
        XchgView a_xv{};
        a_xv.peers = nullptr; a_xv.world = 1;
        const unsigned int* a_run = d_miss;
        void* args[] = {(void*)&a_maxT, (void*)&a_m, (void*)&a_npad, (void*)&a_out, (void*)&a_xv, (void*)&a_run};
        PB_CUDA(cudaLaunchCooperativeKernel((const void*)k4_sort_coop_kernel, dim3((unsigned)grid), dim3(1024), args, 0, st));
        ctx->tm.n_launches += 4;
    } else {
        const unsigned long long* a_maxT = ctx->d_maxT;
        int64_t a_m = m, a_npad = n_pad;
        double* a_out = ctx->d_maxT_sorted;
        XchgView a_xv = pbx_view(ctx);
        const unsigned int* a_run = nullptr;
        void* args[] = {(void*)&a_maxT, (void*)&a_m, (void*)&a_npad, (void*)&a_out, (void*)&a_xv, (void*)&a_run};
        PB_CUDA(cudaLaunchCooperativeKernel((const void*)k4_sort_coop_kernel, dim3((unsigned)grid), dim3(1024), args, 0, st));
        ctx->tm.n_launches += 1;
    }
    k4_stats_kernel<<<(unsigned)std::max<int64_t>(1, (ctx->S_inf + T - 1) / T), T, 0, st>>>(ctx->d_inf_site, ctx->S_inf, ctx->d_alleles, ctx->d_pobs, ctx->d_r,
                                                                      ctx->d_maxT_sorted, m, ctx->cfg.site_offset, ctx->site_index_base, ctx->d_stats);
    ctx->tm.n_launches += 1;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}
