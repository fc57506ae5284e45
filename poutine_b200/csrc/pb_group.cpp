// pb_group — all GPUs of one box behind ONE host thread (SURVEY §8b "Threading": the reference calls
// count_all_homoplasy_events / assoc from the main thread of one JVM, HC.java:299 / 314, and cannot launch itself once per
// GPU).  A group owns one context per device; sites are sharded into contiguous 64-site-word ranges (shard g = words
// [W*g/G, W*(g+1)/G)), every context draws the same Philox permutations, and the only exchange — the minimum of maxT[m] over
// the shards — runs over peer memory inside K4's sort kernel (pb_comm.cu).  Each context is driven by its own persistent
// worker thread, so the ~40 launches of a pass are enqueued on all GPUs at once; the caller's thread only posts the call
// and merges the outputs.  No allocation or free happens between a shard's ready signal and the end of its pass, so a
// worker can never block a peer that is waiting inside the exchange.
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "pb_internal.h"

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define PB_CPU_RELAX() _mm_pause()
#else
#define PB_CPU_RELAX() std::this_thread::yield()
#endif

int pb_assoc_prepare_internal(pb_ctx* ctx);     // pb_api.cu
int pb_labels_alloc_internal(pb_ctx* ctx);      // pb_api.cu

namespace {

// One persistent thread per device.  A call is handed over through atomics the worker spins on for a short while before it
// goes to sleep on the condition variable: back-to-back calls (a pass is four hand-overs) then cost microseconds, not the
// tens of microseconds of a futex wake-up, and an idle group costs nothing.
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    std::atomic<int> has_job{0}, done{0};
    bool quit = false;
    int rc = 0;

    static constexpr int SPIN = 20000;     // ~100-200 us of polling

    void loop() {
        for (;;) {
            bool got = false;
            for (int i = 0; i < SPIN && !got; i++) {
                if (has_job.load(std::memory_order_acquire)) got = true;
                else PB_CPU_RELAX();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has_job.load(std::memory_order_acquire) || quit; });
                if (quit && !has_job.load(std::memory_order_acquire)) return;
            }
            std::function<int()> j;
            {
                std::lock_guard<std::mutex> lk(mu);
                j = std::move(job);
                has_job.store(0, std::memory_order_relaxed);
            }
            rc = j();
            {
                std::lock_guard<std::mutex> lk(mu);
                done.store(1, std::memory_order_release);
            }
            cv.notify_all();
        }
    }
    void post(std::function<int()> j) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(j);
            done.store(0, std::memory_order_relaxed);
            has_job.store(1, std::memory_order_release);
        }
        cv.notify_all();
    }
    int wait() {
        for (int i = 0; i < SPIN; i++) {
            if (done.load(std::memory_order_acquire)) return rc;
            PB_CPU_RELAX();
        }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done.load(std::memory_order_acquire) != 0; });
        return rc;
    }
};

}  // namespace

struct pb_group {
    std::vector<pb_ctx*> ctx;
    std::vector<Worker*> workers;
    pb_config cfg{};
    int64_t S = 0, W = 0;
    std::vector<int64_t> w0;          // G+1 shard boundaries (site words)
    std::vector<int64_t> n_inf;       // informative sites per shard after run_events
    std::string err;
    bool events_done = false;

    int G() const { return (int)ctx.size(); }
    // run f(g) on every worker, return the first failure (message copied from that shard's context)
    int all(const std::function<int(int)>& f) {
        for (int g = 0; g < G(); g++) workers[g]->post([f, g] { return f(g); });
        int first = 0;
        for (int g = 0; g < G(); g++) {
            const int rc = workers[g]->wait();
            if (rc != 0 && first == 0) { first = rc; err = "shard " + std::to_string(g) + " (device " + std::to_string(ctx[g]->device) + "): " + pb_last_error(ctx[g]); }
        }
        return first;
    }
};

static thread_local std::string g_group_err;

extern "C" const char* pb_group_last_error(pb_group* g) { return g ? g->err.c_str() : g_group_err.c_str(); }

extern "C" void pb_group_destroy(pb_group* g) {
    if (!g) return;
    for (Worker* w : g->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    // every stream drains before any block that a peer may still be reading goes away
    for (pb_ctx* c : g->ctx) if (c) { cudaSetDevice(c->device); cudaDeviceSynchronize(); }
    for (pb_ctx* c : g->ctx) pb_destroy(c);
    delete g;
}

extern "C" int pb_group_create(const pb_config* cfg, const int32_t* devices, int32_t n_devices, pb_group** out) {
    if (!cfg || !devices || !out || n_devices < 1 || n_devices > PB_XCHG_MAX_WORLD) {
        g_group_err = "pb_group_create: need 1..64 device ordinals";
        return PB_ERR_INVALID;
    }
    *out = nullptr;
    // test hook: PB_GROUP_ALLOW_SAME_DEVICE lets several shards share one GPU, so that the exchange protocol and the merge
    // logic can be exercised on a one-GPU box (never useful in production: the shards then compete for the same SMs)
    const bool allow_dup = getenv("PB_GROUP_ALLOW_SAME_DEVICE") != nullptr;
    for (int i = 0; i < n_devices && !allow_dup; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) { g_group_err = "pb_group_create: a device is listed twice"; return PB_ERR_INVALID; }
    pb_group* g = new pb_group();
    g->cfg = *cfg;
    for (int i = 0; i < n_devices; i++) {
        pb_config c = *cfg;
        c.device = devices[i];
        pb_ctx* ctx = nullptr;
        const int rc = pb_create(&c, &ctx);
        if (rc != 0) {
            g_group_err = std::string("pb_group_create (device ") + std::to_string(devices[i]) + "): " + pb_last_error(nullptr);
            pb_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(ctx);
    }
    if (n_devices > 1) {
        const int rc = pbx_connect_local(g->ctx.data(), n_devices);
        if (rc != 0) {
            for (pb_ctx* c : g->ctx) if (!c->err.empty()) { g_group_err = "pb_group_create: " + c->err; break; }
            pb_group_destroy(g);
            return rc;
        }
    }
    for (int i = 0; i < n_devices; i++) {
        Worker* w = new Worker();
        w->th = std::thread([w] { w->loop(); });
        g->workers.push_back(w);
    }
    *out = g;
    return PB_OK;
}

extern "C" int32_t pb_group_size(pb_group* g) { return g ? g->G() : 0; }
extern "C" pb_ctx* pb_group_ctx(pb_group* g, int32_t i) { return (g && i >= 0 && i < g->G()) ? g->ctx[(size_t)i] : nullptr; }

extern "C" int pb_group_set_tree(pb_group* g, int32_t n_nodes, const int32_t* parent, const uint8_t* is_leaf) {
    if (!g) return PB_ERR_INVALID;
    g->events_done = false;
    const int rc = g->all([&](int i) { return pb_set_tree(g->ctx[(size_t)i], n_nodes, parent, is_leaf); });
    if (rc == 0 && g->S) {   // pb_set_tree re-ran pb_set_sites with the shard's own count: nothing to redo, the shards are unchanged
    }
    return rc;
}

extern "C" int pb_group_set_labels(pb_group* g, int32_t n_samples, const int32_t* sample_node, const uint8_t* label) {
    if (!g) return PB_ERR_INVALID;
    g->events_done = false;
    return g->all([&](int i) { return pb_set_labels(g->ctx[(size_t)i], n_samples, sample_node, label); });
}

extern "C" int pb_group_set_sites(pb_group* g, int64_t n_sites) {
    if (!g) return PB_ERR_INVALID;
    const int G = g->G();
    const int64_t W = (n_sites + 63) / 64;
    if (n_sites < 1 || W < G) { g->err = "pb_group_set_sites: need at least one 64-site word per device"; return PB_ERR_INVALID; }
    g->S = n_sites; g->W = W;
    g->w0.assign((size_t)G + 1, 0);
    for (int i = 0; i <= G; i++) g->w0[(size_t)i] = W * i / G;
    g->events_done = false;
    return g->all([&](int i) {
        pb_ctx* c = g->ctx[(size_t)i];
        const int64_t s0 = 64 * g->w0[(size_t)i], s1 = std::min<int64_t>(n_sites, 64 * g->w0[(size_t)i + 1]);
        c->cfg.site_offset = g->cfg.site_offset + s0;     // segsite_id stays alignment-global
        c->site_index_base = s0;                          // so does pb_stat_out.site_index
        return pb_set_sites(c, s1 - s0);
    });
}

extern "C" int pb_group_shard_range(pb_group* g, int32_t i, int64_t* site_begin, int64_t* site_end) {
    if (!g || i < 0 || i >= g->G() || g->S == 0) return PB_ERR_INVALID;
    if (site_begin) *site_begin = 64 * g->w0[(size_t)i];
    if (site_end) *site_end = std::min<int64_t>(g->S, 64 * g->w0[(size_t)i + 1]);
    return PB_OK;
}

// a column tile in alignment-global word coordinates goes to every shard it overlaps (the copies run in parallel)
template <typename F>
static int put_tile(pb_group* g, int64_t word_begin, int64_t n_words, const char* what, F&& f) {
    if (!g) return PB_ERR_INVALID;
    if (g->S == 0) { g->err = std::string(what) + ": call pb_group_set_sites first"; return PB_ERR_INVALID; }
    if (word_begin < 0 || n_words < 1 || word_begin + n_words > g->W) { g->err = std::string(what) + ": tile out of range"; return PB_ERR_INVALID; }
    g->events_done = false;
    return g->all([&](int i) {
        const int64_t a = std::max(word_begin, g->w0[(size_t)i]), b = std::min(word_begin + n_words, g->w0[(size_t)i + 1]);
        if (a >= b) return 0;
        return f(g->ctx[(size_t)i], a - g->w0[(size_t)i], b - a, a - word_begin);
    });
}

extern "C" int pb_group_put_planes(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* B0,
                                   const uint64_t* B1, const uint64_t* V) {
    return put_tile(g, word_begin, n_words, "pb_group_put_planes", [&](pb_ctx* c, int64_t local_w, int64_t nw, int64_t src_off) {
        return pb_put_planes(c, local_w, nw, src_pitch_words, B0 + src_off, B1 + src_off, V + src_off);
    });
}

extern "C" int pb_group_put_planes_biallelic(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* X,
                                             const uint64_t* V, const uint64_t* colmask) {
    return put_tile(g, word_begin, n_words, "pb_group_put_planes_biallelic", [&](pb_ctx* c, int64_t local_w, int64_t nw, int64_t src_off) {
        return pb_put_planes_biallelic(c, local_w, nw, src_pitch_words, X + src_off, V ? V + src_off : nullptr, colmask + 4 * src_off);
    });
}

// asynchronous forms: every shard's copies are queued (each on its own GPU's upload stream) and the call returns; the host
// buffers stay in use until pb_group_upload_sync.  One host thread can thus keep all PCIe links busy at once.
extern "C" int pb_group_put_planes_async(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* B0,
                                         const uint64_t* B1, const uint64_t* V) {
    return put_tile(g, word_begin, n_words, "pb_group_put_planes_async", [&](pb_ctx* c, int64_t local_w, int64_t nw, int64_t src_off) {
        return pb_put_planes_async(c, local_w, nw, src_pitch_words, B0 + src_off, B1 + src_off, V + src_off);
    });
}

extern "C" int pb_group_put_planes_biallelic_async(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                                                   const uint64_t* X, const uint64_t* V, const uint64_t* colmask) {
    return put_tile(g, word_begin, n_words, "pb_group_put_planes_biallelic_async", [&](pb_ctx* c, int64_t local_w, int64_t nw, int64_t src_off) {
        return pb_put_planes_biallelic_async(c, local_w, nw, src_pitch_words, X + src_off, V ? V + src_off : nullptr, colmask + 4 * src_off);
    });
}

extern "C" int pb_group_upload_sync(pb_group* g) {
    if (!g) return PB_ERR_INVALID;
    return g->all([&](int i) { return pb_upload_sync(g->ctx[(size_t)i]); });
}

// pb_expect_site_output for every shard: shard i's records belong at per_site + (its first site), where pb_group_run_events
// delivers them anyway
extern "C" int pb_group_expect_site_output(pb_group* g, pb_site_out* per_site) {
    if (!g) return PB_ERR_INVALID;
    if (g->S == 0) { g->err = "pb_group_expect_site_output: call pb_group_set_sites first"; return PB_ERR_INVALID; }
    return g->all([&](int i) { return pb_expect_site_output(g->ctx[(size_t)i], per_site ? per_site + 64 * g->w0[(size_t)i] : nullptr); });
}

extern "C" int pb_group_run_events(pb_group* g, pb_site_out* per_site, pb_class_counts* counts) {
    if (!g) return PB_ERR_INVALID;
    if (g->S == 0) { g->err = "pb_group_run_events: tree and planes must be set"; return PB_ERR_INVALID; }
    const int G = g->G();
    if (G > 1) {
        // label sharing (pb_comm.cu): the shards sweep the same label matrix, so each draws 1/G of it and pulls the rest from
        // its peers.  The matrices are allocated on every shard first, then mapped across the group (again only if one moved).
        auto all_mapped = [&] {
            for (int i = 0; i < G; i++) {
                const pb_ctx* c = g->ctx[(size_t)i];
                if (!c->xchg.lab_active || c->xchg.lab_c[i] != c->d_Xcase || !c->d_Xcase) return false;
            }
            return true;
        };
        if (!all_mapped() && g->ctx[0]->P > 0 && g->cfg.mode == PB_MODE_PHILOX) {     // (once; again only if a matrix had to grow)
            { const int rc0 = g->all([&](int i) { return pb_labels_alloc_internal(g->ctx[(size_t)i]); }); if (rc0) return rc0; }
            const int rc0 = pbx_labels_connect_local(g->ctx.data(), G);
            if (rc0) { for (pb_ctx* c : g->ctx) if (!c->err.empty()) { g->err = c->err; break; } return rc0; }
        }
    }
    std::vector<pb_class_counts> cc((size_t)G);
    const int rc = g->all([&](int i) {
        return pb_run_events(g->ctx[(size_t)i], per_site ? per_site + 64 * g->w0[(size_t)i] : nullptr, &cc[(size_t)i]);
    });
    if (rc) return rc;
    g->n_inf.assign((size_t)G, 0);
    pb_class_counts sum{};
    for (int i = 0; i < G; i++) {
        g->n_inf[(size_t)i] = g->ctx[(size_t)i]->S_inf;
        sum.n_no_allele += cc[(size_t)i].n_no_allele; sum.n_mono += cc[(size_t)i].n_mono; sum.n_bi += cc[(size_t)i].n_bi;
        sum.n_tri += cc[(size_t)i].n_tri; sum.n_quad += cc[(size_t)i].n_quad; sum.n_raw_events += cc[(size_t)i].n_raw_events;
    }
    if (counts) *counts = sum;
    g->events_done = true;
    return PB_OK;
}

extern "C" int pb_group_run_assoc(pb_group* g, const uint32_t* replay_perms, pb_stat_out* per_inf_site, int64_t cap, int64_t* n_inf,
                                  double* maxT_sorted) {
    if (!g) return PB_ERR_INVALID;
    if (!g->events_done) { g->err = "pb_group_run_assoc: call pb_group_run_events first"; return PB_ERR_INVALID; }
    const int G = g->G();
    std::vector<int64_t> off((size_t)G + 1, 0);
    for (int i = 0; i < G; i++) off[(size_t)i + 1] = off[(size_t)i] + g->n_inf[(size_t)i];
    const int64_t total = off[(size_t)G];
    if (n_inf) *n_inf = total;
    // subset_by_min_hcount over the WHOLE alignment (HC.java:3866-3869): a shard may be empty, the run may not
    if (total == 0) { g->err = "min_hcount is set too high and has filtered out all segregating sites"; return PB_ERR_NO_INFORMATIVE; }
    if (per_inf_site && cap < total) { g->err = "pb_group_run_assoc: per_inf_site too small"; return PB_ERR_INVALID; }
    // round 1: every allocation of the pass, on every shard; round 2: the pass.  A shard's last kernel waits for its peers
    // inside the exchange, and an allocation may wait for its device: none may be pending when the first shard gets there.
    // (skipped when every shard's buffers were already sized for at least this pass: they only grow)
    bool sized = true;
    for (int i = 0; i < G; i++) {
        const pb_ctx* c = g->ctx[(size_t)i];
        sized = sized && c->prep_P == c->P && c->prep_sinf >= c->S_inf && c->prep_nmax >= c->n_max;
    }
    if (!sized) { const int rc = g->all([&](int i) { return pb_assoc_prepare_internal(g->ctx[(size_t)i]); }); if (rc) return rc; }
    return g->all([&](int i) {
        int64_t ni = 0;
        return pb_run_assoc(g->ctx[(size_t)i], replay_perms, per_inf_site ? per_inf_site + off[(size_t)i] : nullptr, g->n_inf[(size_t)i], &ni,
                            i == 0 ? maxT_sorted : nullptr);      // every shard holds the same pooled, sorted maxT
    });
}

extern "C" int pb_group_run_extras(pb_group* g, double* exact_pointwise, double* fisher_p) {
    if (!g) return PB_ERR_INVALID;
    const int G = g->G();
    std::vector<int64_t> off((size_t)G + 1, 0);
    for (int i = 0; i < G; i++) off[(size_t)i + 1] = off[(size_t)i] + (i < (int)g->n_inf.size() ? g->n_inf[(size_t)i] : 0);
    return g->all([&](int i) {
        if (g->n_inf[(size_t)i] == 0) return 0;
        return pb_run_extras(g->ctx[(size_t)i], exact_pointwise ? exact_pointwise + 2 * off[(size_t)i] : nullptr, fisher_p ? fisher_p + off[(size_t)i] : nullptr);
    });
}

// times: the slowest shard; counts: summed
extern "C" int pb_group_get_timings(pb_group* g, pb_timings* out) {
    if (!g || !out) return PB_ERR_INVALID;
    pb_timings t{};
    for (int i = 0; i < g->G(); i++) {
        const pb_timings& s = g->ctx[(size_t)i]->tm;
        t.ms_events_kernel = std::max(t.ms_events_kernel, s.ms_events_kernel); t.ms_events_total = std::max(t.ms_events_total, s.ms_events_total);
        t.ms_tally = std::max(t.ms_tally, s.ms_tally); t.ms_ptable = std::max(t.ms_ptable, s.ms_ptable);
        t.ms_null_gen = std::max(t.ms_null_gen, s.ms_null_gen); t.ms_null_count = std::max(t.ms_null_count, s.ms_null_count);
        t.ms_allreduce = std::max(t.ms_allreduce, s.ms_allreduce); t.ms_pvalues = std::max(t.ms_pvalues, s.ms_pvalues);
        t.ms_assoc_total = std::max(t.ms_assoc_total, s.ms_assoc_total);
        t.n_launches += s.n_launches; t.k1_algorithmic_bytes += s.k1_algorithmic_bytes; t.n_informative += s.n_informative;
        t.n_alleles_in_use += s.n_alleles_in_use; t.n_extant_events += s.n_extant_events; t.n_fast_alleles += s.n_fast_alleles;
        t.n_maxt_fallback_reps += s.n_maxt_fallback_reps; t.n_swept_rows += s.n_swept_rows; t.eager_words += s.eager_words; t.tailed_words += s.tailed_words; t.early_sites += s.early_sites;
        t.ms_gen_kernel = std::max(t.ms_gen_kernel, s.ms_gen_kernel); t.gen_shard_frac = std::max(t.gen_shard_frac, s.gen_shard_frac);
        t.plane_mode = std::max(t.plane_mode, s.plane_mode);
    }
    *out = t;
    return PB_OK;
}

// Stopwatch over all shards: marks on every context's stream; the time is the slowest shard's.
extern "C" int pb_group_stopwatch_mark(pb_group* g, int32_t which) {
    if (!g) return PB_ERR_INVALID;
    for (pb_ctx* c : g->ctx) { const int rc = pb_stopwatch_mark(c, which); if (rc) { g->err = pb_last_error(c); return rc; } }
    return PB_OK;
}
extern "C" int pb_group_stopwatch_ms(pb_group* g, float* ms) {
    if (!g || !ms) return PB_ERR_INVALID;
    float best = 0;
    for (pb_ctx* c : g->ctx) { float t = 0; const int rc = pb_stopwatch_ms(c, &t); if (rc) { g->err = pb_last_error(c); return rc; } best = std::max(best, t); }
    *ms = best;
    return PB_OK;
}
