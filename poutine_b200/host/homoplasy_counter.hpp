// C++ host mirror of the two reference entry points libpoutine_b200.so replaces (header only, over the C ABI of
// include/poutine_b200.h):
//
//     Homoplasy_Counter.count_all_homoplasy_events(tree, seg_sites, physical_poss)   HC.java:1402 (called HC.java:299)
//     Homoplasy_Counter.assoc(all_events, phenos)                                    HC.java:2029 (called HC.java:314)
//
// Same names, argument meaning and error behaviour as the reference: the reference prints a message and calls
// System.exit(-1) (e.g. HC.java:3267-3271, 3866-3869); here the failure is an exception the caller maps onto that
// (poutine_main.cpp does).  The reference is Java and no JVM exists in this image, so this is the compiled-language
// host; poutine_b200/counter.py is the same mirror over ctypes.  All compute happens on the GPU inside the library:
// there is no CPU fallback, creating a counter without an sm_100 device throws.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "poutine_b200.h"

namespace poutine {

struct PoutineError : std::runtime_error {
    int status;
    PoutineError(int st, const std::string& msg) : std::runtime_error("[pb status " + std::to_string(st) + "] " + msg), status(st) {}
};

// "min_hcount = .. is set too high and has filtered out all segregating sites" (HC.java:3866-3869)
struct NoInformativeSites : PoutineError {
    using PoutineError::PoutineError;
};

struct Options {
    int64_t num_replicates = 100000;   // HC.java:89
    int32_t min_hcount = 1;            // HC.java:98
    std::vector<int32_t> devices{0};   // CUDA ordinals: more than one = the sites are sharded over the GPUs of the box (pb_group)
    uint64_t seed = 0;
    bool replay = false;               // PB_MODE_REPLAY: assoc() consumes recorded permutations
    int64_t site_offset = 0;           // 64-aligned global index of the first site (a host that shards over processes itself)
    uint32_t flags = 0;                // PB_FLAG_*
};

// page-locked host array (cudaHostAlloc through the library): upload tiles and result buffers
template <class T>
class PinnedArray {
  public:
    PinnedArray() = default;
    explicit PinnedArray(size_t n) { reset(n); }
    PinnedArray(const PinnedArray&) = delete;
    PinnedArray& operator=(const PinnedArray&) = delete;
    PinnedArray(PinnedArray&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    PinnedArray& operator=(PinnedArray&& o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~PinnedArray() { release(); }
    void reset(size_t n) {
        release();
        void* p = nullptr;
        int rc = pb_host_alloc(&p, (n ? n : 1) * sizeof(T));
        if (rc != 0) throw PoutineError(rc, std::string("pb_host_alloc: ") + pb_last_error(nullptr));
        p_ = static_cast<T*>(p);
        n_ = n;
    }
    T* data() { return p_; }
    const T* data() const { return p_; }
    size_t size() const { return n_; }
    T& operator[](size_t i) { return p_[i]; }
    const T& operator[](size_t i) const { return p_[i]; }

  private:
    void release() {
        if (p_) pb_host_free(p_);
        p_ = nullptr;
        n_ = 0;
    }
    T* p_ = nullptr;
    size_t n_ = 0;
};

// Result of assoc(): one Binomial_Test_Stat (+ obs_counts) per homoplasically informative site, and the sorted maxT null.
struct AssocResult {
    std::vector<pb_stat_out> stats;
    std::vector<double> maxT_sorted;
};

// One object = the whole box: the library's pb_group shards the sites over Options::devices, drives each GPU from its own
// worker thread and merges the outputs, so the host stays the single-threaded call() the reference is (HC.java:230-319).
class HomoplasyCounter {
  public:
    explicit HomoplasyCounter(const Options& o) : m_(o.num_replicates) {
        pb_config cfg{};
        cfg.device = 0;
        cfg.mode = o.replay ? PB_MODE_REPLAY : PB_MODE_PHILOX;
        cfg.n_replicates = o.num_replicates;
        cfg.min_hcount = o.min_hcount;
        cfg.seed = o.seed;
        cfg.site_offset = o.site_offset;
        cfg.flags = o.flags;
        if (o.devices.empty()) throw PoutineError(PB_ERR_INVALID, "Options::devices is empty");
        int rc = pb_group_create(&cfg, o.devices.data(), (int32_t)o.devices.size(), &ctx_);
        if (rc != 0) throw PoutineError(rc, pb_group_last_error(nullptr));
    }
    HomoplasyCounter(const HomoplasyCounter&) = delete;
    HomoplasyCounter& operator=(const HomoplasyCounter&) = delete;
    ~HomoplasyCounter() {
        if (ctx_) pb_group_destroy(ctx_);
    }

    // ---- inputs: the reference's NewickTree / phenos / seg_sites, flattened --------------------------------------
    void set_tree(const std::vector<int32_t>& parent, const std::vector<uint8_t>& is_leaf) {
        if (parent.size() != is_leaf.size()) throw PoutineError(PB_ERR_INVALID, "set_tree: parent / is_leaf sizes differ");
        check(pb_group_set_tree(ctx_, (int32_t)parent.size(), parent.data(), is_leaf.data()));
        n_nodes_ = (int64_t)parent.size();
    }
    void set_phenos(const std::vector<int32_t>& sample_node, const std::vector<uint8_t>& label) {
        if (sample_node.size() != label.size()) throw PoutineError(PB_ERR_INVALID, "set_phenos: sample_node / label sizes differ");
        check(pb_group_set_labels(ctx_, (int32_t)label.size(), sample_node.data(), label.data()));
        n_samples_ = (int64_t)label.size();
    }
    void set_sites(int64_t n_sites) {
        check(pb_group_set_sites(ctx_, n_sites));
        n_sites_ = n_sites;
        // the per-site records of tiles whose sweep and tail run underneath the upload come back underneath it as well
        if (sites_.size() < (size_t)n_sites_ || !sites_.data()) sites_.reset((size_t)n_sites_);
        check(pb_group_expect_site_output(ctx_, sites_.data()));
    }
    void put_planes(int64_t word_begin, int64_t n_words, int64_t pitch_words, const uint64_t* B0, const uint64_t* B1, const uint64_t* V) {
        check(pb_group_put_planes(ctx_, word_begin, n_words, pitch_words, B0, B1, V));
    }
    void put_planes_biallelic(int64_t word_begin, int64_t n_words, int64_t pitch_words, const uint64_t* X, const uint64_t* V,
                              const uint64_t* colmask) {
        check(pb_group_put_planes_biallelic(ctx_, word_begin, n_words, pitch_words, X, V, colmask));
    }
    // seg_sites as chars (node-major, row v at seq + v*pitch): packed on the host, compact upload when the alignment qualifies
    void set_seg_sites(const char* seq, int64_t n_nodes, int64_t n_sites, int64_t pitch) {
        const int64_t W = (n_sites + 63) / 64;
        std::vector<uint64_t> B0((size_t)(n_nodes * W)), B1(B0.size()), V(B0.size()), X(B0.size()), cm((size_t)(4 * W));
        int rc = pb_pack_chars(seq, n_nodes, n_sites, pitch, B0.data(), B1.data(), V.data(), W);
        if (rc != 0) throw PoutineError(rc, "pb_pack_chars: invalid arguments");
        set_sites(n_sites);
        int32_t all_valid = 0;
        rc = pb_compact_planes(B0.data(), B1.data(), V.data(), n_nodes, n_sites, W, X.data(), W, cm.data(), &all_valid, 0);
        if (rc < 0) throw PoutineError(rc, "pb_compact_planes: invalid arguments");
        if (rc == 1)
            put_planes_biallelic(0, W, W, X.data(), all_valid ? nullptr : V.data(), cm.data());
        else
            put_planes(0, W, W, B0.data(), B1.data(), V.data());
    }

    // ---- the two reference entry points -----------------------------------------------------------------------------
    // -> one record per site of this context, in site order (site_class == PB_SITE_BI marks the reference's
    // ArrayList<Homoplasy_Events> members); counts = the class counters the reference logs (HC.java:1477-1480).
    // The returned pointer is a page-locked buffer owned by the counter, overwritten by the next call.
    const pb_site_out* count_all_homoplasy_events(pb_class_counts* counts = nullptr) {
        if (sites_.size() < (size_t)n_sites_ || !sites_.data()) {
            sites_.reset((size_t)n_sites_);
            check(pb_group_expect_site_output(ctx_, sites_.data()));
        }
        check(pb_group_run_events(ctx_, sites_.data(), counts));
        return sites_.data();
    }
    AssocResult assoc(const uint32_t* replay_perms = nullptr, bool want_maxT = true) {
        pb_timings tm{};
        check(pb_group_get_timings(ctx_, &tm));
        const int64_t cap = tm.n_informative > 0 ? tm.n_informative : 1;
        if (stats_.size() < (size_t)cap || !stats_.data()) stats_.reset((size_t)cap);
        if (want_maxT && (maxT_.size() < (size_t)m_ || !maxT_.data())) maxT_.reset((size_t)m_);
        int64_t n_inf = 0;
        check(pb_group_run_assoc(ctx_, replay_perms, stats_.data(), cap, &n_inf, want_maxT ? maxT_.data() : nullptr));
        AssocResult r;
        r.stats.assign(stats_.data(), stats_.data() + n_inf);
        if (want_maxT) r.maxT_sorted.assign(maxT_.data(), maxT_.data() + m_);
        n_informative_ = n_inf;
        return r;
    }
    // exact point-wise null and two-sided Fisher (not reference outputs; parity unpinned — see the header)
    void extras(std::vector<double>& exact_pointwise, std::vector<double>& fisher_p) {
        exact_pointwise.assign((size_t)(2 * n_informative_), 0.0);
        fisher_p.assign((size_t)n_informative_, 0.0);
        check(pb_group_run_extras(ctx_, exact_pointwise.data(), fisher_p.data()));
    }

    pb_timings timings() {
        pb_timings tm{};
        check(pb_group_get_timings(ctx_, &tm));
        return tm;
    }
    int n_gpus() const { return (int)pb_group_size(ctx_); }
    int64_t n_sites() const { return n_sites_; }
    int64_t n_nodes() const { return n_nodes_; }
    int64_t n_informative() const { return n_informative_; }
    pb_group* handle() { return ctx_; }

  private:
    void check(int rc) {
        if (rc == 0) return;
        std::string msg = pb_group_last_error(ctx_);
        if (rc == PB_ERR_NO_INFORMATIVE) throw NoInformativeSites(rc, msg);
        throw PoutineError(rc, msg);
    }
    pb_group* ctx_ = nullptr;
    int64_t m_ = 0, n_sites_ = 0, n_nodes_ = 0, n_samples_ = 0, n_informative_ = 0;
    PinnedArray<pb_site_out> sites_;
    PinnedArray<pb_stat_out> stats_;
    PinnedArray<double> maxT_;
};

}  // namespace poutine
