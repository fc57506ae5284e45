// C ABI of libpoutine_b200.so — context, uploads, orchestration (see include/poutine_b200.h).
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>

#include "pb_internal.h"

static thread_local std::string g_create_err;

#define CHECK_CTX(ctx) do { if (!(ctx)) return PB_ERR_INVALID; } while (0)
#define FAIL(ctx, code, msg) do { (ctx)->err = (msg); return (code); } while (0)

extern "C" int pb_version(void) { return 100; }

extern "C" const char* pb_last_error(pb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static void free_all(pb_ctx* c) {
    void* ptrs[] = {c->d_ops, c->d_chunk_op, c->d_ops_x, c->d_chunk_op_x, c->d_node_sample, c->d_leaves_below, c->d_label, c->d_planes,
                    c->d_census, c->d_census_total, c->d_raw, c->d_raw_cursor, c->d_cnt, c->d_sites, c->d_site_flags, c->d_site_scan, c->d_class_counts,
                    c->d_inf_site, c->d_alleles, c->d_csr, c->d_csr_cursor, c->d_ptable, c->d_thr, c->d_pobs, c->d_r, c->d_maxT_own,
                    c->d_maxT_sorted, c->d_Xcase, c->d_Xmiss, c->d_perms, c->d_stats, c->d_fb_reps, c->d_fb_count, c->d_work, c->d_colpop, c->d_colmask, c->d_extras, c->d_X, c->d_xmask, c->d_xvalid, c->d_scan_state, c->d_tabkeys, c->d_binhist, c->d_bincum, c->d_pooled_own, c->d_rec_range};
    for (void* p : ptrs) if (p) cudaFree(p);
}

extern "C" int pb_create(const pb_config* cfg, pb_ctx** out) {
    if (!cfg || !out) { g_create_err = "pb_create: null argument"; return PB_ERR_INVALID; }
    *out = nullptr;
    if (cfg->n_replicates < 1 || cfg->min_hcount < 0 || (cfg->mode != PB_MODE_PHILOX && cfg->mode != PB_MODE_REPLAY) ||
        (cfg->site_offset & 63) != 0) {
        g_create_err = "pb_create: invalid config (n_replicates >= 1, min_hcount >= 0, mode, site_offset % 64 == 0)";
        return PB_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("pb_create: no CUDA device (") + cudaGetErrorString(e) + ") — there is no CPU fallback";
        return PB_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "pb_create: device ordinal out of range"; return PB_ERR_INVALID; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) { g_create_err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e); return PB_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_err = "pb_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                       ", this library carries sm_100a code only — there is no fallback";
        return PB_ERR_CUDA;
    }
    pb_ctx* ctx = new pb_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    ctx->sm_count = prop.multiProcessorCount;
    auto fail = [&](const char* what, cudaError_t ce) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(ce);
        pb_destroy(ctx);      // streams and events created so far go too (every handle starts out null)
        return PB_ERR_CUDA;
    };
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail("cudaSetDevice", e);
    {
        // Priorities.  The label generator is the long pole between K1 and the replicate counts: when it is ONE wave of
        // blocks (m <= ~190k replicates) it runs on the high-priority stream2, so that its blocks become resident first and
        // the short tail kernels of the main stream fill in around them.  When it is many waves (C4: m = 1e7) high priority
        // would starve the tail behind ~11 ms of generator blocks (round 1: ms_events_total 11 ms at C4): it then runs on the
        // low-priority stream2_lo and the tail's few blocks are scheduled as generator blocks retire.
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const int prio_mid = prio_hi < prio_lo - 1 ? prio_lo - 1 : prio_lo;      // numerically lower = more urgent
        int p_main = prio_mid, p_gen = prio_hi;
        if (getenv("PB_GEN_LOW_PRIORITY")) std::swap(p_main, p_gen);   // experiment knob
        if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, p_main)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, p_gen)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithPriority(&ctx->stream2_lo, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithPriority(&ctx->stream3, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithFlags(&ctx->upl, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithFlags(&ctx->dl, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    }
    if ((e = cudaEventCreateWithFlags(&ctx->ev_tail_tile, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    ctx->eager_tail_ok = getenv("PB_NO_EAGER_TAIL") == nullptr;
    if ((e = cudaEventCreateWithFlags(&ctx->ev_upl, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_gen_t0)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_gen_t1)) != cudaSuccess) return fail("cudaEventCreate", e);
    ctx->eager_ok = getenv("PB_NO_EAGER_SWEEP") == nullptr;
    if ((e = cudaEventCreateWithFlags(&ctx->ev_side_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_side_join, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_zero_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    for (auto& ev : ctx->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_k1_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_gen_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_gen_started, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    for (auto& ev : ctx->ev_gen_b) if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    for (auto& ev : ctx->ev_cnt_b) if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_raw_cursor, 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_class_counts, 16 * 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_fb_count, 4)) != cudaSuccess) return fail("cudaMalloc", e);
    ctx->trace_on = getenv("PB_TRACE") != nullptr;
    *out = ctx;
    return PB_OK;
}

extern "C" void pb_destroy(pb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    pbn_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pbx_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    if (ctx->stream3) cudaStreamSynchronize(ctx->stream3);
    if (ctx->stream2_lo) cudaStreamSynchronize(ctx->stream2_lo);
    if (ctx->upl) cudaStreamSynchronize(ctx->upl);
    if (ctx->dl) cudaStreamSynchronize(ctx->dl);
    free_all(ctx);
    if (ctx->dl) cudaStreamDestroy(ctx->dl);
    if (ctx->ev_tail_tile) cudaEventDestroy(ctx->ev_tail_tile);
    if (ctx->stream2_lo) cudaStreamDestroy(ctx->stream2_lo);
    if (ctx->ev_upl) cudaEventDestroy(ctx->ev_upl);
    if (ctx->ev_gen_t0) cudaEventDestroy(ctx->ev_gen_t0);
    if (ctx->ev_gen_t1) cudaEventDestroy(ctx->ev_gen_t1);
    if (ctx->upl) cudaStreamDestroy(ctx->upl);
    for (auto& t : ctx->trace) cudaEventDestroy(t.second);
    if (ctx->ev_side_fork) cudaEventDestroy(ctx->ev_side_fork);
    if (ctx->ev_side_join) cudaEventDestroy(ctx->ev_side_join);
    if (ctx->ev_zero_done) cudaEventDestroy(ctx->ev_zero_done);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    if (ctx->ev_k1_done) cudaEventDestroy(ctx->ev_k1_done);
    if (ctx->ev_gen_done) cudaEventDestroy(ctx->ev_gen_done);
    if (ctx->ev_gen_started) cudaEventDestroy(ctx->ev_gen_started);
    for (auto& ev : ctx->ev_gen_b) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_cnt_b) if (ev) cudaEventDestroy(ev);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

template <typename T>
static int upload(pb_ctx* ctx, T** dptr, const std::vector<T>& h) {
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    PB_CUDA(cudaMalloc(dptr, bytes));
    if (!h.empty()) PB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PB_OK;
}

// Logical edge list of K1: every non-root node once, as (child, parent), in a DFS preorder that lists a
// node's leaf children first and its internal children by ascending subtree size.  A child row is
// visited right after its parent row, so the kernel keeps the parent in registers; coming back from a
// (small) first subtree the parent row is re-read — from L1/L2, because the subtree entered first is the
// smaller one.  Iterative: caterpillar trees are as deep as they are large.
static int build_schedule(pb_ctx* ctx) {
    const int32_t N = ctx->N;
    std::vector<int32_t> nchild(N, 0), first(N + 1, 0), kids(N > 1 ? N - 1 : 0);
    for (int32_t v = 0; v < N; v++) if (ctx->parent[v] >= 0) nchild[ctx->parent[v]]++;
    for (int32_t v = 0; v < N; v++) first[v + 1] = first[v] + nchild[v];
    {
        std::vector<int32_t> fill(first.begin(), first.end() - 1);
        for (int32_t v = 0; v < N; v++) if (ctx->parent[v] >= 0) kids[fill[ctx->parent[v]]++] = v;
    }
    std::vector<int32_t> order; order.reserve(N);
    order.push_back(ctx->root);
    for (size_t i = 0; i < order.size(); i++) {
        const int32_t u = order[i];
        for (int32_t j = first[u]; j < first[u + 1]; j++) order.push_back(kids[j]);
    }
    if ((int32_t)order.size() != N) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: parent[] does not describe one tree (cycle or several roots)");
    std::vector<int32_t> sz(N, 1);
    for (int32_t i = N - 1; i > 0; i--) sz[ctx->parent[order[i]]] += sz[order[i]];
    {   // leaves in each subtree: the weights of the one-plane mode's census sum (xrec_t, pb_internal.h)
        std::vector<int32_t> lb(N, 0);
        for (int32_t v = 0; v < N; v++) lb[v] = ctx->is_leaf[v] ? 1 : 0;
        for (int32_t i = N - 1; i > 0; i--) lb[ctx->parent[order[i]]] += lb[order[i]];
        int rc = upload(ctx, &ctx->d_leaves_below, lb);
        if (rc) return rc;
    }
    for (int32_t u = 0; u < N; u++)
        std::sort(kids.begin() + first[u], kids.begin() + first[u + 1], [&](int32_t a, int32_t b) {
            if (sz[a] != sz[b]) return sz[a] < sz[b];   // leaves (size 1) first, then ascending subtree size
            return a < b;
        });
    ctx->lop_child.clear(); ctx->lop_parent.clear();
    ctx->lop_child.reserve(N); ctx->lop_parent.reserve(N);
    std::vector<std::pair<int32_t, int32_t>> stack;
    stack.push_back({ctx->root, first[ctx->root]});
    while (!stack.empty()) {
        auto& top = stack.back();
        const int32_t u = top.first;
        if (top.second == first[u + 1]) { stack.pop_back(); continue; }
        const int32_t c = kids[top.second++];
        ctx->lop_child.push_back(c);
        ctx->lop_parent.push_back(u);
        if (!ctx->is_leaf[c]) stack.push_back({c, first[c]});
    }
    return PB_OK;
}

// Cut the logical edge list into n_chunks equal pieces (grid.y) and lower each piece to the kernel's op
// stream: SETCUR reloads where the tracked parent row is not the edge's parent, census segments closed
// by FLUSH ops (leaf count padded to a multiple of 8 with null leaves: the carry-save tree works in
// groups of 8), chunks padded to a multiple of 4 ops.
// xprog: the lowering k1x_events_kernel runs (one-plane mode) — no census, so no LEAF padding and no flushes; every edge
// is tested, those hanging off the root as PB_OPX_NOMRCA (their records feed the census sum only); 28-bit node field.
static int build_program(pb_ctx* ctx, bool xprog, int64_t cols_per_block, int64_t blocks_per_sm, uint32_t** d_ops, int32_t** d_chunk_op,
                         int32_t* n_chunks_out, int32_t* n_flush_out) {
    const int64_t n_lops = (int64_t)ctx->lop_child.size();
    const int64_t blocks_x = (ctx->W + cols_per_block - 1) / cols_per_block;
    // grid.y: ~4 waves of 4 resident blocks per SM, picking the chunk count whose last wave is fullest
    int64_t slots = (int64_t)ctx->sm_count * blocks_per_sm;
    if (const char* e = getenv("PB_K1_TARGET_BLOCKS")) slots = atoll(e);
    int64_t n_chunks = 1;
    {
        const int64_t lo = std::max<int64_t>(1, (3 * slots + blocks_x - 1) / blocks_x), hi = std::max<int64_t>(lo, 5 * slots / blocks_x);
        double best = -1;
        for (int64_t n = lo; n <= hi; n++) {
            const int64_t blocks = blocks_x * n, waves = (blocks + slots - 1) / slots;
            const double eff = (double)blocks / (double)(waves * slots);
            if (eff > best + 1e-9) { best = eff; n_chunks = n; }
        }
    }
    if (const char* e = getenv("PB_K1_CHUNKS")) n_chunks = atoll(e);
    n_chunks = std::min<int64_t>(n_chunks, std::max<int64_t>(1, n_lops / 32));
    n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, 65535));
    *n_chunks_out = (int32_t)n_chunks;

    std::vector<uint32_t> recs;   // 4 words per record
    recs.reserve((size_t)n_lops * 3 + 64 * n_chunks);
    std::vector<int32_t> chunk_op(n_chunks + 1, 0);
    int32_t n_flush = 0;
    for (int64_t c = 0; c < n_chunks; c++) {
        const int64_t l0 = n_lops * c / n_chunks, l1 = n_lops * (c + 1) / n_chunks;
        int32_t cur = -1;       // node whose row the kernel holds as "current parent"
        int32_t seg_leaves = 0;
        const uint32_t null_op = xprog ? PB_OPX_NULL : PB_OP_NULL;
        uint32_t slot[2] = {null_op, null_op};
        int filled = 0;
        auto close = [&](uint32_t flush_id) {
            recs.push_back(slot[0]); recs.push_back(slot[1]); recs.push_back(flush_id); recs.push_back(0u);
            slot[0] = slot[1] = null_op; filled = 0;
        };
        auto push = [&](uint32_t op) {
            slot[filled++] = op;
            if (filled == 2) close(PB_REC_NOFLUSH);
        };
        auto flush = [&]() {
            while (seg_leaves % 8) { push(PB_OP_LEAF | PB_OP_NULL); seg_leaves++; }
            close((uint32_t)n_flush++);   // closes the open record (possibly empty) with the flush attached
            seg_leaves = 0;
        };
        for (int64_t i = l0; i < l1; i++) {
            const int32_t child = ctx->lop_child[i], par = ctx->lop_parent[i];
            if (xprog) {
                // every edge tested; an edge hanging off the root never makes an MRCA (HC.java:1674) but is part of the census sum
                const uint32_t fl = PB_OPX_TEST | (par == ctx->root ? PB_OPX_NOMRCA : 0u);
                if (cur != par) { push(PB_OPX_SETCUR | (uint32_t)par); cur = par; }
                if (ctx->is_leaf[child]) push(PB_OPX_LEAF | fl | (uint32_t)child);
                else { push(PB_OPX_SETCUR | fl | (uint32_t)child); cur = child; }
                continue;
            }
            const bool test = par != ctx->root;     // edges hanging off the root are never tested (HC.java:1674)
            if (test && cur != par) { push(PB_OP_SETCUR | (uint32_t)par); cur = par; }
            if (ctx->is_leaf[child]) {
                if (seg_leaves == PB_K1_SEG_LEAVES) flush();
                push(PB_OP_LEAF | (test ? PB_OP_TEST : 0u) | (uint32_t)child);
                seg_leaves++;
            } else {
                push(PB_OP_SETCUR | (test ? PB_OP_TEST : 0u) | (uint32_t)child);
                cur = child;
            }
        }
        if (!xprog) flush();   // every chunk ends with a flush (possibly of an all-zero census)
        else if (filled) close(PB_REC_NOFLUSH);
        while ((recs.size() / 4) % PB_K1_RPS) close(PB_REC_NOFLUSH);   // whole ring stages (empty records cost nothing)
        chunk_op[c + 1] = (int32_t)(recs.size() / 4);
    }
    const std::vector<uint32_t>& ops = recs;
    *n_flush_out = n_flush;
    int rc;
    if ((rc = upload(ctx, d_ops, ops))) return rc;
    return upload(ctx, d_chunk_op, chunk_op);
}

// two lowerings of the same edge list: one per sweep kernel (their grids and op sets differ)
static int build_chunks(pb_ctx* ctx) {
    if (ctx->N == 0 || ctx->W == 0) return PB_OK;
    int rc;
    if ((rc = build_program(ctx, false, PB_K1_COLS, PB_K1_BLOCKS_PER_SM, &ctx->d_ops, &ctx->d_chunk_op, &ctx->n_chunks, &ctx->n_flush))) return rc;
    if ((rc = build_program(ctx, true, pbk_kx_cols(), pbk_kx_blocks_per_sm(), &ctx->d_ops_x, &ctx->d_chunk_op_x, &ctx->n_chunks_x, &ctx->n_flush_x))) return rc;
    // the census partials belong to the three-plane sweep only: allocated at its first run (ensure_census)
    if (ctx->d_census) { cudaFree(ctx->d_census); ctx->d_census = nullptr; }
    if (ctx->d_census_total) { cudaFree(ctx->d_census_total); ctx->d_census_total = nullptr; }
    return PB_OK;
}

static int ensure_census(pb_ctx* ctx) {
    if (ctx->d_census && ctx->d_census_total) return PB_OK;
    if (!ctx->d_census) PB_CUDA(cudaMalloc(&ctx->d_census, (size_t)4 * ctx->n_flush * PB_K1_KS * ctx->W * 8));
    if (!ctx->d_census_total) PB_CUDA(cudaMalloc(&ctx->d_census_total, (size_t)4 * PB_K1_KT * ctx->W * 8));
    return PB_OK;
}


static int build_node_sample(pb_ctx* ctx) {
    if (ctx->N == 0) return PB_OK;
    std::vector<int32_t> ns(ctx->N, -1);
    for (int32_t j = 0; j < ctx->P; j++) ns[ctx->sample_node[j]] = j;
    return upload(ctx, &ctx->d_node_sample, ns);
}

extern "C" int pb_set_tree(pb_ctx* ctx, int32_t n_nodes, const int32_t* parent, const uint8_t* is_leaf) {
    CHECK_CTX(ctx);
    if (n_nodes < 2 || !parent || !is_leaf) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: need >= 2 nodes, parent[] and is_leaf[]");
    if (n_nodes >= (int32_t)PB_OPX_NODE_MASK) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: too many nodes (limit 2^28-2)");
    PB_CUDA(cudaSetDevice(ctx->device));
    int32_t root = -1, L = 0;
    std::vector<int32_t> nchild(n_nodes, 0);
    for (int32_t v = 0; v < n_nodes; v++) {
        if (parent[v] < 0) { if (root >= 0) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: more than one root"); root = v; }
        else if (parent[v] >= n_nodes || parent[v] == v) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: parent index out of range");
        else nchild[parent[v]]++;
    }
    if (root < 0) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: no root (parent == -1)");
    for (int32_t v = 0; v < n_nodes; v++) {
        if ((nchild[v] == 0) != (is_leaf[v] != 0)) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: is_leaf[] disagrees with parent[]");
        L += is_leaf[v] != 0;
    }
    if (is_leaf[root]) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: the root is a leaf");
    if (L >= (1 << PB_K1_KT)) FAIL(ctx, PB_ERR_INVALID, "pb_set_tree: too many leaves (limit 2^24-1)");
    ctx->N = n_nodes; ctx->L = L; ctx->root = root;
    ctx->parent.assign(parent, parent + n_nodes);
    ctx->is_leaf.assign(is_leaf, is_leaf + n_nodes);
    ctx->events_done = ctx->assoc_done = false;
    if (ctx->stream) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream2) PB_CUDA(cudaStreamSynchronize(ctx->stream2));
    if (ctx->stream2_lo) PB_CUDA(cudaStreamSynchronize(ctx->stream2_lo));
    ctx->pregen_valid = false;
    ctx->swept_words = 0;
    ctx->P = 0; ctx->sample_node.clear(); ctx->label.clear();
    int rc;
    if ((rc = build_schedule(ctx))) return rc;
    if ((rc = build_node_sample(ctx))) return rc;
    if ((rc = build_chunks(ctx))) return rc;
    // planes depend on N: force re-allocation
    if (ctx->S) { const int64_t S = ctx->S; ctx->S = 0; return pb_set_sites(ctx, S); }
    return PB_OK;
}

extern "C" int pb_set_labels(pb_ctx* ctx, int32_t n_samples, const int32_t* sample_node, const uint8_t* label) {
    CHECK_CTX(ctx);
    if (ctx->N == 0) FAIL(ctx, PB_ERR_INVALID, "pb_set_labels: call pb_set_tree first");
    if (n_samples < 1 || !sample_node || !label) FAIL(ctx, PB_ERR_INVALID, "pb_set_labels: empty phenotype table");
    PB_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint8_t> seen(ctx->N, 0);
    int32_t nc = 0, nt = 0, nm = 0;
    for (int32_t j = 0; j < n_samples; j++) {
        const int32_t v = sample_node[j];
        // check_sample_names (HC.java:367-404): every pheno sample must be a leaf of the tree
        if (v < 0 || v >= ctx->N || !ctx->is_leaf[v]) FAIL(ctx, PB_ERR_INVALID, "pb_set_labels: sample_node is not a leaf of the tree");
        if (seen[v]) FAIL(ctx, PB_ERR_INVALID, "pb_set_labels: duplicate sample (HashMap keys are unique)");
        seen[v] = 1;
        if (label[j] == PB_LABEL_CASE) nc++; else if (label[j] == PB_LABEL_CONTROL) nt++; else if (label[j] == PB_LABEL_MISSING) nm++;
        else FAIL(ctx, PB_ERR_INVALID, "pb_set_labels: label code must be 0, 1 or 2");
    }
    if (ctx->stream) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream2) PB_CUDA(cudaStreamSynchronize(ctx->stream2));   // a pre-generated label matrix is stale now
    if (ctx->stream2_lo) PB_CUDA(cudaStreamSynchronize(ctx->stream2_lo));
    ctx->pregen_valid = false;
    ctx->swept_words = 0;
    ctx->P = n_samples; ctx->n_case = nc; ctx->n_ctrl = nt; ctx->n_miss = nm;
    ctx->sample_node.assign(sample_node, sample_node + n_samples);
    ctx->label.assign(label, label + n_samples);
    ctx->events_done = ctx->assoc_done = false;  // the CSR only keeps leaves that have a pheno row
    int rc;
    if ((rc = upload(ctx, &ctx->d_label, ctx->label))) return rc;
    return build_node_sample(ctx);
}

typedef CUresult (*pb_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tensor_map(pb_ctx* ctx, CUtensorMap* out, void* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride, const cuuint32_t* box) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) FAIL(ctx, PB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult cr = ((pb_encode_fn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, base, gdim, gstride, box, estride,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) FAIL(ctx, PB_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)cr) + ")");
    return PB_OK;
}

static void free_planes(pb_ctx* ctx) {
    for (void** p : {(void**)&ctx->d_planes, (void**)&ctx->d_X, (void**)&ctx->d_xmask, (void**)&ctx->d_xvalid})
        if (*p) { cudaFree(*p); *p = nullptr; }
    ctx->d_B0 = ctx->d_B1 = ctx->d_V = nullptr;
    ctx->plane_mode = PB_PLANES_NONE;
}

// the three full planes (zero-filled: a column nobody uploads has no valid genotype) and their 3-D TMA descriptor
static int alloc_full_planes(pb_ctx* ctx) {
    const size_t plane = (size_t)ctx->N * ctx->Wp * 8;
    cudaError_t e = cudaMalloc(&ctx->d_planes, 3 * plane);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc of 3 x ") + std::to_string(plane) + " B planes: " + cudaGetErrorString(e);
        ctx->d_planes = nullptr;
        return PB_ERR_OOM;
    }
    ctx->d_V = ctx->d_planes;
    ctx->d_B0 = ctx->d_planes + (size_t)ctx->N * ctx->Wp;
    ctx->d_B1 = ctx->d_planes + 2 * (size_t)ctx->N * ctx->Wp;
    PB_CUDA(cudaMemsetAsync(ctx->d_planes, 0, 3 * plane, ctx->upl));
    // dims (word column, node, plane); one box = the 3 KB a block needs of one tree row
    const cuuint64_t gdim[3] = {(cuuint64_t)ctx->W, (cuuint64_t)ctx->N, 3};
    const cuuint64_t gstride[2] = {(cuuint64_t)ctx->Wp * 8, (cuuint64_t)plane};
    const cuuint32_t box[3] = {PB_K1_COLS, 1, 3};
    return tensor_map(ctx, &ctx->tmap, ctx->d_planes, 3, gdim, gstride, box);
}

// Moves the context to the requested storage mode.  NONE -> X or FULL allocates; X -> FULL promotes (expands what was
// uploaded so far); FULL stays FULL (a compact tile is then expanded on arrival).
static int ensure_plane_mode(pb_ctx* ctx, int want) {
    if (ctx->plane_mode == want || ctx->plane_mode == PB_PLANES_FULL) return PB_OK;
    int rc;
    if (want == PB_PLANES_X) {
        const size_t plane = (size_t)ctx->N * ctx->Wp * 8;
        cudaError_t e = cudaMalloc(&ctx->d_X, plane);
        if (e != cudaSuccess) { ctx->d_X = nullptr; ctx->err = std::string("cudaMalloc of the X plane: ") + cudaGetErrorString(e); return PB_ERR_OOM; }
        PB_CUDA(cudaMalloc(&ctx->d_xmask, (size_t)ctx->W * 32));
        PB_CUDA(cudaMalloc(&ctx->d_xvalid, (size_t)ctx->W * 8));
        PB_CUDA(cudaMemsetAsync(ctx->d_X, 0, plane, ctx->upl));
        PB_CUDA(cudaMemsetAsync(ctx->d_xmask, 0, (size_t)ctx->W * 32, ctx->upl));
        PB_CUDA(cudaMemsetAsync(ctx->d_xvalid, 0, (size_t)ctx->W * 8, ctx->upl));
        const cuuint64_t gdim[2] = {(cuuint64_t)ctx->W, (cuuint64_t)ctx->N};
        const cuuint64_t gstride[1] = {(cuuint64_t)ctx->Wp * 8};
        const cuuint32_t box[2] = {(cuuint32_t)pbk_kx_cols(), 1};
        if ((rc = tensor_map(ctx, &ctx->tmap_x, ctx->d_X, 2, gdim, gstride, box))) return rc;
        ctx->plane_mode = PB_PLANES_X;
        return PB_OK;
    }
    if ((rc = alloc_full_planes(ctx))) return rc;
    if (ctx->plane_mode == PB_PLANES_X) {
        if ((rc = pbk_promote_planes(ctx))) return rc;
        PB_CUDA(cudaStreamSynchronize(ctx->upl));
        cudaFree(ctx->d_X); cudaFree(ctx->d_xmask); cudaFree(ctx->d_xvalid);
        ctx->d_X = ctx->d_xmask = ctx->d_xvalid = nullptr;
    }
    ctx->plane_mode = PB_PLANES_FULL;
    return PB_OK;
}

extern "C" int pb_set_sites(pb_ctx* ctx, int64_t n_sites) {
    CHECK_CTX(ctx);
    if (ctx->N == 0) FAIL(ctx, PB_ERR_INVALID, "pb_set_sites: call pb_set_tree first");
    // the tail packs (informative-site count << 38 | extant-event count) into one scan word and indexes sites with int32:
    // 2^26 sites per context is the supported limit (shard a longer alignment over several contexts, site_offset)
    if (n_sites < 1 || n_sites > PB_MAX_SITES_PER_CTX) FAIL(ctx, PB_ERR_INVALID, "pb_set_sites: n_sites out of range (1 .. 2^26 per context)");
    PB_CUDA(cudaSetDevice(ctx->device));
    ctx->events_done = ctx->assoc_done = false;
    if (ctx->stream) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->upl) PB_CUDA(cudaStreamSynchronize(ctx->upl));
    if (ctx->dl) PB_CUDA(cudaStreamSynchronize(ctx->dl));
    ctx->swept_words = 0; ctx->tailed_words = 0; ctx->sent_sites = 0;
    if (n_sites == ctx->S && ctx->plane_mode != PB_PLANES_NONE) return PB_OK;   // same shape: the uploaded planes stay
    ctx->out_sites = nullptr;             // a buffer registered for the old site count is not trusted with the new one
    free_planes(ctx);
    for (void** p : {(void**)&ctx->d_cnt, (void**)&ctx->d_sites, (void**)&ctx->d_site_flags, (void**)&ctx->d_site_scan, (void**)&ctx->d_raw})
        if (*p) { cudaFree(*p); *p = nullptr; }
    ctx->S = n_sites;
    ctx->W = (n_sites + 63) / 64;
    ctx->Wp = (ctx->W + 15) & ~15ll;  // rows padded to 128 B
    // The planes themselves are allocated by the first upload: one plane (X) while every tile is biallelic with all
    // genotypes valid, three otherwise.  Make sure the worst case fits now, so that an oversized alignment fails here.
    {
        size_t free_b = 0, total_b = 0;
        PB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t plane = (size_t)ctx->N * ctx->Wp * 8;
        if (plane > free_b) {
            ctx->err = "pb_set_sites: a plane of " + std::to_string(plane) + " B does not fit the device (" + std::to_string(free_b) + " B free)";
            ctx->S = 0;
            return PB_ERR_OOM;
        }
    }
    PB_CUDA(cudaMalloc(&ctx->d_cnt, (size_t)n_sites * PB_CNT_ALLOC_STRIDE * 4));
    PB_CUDA(cudaMalloc(&ctx->d_sites, (size_t)n_sites * sizeof(pb_site_out)));
    PB_CUDA(cudaMalloc(&ctx->d_site_flags, (size_t)n_sites));
    PB_CUDA(cudaMalloc(&ctx->d_site_scan, (size_t)n_sites * 8));
    ctx->raw_cap = std::max<int64_t>(4 * n_sites, 1 << 20);
    PB_CUDA(cudaMalloc(&ctx->d_raw, (size_t)ctx->raw_cap * sizeof(raw_rec_t)));
    return build_chunks(ctx);
}

// every upload runs on the context's upload stream; the compute stream picks the data up through ev_upl
static int upload_done(pb_ctx* ctx, bool async) {
    PB_CUDA(cudaEventRecord(ctx->ev_upl, ctx->upl));
    if (!async) PB_CUDA(cudaStreamSynchronize(ctx->upl));   // host buffers are caller-owned: the plain calls do not retain them past return
    return PB_OK;
}

static int put_planes_impl(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* B0, const uint64_t* B1,
                           const uint64_t* V, bool async) {
    CHECK_CTX(ctx);
    if (ctx->S == 0) FAIL(ctx, PB_ERR_INVALID, "pb_put_planes: call pb_set_sites first");
    if (!B0 || !B1 || !V || word_begin < 0 || n_words < 1 || word_begin + n_words > ctx->W || src_pitch_words < n_words)
        FAIL(ctx, PB_ERR_INVALID, "pb_put_planes: tile out of range");
    PB_CUDA(cudaSetDevice(ctx->device));
    ctx->events_done = ctx->assoc_done = false;
    ctx->swept_words = 0;                              // a three-plane tile: no sweep ahead of pb_run_events
    { int rc = ensure_plane_mode(ctx, PB_PLANES_FULL); if (rc) return rc; }
    const uint64_t* src[3] = {B0, B1, V};
    uint64_t* dst[3] = {ctx->d_B0, ctx->d_B1, ctx->d_V};
    if (word_begin == 0 && n_words == ctx->W && src_pitch_words == ctx->Wp) {
        // the caller's rows already have the device pitch (W rounded up to 16 words): one flat copy per plane — up to the last
        // row's last valid word, not beyond it (the source may be a window into a wider buffer: pb_group passes such)
        for (int i = 0; i < 3; i++)
            PB_CUDA(cudaMemcpyAsync(dst[i], src[i], ((size_t)(ctx->N - 1) * ctx->Wp + (size_t)n_words) * 8, cudaMemcpyHostToDevice, ctx->upl));
    } else {
        for (int i = 0; i < 3; i++)
            PB_CUDA(cudaMemcpy2DAsync(dst[i] + word_begin, (size_t)ctx->Wp * 8, src[i], (size_t)src_pitch_words * 8, (size_t)n_words * 8,
                                      (size_t)ctx->N, cudaMemcpyHostToDevice, ctx->upl));
    }
    if (word_begin + n_words == ctx->W) { int rc = pbk_mask_last_word(ctx); if (rc) return rc; }
    return upload_done(ctx, async);
}

// One-plane mode, tiles arriving in column order and cut at the sweep's block width: the event sweep of a tile is
// launched as soon as the tile has landed, on the compute stream, so that it runs underneath the NEXT tile's copy (the
// copy engine and the SMs are otherwise idle in turn), and the label generator — which needs only the phenotypes and the
// seed — starts with the first tile.  pb_run_events then sweeps what is left (usually nothing) and runs the tail.
// Anything out of order (a re-upload, an unaligned tile, a three-plane tile) simply drops the head start: the records
// written so far are discarded by the full sweep's cursor reset.
static int eager_sweep(pb_ctx* ctx, int64_t word_begin, int64_t n_words) {
    const int64_t cols = pbk_kx_cols();
    // a tile that starts at word 0 always opens a new pass (whatever was swept before is discarded by the cursor reset)
    const bool in_order = ctx->eager_ok && ctx->plane_mode == PB_PLANES_X && (word_begin == 0 || word_begin == ctx->swept_words) && word_begin % cols == 0 &&
                          ((word_begin + n_words) % cols == 0 || word_begin + n_words == ctx->W) && ctx->N > 0;
    if (!in_order) { ctx->swept_words = 0; ctx->tailed_words = 0; return PB_OK; }
    int rc;
    if (word_begin == 0) {
        ctx->tm = pb_timings{};
        if (ctx->dl) PB_CUDA(cudaStreamSynchronize(ctx->dl));     // (a dropped pass may still be copying records out)
        ctx->tailed_words = 0; ctx->tail_tiles = 0; ctx->sent_sites = 0;
        if ((rc = pbk_prepare_tail(ctx))) return rc;
        if (!ctx->pregen_valid && (rc = pbk_pregen_labels(ctx, true))) return rc;
    }
    PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_upl, 0));
    if ((rc = pbk_events_range(ctx, word_begin / cols, (n_words + cols - 1) / cols, word_begin == 0, false))) return rc;
    ctx->swept_words = word_begin + n_words;
    // ... and its tail.  Needs the phenotype rows (which leaves are sampled decides the tallies and the CSR): a host that sets its
    // labels after the upload gets the whole tail in pb_run_events, as before (pb_set_labels drops everything done ahead).
    if (ctx->eager_tail_ok && ctx->P > 0 && (rc = pbk_tail_tile(ctx, word_begin, n_words))) return rc;
    return PB_OK;
}

static int put_planes_biallelic_impl(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* X,
                                     const uint64_t* V, const uint64_t* colmask, bool async) {
    CHECK_CTX(ctx);
    if (ctx->S == 0) FAIL(ctx, PB_ERR_INVALID, "pb_put_planes_biallelic: call pb_set_sites first");
    if (!X || !colmask || word_begin < 0 || n_words < 1 || word_begin + n_words > ctx->W || src_pitch_words < n_words)
        FAIL(ctx, PB_ERR_INVALID, "pb_put_planes_biallelic: tile out of range");
    PB_CUDA(cudaSetDevice(ctx->device));
    ctx->events_done = ctx->assoc_done = false;
    int rc;
    const bool flat = word_begin == 0 && n_words == ctx->W && src_pitch_words == ctx->Wp;
    auto copy_rows = [&](uint64_t* dst, const uint64_t* src) -> cudaError_t {
        if (flat) return cudaMemcpyAsync(dst, src, ((size_t)(ctx->N - 1) * ctx->Wp + (size_t)n_words) * 8, cudaMemcpyHostToDevice, ctx->upl);
        return cudaMemcpy2DAsync(dst + word_begin, (size_t)ctx->Wp * 8, src, (size_t)src_pitch_words * 8, (size_t)n_words * 8, (size_t)ctx->N,
                                 cudaMemcpyHostToDevice, ctx->upl);
    };
    // every genotype valid and no full-plane tile so far: the alignment stays one plane in HBM (k1x_events_kernel)
    const bool keep_x = !V && ctx->plane_mode != PB_PLANES_FULL && !(ctx->cfg.flags & PB_FLAG_FULL_PLANES);
    if (keep_x) {
        if ((rc = ensure_plane_mode(ctx, PB_PLANES_X))) return rc;
        PB_CUDA(cudaMemcpyAsync(ctx->d_xmask + 4 * word_begin, colmask, (size_t)n_words * 32, cudaMemcpyHostToDevice, ctx->upl));
        PB_CUDA(cudaMemsetAsync(ctx->d_xvalid + word_begin, 0xFF, (size_t)n_words * 8, ctx->upl));
        PB_CUDA(copy_rows(ctx->d_X, X));
        if (word_begin + n_words == ctx->W && (rc = pbk_mask_last_word_x(ctx))) return rc;
        PB_CUDA(cudaEventRecord(ctx->ev_upl, ctx->upl));
        if ((rc = eager_sweep(ctx, word_begin, n_words))) return rc;
        return upload_done(ctx, async);
    }
    ctx->swept_words = 0;
    if ((rc = ensure_plane_mode(ctx, PB_PLANES_FULL))) return rc;
    if (n_words > ctx->colmask_cap) {
        if (ctx->d_colmask) { cudaFree(ctx->d_colmask); ctx->d_colmask = nullptr; ctx->colmask_cap = 0; }
        PB_CUDA(cudaMalloc(&ctx->d_colmask, (size_t)n_words * 32));
        ctx->colmask_cap = n_words;
    }
    PB_CUDA(cudaMemcpyAsync(ctx->d_colmask, colmask, (size_t)n_words * 32, cudaMemcpyHostToDevice, ctx->upl));
    PB_CUDA(copy_rows(ctx->d_B0, X));               // X lands in the B0 rows and is expanded in place
    if (V) PB_CUDA(copy_rows(ctx->d_V, V));
    if ((rc = pbk_expand_biallelic(ctx, word_begin, n_words, ctx->d_colmask, V == nullptr))) return rc;
    if (V && word_begin + n_words == ctx->W) { rc = pbk_mask_last_word(ctx); if (rc) return rc; }
    return upload_done(ctx, async);
}

extern "C" int pb_put_planes(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* B0,
                             const uint64_t* B1, const uint64_t* V) {
    return put_planes_impl(ctx, word_begin, n_words, src_pitch_words, B0, B1, V, false);
}
extern "C" int pb_put_planes_async(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* B0,
                                   const uint64_t* B1, const uint64_t* V) {
    return put_planes_impl(ctx, word_begin, n_words, src_pitch_words, B0, B1, V, true);
}
extern "C" int pb_put_planes_biallelic(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* X,
                                       const uint64_t* V, const uint64_t* colmask) {
    return put_planes_biallelic_impl(ctx, word_begin, n_words, src_pitch_words, X, V, colmask, false);
}
extern "C" int pb_put_planes_biallelic_async(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words, const uint64_t* X,
                                             const uint64_t* V, const uint64_t* colmask) {
    return put_planes_biallelic_impl(ctx, word_begin, n_words, src_pitch_words, X, V, colmask, true);
}
extern "C" int pb_upload_sync(pb_ctx* ctx) {
    CHECK_CTX(ctx);
    PB_CUDA(cudaSetDevice(ctx->device));
    PB_CUDA(cudaStreamSynchronize(ctx->upl));
    return PB_OK;
}

extern "C" int pb_expect_site_output(pb_ctx* ctx, pb_site_out* per_site) {
    CHECK_CTX(ctx);
    PB_CUDA(cudaSetDevice(ctx->device));
    if (ctx->dl) PB_CUDA(cudaStreamSynchronize(ctx->dl));     // nothing may still be writing to the previous buffer
    ctx->out_sites = per_site;
    ctx->sent_sites = 0;
    return PB_OK;
}

// defer_sites_copy: the per-site records go to the host on the upload stream's sibling copy queue WITHOUT waiting for them
// (pb_run_events_assoc lets that copy run underneath the association kernels and waits at its own end)
static int run_events_impl(pb_ctx* ctx, pb_site_out* per_site, pb_class_counts* counts, bool defer_sites_copy) {
    CHECK_CTX(ctx);
    if (ctx->S == 0 || ctx->N == 0) FAIL(ctx, PB_ERR_INVALID, "pb_run_events: tree and planes must be set");
    PB_CUDA(cudaSetDevice(ctx->device));
    if (ctx->plane_mode == PB_PLANES_NONE) {   // nothing uploaded: no valid genotype anywhere
        int rc0 = ensure_plane_mode(ctx, PB_PLANES_FULL);
        if (rc0) return rc0;
        PB_CUDA(cudaEventRecord(ctx->ev_upl, ctx->upl));
    }
    PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_upl, 0));     // uploads (async ones included) are ordered before the sweep
    // tiles swept ahead of this call (eager_sweep): their timings / launch counts belong to this pass
    const bool ahead = ctx->swept_words > 0 && ctx->plane_mode == PB_PLANES_X;
    const int64_t swept = ahead ? ctx->swept_words : 0;
    int64_t tailed = ahead ? ctx->tailed_words : 0;        // tiles whose tail ran underneath the upload too
    int64_t sent = (tailed > 0 && per_site && per_site == ctx->out_sites) ? ctx->sent_sites : 0;   // ... and whose records are on their way
    ctx->swept_words = 0; ctx->tailed_words = 0; ctx->sent_sites = 0;
    if (!ahead) ctx->tm = pb_timings{};
    ctx->tm.eager_words = swept;
    ctx->tm.tailed_words = tailed; ctx->tm.early_sites = sent;
    int rc;
    if (ctx->plane_mode == PB_PLANES_FULL && (rc = ensure_census(ctx))) return rc;
    PB_CUDA(cudaEventRecord(ctx->ev[12], ctx->stream));
    for (int attempt = 0; attempt < 3; attempt++) {
        if (attempt == 0 && ahead) {
            const int64_t cols = pbk_kx_cols(), b0 = (swept + cols - 1) / cols, b1 = (ctx->W + cols - 1) / cols;
            if ((rc = pbk_events_range(ctx, b0, b1 - b0, false, true))) return rc;       // what the uploads did not cover (usually nothing)
        } else {
            tailed = 0; sent = 0;      // a full pass: nothing done ahead counts
            ctx->tm.tailed_words = 0; ctx->tm.early_sites = 0;
            if ((rc = pbk_prepare_tail(ctx))) return rc;
            const bool gen_first = getenv("PB_PREGEN_FIRST") != nullptr;   // experiment: label generator resident before the sweep
            if (attempt == 0 && gen_first && !ctx->pregen_valid && (rc = pbk_pregen_labels(ctx, false))) return rc;
            if ((rc = pbk_events(ctx))) return rc;
        }
        // (Tried: starting the generator behind the tail's record-count kernel instead of behind the sweep — that kernel then runs
        // alone (276 -> 58 us at C2) but finalize+compact inherits the squeeze (115 -> 290 us): whichever tail kernel shares the SMs
        // with the generator pays for it (issue slots: capping the generator's registers so that more tail warps fit changed nothing).
        // No gain; not kept.)
        if (attempt == 0 && !ctx->pregen_valid && (rc = pbk_pregen_labels(ctx, false))) return rc;
        unsigned long long cur = 0;
        if ((rc = pbk_finalize_events(ctx, counts, &cur, tailed))) return rc;   // the (rest of the) tail, one host round trip at its end
        if ((int64_t)cur <= ctx->raw_cap && ctx->E <= ctx->csr_cap) { ctx->n_rec = (int64_t)cur; break; }
        // record list (or the CSR of extant events: a record can hold up to 64) larger than its buffer — dense / adversarial
        // input: grow to the exact need and redo the pass
        if ((int64_t)cur > ctx->raw_cap) {
            cudaFree(ctx->d_raw); ctx->d_raw = nullptr;
            ctx->raw_cap = (int64_t)cur + (int64_t)cur / 16 + 1024;
            cudaError_t e = cudaMalloc(&ctx->d_raw, (size_t)ctx->raw_cap * sizeof(raw_rec_t));
            if (e != cudaSuccess) { ctx->err = "pb_run_events: cannot allocate the event list"; ctx->raw_cap = 0; return PB_ERR_OOM; }
        }
        if (ctx->E > ctx->csr_cap) ctx->csr_want = ctx->E + ctx->E / 16 + 1024;
        if (attempt == 2) FAIL(ctx, PB_ERR_CUDA, "pb_run_events: event list kept overflowing");
    }
    PB_CUDA(cudaEventRecord(ctx->ev[13], ctx->stream));
    if (ctx->dl) PB_CUDA(cudaStreamSynchronize(ctx->dl));        // records sent ahead have landed (or, after a repeated pass, are out of the way)
    if (sent > ctx->S) sent = ctx->S;
    if (per_site && defer_sites_copy) {       // the tail has been synchronised already (its one readback): d_sites is final
        PB_CUDA(cudaMemcpyAsync(per_site + sent, ctx->d_sites + sent, (size_t)(ctx->S - sent) * sizeof(pb_site_out), cudaMemcpyDeviceToHost, ctx->upl));
    } else if (per_site && sent < ctx->S) {
        PB_CUDA(cudaMemcpyAsync(per_site + sent, ctx->d_sites + sent, (size_t)(ctx->S - sent) * sizeof(pb_site_out), cudaMemcpyDeviceToHost, ctx->stream));
    }
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    float t;
    cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]); ctx->tm.ms_events_kernel = t;
    cudaEventElapsedTime(&t, ctx->ev[12], ctx->ev[13]); ctx->tm.ms_events_total = t;
    ctx->tm.plane_mode = ctx->plane_mode;
    // what the sweep kernel itself moves: the plane bits once + one record per (row, word) that holds an event
    ctx->tm.k1_algorithmic_bytes = (ctx->plane_mode == PB_PLANES_X ? 1ll : 3ll) * ctx->N * ctx->W * 8 +
                                   (ctx->plane_mode == PB_PLANES_X ? 16ll : 32ll) * ctx->n_rec;
    ctx->tm.n_informative = ctx->S_inf; ctx->tm.n_alleles_in_use = ctx->A_inuse; ctx->tm.n_extant_events = ctx->E;
    ctx->events_done = true;
    ctx->assoc_done = false;
    return PB_OK;
}

extern "C" int pb_run_events(pb_ctx* ctx, pb_site_out* per_site, pb_class_counts* counts) {
    return run_events_impl(ctx, per_site, counts, false);
}

extern "C" int pb_run_events_assoc(pb_ctx* ctx, pb_site_out* per_site, pb_class_counts* counts, const uint32_t* replay_perms,
                                   pb_stat_out* per_inf_site, int64_t cap, int64_t* n_inf, double* maxT_sorted) {
    const bool trace = ctx && ctx->trace_on;
    timespec t0, t1, t2, t3;
    if (trace) clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = run_events_impl(ctx, per_site, counts, true);
    if (rc) return rc;
    if (trace) clock_gettime(CLOCK_MONOTONIC, &t1);
    rc = pb_run_assoc(ctx, replay_perms, per_inf_site, cap, n_inf, maxT_sorted);
    if (trace) clock_gettime(CLOCK_MONOTONIC, &t2);
    const cudaError_t e = cudaStreamSynchronize(ctx->upl);      // per_site has arrived (also when the association failed)
    if (trace) {
        clock_gettime(CLOCK_MONOTONIC, &t3);
        auto us = [](const timespec& a, const timespec& b) { return (b.tv_sec - a.tv_sec) * 1e6 + (b.tv_nsec - a.tv_nsec) * 1e-3; };
        fprintf(stderr, "[poutine_b200 trace] events_assoc host: events %.0f us, assoc %.0f us, wait for per_site %.0f us\n", us(t0, t1), us(t1, t2), us(t2, t3));
    }
    if (rc == 0 && e != cudaSuccess) { ctx->err = std::string("pb_run_events_assoc: ") + cudaGetErrorString(e); return PB_ERR_CUDA; }
    return rc;
}

extern "C" int pb_run_assoc(pb_ctx* ctx, const uint32_t* replay_perms, pb_stat_out* per_inf_site, int64_t cap, int64_t* n_inf,
                            double* maxT_sorted) {
    CHECK_CTX(ctx);
    if (!ctx->events_done) FAIL(ctx, PB_ERR_INVALID, "pb_run_assoc: call pb_run_events first");
    if (ctx->P == 0) FAIL(ctx, PB_ERR_INVALID, "pb_run_assoc: call pb_set_labels first");
    if (ctx->cfg.mode == PB_MODE_REPLAY && !replay_perms) FAIL(ctx, PB_ERR_INVALID, "pb_run_assoc: replay mode needs replay_perms");
    if (ctx->n_case + ctx->n_ctrl == 0) FAIL(ctx, PB_ERR_INVALID, "pb_run_assoc: phenotype file has neither cases nor controls");
    PB_CUDA(cudaSetDevice(ctx->device));
    if (n_inf) *n_inf = ctx->S_inf;
    // subset_by_min_hcount: "min_hcount ... has filtered out all segregating sites" (HC.java:3866-3869).
    // In a sharded run an empty shard is legal: it still has to join the all-reduce.
    if (ctx->S_inf == 0 && ctx->world == 1) FAIL(ctx, PB_ERR_NO_INFORMATIVE, "min_hcount is set too high and has filtered out all segregating sites");
    if (per_inf_site && cap < ctx->S_inf) FAIL(ctx, PB_ERR_INVALID, "pb_run_assoc: per_inf_site too small");
    int rc;
    if ((rc = pbk_assoc_prepare(ctx))) return rc;       // allocations first: none while the pass is in flight
    if ((rc = pbk_assoc(ctx, replay_perms))) return rc;
    if (per_inf_site && ctx->S_inf)
        PB_CUDA(cudaMemcpyAsync(per_inf_site, ctx->d_stats, (size_t)ctx->S_inf * sizeof(pb_stat_out), cudaMemcpyDeviceToHost, ctx->stream));
    if (maxT_sorted)
        PB_CUDA(cudaMemcpyAsync(maxT_sorted, ctx->d_maxT_sorted, (size_t)ctx->cfg.n_replicates * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->assoc_done = true;
    return PB_OK;
}

// internal (pb_group): the allocations of the coming pb_run_assoc, so that a group can finish them on every shard before
// any shard starts its pass
int pb_assoc_prepare_internal(pb_ctx* ctx) {
    if (!ctx || !ctx->events_done || ctx->P == 0) return PB_OK;     // pb_run_assoc reports the call-order error itself
    if (cudaSetDevice(ctx->device) != cudaSuccess) return PB_ERR_CUDA;
    const int rc = pbk_assoc_prepare(ctx);
    if (rc == 0) { ctx->prep_sinf = ctx->S_inf; ctx->prep_nmax = ctx->n_max; ctx->prep_P = ctx->P; }
    return rc;
}

extern "C" int pb_run_extras(pb_ctx* ctx, double* exact_pointwise, double* fisher_p) {
    CHECK_CTX(ctx);
    if (!ctx->assoc_done) FAIL(ctx, PB_ERR_INVALID, "pb_run_extras: call pb_run_assoc first");
    PB_CUDA(cudaSetDevice(ctx->device));
    return pbk_extras(ctx, exact_pointwise, fisher_p);
}

// ---- introspection -------------------------------------------------------------------------------------
extern "C" int pb_get_sites(pb_ctx* ctx, pb_site_out* per_site) {
    CHECK_CTX(ctx);
    if (!ctx->events_done || !per_site) FAIL(ctx, PB_ERR_INVALID, "pb_get_sites: no results");
    PB_CUDA(cudaSetDevice(ctx->device));
    PB_CUDA(cudaMemcpy(per_site, ctx->d_sites, (size_t)ctx->S * sizeof(pb_site_out), cudaMemcpyDeviceToHost));
    return PB_OK;
}

extern "C" int pb_get_raw_events(pb_ctx* ctx, int64_t cap, int32_t* site, int32_t* node, int64_t* n) {
    CHECK_CTX(ctx);
    if (!ctx->events_done) FAIL(ctx, PB_ERR_INVALID, "pb_get_raw_events: no results");
    PB_CUDA(cudaSetDevice(ctx->device));
    if (n) *n = ctx->n_raw;
    if (!site || !node) return PB_OK;
    if (cap < ctx->n_raw) FAIL(ctx, PB_ERR_INVALID, "pb_get_raw_events: buffer too small");
    int64_t k = 0;
    if (ctx->plane_mode == PB_PLANES_X) {   // 16-byte records; those of the root's own edges are not MRCAs
        std::vector<xrec_t> hx((size_t)ctx->n_rec);
        if (ctx->n_rec) PB_CUDA(cudaMemcpy(hx.data(), ctx->d_raw, (size_t)ctx->n_rec * sizeof(xrec_t), cudaMemcpyDeviceToHost));
        for (const xrec_t& r : hx) {
            if (r.tag & PB_XT_NOMRCA) continue;
            for (uint64_t m = r.mask; m; m &= m - 1) {
                if (k >= cap) FAIL(ctx, PB_ERR_INVALID, "pb_get_raw_events: buffer too small");
                site[k] = (int32_t)(r.wcol * 64u + (uint32_t)__builtin_ctzll(m));
                node[k] = (int32_t)(r.tag & PB_XT_NODE_MASK);
                k++;
            }
        }
        return PB_OK;
    }
    std::vector<raw_rec_t> h((size_t)ctx->n_rec);
    if (ctx->n_rec) PB_CUDA(cudaMemcpy(h.data(), ctx->d_raw, (size_t)ctx->n_rec * sizeof(raw_rec_t), cudaMemcpyDeviceToHost));
    for (const raw_rec_t& r : h)
        for (uint64_t m = r.mrca; m; m &= m - 1) {
            if (k >= cap) FAIL(ctx, PB_ERR_INVALID, "pb_get_raw_events: buffer too small");
            site[k] = (int32_t)(r.wcol * 64u + (uint32_t)__builtin_ctzll(m));
            node[k] = (int32_t)(r.tag & PB_EV_NODE_MASK);
            k++;
        }
    return PB_OK;
}

extern "C" int pb_get_event_csr(pb_ctx* ctx, int64_t* offsets, int32_t* sample_index, int64_t cap) {
    CHECK_CTX(ctx);
    if (!ctx->events_done) FAIL(ctx, PB_ERR_INVALID, "pb_get_event_csr: no results");
    PB_CUDA(cudaSetDevice(ctx->device));
    if (!offsets) FAIL(ctx, PB_ERR_INVALID, "pb_get_event_csr: offsets is NULL (2*n_inf+1 entries; sizes are in pb_get_timings)");
    if (sample_index && cap < ctx->E) FAIL(ctx, PB_ERR_INVALID, "pb_get_event_csr: buffer too small");
    std::vector<AlleleMeta> h((size_t)(2 * ctx->S_inf));
    if (ctx->S_inf) PB_CUDA(cudaMemcpy(h.data(), ctx->d_alleles, h.size() * sizeof(AlleleMeta), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < 2 * ctx->S_inf; i++) offsets[i] = h[i].off;
    offsets[2 * ctx->S_inf] = ctx->E;
    if (ctx->E && sample_index) PB_CUDA(cudaMemcpy(sample_index, ctx->d_csr, (size_t)ctx->E * 4, cudaMemcpyDeviceToHost));
    return PB_OK;
}

extern "C" int pb_get_maxT_unsorted(pb_ctx* ctx, double* out) {
    CHECK_CTX(ctx);
    if (!ctx->assoc_done || !out) FAIL(ctx, PB_ERR_INVALID, "pb_get_maxT_unsorted: no results");
    PB_CUDA(cudaSetDevice(ctx->device));
    // sharded over peer memory: the pooled minimum lives in its own buffer (d_maxT is this shard's contribution)
    PB_CUDA(cudaMemcpy(out, ctx->xchg.active ? ctx->xchg.d_pooled : ctx->d_maxT, (size_t)ctx->cfg.n_replicates * 8, cudaMemcpyDeviceToHost));
    return PB_OK;
}

extern "C" int pb_get_ptable(pb_ctx* ctx, int32_t* n_max, double* out, int64_t cap) {
    CHECK_CTX(ctx);
    if (!ctx->assoc_done) FAIL(ctx, PB_ERR_INVALID, "pb_get_ptable: no results");
    PB_CUDA(cudaSetDevice(ctx->device));
    if (n_max) *n_max = ctx->ptable_nmax;
    const int64_t total = (int64_t)(ctx->ptable_nmax + 1) * (ctx->ptable_nmax + 2) / 2;
    if (out) {
        if (cap < total) FAIL(ctx, PB_ERR_INVALID, "pb_get_ptable: buffer too small");
        PB_CUDA(cudaMemcpy(out, ctx->d_ptable, (size_t)total * 8, cudaMemcpyDeviceToHost));
    }
    return PB_OK;
}

extern "C" int pb_get_timings(pb_ctx* ctx, pb_timings* out) {
    CHECK_CTX(ctx);
    if (!out) return PB_ERR_INVALID;
    *out = ctx->tm;
    return PB_OK;
}

extern "C" int pb_stopwatch_mark(pb_ctx* ctx, int32_t which) {
    CHECK_CTX(ctx);
    if (which != 0 && which != 1) FAIL(ctx, PB_ERR_INVALID, "pb_stopwatch_mark: which must be 0 or 1");
    PB_CUDA(cudaSetDevice(ctx->device));
    PB_CUDA(cudaEventRecord(ctx->ev[14 + which], ctx->stream));
    return PB_OK;
}

extern "C" int pb_stopwatch_ms(pb_ctx* ctx, float* ms) {
    CHECK_CTX(ctx);
    if (!ms) return PB_ERR_INVALID;
    PB_CUDA(cudaSetDevice(ctx->device));
    PB_CUDA(cudaEventSynchronize(ctx->ev[15]));
    PB_CUDA(cudaEventElapsedTime(ms, ctx->ev[14], ctx->ev[15]));
    return PB_OK;
}

extern "C" int pb_comm_unique_id(void* id128) {
    std::string err;
    int rc = pbn_unique_id(id128, err);
    if (rc) g_create_err = err;
    return rc;
}

extern "C" int pb_comm_init(pb_ctx* ctx, const void* id128, int32_t rank, int32_t world) {
    CHECK_CTX(ctx);
    if (!id128 || world < 1 || rank < 0 || rank >= world) FAIL(ctx, PB_ERR_INVALID, "pb_comm_init: bad rank/world");
    PB_CUDA(cudaSetDevice(ctx->device));
    ctx->xchg.active = false;      // (the fall-back after a peer-memory connect that did not succeed on every rank)
    return pbn_init(ctx, id128, rank, world);
}

extern "C" int pb_comm_export(pb_ctx* ctx, void* handle64) {
    CHECK_CTX(ctx);
    if (!handle64) FAIL(ctx, PB_ERR_INVALID, "pb_comm_export: null handle buffer");
    if (ctx->xchg.active) FAIL(ctx, PB_ERR_INVALID, "pb_comm_export: the context is already connected");
    PB_CUDA(cudaSetDevice(ctx->device));
    return pbx_export(ctx, handle64);
}

extern "C" int pb_comm_connect(pb_ctx* ctx, const void* handles, int32_t rank, int32_t world) {
    CHECK_CTX(ctx);
    if (!handles || world < 1 || rank < 0 || rank >= world) FAIL(ctx, PB_ERR_INVALID, "pb_comm_connect: bad rank/world");
    if (ctx->xchg.active || ctx->nccl_comm) FAIL(ctx, PB_ERR_INVALID, "pb_comm_connect: the context is already connected");
    PB_CUDA(cudaSetDevice(ctx->device));
    if (world == 1) { ctx->rank = 0; ctx->world = 1; return PB_OK; }
    return pbx_connect_ipc(ctx, handles, rank, world);
}

extern "C" int pb_comm_export_labels(pb_ctx* ctx, void* blob128) {
    CHECK_CTX(ctx);
    if (!blob128) FAIL(ctx, PB_ERR_INVALID, "pb_comm_export_labels: null buffer");
    if (ctx->P == 0) FAIL(ctx, PB_ERR_INVALID, "pb_comm_export_labels: call pb_set_labels first");
    PB_CUDA(cudaSetDevice(ctx->device));
    return pbx_labels_export(ctx, blob128);
}

extern "C" int pb_comm_connect_labels(pb_ctx* ctx, const void* blobs) {
    CHECK_CTX(ctx);
    if (!blobs) FAIL(ctx, PB_ERR_INVALID, "pb_comm_connect_labels: null buffer");
    PB_CUDA(cudaSetDevice(ctx->device));
    return pbx_labels_connect_ipc(ctx, blobs);
}

// internal (pb_group): allocate the label matrices of the current plan so that the group can map them across its shards
int pb_labels_alloc_internal(pb_ctx* ctx) {
    if (!ctx || ctx->P == 0) return PB_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return PB_ERR_CUDA;
    return pbk_labels_alloc(ctx);
}

extern "C" int pb_host_alloc(void** out, uint64_t bytes) {
    if (!out) return PB_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { g_create_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); return PB_ERR_OOM; }
    return PB_OK;
}
extern "C" int pb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? PB_OK : PB_ERR_CUDA; }
