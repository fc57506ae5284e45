// Internal declarations shared by the translation units of libpoutine_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/poutine_b200.h"

// ---- K1 edge program -----------------------------------------------------------------------------
// One 32-bit op per tree row visit, in DFS preorder (pb_api.cu build_chunks); the kernel consumes them as
// 16-byte records {op A, op B, flush id or PB_REC_NOFLUSH, 0}: two rows per ring stage, census flush after them.
//   bits 0..28 node index (PB_OP_NULL = no row), bit 29 TEST (MRCA test against the current parent row),
//   bit 30 SETCUR (this row becomes the current parent row), bit 31 LEAF (leaf census).
//   TEST|SETCUR|LEAF together = FLUSH: write the census partial number <node field> and clear it.
#define PB_OP_NODE_MASK 0x1FFFFFFFu
#define PB_OP_NULL 0x1FFFFFFFu
#define PB_OP_TEST 0x20000000u
#define PB_OP_SETCUR 0x40000000u
#define PB_OP_LEAF 0x80000000u
#define PB_OP_FLUSH 0xE0000000u
#define PB_REC_NOFLUSH 0xFFFFFFFFu

// The one-plane sweep (k1x_events_kernel) runs its own lowering of the same edge list: no census flags (the leaf census
// comes out of the events, see xrec_t), every edge tested — the edges hanging off the root too, marked NOMRCA because
// the reference never makes their child an MRCA (HC.java:1674) — and a 28-bit node field.  The flag bits are laid out so
// that a record's tag is ONE AND of the op: node (0..27), SETCUR (28; the tag's X bit), TEST (29), NOMRCA (30), LEAF (31).
#define PB_OPX_NODE_MASK 0x0FFFFFFFu
#define PB_OPX_NULL 0x0FFFFFFFu
#define PB_OPX_SETCUR 0x10000000u
#define PB_OPX_TEST 0x20000000u
#define PB_OPX_NOMRCA 0x40000000u
#define PB_OPX_LEAF 0x80000000u
#define PB_OPX_TAG_MASK (PB_OPX_NODE_MASK | PB_OPX_NOMRCA | PB_OPX_LEAF)

enum { PB_PLANES_NONE = 0, PB_PLANES_X = 1, PB_PLANES_FULL = 2 };
#define PB_MAX_SITES_PER_CTX (1ll << 26)   // scan word = inf count << 38 | event count; int32 site indices

// raw MRCA record: ALL events of one (tree row, 64-site word): the sweep kernels store the masks as they are and the
// tail kernels peel the bits (a per-event record cost the sweep ~36 warp-instructions per event, at one active lane)
struct __align__(16) raw_rec_t {
    uint32_t wcol;        // site word (context-local): sites 64*wcol ..
    uint32_t tag;         // node | is_leaf << 31
    uint64_t mrca;        // sites of the word where this node is an MRCA
    uint64_t b0, b1;      // the node's base-code planes of the word (one-plane mode: b0 = X, b1 = 0)
};
#define PB_EV_NODE_MASK 0x1FFFFFFFu

// One-plane mode: 16-byte records.  All sites of one (tree row, 64-site word) where X_child != X_parent AND the child
// carries X = x (two records when the word holds edges of both directions: rare).  The X bit is all the tail needs to
// rebuild the child's base code from the column's site masks — and to rebuild the LEAF CENSUS without ever counting leaf
// rows: X(leaf) = X(root) + sum over the edges (p -> c) of its root path of X(c) - X(p), hence
//   #leaves carrying the second base = L * X(root) + sum over edges with X(c) != X(p) of (X(c) ? +1 : -1) * leaves_below(c),
// a sum over exactly the (sparse) records, including those of the root's own edges (NOMRCA: census only).
struct __align__(16) xrec_t {
    uint32_t wcol;        // site word (context-local)
    uint32_t tag;         // node (28 bits) | X << 28 | NOMRCA << 30 | is_leaf << 31
    uint64_t mask;        // sites of the word with X_child != X_parent and X_child == X
};
#define PB_XT_NODE_MASK 0x0FFFFFFFu
#define PB_XT_XBIT 0x10000000u
#define PB_XT_NOMRCA 0x40000000u
#define PB_XT_LEAF 0x80000000u
static_assert(PB_XT_NODE_MASK == PB_OPX_NODE_MASK && PB_XT_NOMRCA == PB_OPX_NOMRCA && PB_XT_LEAF == PB_OPX_LEAF &&
              (PB_OPX_TAG_MASK & PB_XT_XBIT) == 0, "an xrec_t tag is the X op ANDed with PB_OPX_TAG_MASK");

// per-site event counters filled by k1_count_events: [0..3] all MRCAs by own base code,
// [4..7] leaf MRCAs, [8..11] leaf MRCAs that have a pheno row
#define PB_CNT_STRIDE 12
// the allocation holds one more int32 per site behind the [S][12] counters: the one-plane mode's signed census sum
#define PB_CNT_ALLOC_STRIDE 13

// site flags byte written by k1_finalize
#define PB_SF_BI 1u
#define PB_SF_A1_USE 2u
#define PB_SF_A2_USE 4u
#define PB_SF_A1_CODE_SHIFT 3  // 2 bits
#define PB_SF_A1_ALIVE 32u     // the allele's MRCA bucket survived the <2 rule (a_count >= 2)
#define PB_SF_A2_ALIVE 64u
#define PB_SF_INF 128u         // homoplasically informative: some allele in use (subset_by_min_hcount)

// census: 4 per-base leaf counters, bit-sliced: slices 0..2 (ones/twos/fours) in registers,
// PB_K1_KHI high slices in shared memory; a census segment holds at most PB_K1_SEG_LEAVES leaves
#define PB_K1_KHI 5
#define PB_K1_KS (3 + PB_K1_KHI)
#define PB_K1_SEG_LEAVES 248
#define PB_K1_KT 24         // slices of the reduced census (leaf counts < 2^24)
#ifndef PB_K1_COLS
#define PB_K1_COLS 128      // 64-site word columns per block (128 -> 4 resident blocks per SM, 256 -> 2)
#endif
#define PB_K1_BLOCKS_PER_SM (512 / PB_K1_COLS)
#ifndef PB_K1_LANE_BITS
#define PB_K1_LANE_BITS 64   // sites per consumer thread: one word column (must be 64)
#endif
#define PB_K1_CONSUMERS (PB_K1_COLS * 64 / PB_K1_LANE_BITS)
#define PB_K1_THREADS (PB_K1_CONSUMERS + 32)   // consumer warps + 1 TMA producer warp
#ifndef PB_K1_RPS
#define PB_K1_RPS 2         // 16-byte records (row pairs) per ring stage: one full/empty handshake per 2*RPS rows
#endif
#define PB_K1_STAGES (4 / PB_K1_RPS)   // ring = 8 rows (24 KB per 128 columns): 3 KB TMA boxes, 2*RPS of them per stage
#define PB_K1_QCAP 192   // per-block queue of raw records (6 KB)
#ifndef PB_K1_DRAIN_EVERY
#define PB_K1_DRAIN_EVERY 16  // op records (2 rows each) between queue-drain checks (a consumer-wide barrier each)
#endif

// allele flags
#define PB_AF_IN_USE 1u
#define PB_AF_FAST 2u   // {k : f(n,k) <= p_obs} is an up-set: threshold path for replicates that see all n labels
#define PB_AF_MISS_OK 8u // n <= 7 and every row f(n-j, .), j = 0..n, is an up-set w.r.t. p_obs: kpack is valid
#define PB_AF_TIGHT_CAND 16u // like TIGHT (no missing labels, n < 8) but some k can be a maxT candidate: 1 <= ktau[n] <= n
#define PB_AF_TIGHT 4u  // FAST and no replicate of this allele can be a maxT candidate: sweep without per-bit work

struct AlleleMeta {      // one per allele slot (2 per informative site), 32 bytes
    int64_t off;         // CSR offset of its sample list
    int32_t n;           // # extant MRCA leaves with a pheno row
    uint32_t flags;
    int32_t kstar;       // min k with f(n,k) <= p_obs (fast path)
    int32_t obs_case, obs_ctrl;
    uint32_t kpack;      // missing labels in the pheno file, n <= 7: k*_j for n_r = n - j, 4 bits each (j = 0..n)
};

// ---- the path's one exchange, over peer memory (pb_comm.cu) ----------------------------------------------------------
// Site shards pool the per-replicate minimum p-value maxT[m] (HC.java:3239-3243 takes it over ALL sites).  Every rank owns
// one exchange block {ready flags[64], error word, maxT buffer 0, maxT buffer 1} that its peers map (cudaIpc across
// processes, plain peer access inside one process).  K3 writes the epoch's buffer; a signal kernel then stores the epoch
// into every peer's flag word; the sort kernel of K4 waits for all flags and takes the minimum over the peers' buffers
// with NVLink loads WHILE it loads its tiles — the all-reduce is fused into the sort's first pass, no separate collective.
// Buffers alternate by epoch parity: a peer can run at most one epoch ahead (it needs this rank's next flag first), so the
// buffer being read is never the one being refilled.
#define PB_XCHG_MAX_WORLD 64
#define PB_XCHG_HDR_WORDS 256          // maxT-ready flags [0,64), error word 64, label-slice-ready flags [128,192), label-slice-pulled
                                       // flags [192,256) (2 KB), then the two maxT buffers
#define PB_XCHG_GEN_FLAGS 128
#define PB_XCHG_ACK_FLAGS 192
struct pb_xchg {
    bool active = false;
    unsigned long long* block = nullptr;                 // own block (cudaMalloc)
    size_t block_bytes = 0;
    unsigned long long* peer[PB_XCHG_MAX_WORLD] = {};    // every rank's block as seen from this device (own included)
    bool peer_ipc[PB_XCHG_MAX_WORLD] = {};               // mapped with cudaIpcOpenMemHandle (closed in pb_destroy)
    unsigned long long** d_peer_tbl = nullptr;           // device copy of peer[]
    unsigned long long* d_pooled = nullptr;              // [m] pooled (all-rank) minimum, unsorted: pb_get_maxT_unsorted
    unsigned long long epoch = 0;
    // label sharing: every shard needs the SAME label matrix, so each rank draws 1/world of a tile's replicate columns and
    // pulls the other ranks' columns out of their matrices (mapped like the exchange block)
    bool lab_active = false;
    size_t lab_bytes = 0;                                // size of the mapped matrix allocations (the plan they were made for)
    uint64_t* lab_c[PB_XCHG_MAX_WORLD] = {};             // every rank's d_Xcase as seen from this device (own included)
    uint64_t* lab_m[PB_XCHG_MAX_WORLD] = {};             // ... d_Xmiss (null without missing labels)
    bool lab_ipc[PB_XCHG_MAX_WORLD] = {};
    uint64_t** d_lab_c = nullptr;                        // device copies of the two tables
    uint64_t** d_lab_m = nullptr;
    unsigned long long gen_epoch = 0, buf_epoch[2] = {0, 0};   // one epoch per generated tile; the epoch that last filled each buffer
};
// what the fused sort kernel needs (by value)
struct XchgView {
    unsigned long long* const* peers;   // nullptr: single rank, plain load
    unsigned long long* pooled;
    int world, rank;
    unsigned long long epoch;
    long long buf_off;                  // word offset of the epoch's maxT buffer inside a block
};

struct pb_ctx {
    pb_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // label-matrix pre-generation (K3a) alongside the K1 tail (high priority: one-wave generator)
    cudaStream_t stream2_lo = nullptr;       // the same at the lowest priority: a many-wave generator must not starve the tail
    cudaStream_t stream3 = nullptr;          // K3b side stream: the few heavy short-list/long-list sweeps run next to the tight sweep
    cudaStream_t upl = nullptr;              // uploads (H2D copies, plane zero-fills, expansion kernels): the copy engine's own queue
    cudaEvent_t ev_upl = nullptr;            // last upload enqueued; the compute stream waits on it
    bool eager_ok = true;                    // PB_NO_EAGER_SWEEP unset
    int64_t swept_words = 0;                 // one-plane mode: site words [0, swept_words) were swept while later tiles uploaded
    // the tail of those tiles, run underneath the upload as well (pbk_tail_tile), and their per-site records on the way back
    int64_t tailed_words = 0;                // site words [0, tailed_words) have had their tail (<= swept_words)
    int tail_tiles = 0;                      // tiles that make them up: d_rec_range[t] = record cursor before tile t's sweep
    unsigned long long* d_rec_range = nullptr;
    cudaStream_t dl = nullptr;               // downloads ahead of pb_run_events (the D2H copy engine is idle during the upload)
    cudaEvent_t ev_tail_tile = nullptr;
    pb_site_out* out_sites = nullptr;        // pb_expect_site_output: where pb_run_events will be asked to put the per-site records
    int64_t sent_sites = 0;                  // records [0, sent_sites) are on their way there already
    bool eager_tail_ok = true;               // PB_NO_EAGER_TAIL unset
    cudaEvent_t ev_side_fork = nullptr, ev_side_join = nullptr, ev_zero_done = nullptr;
    cudaEvent_t ev_k1_done = nullptr, ev_gen_started = nullptr, ev_gen_done = nullptr, ev_gen_b[2] = {nullptr, nullptr}, ev_cnt_b[2] = {nullptr, nullptr};
    bool pregen_valid = false; int64_t pregen_tile = 0;
    cudaEvent_t ev_gen_t0 = nullptr, ev_gen_t1 = nullptr; bool gen_timed = false;   // the label generator's own time (first tile)
    std::string err;
    int sm_count = 148;

    // tree
    int32_t N = 0, L = 0, root = -1;
    std::vector<int32_t> parent;
    std::vector<uint8_t> is_leaf;
    std::vector<int32_t> lop_child, lop_parent;  // logical edge list in DFS preorder (host)
    uint32_t* d_ops = nullptr;      // K1 program
    int32_t* d_chunk_op = nullptr;  // n_chunks+1 record boundaries
    int32_t n_chunks = 0, n_flush = 0;
    uint32_t* d_ops_x = nullptr;    // the same edge list cut for k1x_events_kernel's grid (wider blocks, other residency)
    int32_t* d_chunk_op_x = nullptr;
    int32_t n_chunks_x = 0, n_flush_x = 0;
    int32_t* d_node_sample = nullptr;  // N: sample index of a leaf with a pheno row, else -1
    int32_t* d_leaves_below = nullptr; // N: leaves in the node's subtree (1 for a leaf): weights of the one-plane census sum

    // labels
    int32_t P = 0, n_case = 0, n_ctrl = 0, n_miss = 0;
    std::vector<int32_t> sample_node;
    std::vector<uint8_t> label;
    uint8_t* d_label = nullptr;

    // sites / planes
    int64_t S = 0, W = 0, Wp = 0;
    uint64_t* d_planes = nullptr;   // one allocation [3][N][Wp]: plane 0 = V, 1 = B0, 2 = B1
    uint64_t *d_B0 = nullptr, *d_B1 = nullptr, *d_V = nullptr;   // views into d_planes
    CUtensorMap tmap{};             // 3-D TMA descriptor over d_planes: box = 128 words x 1 node x 3 planes
    uint64_t* d_colmask = nullptr; int64_t colmask_cap = 0;   // [n_words][4] site masks of a compact (biallelic) tile upload
    // Plane storage mode.  PB_PLANES_X: every tile uploaded so far was biallelic with all genotypes valid, so the alignment is
    // held as ONE plane X[N][Wp] + per-column site masks, and events run k1x_events_kernel over a third of the bytes.  The
    // first tile that does not qualify promotes the context to the three full planes (pbk_promote_planes).
    int plane_mode = 0;             // PB_PLANES_*
    uint64_t* d_X = nullptr;        // [N][Wp]
    uint64_t* d_xmask = nullptr;    // [W][4] {A0, A1, D0, D1} of every uploaded column (zero where nothing was uploaded)
    uint64_t* d_xvalid = nullptr;   // [W] sites that exist in an uploaded column (padding bits of word W-1 excluded)
    CUtensorMap tmap_x{};           // 2-D TMA descriptor over d_X: box = 128 words x 1 node

    // K1 outputs
    uint64_t* d_census = nullptr;      // [n_flush][4][KS][W]   (three-plane mode only)
    uint64_t* d_census_total = nullptr;  // [4][KT][W]
    raw_rec_t* d_raw = nullptr;        // raw_cap records of 32 bytes; the one-plane mode writes xrec_t (16 bytes) into the same buffer
    int64_t raw_cap = 0, n_rec = 0;   // capacity / length of the record list (n_raw = MRCA events = set bits over all MRCA records)
    unsigned long long* d_raw_cursor = nullptr;
    int64_t n_raw = 0;
    uint32_t* d_cnt = nullptr;         // [S][PB_CNT_STRIDE]
    pb_site_out* d_sites = nullptr;    // [S]
    uint8_t* d_site_flags = nullptr;   // [S]
    uint64_t* d_site_scan = nullptr;   // [S] exclusive prefix: inf count << 38 | event count
    unsigned long long* d_scan_state = nullptr; int64_t scan_state_cap = 0;   // single-pass scan of the tail: ticket + 3 words per tile
    unsigned long long* d_class_counts = nullptr;  // 5 + n_max + scan totals ...
    bool events_done = false;

    // informative sites / CSR
    int64_t S_inf = 0, E = 0, A_inuse = 0;
    int32_t n_max = 0;
    int32_t* d_inf_site = nullptr;     // [S_inf]
    AlleleMeta* d_alleles = nullptr;   // [2*S_inf]
    int32_t* d_csr = nullptr;          // [E]
    uint32_t* d_csr_cursor = nullptr;  // [2*S_inf]
    int64_t inf_cap = 0, csr_cap = 0, csr_want = 0;   // csr_want: E overflowed the CSR buffer, grow to this on the next pass

    // assoc
    double p_success = 0;
    double* d_ptable = nullptr;  int64_t ptable_cap = 0;  int32_t ptable_nmax = -1;
    void* d_thr = nullptr; int64_t thr_cap = 0;  // threshold scratch (expect, tau, ktau, n-histogram)
    int32_t* d_ktau = nullptr;   // [n_max+1]   (inside d_thr)
    unsigned long long* d_nfast = nullptr;     // (inside d_thr) followed by the 4 list counts and the fallback total
    unsigned int* d_lcount = nullptr;          // (inside d_thr) lengths of the 4 K3b work lists
    unsigned long long* d_fb_total = nullptr;  // (inside d_thr) replicates recomputed by the maxT fallback
    double* d_tau = nullptr;     // 1           (inside d_thr)
    double* d_pobs = nullptr;    // [2*S_inf]
    uint32_t* d_r = nullptr;     // [2*S_inf]
    unsigned long long* d_maxT = nullptr;  // [m] (double bits): the private allocation below, or the epoch's buffer of the exchange block
    unsigned long long* d_maxT_own = nullptr;
    pb_xchg xchg;
    int64_t prep_sinf = -1; int32_t prep_nmax = -1, prep_P = -1;   // what pbk_assoc_prepare last sized the buffers for (buffers only grow)
    int64_t site_index_base = 0;           // added to pb_stat_out.site_index (a pb_group reports group-global indices)
    double* d_maxT_sorted = nullptr;       // [pow2 >= m]
    // K4 counting path: sorted table keys (+inf last), bin histogram (+ miss flag), cumulative counts, a copy of maxT
    unsigned long long* d_tabkeys = nullptr; unsigned int* d_binhist = nullptr; unsigned long long* d_bincum = nullptr;
    unsigned long long* d_pooled_own = nullptr; int64_t pooled_own_cap = 0;
    bool tabkeys_valid = false;
    int64_t maxT_cap = 0, sorted_cap = 0;
    uint64_t *d_Xcase = nullptr, *d_Xmiss = nullptr;  // [P][Rw]
    int64_t X_cap = 0, Xm_cap = 0, fb_cap = 0;
    uint32_t* d_perms = nullptr; int64_t perms_cap = 0;
    pb_stat_out* d_stats = nullptr; int64_t stats_cap = 0;
    int32_t* d_work = nullptr; int64_t work_cap = 0;   // 4 work lists of slot ids (k4_classify_kernel), n_slots each
    int64_t n_work[4] = {0, 0, 0, 0};
    int sort_blocks_per_sm = -1;                                 // occupancy of k4_sort_coop_kernel on this context's device
    bool k1_attr_set = false, k1x_attr_set = false;              // cudaFuncSetAttribute(max dynamic shared memory) done on this context's device
    uint32_t* d_colpop = nullptr; int64_t colpop_cap = 0;  // [2][P] per-sample case / missing-label popcounts of the current tile
    int32_t* d_fb_reps = nullptr;  // fallback replicate list
    unsigned int* d_fb_count = nullptr;
    double* d_extras = nullptr; int64_t extras_cap = 0;   // [3*S_inf] exact point-wise (2 per site) + Fisher (k5_extras.cu)
    bool assoc_done = false;

    // nccl
    void* nccl_comm = nullptr;
    int32_t rank = 0, world = 1;

    // timing
    pb_timings tm{};
    cudaEvent_t ev[16]{};
    // PB_TRACE=1: an event after every main-stream kernel of the tail / assoc, printed by pb_trace_dump (debug aid)
    bool trace_on = false;
    std::vector<std::pair<const char*, cudaEvent_t>> trace;
    size_t trace_used = 0;
};

#define PB_CUDA(call)                                                                             \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
            return PB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// ---- kernel launchers (defined in the .cu files) ---------------------------------------------
int pbk_events(pb_ctx* ctx);                                   // k1_events.cu
int pbk_events_range(pb_ctx* ctx, int64_t block0, int64_t n_blocks, bool first, bool last);   // k1_events.cu: one-plane sweep of a column-block range
int pbk_mask_last_word(pb_ctx* ctx);                           // k1_events.cu (clears the padding bits of site word W-1 in V)
int pbk_kx_cols();                                             // k1_events.cu: word columns per k1x block
int pbk_kx_blocks_per_sm();
int pbk_promote_planes(pb_ctx* ctx);                           // k1_events.cu: X + masks -> the three full planes (already allocated)
int pbk_mask_last_word_x(pb_ctx* ctx);                         // k1_events.cu
int pbk_expand_biallelic(pb_ctx* ctx, int64_t word_begin, int64_t n_words, const uint64_t* d_colmask, bool fill_v);  // k1_events.cu
int pbk_prepare_tail(pb_ctx* ctx);                             // k1_events.cu (tail buffers + zero-fill on stream2, before K1)
int pbk_finalize_events(pb_ctx* ctx, pb_class_counts* counts, unsigned long long* cursor_out, int64_t tailed_words); // k1_events.cu
int pbk_tail_tile(pb_ctx* ctx, int64_t word_begin, int64_t n_words);   // k1_events.cu: the tail of one eagerly swept tile (+ its records' copy back)
#define PB_MAX_TAIL_TILES 4096
int pbk_assoc(pb_ctx* ctx, const uint32_t* replay_perms);      // k3_null.cu (+ k4)
int pbk_pregen_labels(pb_ctx* ctx, bool immediately);          // k3_null.cu
int pbk_ptable(pb_ctx* ctx, bool alloc_only = false);          // k4_pvalues.cu
int pbk_thresholds(pb_ctx* ctx, bool alloc_only = false);      // k4_pvalues.cu
int pbk_pvalues(pb_ctx* ctx, bool alloc_only = false);         // k4_pvalues.cu
int pbk_assoc_prepare(pb_ctx* ctx);                            // k3_null.cu: every allocation of pb_run_assoc, up front
int pbk_extras(pb_ctx* ctx, double* h_exact_pointwise, double* h_fisher);   // k5_extras.cu

// pb_comm.cu: peer-memory exchange
int pbx_export(pb_ctx* ctx, void* handle64);
int pbx_connect_ipc(pb_ctx* ctx, const void* handles, int rank, int world);
int pbx_connect_local(pb_ctx** ctxs, int n);     // contexts of ONE process, one per GPU
int pbx_begin_epoch(pb_ctx* ctx);                // points d_maxT at the epoch's buffer
int pbx_signal(pb_ctx* ctx);                     // after K3: tell every peer this rank's buffer of the epoch is complete
XchgView pbx_view(pb_ctx* ctx);
int pbx_check(pb_ctx* ctx);                      // after the stream synchronised: did a peer fail to arrive?
bool pbx_labels_shared(pb_ctx* ctx, size_t need_bytes);    // label sharing connected and the matrices are still the mapped ones
int pbx_labels_begin(pb_ctx* ctx, int buf, cudaStream_t st);    // before a tile is drawn into `buf`: peers have pulled its previous content
int pbx_labels_exchange(pb_ctx* ctx, int buf, int64_t buf_off_words, int32_t P, int64_t Rw, int64_t total_blocks, int64_t words_per_block,
                        bool has_miss, cudaStream_t st);    // after this rank's slice is drawn: signal, pull the peers' slices, acknowledge
int pbx_labels_export(pb_ctx* ctx, void* blob128);
int pbx_labels_connect_ipc(pb_ctx* ctx, const void* blobs);
int pbx_labels_connect_local(pb_ctx** ctxs, int n);
int pbk_labels_alloc(pb_ctx* ctx);               // k3_null.cu: the label matrices of the current plan (P, m), allocated now
void pbx_destroy(pb_ctx* ctx);

// nccl_dyn.cpp
int pbn_unique_id(void* id128, std::string& err);
int pbn_init(pb_ctx* ctx, const void* id128, int rank, int world);
int pbn_allreduce_min_f64(pb_ctx* ctx, double* buf, int64_t n);
void pbn_destroy(pb_ctx* ctx);

// debug timeline (PB_TRACE=1): mark = event on the main stream; dump after a synchronize
static inline void pb_trace_mark(pb_ctx* ctx, const char* label) {
    if (!ctx->trace_on) return;
    if (ctx->trace_used == ctx->trace.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->trace.push_back({label, e}); }
    ctx->trace[ctx->trace_used].first = label;
    cudaEventRecord(ctx->trace[ctx->trace_used].second, ctx->stream);
    ctx->trace_used++;
}
static inline void pb_trace_dump(pb_ctx* ctx, const char* what) {
    if (!ctx->trace_on || ctx->trace_used < 2) { ctx->trace_used = 0; return; }
    fprintf(stderr, "[poutine_b200 trace] %s:", what);
    for (size_t i = 1; i < ctx->trace_used; i++) {
        float t = 0; cudaEventElapsedTime(&t, ctx->trace[i - 1].second, ctx->trace[i].second);
        fprintf(stderr, " %s %.1f", ctx->trace[i].first, t * 1000.f);
    }
    float tt = 0; cudaEventElapsedTime(&tt, ctx->trace[0].second, ctx->trace[ctx->trace_used - 1].second);
    fprintf(stderr, " | total %.1f us\n", tt * 1000.f);
    ctx->trace_used = 0;
}
