/*
 * poutine_b200.h — C ABI of libpoutine_b200.so
 *
 * B200-native (sm_100a) replacement for the ONE hot path of POUTINE's Homoplasy_Counter:
 *     count_all_homoplasy_events(tree, seg_sites, physical_poss)   HC.java:1402  (called HC.java:299)
 *     assoc(all_events, phenos)                                    HC.java:2029  (called HC.java:314)
 * ("HC.java" = /root/reference/src/Homoplasy_Counter.java).  The reference has no plugin/FFI
 * surface (both are private methods of one class), so the boundary is cut at those two calls:
 * the Java host keeps CLI, loaders and the TSV writer, flattens its NewickTree / Fasta records
 * into the plain arrays below and binds these entry points through FFM or JNI
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C, no callbacks, no C++/torch types; every pointer is caller-owned host memory,
 *     copied during the call and never retained (exception: none).
 *   - return 0 on success, negative pb_status on failure; pb_last_error() gives the message.
 *     The reference's convention is "print, stack trace to the log, System.exit(-1)"
 *     (e.g. HC.java:3267-3271); the host maps a negative status onto that.
 *   - there is NO CPU fallback: a missing sm_100 device or a failed launch is an error.
 *   - one context per host thread / per GPU; calls on a context are serialised by the caller.
 */
#ifndef POUTINE_B200_H
#define POUTINE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

typedef enum {
    PB_OK = 0,
    PB_ERR_INVALID = -1,         /* bad argument / call order */
    PB_ERR_CUDA = -2,            /* CUDA runtime error (incl. no sm_100 device) */
    PB_ERR_NO_INFORMATIVE = -3,  /* min_hcount filtered out every site — HC.java:3866-3869 */
    PB_ERR_NCCL = -4,
    PB_ERR_OOM = -5
} pb_status;

enum { PB_MODE_PHILOX = 0, PB_MODE_REPLAY = 1 };

/* label codes of pb_set_labels (pheno file column 2, HC.java:3747-3751) */
enum { PB_LABEL_CONTROL = 0, PB_LABEL_CASE = 1, PB_LABEL_MISSING = 2 };

/* site classes (switch at HC.java:1443-1473) */
enum { PB_SITE_NO_ALLELE = 0, PB_SITE_MONO = 1, PB_SITE_BI = 2, PB_SITE_TRI = 3, PB_SITE_QUAD = 4 };

typedef struct {
    int32_t device;          /* CUDA ordinal */
    int32_t mode;            /* PB_MODE_PHILOX | PB_MODE_REPLAY */
    int64_t n_replicates;    /* m: --num-replicates, default 100000 (HC.java:89) */
    int32_t min_hcount;      /* default 1 (HC.java:98) */
    int32_t reserved0;
    uint64_t seed;           /* Philox key; every rank of a sharded run must pass the same seed */
    int64_t site_offset;     /* global index of this context's first site (site sharding, 64-aligned) */
    double maxt_tau;         /* 0 = auto.  Candidate threshold of the maxT pass (test hook) */
    uint32_t flags;          /* PB_FLAG_* */
    uint32_t reserved1;
} pb_config;

enum {
    PB_FLAG_FORCE_GENERAL_NULL = 1u,  /* disable the all-labelled bit-sliced fast path (test hook) */
    PB_FLAG_KEEP_RAW_EVENTS = 2u,     /* reserved: the raw MRCA event list is always kept until the next pb_run_events */
    PB_FLAG_FULL_PLANES = 4u          /* always hold three planes in HBM, even when every uploaded tile is compact */
};

/* One record per site of this context, field-for-field Homoplasy_Events (HC.java:5531-5550);
 * physical_pos stays on the host (it is only echoed into the output rows). */
typedef struct {
    int32_t segsite_id;   /* 1-based global id (HC.java:1419) */
    int32_t site_class;   /* PB_SITE_*; all fields below are 0 unless PB_SITE_BI */
    int32_t allele1;      /* char: major allele (HC.java:1712-1743) */
    int32_t allele2;      /* char: minor allele */
    int32_t a1_count;     /* |allele1_MRCAs| after the <2 clearing (HC.java:1954-1962) */
    int32_t a2_count;
    int32_t a1_extant;    /* a1_count_extant_only (HC.java:5586-5600) */
    int32_t a2_extant;
} pb_site_out;

/* # bi-allelic / monomorphic / tri / quad sites, HC.java:1477-1480 (+ the ">4"/no-allele default arm) */
typedef struct {
    int64_t n_no_allele, n_mono, n_bi, n_tri, n_quad;
    int64_t n_raw_events;     /* MRCA nodes found over all sites, before the <2 clearing, root excluded */
} pb_class_counts;

/* One record per homoplasically informative site, Binomial_Test_Stat (HC.java:5876-5887)
 * + obs_counts (HC.java:3042); sentinels -1 / NaN as in its constructor (HC.java:5891-5906). */
typedef struct {
    int32_t site_index;       /* 0-based index into this context's sites */
    int32_t segsite_id;
    int32_t a1_in_use, a2_in_use;
    int32_t obs_counts[4];    /* a1case, a1ctrl, a2case, a2ctrl */
    int32_t r_a1, r_a2, r_maxT_a1, r_maxT_a2;
    double obs_p_a1, obs_p_a2, pointwise_a1, pointwise_a2, familywise_a1, familywise_a2;
} pb_stat_out;

/* Device time of the last pb_run_events / pb_run_assoc, CUDA events on the library's stream. */
typedef struct {
    float ms_events_kernel;   /* K1 main kernel only (the HBM-bound plane sweep) */
    float ms_events_total;    /* K1 + count + finalize + compaction */
    float ms_tally;           /* K2 */
    float ms_ptable;          /* p-value table + thresholds */
    float ms_null_gen;        /* K3a label bit-matrix generation as seen by the main stream (~0 when the tile was
                                 pre-generated on the second stream during pb_run_events) */
    float ms_null_count;      /* K3b replicate counting */
    float ms_null_fallback;   /* (the maxT fallback pass now runs inside the count tiles: included in ms_null_count) */
    float ms_allreduce;       /* NCCL all-reduce(min) of maxT */
    float ms_pvalues;         /* K4: sort maxT + pointwise/familywise */
    float ms_assoc_total;
    int64_t n_launches;       /* kernels launched by the last events+assoc pair */
    int64_t k1_algorithmic_bytes;  /* what the sweep kernel moves: planes*N*W*8 plane bytes (planes = 3, or 1 in one-plane mode)
                                      + one record (32 B; one-plane mode 16 B) per (tree row, site word) holding an event */
    int64_t n_informative, n_alleles_in_use, n_extant_events;  /* S_inf, A, E of SURVEY §8 */
    int64_t n_fast_alleles;   /* alleles handled by the bit-sliced threshold path */
    int64_t n_maxt_fallback_reps;
    int64_t plane_mode;       /* how the alignment was held in HBM for the last pb_run_events: 1 = one plane X + site masks
                                 (every tile compact, all genotypes valid; k1x_events_kernel), 2 = three planes (k1_events_kernel) */
    int64_t eager_words;      /* site words whose sweep ran ahead of pb_run_events, underneath the uploads (0: none) */
    int64_t n_swept_rows;     /* K3b: label-matrix rows read per replicate word, summed over the swept alleles (x m/8 B = its L2 traffic) */
    float ms_gen_kernel;      /* K3a: the label generator's own time for the first replicate tile (events on its stream) */
    float gen_shard_frac;     /* fraction of the label matrix this context drew itself: 1, or 1/world when the shards share it */
    int64_t tailed_words;     /* site words whose tail (counts, per-site records, compaction, CSR) ran ahead of pb_run_events as well (<= eager_words) */
    int64_t early_sites;      /* per-site records that were already on their way to the host when pb_run_events was called (pb_expect_site_output) */
} pb_timings;

int pb_create(const pb_config* cfg, pb_ctx** out);
void pb_destroy(pb_ctx* ctx);
/* ctx may be NULL: returns the calling thread's last pb_create error. */
const char* pb_last_error(pb_ctx* ctx);

/* NewickTree flattened by the host: parent index per node (root = -1), leaf flags.
 * Any node order; the host keeps the name table.  Replaces the tree argument of HC.java:1402.
 * Limits: 2 <= n_nodes <= 2^28 - 2, fewer than 2^24 leaves. */
int pb_set_tree(pb_ctx* ctx, int32_t n_nodes, const int32_t* parent, const uint8_t* is_leaf);

/* `phenos` (HC.java:1972-2019) flattened: one entry per pheno-file row in file order:
 * the leaf's node index and its label code.  Leaves without a row are simply absent. */
int pb_set_labels(pb_ctx* ctx, int32_t n_samples, const int32_t* sample_node, const uint8_t* label);

/* Number of sites this context holds (1 .. 2^26; a longer alignment is sharded over contexts / a pb_group);
 * (re)allocates the device planes. */
int pb_set_sites(pb_ctx* ctx, int64_t n_sites);

/* Node-major bit planes of `seg_sites` (HC.java:1241-1306): 2-bit base code A0 C1 G2 T3 in
 * (B1,B0), V = genotype is an upper-case ACGT.  Bit j of word w = site 64*w+j.  Uploads the
 * column tile [word_begin, word_begin+n_words) of all nodes; row v of each source starts at
 * src + v*src_pitch_words.  Streamable: call once per tile.  A whole-alignment upload whose src_pitch_words equals
 * the device pitch (W rounded up to a multiple of 16 words) is done as one flat copy per plane. */
int pb_put_planes(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                  const uint64_t* B0, const uint64_t* B1, const uint64_t* V);

/* Compact upload of a column tile in which every site carries at most TWO distinct bases over all nodes (the usual
 * biallelic SNP alignment): one plane instead of three crosses PCIe, the library expands it on the device into the same
 * B0/B1/V rows pb_put_planes would have written (identical results; under V = 0 the B bits are don't-care).
 *   X        bit = the node's genotype is the site's second base (0 = its first base, or not a valid base)
 *   V        as in pb_put_planes, or NULL when every genotype of the tile is an upper-case ACGT
 *   colmask  n_words x 4 words {A0, A1, D0, D1}: per site the 2-bit code of the first base (A1:A0) and first XOR second
 *            base (D1:D0); the device computes B0 = A0 ^ (X & D0), B1 = A1 ^ (X & D1)
 * pb_compact_planes (host helper below) derives X / colmask from full planes and says whether the tile qualifies. */
int pb_put_planes_biallelic(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                            const uint64_t* X, const uint64_t* V, const uint64_t* colmask);

/* Asynchronous uploads: as above, but the call returns once the copies are QUEUED; the host buffers must stay valid and
 * unchanged until pb_upload_sync (or the next pb_run_events, which orders itself behind every queued upload but does not
 * wait for the host's sake).  Lets one host thread keep several GPUs' copy engines busy at once (pb_group does this), or
 * pack tile t+1 while tile t is in flight without a helper thread.
 * One-plane tiles that arrive in column order and are cut at multiples of 256 site words (the sweep's block width; the
 * last tile may be shorter) are swept right after they land, underneath the next tile's copy, and the label generator
 * starts with the first tile when the labels are already set: pb_run_events then only runs the tail. */
int pb_put_planes_async(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                        const uint64_t* B0, const uint64_t* B1, const uint64_t* V);
int pb_put_planes_biallelic_async(pb_ctx* ctx, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                                  const uint64_t* X, const uint64_t* V, const uint64_t* colmask);
int pb_upload_sync(pb_ctx* ctx);
/* Optional, before the streamed upload of a pass: where the coming pb_run_events will be asked to put its per_site records
 * (n_sites records; page-locked memory, pb_host_alloc, if the copies are to overlap anything).  One-plane tiles that are swept
 * underneath the upload then have their tail — event counts, Homoplasy_Events fields, the informative-site compaction, the
 * CSR of extant events — run there as well (needs pb_set_labels before the upload), and their finished records travel back
 * while later tiles are still going up: PCIe is full duplex.  pb_run_events(ctx, the same pointer, ...) then handles only
 * what is left (usually the last tile) and returns when every record has landed.  The registration stays until it is replaced
 * (NULL = none) or pb_set_sites changes the site count; the buffer must stay valid until the pb_run_events that delivers into
 * it has returned.  Results are identical with and without it. */
int pb_expect_site_output(pb_ctx* ctx, pb_site_out* per_site);

/* count_all_homoplasy_events (HC.java:1402).  per_site (n_sites records) and counts may be NULL.
 * If pb_set_labels was called before (Philox mode), the label matrix of the first replicate tile is generated on a
 * second stream behind the event kernel, so that pb_run_assoc finds it ready: set labels first, then run events. */
int pb_run_events(pb_ctx* ctx, pb_site_out* per_site, pb_class_counts* counts);

/* assoc (HC.java:2029 -> 3013-3186): observed tallies, permutation null, point-wise and
 * family-wise p-values.  replay_perms: m*P uint32 (required iff PB_MODE_REPLAY): in replicate r
 * sample j receives label[replay_perms[r*P+j]].  per_inf_site has room for `cap` records
 * (n_inf is always written; PB_ERR_INVALID if cap is too small).  maxT_sorted (m doubles) may be NULL. */
int pb_run_assoc(pb_ctx* ctx, const uint32_t* replay_perms, pb_stat_out* per_inf_site, int64_t cap,
                 int64_t* n_inf, double* maxT_sorted);

/* The two calls back to back, as Homoplasy_Counter.call() makes them (HC.java:299, 314), in one entry: the per-site
 * records travel to the host underneath the association kernels instead of in front of them.  per_inf_site must have room
 * for the worst case known before the call (cap >= n_sites always suffices).  Same results and errors as the two calls. */
int pb_run_events_assoc(pb_ctx* ctx, pb_site_out* per_site, pb_class_counts* counts, const uint32_t* replay_perms,
                        pb_stat_out* per_inf_site, int64_t cap, int64_t* n_inf, double* maxT_sorted);

/* Two per-site statistics outside the reference's LIVE path, on request after pb_run_assoc (either pointer may be NULL).
 * PARITY UNPINNED: the reference has no runnable counterpart; checked against exact rational arithmetic (1e-12 relative).
 *   exact_pointwise[2*i + a]  the m -> infinity limit of allele a's point-wise estimate r/m (HC.java:4812-4838) for
 *                             informative site i: the exact probability, under a uniform permutation of the pheno
 *                             labels (HC.java:4890-4923), that a replicate's f(n_r, k_r) <= p_obs (HC.java:4239);
 *                             NaN when the allele is not in use.  
 *   fisher_p[i]               two-sided Fisher's exact p-value of {a1case, a1ctrl; a2case, a2ctrl} (= obs_counts), R's
 *                             fisher.test definition; the reference's dead fishers_exact path (HC.java:5008-5093) pipes
 *                             the same table to an R script that is not in its repository. */
int pb_run_extras(pb_ctx* ctx, double* exact_pointwise /* 2*n_inf */, double* fisher_p /* n_inf */);

/* Introspection used by the parity tests. */
int pb_get_sites(pb_ctx* ctx, pb_site_out* per_site);
int pb_get_raw_events(pb_ctx* ctx, int64_t cap, int32_t* site, int32_t* node, int64_t* n);
int pb_get_event_csr(pb_ctx* ctx, int64_t* offsets /* 2*n_inf+1 */, int32_t* sample_index /* E */, int64_t cap);
int pb_get_maxT_unsorted(pb_ctx* ctx, double* out /* m */);
int pb_get_ptable(pb_ctx* ctx, int32_t* n_max, double* out, int64_t cap);
int pb_get_timings(pb_ctx* ctx, pb_timings* out);
/* Stopwatch on the library's own stream (CUDA events): which = 0 records the start mark, 1 the stop
 * mark; pb_stopwatch_ms synchronises on the stop mark and returns the device time between them. */
int pb_stopwatch_mark(pb_ctx* ctx, int32_t which);
int pb_stopwatch_ms(pb_ctx* ctx, float* ms);

/* Multi-GPU: one context per GPU, sites sharded (pb_config.site_offset); every context must be given the same tree,
 * labels, seed and m.  The only exchange is the minimum of maxT[m] over the shards inside pb_run_assoc (the reference
 * takes that minimum over all sites of the alignment, HC.java:3239-3243); every rank must call pb_run_assoc the same
 * number of times.  Three ways to connect the contexts:
 *
 *  (1) one process, one host thread: pb_group below — the library drives all GPUs of the box itself.
 *  (2) one process per GPU, peer memory (NVLink / NVSwitch): every rank calls pb_comm_export, the host gathers the 64-byte
 *      handles of all ranks (any transport) and hands the rank-ordered array to pb_comm_connect.  The reduction is fused
 *      into K4's sort kernel (peer loads over NVLink), there is no separate collective.  Before any rank calls pb_destroy,
 *      every rank must have returned from its last pb_run_assoc (host-side barrier).
 *  (3) one process per GPU, NCCL all-reduce(min): rank 0 creates the id, the host broadcasts its 128 bytes, pb_comm_init.
 *      For boxes where (2) is not available (cudaIpcOpenMemHandle refused: pb_comm_connect returns PB_ERR_NCCL). */
int pb_comm_export(pb_ctx* ctx, void* handle64);
int pb_comm_connect(pb_ctx* ctx, const void* handles /* world x 64 bytes, rank order */, int32_t rank, int32_t world);
/* Optional, after pb_set_labels and pb_comm_connect (Philox mode): map the ranks' label matrices as well.  Every shard
 * sweeps the SAME label matrix, so each rank then draws only 1/world of its replicate columns and pulls the rest from its
 * peers over NVLink (pb_timings.gen_shard_frac).  Same results.  Gather the 128-byte blobs in rank order like the handles.
 * (A pb_group does this by itself.) */
int pb_comm_export_labels(pb_ctx* ctx, void* blob128);
int pb_comm_connect_labels(pb_ctx* ctx, const void* blobs /* world x 128 bytes, rank order */);
int pb_comm_unique_id(void* id128);
int pb_comm_init(pb_ctx* ctx, const void* id128, int32_t rank, int32_t world);

/* pb_group: all GPUs of one box behind the ONE host thread that Homoplasy_Counter.call() is (HC.java:230-319; SURVEY
 * section 8b "Threading").  Same calls as a single context; the group shards the sites into contiguous 64-site-word
 * ranges (shard g = words [W*g/G, W*(g+1)/G)), drives each device from its own worker thread and merges the outputs:
 * per_site / per_inf_site come back in alignment order with alignment-global segsite_id and site_index, class counts are
 * summed, maxT_sorted is the pooled null.  Results are identical to one context holding the whole alignment.
 * cfg->device is ignored (devices[] lists the CUDA ordinals), cfg->site_offset is the offset of the group's first site.
 * PB_ERR_NO_INFORMATIVE is judged over the whole alignment (a shard may be empty). */
typedef struct pb_group pb_group;
int pb_group_create(const pb_config* cfg, const int32_t* devices, int32_t n_devices, pb_group** out);
void pb_group_destroy(pb_group* g);
const char* pb_group_last_error(pb_group* g);       /* g may be NULL: the calling thread's last pb_group_create error */
int32_t pb_group_size(pb_group* g);
pb_ctx* pb_group_ctx(pb_group* g, int32_t i);       /* shard i's context (introspection: pb_get_timings, pb_get_raw_events ...) */
int pb_group_set_tree(pb_group* g, int32_t n_nodes, const int32_t* parent, const uint8_t* is_leaf);
int pb_group_set_labels(pb_group* g, int32_t n_samples, const int32_t* sample_node, const uint8_t* label);
int pb_group_set_sites(pb_group* g, int64_t n_sites);
int pb_group_shard_range(pb_group* g, int32_t i, int64_t* site_begin, int64_t* site_end);
/* column tiles in alignment-global word coordinates; a tile may straddle shards (each part goes to its device, in parallel) */
int pb_group_put_planes(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                        const uint64_t* B0, const uint64_t* B1, const uint64_t* V);
int pb_group_put_planes_biallelic(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                                  const uint64_t* X, const uint64_t* V, const uint64_t* colmask);
int pb_group_put_planes_async(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                              const uint64_t* B0, const uint64_t* B1, const uint64_t* V);
int pb_group_put_planes_biallelic_async(pb_group* g, int64_t word_begin, int64_t n_words, int64_t src_pitch_words,
                                        const uint64_t* X, const uint64_t* V, const uint64_t* colmask);
int pb_group_upload_sync(pb_group* g);
int pb_group_expect_site_output(pb_group* g, pb_site_out* per_site /* n_sites records, alignment order; after pb_group_set_sites */);
int pb_group_run_events(pb_group* g, pb_site_out* per_site /* n_sites, may be NULL */, pb_class_counts* counts);
int pb_group_run_assoc(pb_group* g, const uint32_t* replay_perms, pb_stat_out* per_inf_site, int64_t cap, int64_t* n_inf,
                       double* maxT_sorted);
int pb_group_run_extras(pb_group* g, double* exact_pointwise, double* fisher_p);
int pb_group_get_timings(pb_group* g, pb_timings* out);   /* times: slowest shard; counts: summed over shards */
int pb_group_stopwatch_mark(pb_group* g, int32_t which);
int pb_group_stopwatch_ms(pb_group* g, float* ms);        /* slowest shard */

/* Host-side helpers (no GPU work). */
/* Pack `seg_sites` chars (node-major, row v at seq + v*seq_pitch) into bit planes — the
 * Fasta_Record -> plane step of the host (HC.java:1252-1266). */
int pb_pack_chars(const char* seq, int64_t n_nodes, int64_t n_sites, int64_t seq_pitch,
                  uint64_t* B0, uint64_t* B1, uint64_t* V, int64_t pitch_words);
/* Tries to compact a tile of full planes (n_sites = sites of the TILE, rows at pitch_words) for pb_put_planes_biallelic.
 * Returns 1 and fills X (rows at x_pitch_words), colmask (4 words per word column) and *all_valid (1: V may be passed as
 * NULL) when every site of the tile has at most two distinct valid bases over all nodes; returns 0 (outputs undefined)
 * when some site has three or more: upload that tile with pb_put_planes.  n_threads 0 = all host threads. */
int pb_compact_planes(const uint64_t* B0, const uint64_t* B1, const uint64_t* V, int64_t n_nodes, int64_t n_sites,
                      int64_t pitch_words, uint64_t* X, int64_t x_pitch_words, uint64_t* colmask, int32_t* all_valid,
                      int32_t n_threads);
/* FASTA -> planes at speed (host side of HC.java:1241-1306; record semantics of compiled/Fasta_Manager.class next()):
 * the file is mmap'ed and indexed once; pb_fasta_pack_tile packs the column tile [site_begin, site_begin+n_sites_tile)
 * (site_begin % 64 == 0) of every record into row node_of_record[i] (-1 = skip) of tile-local planes, multi-threaded,
 * so a host can pack tile t+1 while pb_put_planes uploads tile t.  Header text is returned trimmed, NOT NUL-terminated. */
typedef struct pb_fasta pb_fasta;
int pb_fasta_open(const char* path, pb_fasta** out);
void pb_fasta_close(pb_fasta* f);
int64_t pb_fasta_num_records(pb_fasta* f);
int pb_fasta_record(pb_fasta* f, int64_t i, const char** name, int64_t* name_len, int64_t* seq_len);
int pb_fasta_pack_tile(pb_fasta* f, const int32_t* node_of_record, int64_t site_begin, int64_t n_sites_tile,
                       uint64_t* B0, uint64_t* B1, uint64_t* V, int64_t pitch_words, int32_t n_threads);
/* One pass from the FASTA text to the compact upload format of pb_put_planes_biallelic (V = NULL): X bit = the node's
 * character differs from the first mapped record's at that site, colmask = {first base, first XOR second base} codes.
 * Returns 1 when the tile qualifies — every node has exactly one mapped record covering the tile, every site carries at
 * most two distinct characters and both are upper-case A/C/G/T (so every genotype is valid) — else 0 (outputs undefined:
 * pack that tile with pb_fasta_pack_tile, optionally followed by pb_compact_planes).  No three-plane intermediate, no
 * compaction pass: per 32 sites one byte compare and one movemask. */
int pb_fasta_pack_tile_biallelic(pb_fasta* f, const int32_t* node_of_record, int64_t n_nodes, int64_t site_begin, int64_t n_sites_tile,
                                 uint64_t* X, int64_t pitch_words, uint64_t* colmask, int32_t n_threads);
/* the same from a character matrix (node-major, row v at seq + v*seq_pitch; the reference row is row 0) */
int pb_pack_chars_biallelic(const char* seq, int64_t n_nodes, int64_t n_sites, int64_t seq_pitch, uint64_t* X, int64_t x_pitch_words,
                            uint64_t* colmask, int32_t n_threads);
const char* pb_fasta_last_error(void);
/* Pinned host allocations for streaming uploads. */
int pb_host_alloc(void** out, uint64_t bytes);
int pb_host_free(void* p);
int pb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* POUTINE_B200_H */
