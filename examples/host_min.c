/* host_min.c — the C ABI of libpoutine_b200.so driven from plain C, no Python, no torch.
 *
 * The same call sequence a POUTINE host performs around the two replaced methods:
 *   tree -> pb_set_tree, phenos -> pb_set_labels, seg_sites chars -> pb_pack_chars -> pb_compact_planes ->
 *   pb_put_planes_biallelic | pb_put_planes, pb_run_events (HC.java:1402), pb_run_assoc (HC.java:2029), pb_run_extras.
 *
 *   gcc -std=c99 -Wall -Iinclude examples/host_min.c -o host_min -Lpoutine_b200 -lpoutine_b200 -Wl,-rpath,$PWD/poutine_b200
 *   ./host_min [n_replicates] [seed]
 *
 * Input: a 9-leaf tree (17 nodes: root 0 with three children, treetime-style; node 7 is unary) and 70 sites written out below; site 64 and
 * on exercise the second word.  Output: one line per biallelic site / informative site, machine-readable (tests/test_abi.py
 * compares it with the Python mirror's results for the same input).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "poutine_b200.h"

#define N_NODES 17
#define N_SITES 70

/* parent index per node (root = -1); leaves are the nodes without children */
static const int32_t PARENT[N_NODES] = {-1, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7};

static void die(pb_ctx* ctx, const char* what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, pb_last_error(ctx));
    exit(1);
}

int main(int argc, char** argv) {
    const int64_t m = argc > 1 ? atoll(argv[1]) : 1000;
    const uint64_t seed = argc > 2 ? (uint64_t)atoll(argv[2]) : 7;

    uint8_t is_leaf[N_NODES];
    int n_child[N_NODES] = {0};
    for (int v = 0; v < N_NODES; v++) if (PARENT[v] >= 0) n_child[PARENT[v]]++;
    for (int v = 0; v < N_NODES; v++) is_leaf[v] = n_child[v] == 0;

    /* seg_sites: node-major chars.  Deterministic pattern: every site biallelic (A/G or C/T), a few leaves flipped per site so
     * that independent mutations (homoplasies) occur; node 16 carries an 'N' at sites 5 and 66 (a V plane is then needed). */
    static char chars[N_NODES][N_SITES];
    for (int s = 0; s < N_SITES; s++) {
        const char a = (s & 1) ? 'C' : 'A', b = (s & 1) ? 'T' : 'G';
        for (int v = 0; v < N_NODES; v++) chars[v][s] = a;
        for (int v = 0; v < N_NODES; v++)
            if (is_leaf[v] && ((v * 7 + s * 3) % 5 == 0 || (v + s) % 11 == 0)) chars[v][s] = b;
        if (s % 9 == 4) { chars[4][s] = b; chars[10][s] = b; chars[11][s] = b; }   /* a clade carrying the second base */
    }
    chars[16][5] = 'N';
    chars[16][66] = 'N';

    /* phenos: every leaf has a row; label = case for odd node ids, one NA */
    int32_t sample_node[N_NODES];
    uint8_t label[N_NODES];
    int32_t P = 0;
    for (int v = 0; v < N_NODES; v++)
        if (is_leaf[v]) { sample_node[P] = v; label[P] = (v == 12) ? PB_LABEL_MISSING : (v & 1) ? PB_LABEL_CASE : PB_LABEL_CONTROL; P++; }

    pb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0; cfg.mode = PB_MODE_PHILOX; cfg.n_replicates = m; cfg.min_hcount = 1; cfg.seed = seed;
    pb_ctx* ctx = NULL;
    int rc = pb_create(&cfg, &ctx);
    if (rc) die(NULL, "pb_create", rc);
    if ((rc = pb_set_tree(ctx, N_NODES, PARENT, is_leaf))) die(ctx, "pb_set_tree", rc);
    if ((rc = pb_set_labels(ctx, P, sample_node, label))) die(ctx, "pb_set_labels", rc);
    if ((rc = pb_set_sites(ctx, N_SITES))) die(ctx, "pb_set_sites", rc);

    enum { W = (N_SITES + 63) / 64 };
    static uint64_t B0[N_NODES][W], B1[N_NODES][W], V[N_NODES][W], X[N_NODES][W], colmask[W][4];
    if ((rc = pb_pack_chars(&chars[0][0], N_NODES, N_SITES, N_SITES, &B0[0][0], &B1[0][0], &V[0][0], W))) die(ctx, "pb_pack_chars", rc);
    int32_t all_valid = 0;
    const int compact = pb_compact_planes(&B0[0][0], &B1[0][0], &V[0][0], N_NODES, N_SITES, W, &X[0][0], W, &colmask[0][0], &all_valid, 1);
    if (compact == 1) rc = pb_put_planes_biallelic(ctx, 0, W, W, &X[0][0], all_valid ? NULL : &V[0][0], &colmask[0][0]);
    else rc = pb_put_planes(ctx, 0, W, W, &B0[0][0], &B1[0][0], &V[0][0]);
    if (rc) die(ctx, "pb_put_planes*", rc);
    printf("upload compact=%d all_valid=%d\n", compact, (int)all_valid);

    static pb_site_out sites[N_SITES];
    pb_class_counts cc;
    if ((rc = pb_run_events(ctx, sites, &cc))) die(ctx, "pb_run_events", rc);
    printf("classes mono=%lld bi=%lld tri=%lld quad=%lld none=%lld raw_events=%lld\n", (long long)cc.n_mono, (long long)cc.n_bi,
           (long long)cc.n_tri, (long long)cc.n_quad, (long long)cc.n_no_allele, (long long)cc.n_raw_events);
    for (int s = 0; s < N_SITES; s++)
        if (sites[s].site_class == PB_SITE_BI)
            printf("site %d %c %c %d %d %d %d\n", sites[s].segsite_id, (char)sites[s].allele1, (char)sites[s].allele2, sites[s].a1_count,
                   sites[s].a2_count, sites[s].a1_extant, sites[s].a2_extant);

    static pb_stat_out stats[N_SITES];
    static double maxT[1 << 20];
    if (m > (1 << 20)) { fprintf(stderr, "n_replicates too large for this demo\n"); return 1; }
    int64_t n_inf = 0;
    rc = pb_run_assoc(ctx, NULL, stats, N_SITES, &n_inf, maxT);
    if (rc == PB_ERR_NO_INFORMATIVE) { printf("no informative sites\n"); pb_destroy(ctx); return 0; }
    if (rc) die(ctx, "pb_run_assoc", rc);
    static double exact[2 * N_SITES], fisher[N_SITES];
    if ((rc = pb_run_extras(ctx, exact, fisher))) die(ctx, "pb_run_extras", rc);
    for (int64_t i = 0; i < n_inf; i++)
        printf("stat %d [%d, %d, %d, %d] r=%d,%d rmaxT=%d,%d obs_p=%.17g,%.17g pointwise=%.17g,%.17g exact=%.17g,%.17g fisher=%.17g\n",
               stats[i].segsite_id, stats[i].obs_counts[0], stats[i].obs_counts[1], stats[i].obs_counts[2], stats[i].obs_counts[3],
               stats[i].r_a1, stats[i].r_a2, stats[i].r_maxT_a1, stats[i].r_maxT_a2, stats[i].obs_p_a1, stats[i].obs_p_a2,
               stats[i].pointwise_a1, stats[i].pointwise_a2, exact[2 * i], exact[2 * i + 1], fisher[i]);
    printf("maxT[0]=%.17g maxT[m-1]=%.17g\n", maxT[0], maxT[m - 1]);
    pb_timings tm;
    pb_get_timings(ctx, &tm);
    printf("plane_mode=%lld launches=%lld\n", (long long)tm.plane_mode, (long long)tm.n_launches);
    pb_destroy(ctx);
    return 0;
}
