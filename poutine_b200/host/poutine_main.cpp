// poutine_b200_host — the reference's command line for its --use-precomputed-anc-recon mode, as a native host around
// libpoutine_b200.so:
//
//     poutine_b200_host -u -f anc.fasta -t anc.newick|annotated_tree.nexus -p pheno.tsv -m sites.map [-r m] [-c min_hcount]
//                       [-T threads] [-d out_dir] [-o out] [-X] [--seed s] [--devices d0[,d1,...]] [--extras]
//
// = Homoplasy_Counter.call() (HC.java:230-319) with the picocli flags of HC.java:59-154: loaders (hostio.cpp), the mmap'ed
// FASTA packer streamed tile by tile into pb_put_planes / pb_put_planes_biallelic, the two hot entry points through
// HomoplasyCounter, the three result files.  The non -u mode (running treetime, HC.java:817-857) stays with the Java host.
// Exit status 255 on error like the reference's System.exit(-1).
//
// The "--x-..." sub-commands print what the loaders / formatter produce; tests/test_host_cpp.py compares them with the
// Python host (poutine_b200/hostio.py) — no GPU involved.
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "homoplasy_counter.hpp"
#include "hostio.hpp"

using namespace poutine;

namespace {

struct Args {
    std::string fasta, phenos, tree, map, out_dir, out;
    int64_t replicates = 100000;
    int32_t min_hcount = 1, threads = 0;
    std::vector<int32_t> devices{0};     // --devices d0[,d1,...]: the sites are sharded over these GPUs (pb_group), all from this process
    bool precomputed = false, overwrite = false, extras = false, have_seed = false;
    uint64_t seed = 0;
    int64_t tile_sites = 1 << 14;   // profiles/r1_native_host_sweep.json: page-locked buffer set-up vs per-tile overhead
    bool compact_upload = true;
    bool timing = false;       // --timing: phase wall clocks on stderr
};

[[noreturn]] void usage(const char* why) {
    if (why) std::fprintf(stderr, "poutine_b200_host: %s\n", why);
    std::fprintf(stderr,
                 "usage: poutine_b200_host -u -f <ancestral fasta> -t <ancestral newick | annotated_tree.nexus> -p <phenos> -m <map>\n"
                 "       [-r <# replicates>=100000] [-c <min_hcount>=1] [-T <# packer threads>] [-d <out dir>] [-o <out file>] [-X]\n"
                 "       [--seed <n>] [--devices <ordinal>[,<ordinal>...]] [--extras] [--tile-sites <n>=16384] [--no-compact-upload] [--timing]\n");
    std::exit(2);
}

bool file_exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}

void make_dirs(const std::string& p) {
    for (size_t i = 1; i <= p.size(); i++)
        if (i == p.size() || p[i] == '/') {
            std::string sub = p.substr(0, i);
            if (!sub.empty() && !file_exists(sub) && ::mkdir(sub.c_str(), 0777) != 0 && errno != EEXIST)
                throw DataFormatError("cannot create directory " + sub + ": " + std::strerror(errno));
        }
}

struct FastaFile {     // build_seg_sites (HC.java:1241-1306) source: indexed once, packed per column tile
    pb_fasta* h = nullptr;
    std::vector<std::string> names;
    std::vector<int64_t> lengths;
    explicit FastaFile(const std::string& path) {
        if (pb_fasta_open(path.c_str(), &h) != 0) throw DataFormatError(pb_fasta_last_error());
        int64_t n = pb_fasta_num_records(h);
        for (int64_t i = 0; i < n; i++) {
            const char* name = nullptr;
            int64_t nlen = 0, slen = 0;
            pb_fasta_record(h, i, &name, &nlen, &slen);
            names.emplace_back(name ? name : "", (size_t)nlen);
            lengths.push_back(slen);
        }
    }
    ~FastaFile() {
        if (h) pb_fasta_close(h);
    }
};

struct RunResult {
    pb_class_counts counts{};
    int64_t n_informative = 0, n_sites = 0, n_tiles = 0, n_compact_tiles = 0;
};

struct PhaseClock {
    bool on;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        auto now = std::chrono::steady_clock::now();
        if (on) std::fprintf(stderr, "[timing] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// The part of call() in front of the two entry points (HC.java:230-296): loaders + cross-file checks, flattened to what the
// C ABI takes.
struct Inputs {
    Tree tree;
    std::unique_ptr<FastaFile> fa;
    std::vector<int32_t> physical_pos, node_of_record, sample_node;
    std::vector<uint8_t> label;
    int64_t S = 0;
};

Inputs load_inputs(const Args& a) {
    Inputs in;
    const std::string text = read_file(a.tree);
    size_t k = 0;
    while (k < text.size() && std::isspace((unsigned char)text[k])) k++;
    const bool nexus = text.size() >= k + 6 && ::strncasecmp(text.c_str() + k, "#NEXUS", 6) == 0;
    in.tree = parse_newick(nexus ? nexus_to_newick(a.tree) : text);
    const Tree& tree = in.tree;
    const int64_t N = tree.n_nodes();
    std::unordered_map<std::string, int32_t> node_of_name;
    for (int64_t v = 0; v < N; v++) {
        if (!tree.has_name[v])
            throw DataFormatError("every tree node needs a label (treetime writes NODE_xxxxxxx for internal nodes)");
        if (!node_of_name.emplace(tree.names[v], (int32_t)v).second) throw DataFormatError("duplicate node label '" + tree.names[v] + "'");
    }
    in.fa.reset(new FastaFile(a.fasta));
    const FastaFile& fa = *in.fa;
    in.physical_pos = read_map(a.map);
    // The reference loops over the map rows (HC.java:1418) and reads snps[i] of EVERY record (HC.java:1383-1394): a shorter
    // record is an ArrayIndexOutOfBoundsException there, and check_num_seg_sites (HC.java:327-365) exits unless the first
    // record in its HashMap's iteration order has exactly as many sites as the map file has rows.  Here: every record must.
    const int64_t S = in.S = (int64_t)in.physical_pos.size();
    for (size_t i = 0; i < fa.names.size(); i++)
        if (fa.lengths[i] != S)
            throw DataFormatError("# sites of FASTA record '" + fa.names[i] + "' (" + std::to_string(fa.lengths[i]) +
                                  ") != # rows in the map file (" + std::to_string(S) + ")");
    in.node_of_record.assign(fa.names.size(), -1);
    std::vector<uint8_t> seen((size_t)N, 0);
    std::unordered_map<std::string, size_t> last_record;
    for (size_t i = 0; i < fa.names.size(); i++) {
        auto dup = last_record.find(fa.names[i]);           // seg_sites.put(node_name, seq): a repeated name keeps its LAST record
        if (dup != last_record.end()) in.node_of_record[dup->second] = -1;
        last_record[fa.names[i]] = i;
        auto it = node_of_name.find(fa.names[i]);
        if (it != node_of_name.end()) {
            in.node_of_record[i] = it->second;
            seen[it->second] = 1;
        }
    }
    for (int64_t v = 0; v < N; v++)
        if (!seen[v]) throw DataFormatError("tree node '" + tree.names[v] + "' has no FASTA record");
    Phenos ph = read_phenos(a.phenos);
    std::unordered_set<std::string> leaf_names;
    for (int32_t v : tree.leaf_nodes) leaf_names.insert(tree.names[v]);
    for (size_t i = 0; i < ph.ids.size(); i++) {                      // check_sample_names (HC.java:367-404)
        auto it = node_of_name.find(ph.ids[i]);
        if (it == node_of_name.end() || !leaf_names.count(ph.ids[i]))
            throw DataFormatError("phenotype sample '" + ph.ids[i] + "' is not a leaf of the tree / not in the FASTA");
        in.sample_node.push_back(it->second);
        in.label.push_back(ph.labels[i] == "1" ? PB_LABEL_CASE : ph.labels[i] == "0" ? PB_LABEL_CONTROL : PB_LABEL_MISSING);
    }
    return in;
}

RunResult run_homoplasy_counter(const Args& a, const std::string& out_path) {
    PhaseClock clk{a.timing};
    Inputs in = load_inputs(a);
    const Tree& tree = in.tree;
    const FastaFile& fa = *in.fa;
    const std::vector<int32_t>&physical_pos = in.physical_pos, &node_of_record = in.node_of_record, &sample_node = in.sample_node;
    const std::vector<uint8_t>& label = in.label;
    const int64_t N = tree.n_nodes(), S = in.S;

    clk.lap("loaders + fasta index");
    Options o;
    o.num_replicates = a.replicates;
    o.min_hcount = a.min_hcount;
    o.devices = a.devices;
    o.seed = a.seed;
    HomoplasyCounter hc(o);
    hc.set_tree(tree.parent, tree.is_leaf);
    hc.set_phenos(sample_node, label);
    hc.set_sites(S);
    clk.lap("context, tree, phenos, planes");

    // stream the alignment: tile t+1 is packed (and compacted, when all of its sites carry at most two bases) on a worker
    // thread while tile t uploads from the other page-locked buffer set
    const int64_t tile_sites = std::max<int64_t>(64, std::min(a.tile_sites, S + 63) / 64 * 64);
    const int64_t tw = tile_sites / 64;
    struct Buf {
        PinnedArray<uint64_t> B0, B1, V, X;
        std::vector<uint64_t> colmask;
        int compact = 0;
        int32_t all_valid = 0;
        std::string error;
    } bufs[2];
    for (Buf& b : bufs) {
        b.B0.reset((size_t)(N * tw));
        b.B1.reset((size_t)(N * tw));
        b.V.reset((size_t)(N * tw));
        b.X.reset((size_t)(N * tw));
        b.colmask.resize((size_t)(4 * tw));
    }
    const int64_t n_tiles = (S + tile_sites - 1) / tile_sites;
    clk.lap("page-locked tile buffers");
    auto pack = [&](int64_t t) {
        Buf& b = bufs[t & 1];
        const int64_t s0 = t * tile_sites, ns = std::min(tile_sites, S - s0);
        b.error.clear();
        b.compact = 0;
        if (a.compact_upload) {
            // the usual alignment (every genotype a base, at most two bases per site): text -> compact format in one pass
            const int rc = pb_fasta_pack_tile_biallelic(fa.h, node_of_record.data(), N, s0, ns, b.X.data(), tw, b.colmask.data(), a.threads);
            if (rc < 0) { b.error = pb_fasta_last_error(); return; }
            if (rc == 1) { b.compact = 1; b.all_valid = 1; return; }
        }
        if (pb_fasta_pack_tile(fa.h, node_of_record.data(), s0, ns, b.B0.data(), b.B1.data(), b.V.data(), tw, a.threads) != 0) {
            b.error = pb_fasta_last_error();
            return;
        }
        b.compact = 0;
        if (a.compact_upload) {
            int rc = pb_compact_planes(b.B0.data(), b.B1.data(), b.V.data(), N, ns, tw, b.X.data(), tw, b.colmask.data(), &b.all_valid,
                                       a.threads);
            if (rc < 0) b.error = "pb_compact_planes: invalid arguments";
            b.compact = rc == 1;
        }
    };
    RunResult res;
    res.n_sites = S;
    res.n_tiles = n_tiles;
    if (n_tiles) pack(0);
    for (int64_t t = 0; t < n_tiles; t++) {
        std::thread worker;
        if (t + 1 < n_tiles) worker = std::thread(pack, t + 1);
        Buf& b = bufs[t & 1];
        const int64_t s0 = t * tile_sites, ns = std::min(tile_sites, S - s0), nw = (ns + 63) / 64;
        try {
            if (!b.error.empty()) throw DataFormatError(b.error);
            if (b.compact) {
                hc.put_planes_biallelic(s0 / 64, nw, tw, b.X.data(), b.all_valid ? nullptr : b.V.data(), b.colmask.data());
                res.n_compact_tiles++;
            } else {
                hc.put_planes(s0 / 64, nw, tw, b.B0.data(), b.B1.data(), b.V.data());
            }
        } catch (...) {
            if (worker.joinable()) worker.join();
            throw;
        }
        if (worker.joinable()) worker.join();
    }
    clk.lap("pack + upload (pipelined)");
    const pb_site_out* all_events = hc.count_all_homoplasy_events(&res.counts);     // HC.java:299
    clk.lap("count_all_homoplasy_events");
    AssocResult as = hc.assoc();                                                    // HC.java:314
    clk.lap("assoc");
    res.n_informative = (int64_t)as.stats.size();
    write_outputs(out_path, all_events, as.stats, physical_pos);
    clk.lap("write_outputs");
    if (a.extras) {      // not a reference file: exact point-wise null + two-sided Fisher per informative site
        std::vector<double> ep, fp;
        hc.extras(ep, fp);
        FILE* f = std::fopen((out_path + ".extras").c_str(), "wb");
        if (!f) throw DataFormatError("cannot write " + out_path + ".extras");
        std::fputs("segsite_ID\tpos\texact_pointwise_a1\texact_pointwise_a2\tfisher_two_sided\n", f);
        for (size_t i = 0; i < as.stats.size(); i++)
            std::fprintf(f, "%d\t%d\t%s\t%s\t%s\n", as.stats[i].segsite_id, physical_pos[(size_t)as.stats[i].site_index],
                         java_double_str(ep[2 * i]).c_str(), java_double_str(ep[2 * i + 1]).c_str(), java_double_str(fp[i]).c_str());
        std::fclose(f);
    }
    return res;
}

// ---- loader / formatter probes for the CPU tests ---------------------------------------------------------------------
int probe(int argc, char** argv) {
    const std::string cmd = argv[1];
    if (cmd == "--x-parse-tree" && argc == 3) {
        Tree t = parse_newick(read_file(argv[2]));
        std::printf("%" PRId64 "\n", t.n_nodes());
        for (int64_t v = 0; v < t.n_nodes(); v++) {
            uint32_t bits;
            std::memcpy(&bits, &t.dist_to_parent[v], 4);
            std::printf("%d\t%d\t%08x\t%d\t%s\n", t.parent[v], (int)t.is_leaf[v], bits, (int)t.has_name[v], t.names[v].c_str());
        }
        for (int32_t v : t.leaf_nodes) std::printf("%d ", v);
        std::printf("\n");
        return 0;
    }
    if (cmd == "--x-nexus-to-newick" && argc == 3) {
        std::printf("%s\n", nexus_to_newick(argv[2]).c_str());
        return 0;
    }
    if (cmd == "--x-read-phenos" && argc == 3) {
        Phenos p = read_phenos(argv[2]);
        std::printf("%" PRId64 "\t%" PRId64 "\t%" PRId64 "\n", p.cases, p.controls, p.missing);
        for (size_t i = 0; i < p.ids.size(); i++) std::printf("%s\t%s\n", p.ids[i].c_str(), p.labels[i].c_str());
        return 0;
    }
    if (cmd == "--x-read-map" && argc == 3) {
        for (int32_t p : read_map(argv[2])) std::printf("%d\n", p);
        return 0;
    }
    if (cmd == "--x-flatten" && argc == 6) {             // fasta tree phenos map -> what the C ABI would be given
        Args a;
        a.fasta = argv[2]; a.tree = argv[3]; a.phenos = argv[4]; a.map = argv[5];
        Inputs in = load_inputs(a);
        auto dump = [](const char* what, const auto& v) {
            std::printf("%s", what);
            for (auto x : v) std::printf(" %d", (int)x);
            std::printf("\n");
        };
        std::printf("n_nodes %" PRId64 " n_sites %" PRId64 "\n", in.tree.n_nodes(), in.S);
        dump("parent", in.tree.parent);
        dump("is_leaf", in.tree.is_leaf);
        dump("node_of_record", in.node_of_record);
        dump("sample_node", in.sample_node);
        dump("label", in.label);
        dump("physical_pos", in.physical_pos);
        return 0;
    }
    if (cmd == "--x-format-doubles" && argc == 2) {      // stdin: one 64-bit pattern in hex per line
        char line[64];
        while (std::fgets(line, sizeof line, stdin)) {
            uint64_t bits = std::strtoull(line, nullptr, 16);
            double x;
            std::memcpy(&x, &bits, 8);
            std::printf("%s\n", java_double_str(x).c_str());
        }
        return 0;
    }
    if (cmd == "--x-write-records" && argc == 6) {       // sites.bin stats.bin map out
        std::string sb = read_file(argv[2]), tb = read_file(argv[3]);
        if (sb.size() % sizeof(pb_site_out) || tb.size() % sizeof(pb_stat_out)) throw DataFormatError("record files: bad size");
        std::vector<pb_site_out> sites(sb.size() / sizeof(pb_site_out));
        std::vector<pb_stat_out> stats(tb.size() / sizeof(pb_stat_out));
        std::memcpy(sites.data(), sb.data(), sb.size());
        std::memcpy(stats.data(), tb.data(), tb.size());
        for (const pb_stat_out& t : stats)
            if (t.site_index < 0 || (size_t)t.site_index >= sites.size()) throw DataFormatError("record files: site_index out of range");
        write_outputs(argv[5], sites.data(), stats, read_map(argv[4]));
        return 0;
    }
    usage("unknown probe");
}

}  // namespace

int main(int argc, char** argv) {
    try {
        if (argc >= 2 && std::strncmp(argv[1], "--x-", 4) == 0) return probe(argc, argv);
        Args a;
        for (int i = 1; i < argc; i++) {
            const std::string f = argv[i];
            auto val = [&]() -> const char* {
                if (i + 1 >= argc) usage(("missing value for " + f).c_str());
                return argv[++i];
            };
            if (f == "-f" || f == "--fasta") a.fasta = val();
            else if (f == "-p" || f == "--phenos") a.phenos = val();
            else if (f == "-t" || f == "--tree") a.tree = val();
            else if (f == "-m" || f == "--map") a.map = val();
            else if (f == "-r" || f == "--replicates") a.replicates = std::atoll(val());
            else if (f == "-c" || f == "--min_hcount") a.min_hcount = std::atoi(val());
            else if (f == "-u" || f == "--use-precomputed-anc-recon") a.precomputed = true;
            else if (f == "-T" || f == "--threads") a.threads = std::atoi(val());
            else if (f == "-d" || f == "--out-dir") a.out_dir = val();
            else if (f == "-o" || f == "--out") a.out = val();
            else if (f == "-X" || f == "--force-overwrite") a.overwrite = true;
            else if (f == "--seed") { a.seed = std::strtoull(val(), nullptr, 10); a.have_seed = true; }
            else if (f == "--device" || f == "--devices") {      // one site shard per listed GPU, all driven from this one process
                a.devices.clear();
                for (const char* p = val(); *p;) {
                    char* end = nullptr;
                    const long d = std::strtol(p, &end, 10);
                    if (end == p || d < 0 || (*end && *end != ',')) usage("--devices expects a comma-separated list of CUDA ordinals");
                    a.devices.push_back((int32_t)d);
                    p = *end == ',' ? end + 1 : end;
                }
                if (a.devices.empty()) usage("--devices expects a comma-separated list of CUDA ordinals");
            }
            else if (f == "--extras") a.extras = true;
            else if (f == "--tile-sites") a.tile_sites = std::atoll(val());
            else if (f == "--no-compact-upload") a.compact_upload = false;
            else if (f == "--timing") a.timing = true;
            else if (f == "-h" || f == "--help") usage(nullptr);
            else usage(("unknown option " + f).c_str());
        }
        if (a.fasta.empty() || a.phenos.empty() || a.tree.empty() || a.map.empty()) usage("-f, -p, -t and -m are required");
        if (!a.precomputed)
            usage("only --use-precomputed-anc-recon (-u) mode is supported: ancestral reconstruction stays with the Java host / treetime");
        char stamp[32];
        std::time_t now = std::time(nullptr);
        std::strftime(stamp, sizeof stamp, "%Y_%m_%d_%H_%M_%S", std::localtime(&now));
        const std::string session = std::string("poutine_session_") + stamp;
        const std::string out_dir = a.out_dir.empty() ? session : a.out_dir;
        make_dirs(out_dir);
        const std::string out = out_dir + "/" + (a.out.empty() ? session + ".out" : a.out);
        if (file_exists(out) && !a.overwrite) {
            std::fprintf(stderr, "output file %s exists (use -X to overwrite)\n", out.c_str());
            return 1;
        }
        if (!a.have_seed)      // the reference's RNG is unseeded (HC.java:4911)
            a.seed = (uint64_t)std::chrono::system_clock::now().time_since_epoch().count() & 0xFFFFFFFFFFFFull;
        auto t0 = std::chrono::steady_clock::now();
        RunResult res;
        try {
            res = run_homoplasy_counter(a, out);
        } catch (const NoInformativeSites&) {
            std::printf("min_hcount = %d is set too high and has filtered out all segregating sites\n", a.min_hcount);   // HC.java:3866-3869
            return 255;
        }
        std::printf("# bi-allelic sites = %" PRId64 ", monomorphic = %" PRId64 ", tri-allelic = %" PRId64 ", quad-allelic = %" PRId64 "\n",
                    res.counts.n_bi, res.counts.n_mono, res.counts.n_tri, res.counts.n_quad);
        std::printf("# homoplasically informative sites = %" PRId64 "\n", res.n_informative);
        std::printf("results: %s (+ .sorted_by_a1_maxT, .sorted_by_a2_maxT); %" PRId64 " tiles (%" PRId64 " compact); wall clock %.2f s\n",
                    out.c_str(), res.n_tiles, res.n_compact_tiles,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 255;
    }
}
