// Host side of Homoplasy_Counter.call() around the two hot entry points (SURVEY §8f "next" rows), in C++:
// loaders with the reference's semantics and the three result files in the reference's format.
//
//     build_tree             HC.java:1144   -> parse_newick   (coevolution.jar NewickTreeTokenizer / NewickTreeReader)
//     nexus_to_newick        HC.java:866    -> nexus_to_newick
//     get_physical_positions HC.java:1207   -> read_map
//     read_phenos            HC.java:1972   -> read_phenos
//     output_significance_assessments_binom HC.java:2137 -> write_outputs
//
// build_seg_sites (HC.java:1241) is the mmap'ed FASTA packer inside the library (csrc/fasta.cpp, pb_fasta_*).
// Nothing here computes a statistic.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "poutine_b200.h"

namespace poutine {

// org.gersteinlab.coevolution.core.data.DataFormatException, and the reference's "malformed ... file" exits
struct DataFormatError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Tree {
    std::vector<int32_t> parent;        // -1 for the root (node 0: the outermost bracket pair)
    std::vector<uint8_t> is_leaf;
    std::vector<int32_t> leaf_nodes;    // NewickTree.buildCaches: left-to-right DFS order
    std::vector<float> dist_to_parent;  // parsed like the reference does (NaN when absent), never used by POUTINE
    std::vector<std::string> names;     // node ids; has_name[v] == 0 for an unnamed node
    std::vector<uint8_t> has_name;
    int64_t n_nodes() const { return (int64_t)parent.size(); }
};

// The reader's state machine (S node start, I id, C colon, D dist, E node end, F done); throws DataFormatError where
// the reference's reader throws DataFormatException.
Tree parse_newick(const std::string& text);

// HC.java:866-914: the line after "Begin Trees;" (equalsIgnoreCase), cut before the first '(', every "[...]" deleted.
std::string nexus_to_newick(const std::string& path);

// HC.java:1972-2019.  Rows in file order; a repeated id keeps its first position and its last label (HashMap.put).
struct Phenos {
    std::vector<std::string> ids;
    std::vector<std::string> labels;
    int64_t cases = 0, controls = 0, missing = 0;   // tot_num_sample_cases / controls (HC.java:1999-2007)
};
Phenos read_phenos(const std::string& path);

// HC.java:1207-1235: Integer.parseInt of the last tab-separated column of every row.
std::vector<int32_t> read_map(const std::string& path);

// java.lang.Double.toString (shortest digits that round-trip, JDK >= 19)
std::string java_double_str(double x);

// Homoplasy_Events.toString + "\t" + Binomial_Test_Stat.toString (HC.java:5574-5578, 5918-5922)
std::string format_row(const pb_site_out& e, const pb_stat_out& t, int32_t physical_pos);

// <out>, <out>.sorted_by_a2_maxT, <out>.sorted_by_a1_maxT (HC.java:2137-2210): the a1 file is a STABLE sort of the
// a2-sorted rows (the reference sorts the same array twice), rows whose key is NaN are not written.
// sites = the per-site records of this context (indexed by pb_stat_out.site_index).
void write_outputs(const std::string& out_path, const pb_site_out* sites, const std::vector<pb_stat_out>& stats,
                   const std::vector<int32_t>& physical_pos);

std::string read_file(const std::string& path);
// BufferedReader.readLine: lines end at \n, \r or \r\n; no empty last line after a final terminator
std::vector<std::string> java_lines(const std::string& text);
// String.split("\t"): trailing empty strings removed
std::vector<std::string> java_split_tabs(const std::string& line);

}  // namespace poutine
