// Loaders and result writer of the C++ host (see hostio.hpp).  Reference lines are cited per function.
#include "hostio.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <unordered_map>

namespace poutine {

std::string read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw DataFormatError("cannot open " + path + ": " + std::strerror(errno));
    std::string s;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    std::fclose(f);
    return s;
}

std::vector<std::string> java_lines(const std::string& t) {
    std::vector<std::string> out;
    size_t i = 0, n = t.size();
    while (i < n) {
        size_t j = i;
        while (j < n && t[j] != '\n' && t[j] != '\r') j++;
        out.emplace_back(t, i, j - i);
        if (j < n && t[j] == '\r' && j + 1 < n && t[j + 1] == '\n') j++;
        i = j + 1;
    }
    return out;
}

std::vector<std::string> java_split_tabs(const std::string& line) {
    std::vector<std::string> cols;
    size_t i = 0;
    for (;;) {
        size_t j = line.find('\t', i);
        if (j == std::string::npos) {
            cols.emplace_back(line, i);
            break;
        }
        cols.emplace_back(line, i, j - i);
        i = j + 1;
    }
    if (cols.size() > 1 || !line.empty())          // "" splits into [""]; otherwise trailing empty strings are dropped
        while (!cols.empty() && cols.back().empty()) cols.pop_back();
    return cols;
}

// -----------------------------------------------------------------------------------------------------------------
// Newick: NewickTreeTokenizer.nextToken / NewickTreeReader.readTree (coevolution.jar, read from the bytecode)
// -----------------------------------------------------------------------------------------------------------------
namespace {

inline bool is_delim(char c) { return c == '(' || c == ')' || c == ',' || c == ':' || c == ';'; }
// Character.isWhitespace over the byte range: \t \n VT FF \r, FS GS RS US, space (0xA0 is not whitespace in Java)
inline bool is_java_ws(unsigned char c) { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); }

struct Tokenizer {
    const std::string& s;
    size_t i = 0;
    explicit Tokenizer(const std::string& text) : s(text) {}
    // false at end of data; a delimiter is a token of its own, whitespace ends a token and is never part of one
    bool next(std::string& tok) {
        while (i < s.size() && is_java_ws((unsigned char)s[i])) i++;
        if (i >= s.size()) return false;
        if (is_delim(s[i])) {
            tok.assign(1, s[i++]);
            return true;
        }
        size_t j = i;
        while (j < s.size() && !is_delim(s[j]) && !is_java_ws((unsigned char)s[j])) j++;
        tok.assign(s, i, j - i);
        i = j;
        return true;
    }
};

// NewickTreeReader.reportFormatError's message
[[noreturn]] void newick_error(char state, const std::string& tok) {
    const char* name = state == 'S' ? "NODE_START" : state == 'I' ? "ID" : state == 'C' ? "COLON" : state == 'D' ? "DIST"
                     : state == 'E' ? "NODE_END" : "SEMI_COLON";
    throw DataFormatError("Unexpected token [" + tok + "] encountered in the " + name + " state.");
}

// Float.parseFloat on a token (no blanks inside): optional sign, decimal digits / '.' / exponent, optional f|F|d|D suffix;
// "NaN" / "Infinity".  Anything else is the reference's NumberFormatException (which build_tree, HC.java:1155, does not
// catch: the run dies with that exception).
float parse_java_float(const std::string& tok) {
    std::string t = tok;
    if (t.size() > 1 && (t.back() == 'f' || t.back() == 'F' || t.back() == 'd' || t.back() == 'D')) t.pop_back();
    bool ok = !t.empty() && t != "+" && t != "-";
    for (char c : t) ok = ok && ((c >= '0' && c <= '9') || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-');
    const size_t sign = !t.empty() && (t[0] == '+' || t[0] == '-') ? 1 : 0;
    if (t.compare(sign, std::string::npos, "NaN") == 0) return std::nanf("");
    if (t.compare(sign, std::string::npos, "Infinity") == 0) return t[0] == '-' ? -HUGE_VALF : HUGE_VALF;
    if (ok) {
        char* end = nullptr;
        double v = std::strtod(t.c_str(), &end);
        if (end == t.c_str() + t.size()) return (float)v;
    }
    throw DataFormatError("For input string: \"" + tok + "\"");
}

}  // namespace

Tree parse_newick(const std::string& text) {
    Tree tr;
    auto new_node = [&](int32_t par) {
        tr.parent.push_back(par);
        tr.names.emplace_back();
        tr.has_name.push_back(0);
        tr.dist_to_parent.push_back(std::nanf(""));     // NewickTreeNode.<init>: Float.NaN until a ':dist' is read
        return (int32_t)tr.parent.size() - 1;
    };
    auto set_id = [&](int32_t v, const std::string& id) {
        tr.names[v] = id;
        tr.has_name[v] = 1;
    };
    Tokenizer tk(text);
    std::string tok;
    char state = 'S';
    int32_t cur = -1;
    while (state != 'F') {
        if (!tk.next(tok)) throw DataFormatError("Unexpected end of data.");
        const bool delim = tok.size() == 1 && is_delim(tok[0]);
        const char c = delim ? tok[0] : 0;
        switch (state) {
        case 'S':
            if (c == '(') {
                cur = new_node(cur);
            } else if (c == ')') {
                if (cur < 0) newick_error(state, tok);
                new_node(cur);                       // empty last child
                state = 'E';
            } else if (c == ',') {
                if (cur < 0) newick_error(state, tok);
                new_node(cur);                       // empty child
            } else if (c == ':') {
                cur = new_node(cur);
                state = 'C';
            } else if (!delim) {
                cur = new_node(cur);
                set_id(cur, tok);
                state = 'I';
            } else {
                newick_error(state, tok);
            }
            break;
        case 'I':
        case 'D':
        case 'E':
            if (c == ')') {
                cur = tr.parent[cur];
                if (cur < 0) newick_error(state, tok);
                state = 'E';
            } else if (c == ',') {
                cur = tr.parent[cur];
                if (cur < 0) newick_error(state, tok);
                state = 'S';
            } else if (c == ':' && state != 'D') {
                state = 'C';
            } else if (c == ';') {
                state = 'F';
            } else if (state == 'E' && !delim) {
                set_id(cur, tok);                    // the id of an internal node follows its ')'
                state = 'I';
            } else {
                newick_error(state, tok);
            }
            break;
        case 'C':
            if (delim) newick_error(state, tok);
            tr.dist_to_parent[cur] = parse_java_float(tok);
            state = 'D';
            break;
        }
    }
    if (cur < 0) throw DataFormatError("More close brackets then open brackets");
    if (tr.parent[cur] >= 0) throw DataFormatError("More open brackets then close brackets");
    const size_t n = tr.parent.size();
    tr.is_leaf.assign(n, 1);
    for (size_t v = 0; v < n; v++)
        if (tr.parent[v] >= 0) tr.is_leaf[tr.parent[v]] = 0;
    // children are created left to right, depth first: creation order of the childless nodes is the DFS leaf order
    for (size_t v = 0; v < n; v++)
        if (tr.is_leaf[v]) tr.leaf_nodes.push_back((int32_t)v);
    return tr;
}

std::string nexus_to_newick(const std::string& path) {
    std::vector<std::string> lines = java_lines(read_file(path));
    size_t i = 0;
    auto equals_ignore_case = [](const std::string& a, const char* b) {
        size_t n = std::strlen(b);
        if (a.size() != n) return false;
        for (size_t k = 0; k < n; k++)
            if (std::tolower((unsigned char)a[k]) != std::tolower((unsigned char)b[k])) return false;
        return true;
    };
    while (i < lines.size() && !equals_ignore_case(lines[i], "Begin Trees;")) i++;
    if (i + 1 >= lines.size())      // the reference dies here: NullPointerException on readLine() == null
        throw DataFormatError("no tree line after 'Begin Trees;' in " + path);
    std::string s = lines[i + 1];
    size_t k = s.find('(');
    if (k == std::string::npos)     // StringBuilder.delete(0, -1) throws StringIndexOutOfBoundsException
        throw DataFormatError("the tree line of " + path + " holds no '('");
    s.erase(0, k);
    for (;;) {
        size_t a = s.find('[');
        if (a == std::string::npos) break;
        size_t b = s.find(']');
        if (b == std::string::npos || b < a)   // delete(start, end + 1) with end + 1 < start throws in the reference
            throw DataFormatError("unbalanced '[' ... ']' annotation in the tree line of " + path);
        s.erase(a, b + 1 - a);
    }
    return s;
}

// -----------------------------------------------------------------------------------------------------------------
// pheno / map files
// -----------------------------------------------------------------------------------------------------------------
Phenos read_phenos(const std::string& path) {
    Phenos ph;
    std::unordered_map<std::string, size_t> pos;
    for (const std::string& line : java_lines(read_file(path))) {
        std::vector<std::string> cols = java_split_tabs(line);
        if (cols.size() < 2)      // ArrayIndexOutOfBoundsException -> "malformed phenotype file" exit (HC.java:1983-1987)
            throw DataFormatError("malformed phenotype file " + path + ": a row has fewer than two tab-separated columns");
        auto it = pos.find(cols[0]);
        if (it == pos.end()) {
            pos.emplace(cols[0], ph.ids.size());
            ph.ids.push_back(cols[0]);
            ph.labels.push_back(cols[1]);
        } else {
            ph.labels[it->second] = cols[1];
        }
    }
    for (const std::string& l : ph.labels) {
        if (l == "1") ph.cases++;
        else if (l == "0") ph.controls++;
        else ph.missing++;
    }
    return ph;
}

namespace {
// Integer.parseInt: [+-]?[0-9]+ within int32, nothing else (no blanks)
bool parse_java_int(const std::string& s, int32_t& out) {
    size_t i = 0;
    bool neg = false;
    if (i < s.size() && (s[i] == '-' || s[i] == '+')) neg = s[i++] == '-';
    if (i >= s.size()) return false;
    int64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > (int64_t)1 << 31) return false;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return false;
    out = (int32_t)v;
    return true;
}
}  // namespace

std::vector<int32_t> read_map(const std::string& path) {
    std::vector<int32_t> out;
    for (const std::string& line : java_lines(read_file(path))) {
        std::vector<std::string> cols = java_split_tabs(line);
        int32_t p;
        if (cols.empty() || !parse_java_int(cols.back(), p))     // -> "malformed map file" exit (HC.java:1224-1228)
            throw DataFormatError("malformed map file " + path + ": '" + (cols.empty() ? std::string() : cols.back()) +
                                  "' is not an integer position");
        out.push_back(p);
    }
    return out;
}

// -----------------------------------------------------------------------------------------------------------------
// result files
// -----------------------------------------------------------------------------------------------------------------
std::string java_double_str(double x) {
    if (x != x) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    std::string out = x < 0 ? "-" : "";
    const double ax = std::fabs(x);
    // shortest digits that round-trip, as d[.ddd]e[+-]XX
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, ax, std::chars_format::scientific);
    std::string sci(buf, res.ptr);
    size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; i++)
        if (sci[i] != '.') digits.push_back(sci[i]);
    int e10 = std::atoi(sci.c_str() + epos + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    if (digits.size() == 1) {
        // Double.toString renders at least two digits and picks the two-digit decimal closest to x; that differs from
        // "shortest digit + 0" only for subnormals with a few bits (Double.MIN_VALUE prints as 4.9E-324)
        char two[32];
        std::snprintf(two, sizeof two, "%.1e", ax);
        if (std::strtod(two, nullptr) == ax) {
            digits.assign(1, two[0]);
            if (two[2] != '0') digits.push_back(two[2]);
            e10 = std::atoi(std::strchr(two, 'e') + 1);
        }
    }
    if (ax >= 1e-3 && ax < 1e7) {
        if (e10 >= 0) {
            std::string ipart = digits.substr(0, std::min(digits.size(), (size_t)e10 + 1));
            ipart.append((size_t)e10 + 1 - ipart.size(), '0');
            std::string fpart = digits.size() > (size_t)e10 + 1 ? digits.substr((size_t)e10 + 1) : "0";
            return out + ipart + "." + fpart;
        }
        return out + "0." + std::string((size_t)(-e10 - 1), '0') + digits;
    }
    return out + digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(e10);
}

std::string format_row(const pb_site_out& e, const pb_stat_out& t, int32_t physical_pos) {
    std::string r;
    r.reserve(256);
    auto add_int = [&](long long v) { r += std::to_string(v); r += '\t'; };
    add_int(e.segsite_id);
    add_int(physical_pos);
    r += (char)e.allele1; r += '\t';
    r += (char)e.allele2; r += '\t';
    add_int(e.a1_count);
    add_int(e.a2_count);
    add_int(e.a1_extant);
    add_int(e.a2_extant);
    r += "[" + std::to_string(t.obs_counts[0]) + ", " + std::to_string(t.obs_counts[1]) + ", " + std::to_string(t.obs_counts[2]) +
         ", " + std::to_string(t.obs_counts[3]) + "]\t";                                     // Arrays.toString(int[])
    add_int(t.r_a1);
    add_int(t.r_a2);
    r += java_double_str(t.pointwise_a1); r += '\t';
    r += java_double_str(t.pointwise_a2); r += '\t';
    r += java_double_str(t.obs_p_a1); r += '\t';
    r += java_double_str(t.obs_p_a2); r += '\t';
    add_int(t.r_maxT_a1);
    add_int(t.r_maxT_a2);
    r += java_double_str(t.familywise_a1); r += '\t';
    r += java_double_str(t.familywise_a2);
    return r;
}

namespace {
const char* const kHeader =
    "segsite_ID\tphysical_pos\tallele1\tallele2\ta1_count\ta2_count\ta1_count_extant_only\ta2_count_extant_only\tobs_homoplasy_counts\t"
    "r_a1\tr_a2\tpointwise_pvalue_a1\tpointwise_pvalue_a2\tobs_binom_pvalue_a1\tobs_binom_pvalue_a2\tr_maxT_a1\tr_maxT_a2\t"
    "familywise_pvalue_a1\tfamilywise_pvalue_a2";

// Double.compare: -0.0 below 0.0, NaN above everything (and equal to itself)
bool java_double_less(double a, double b) {
    const bool an = a != a, bn = b != b;
    if (an || bn) return !an && bn;
    if (a == 0 && b == 0) return std::signbit(a) && !std::signbit(b);
    return a < b;
}

void write_file(const std::string& path, const std::vector<std::string>& rows, const std::vector<size_t>& order,
                const std::vector<uint8_t>* skip) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw DataFormatError("cannot write " + path + ": " + std::strerror(errno));
    std::fputs(kHeader, f);
    std::fputc('\n', f);
    for (size_t i : order) {
        if (skip && (*skip)[i]) continue;
        std::fwrite(rows[i].data(), 1, rows[i].size(), f);
        std::fputc('\n', f);
    }
    if (std::fclose(f) != 0) throw DataFormatError("error while writing " + path);
}
}  // namespace

void write_outputs(const std::string& out_path, const pb_site_out* sites, const std::vector<pb_stat_out>& stats,
                   const std::vector<int32_t>& physical_pos) {
    const size_t n = stats.size();
    std::vector<std::string> rows(n);
    std::vector<uint8_t> nan1(n), nan2(n);
    for (size_t i = 0; i < n; i++) {
        const pb_site_out& e = sites[stats[i].site_index];
        if (e.segsite_id < 1 || (size_t)e.segsite_id > physical_pos.size())
            throw DataFormatError("segsite_ID " + std::to_string(e.segsite_id) + " has no row in the map file");
        rows[i] = format_row(e, stats[i], physical_pos[(size_t)e.segsite_id - 1]);
        nan1[i] = stats[i].familywise_a1 != stats[i].familywise_a1;
        nan2[i] = stats[i].familywise_a2 != stats[i].familywise_a2;
    }
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), (size_t)0);
    write_file(out_path, rows, order, nullptr);
    std::stable_sort(order.begin(), order.end(),
                     [&](size_t a, size_t b) { return java_double_less(stats[a].familywise_a2, stats[b].familywise_a2); });
    write_file(out_path + ".sorted_by_a2_maxT", rows, order, &nan2);
    std::stable_sort(order.begin(), order.end(),
                     [&](size_t a, size_t b) { return java_double_less(stats[a].familywise_a1, stats[b].familywise_a1); });
    write_file(out_path + ".sorted_by_a1_maxT", rows, order, &nan1);
}

}  // namespace poutine
