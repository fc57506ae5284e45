"""Host side of the reference's call() around the two hot entry points (SURVEY §8f "next" rows): loaders with the
reference's semantics, the FASTA -> bit-plane packer at speed, and the three result files in the reference's
format.  A JVM-less stand-in for Homoplasy_Counter.call() with --use-precomputed-anc-recon (HC.java:230-319):

    build_tree             HC.java:1144   -> parse_newick   (coevolution.jar NewickTreeTokenizer / NewickTreeReader state machine)
    nexus_to_newick        HC.java:866    -> nexus_to_newick
    build_seg_sites        HC.java:1241   -> FastaFile      (compiled/Fasta_Manager.class next(); csrc/fasta.cpp)
    get_physical_positions HC.java:1207   -> read_map
    read_phenos            HC.java:1972   -> read_phenos
    check_* (327-404)                     -> run_homoplasy_counter
    output_significance_assessments_binom HC.java:2137 -> write_outputs

All numerics still run in libpoutine_b200.so on the GPU; nothing here computes a statistic.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import threading
from decimal import Decimal

import numpy as np

from . import _lib
from ._lib import ptr
from .counter import HomoplasyCounter, compact_planes
from .synth import Tree

DELIMS = "(),:;"


class DataFormatError(ValueError):
    """org.gersteinlab.coevolution.core.data.DataFormatException"""


# -------------------------------------------------------------------------------------------------
# Newick (bytecode-faithful restatement of NewickTreeTokenizer.nextToken / NewickTreeReader.readTree)
# -------------------------------------------------------------------------------------------------
_STATE_NAMES = {"S": "NODE_START", "I": "ID", "C": "COLON", "D": "DIST", "E": "NODE_END", "F": "SEMI_COLON"}


def _java_is_whitespace(ch: str) -> bool:
    """java.lang.Character.isWhitespace: \\t \\n VT FF \\r FS GS RS US and the Unicode space / line / paragraph separators
    except the no-break spaces (str.isspace() differs: it accepts U+00A0 and U+0085)"""
    if ch in "\t\n\x0b\x0c\r\x1c\x1d\x1e\x1f ":
        return True
    if ord(ch) < 0x80:
        return False
    import unicodedata
    return unicodedata.category(ch) in ("Zs", "Zl", "Zp") and ch not in "\u00a0\u2007\u202f"


def _tokens(text: str):
    buf = []
    for ch in text:
        if ch in DELIMS:
            if buf:
                yield "".join(buf)
                buf = []
            yield ch
        elif _java_is_whitespace(ch):           # Character.isWhitespace: ends a token, never part of one
            if buf:
                yield "".join(buf)
                buf = []
        else:
            buf.append(ch)
    if buf:
        yield "".join(buf)


def _java_parse_float(tok: str) -> float:
    """Float.parseFloat on a token (no blanks inside): optional sign, decimal digits / '.' / exponent, optional f|F|d|D
    suffix; "NaN" / "Infinity"; anything else is the reference's NumberFormatException (which build_tree, HC.java:1155, does
    not catch: the run dies with that exception)."""
    body = tok[:-1] if len(tok) > 1 and tok[-1] in "fFdD" else tok
    ok = body not in ("", "+", "-") and all(c in "0123456789.eE+-" for c in body)
    if body.lstrip("+-") in ("NaN", "Infinity") and len(body) - len(body.lstrip("+-")) <= 1:
        body, ok = body.replace("Infinity", "inf").replace("NaN", "nan"), True
    if ok:
        try:
            return float(np.float32(float(body)))
        except ValueError:
            pass
    raise DataFormatError('For input string: "%s"' % tok)


def parse_newick(text: str) -> Tree:
    """States S (node start), I (id), C (colon), D (dist), E (node end), F (done) exactly as the reader's switch;
    empty children ("(,)") become unnamed nodes; the root is the outermost node; there is no quoted-label
    support and whitespace splits labels — as in the reference's parser."""
    parent, ids, dist = [], [], []

    def new_node(par):
        parent.append(par)
        ids.append(None)
        dist.append(float("nan"))              # NewickTreeNode.<init>: distToParent = Float.NaN until a ':dist' is read
        return len(parent) - 1

    def err(state, tok):
        # NewickTreeReader.reportFormatError's message
        raise DataFormatError("Unexpected token [%s] encountered in the %s state." % (tok, _STATE_NAMES[state]))

    state, cur = "S", -1
    it = _tokens(text)
    while state != "F":
        tok = next(it, None)
        if tok is None:
            raise DataFormatError("Unexpected end of data.")
        if state == "S":
            if tok == "(":
                cur = new_node(cur)
            elif tok == ")":
                if cur < 0:
                    err(state, tok)
                new_node(cur)
                state = "E"
            elif tok == ",":
                if cur < 0:
                    err(state, tok)
                new_node(cur)
            elif tok == ":":
                cur = new_node(cur)
                state = "C"
            elif tok not in DELIMS:
                cur = new_node(cur)
                ids[cur] = tok
                state = "I"
            else:
                err(state, tok)
        elif state in ("I", "D", "E"):
            if tok == ")":
                cur = parent[cur]
                if cur < 0:
                    err(state, tok)
                state = "E"
            elif tok == ",":
                cur = parent[cur]
                if cur < 0:
                    err(state, tok)
                state = "S"
            elif tok == ":" and state != "D":
                state = "C"
            elif tok == ";":
                state = "F"
            elif state == "E" and tok not in DELIMS:
                ids[cur] = tok
                state = "I"
            else:
                err(state, tok)
        elif state == "C":
            if tok in DELIMS:
                err(state, tok)
            dist[cur] = _java_parse_float(tok)
            state = "D"
    if cur < 0:
        raise DataFormatError("More close brackets then open brackets")
    if parent[cur] >= 0:
        raise DataFormatError("More open brackets then close brackets")
    n = len(parent)
    par = np.asarray(parent, dtype=np.int32)
    nchild = np.bincount(par[par >= 0], minlength=n)
    is_leaf = (nchild == 0).astype(np.uint8)
    # NewickTree.buildCaches: leaves in left-to-right DFS order == creation order of childless nodes
    leaves = np.flatnonzero(is_leaf).astype(np.int32)
    return Tree(parent=par, is_leaf=is_leaf, leaf_nodes=leaves, branch_len=np.asarray(dist, dtype=np.float64), names=ids)


def nexus_to_newick(path: str) -> str:
    """HC.java:866-914: the line after 'Begin Trees;' (case-insensitive, exact), cut before the first '(',
    every '[...]' annotation deleted."""
    lines = _java_lines(path)
    for i, line in enumerate(lines):
        if line.lower() == "begin trees;":
            break
    else:
        raise DataFormatError("no 'Begin Trees;' line in " + path)    # the reference dies with an NPE here
    if i + 1 >= len(lines):
        raise DataFormatError("no tree line after 'Begin Trees;' in " + path)
    s = lines[i + 1]
    k = s.find("(")
    if k < 0:                        # StringBuilder.delete(0, -1) throws in the reference
        raise DataFormatError("the tree line of %s holds no '('" % path)
    s = s[k:]
    while True:
        a = s.find("[")
        if a < 0:
            break
        b = s.find("]")
        if b < a:      # StringBuilder.delete(start, end + 1) with end + 1 < start throws in the reference
            raise DataFormatError("unbalanced '[' ... ']' annotation in the tree line of " + path)
        s = s[:a] + s[b + 1:]
    return s


# -------------------------------------------------------------------------------------------------
# pheno / map files
# -------------------------------------------------------------------------------------------------
def _java_lines(path: str):
    """BufferedReader.readLine: a line ends at \\n, \\r or \\r\\n (str.splitlines also splits at VT, FF, FS, GS, RS, NEL, LS, PS)"""
    with open(path, newline="", encoding="utf-8", errors="replace") as f:     # FileReader: default charset, malformed -> U+FFFD
        text = f.read()
    lines = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def _java_split_tabs(line: str):
    """String.split("\\t"): trailing empty strings are removed ("a\\t\\t" -> ["a"]); "" -> [""]"""
    cols = line.split("\t")
    if len(cols) > 1:
        while cols and cols[-1] == "":
            cols.pop()
    return cols


def _java_parse_int(s: str) -> int:
    """Integer.parseInt: [+-]?[0-9]+ within int32 — no blanks, no underscores (int() accepts both)"""
    body = s[1:] if s[:1] in "+-" else s
    if not body or not body.isascii() or not body.isdigit() or not -(1 << 31) <= int(s) < (1 << 31):
        raise ValueError('For input string: "%s"' % s)
    return int(s)


def read_phenos(path: str):
    """HC.java:1972-2019 -> ordered dict id -> label string (a HashMap: a repeated id keeps the last label).  A row with
    fewer than two columns is the reference's "malformed phenotype file" exit (HC.java:1983-1987)."""
    phenos = {}
    for line in _java_lines(path):
        cols = _java_split_tabs(line)
        if len(cols) < 2:
            raise DataFormatError("malformed phenotype file %s: a row has fewer than two tab-separated columns" % path)
        phenos[cols[0]] = cols[1]
    return phenos


def read_map(path: str) -> np.ndarray:
    """HC.java:1207-1235: Integer.parseInt of the last tab-separated column of every row ("malformed map file" exit otherwise)."""
    out = []
    for line in _java_lines(path):
        cols = _java_split_tabs(line)
        try:
            out.append(_java_parse_int(cols[-1]))
        except (ValueError, IndexError):
            raise DataFormatError("malformed map file %s: %r is not an integer position" % (path, cols[-1] if cols else "")) from None
    return np.asarray(out, dtype=np.int64)


# -------------------------------------------------------------------------------------------------
# FASTA -> planes
# -------------------------------------------------------------------------------------------------
class FastaFile:
    """mmap'ed multi-FASTA indexed by csrc/fasta.cpp (Fasta_Manager.next() semantics)."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        L = _lib.lib()
        rc = L.pb_fasta_open(path.encode(), C.byref(self._h))
        if rc != 0:
            raise OSError(L.pb_fasta_last_error().decode())
        n = L.pb_fasta_num_records(self._h)
        self.names, self.lengths = [], np.empty(n, dtype=np.int64)
        name, nlen, slen = C.c_void_p(), C.c_int64(), C.c_int64()
        for i in range(n):
            L.pb_fasta_record(self._h, i, C.byref(name), C.byref(nlen), C.byref(slen))
            # the same charset as the tree / pheno readers, so that non-ASCII sample names match across the files
            self.names.append(C.string_at(name.value, nlen.value).decode("utf-8", errors="replace") if nlen.value else "")
            self.lengths[i] = slen.value

    def close(self):
        if self._h:
            _lib.lib().pb_fasta_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack_tile(self, node_of_record, site_begin, n_sites_tile, B0, B1, V, n_threads=0):
        """columns [site_begin, site_begin + n_sites_tile) of every record -> rows node_of_record[i] of the tile planes"""
        nor = np.ascontiguousarray(node_of_record, dtype=np.int32)
        assert B0.strides == B1.strides == V.strides and B0.strides[1] == 8
        rc = _lib.lib().pb_fasta_pack_tile(self._h, ptr(nor), int(site_begin), int(n_sites_tile), ptr(B0), ptr(B1), ptr(V),
                                           B0.strides[0] // 8, int(n_threads))
        if rc != 0:
            raise ValueError(_lib.lib().pb_fasta_last_error().decode())

    def pack_tile_biallelic(self, node_of_record, n_nodes, site_begin, n_sites_tile, X, colmask, n_threads=0):
        """one pass from the text to the compact upload format (X plane + per-column site masks): True when the tile
        qualifies (every genotype an upper-case ACGT, at most two bases per site), else False (outputs undefined)"""
        nor = np.ascontiguousarray(node_of_record, dtype=np.int32)
        assert X.dtype == np.uint64 and X.strides[1] == 8 and colmask.dtype == np.uint64 and colmask.flags.c_contiguous
        assert colmask.size >= 4 * ((n_sites_tile + 63) // 64)
        rc = _lib.lib().pb_fasta_pack_tile_biallelic(self._h, ptr(nor), int(n_nodes), int(site_begin), int(n_sites_tile), ptr(X),
                                                     X.strides[0] // 8, ptr(colmask), int(n_threads))
        if rc < 0:
            raise ValueError(_lib.lib().pb_fasta_last_error().decode())
        return rc == 1


# -------------------------------------------------------------------------------------------------
# result files (HC.java:2137-2210; Homoplasy_Events.toString 5574-5578; Binomial_Test_Stat.toString 5918-5922)
# -------------------------------------------------------------------------------------------------
EVENT_COLS = "segsite_ID\tphysical_pos\tallele1\tallele2\ta1_count\ta2_count\ta1_count_extant_only\ta2_count_extant_only\tobs_homoplasy_counts"
STAT_COLS = ("r_a1\tr_a2\tpointwise_pvalue_a1\tpointwise_pvalue_a2\tobs_binom_pvalue_a1\tobs_binom_pvalue_a2\tr_maxT_a1\tr_maxT_a2\t"
             "familywise_pvalue_a1\tfamilywise_pvalue_a2")


def java_double_str(x: float) -> str:
    if x != x:                                          # NaN is not a usable cache key (nan != nan) ...
        return "NaN"
    if x == 0:                                          # ... and -0.0 / 0.0 are the same key
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    return _java_double_str(x)


@functools.lru_cache(maxsize=1 << 20)     # p-values repeat massively: (r+1)/(m+1) and table entries
def _java_double_str(x: float) -> str:
    """java.lang.Double.toString: shortest digits that round-trip (JDK 19+; earlier JDKs print one digit more in
    rare cases), plain decimal for 1e-3 <= |x| < 1e7, otherwise d.dddE[-]n; always at least one fraction digit."""
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    _, dg, ex = Decimal(repr(abs(x))).as_tuple()          # repr(): shortest digits that round-trip
    digits = "".join(map(str, dg)).lstrip("0")
    e10 = len(digits) - 1 + ex                            # decimal exponent of the first digit
    digits = digits.rstrip("0") or "0"
    if len(digits) == 1:
        # at least two digits are rendered and the two-digit decimal closest to x is chosen: differs from "shortest digit
        # + 0" only for subnormals with a few bits (Double.MIN_VALUE prints as 4.9E-324)
        two = "%.1e" % abs(x)
        if float(two) == abs(x):
            digits = (two[0] + two[2]).rstrip("0") or "0"
            e10 = int(two.split("e")[1])
    if 1e-3 <= abs(x) < 1e7:
        if e10 >= 0:
            ipart = digits[:e10 + 1].ljust(e10 + 1, "0")
            fpart = digits[e10 + 1:] or "0"
        else:
            ipart = "0"
            fpart = "0" * (-e10 - 1) + digits
        return sign + ipart + "." + fpart
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)


def format_rows(sites, stats, physical_pos):
    """-> list of row strings in site order (one per informative site)."""
    ev = sites[stats["site_index"]]
    rows = []
    for e, t in zip(ev, stats):
        sid = int(e["segsite_id"])
        left = "%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t[%d, %d, %d, %d]" % (
            sid, int(physical_pos[sid - 1]), chr(e["allele1"]), chr(e["allele2"]), e["a1_count"], e["a2_count"], e["a1_extant"],
            e["a2_extant"], *[int(c) for c in t["obs_counts"]])
        right = "\t".join([str(int(t["r_a1"])), str(int(t["r_a2"])), java_double_str(float(t["pointwise_a1"])),
                           java_double_str(float(t["pointwise_a2"])), java_double_str(float(t["obs_p_a1"])),
                           java_double_str(float(t["obs_p_a2"])), str(int(t["r_maxT_a1"])), str(int(t["r_maxT_a2"])),
                           java_double_str(float(t["familywise_a1"])), java_double_str(float(t["familywise_a2"]))])
        rows.append(left + "\t" + right)
    return rows


def _java_compare_key(x):
    # Double.compare: -0.0 < 0.0 and NaN above everything
    return (1, 0.0) if x != x else (0, x)


def write_outputs(out_path: str, sites, stats, physical_pos):
    """<out>, <out>.sorted_by_a2_maxT, <out>.sorted_by_a1_maxT.  The a1 file is a STABLE sort of the a2-sorted rows,
    as the reference sorts the same array twice (Arrays.parallelSort on objects is a stable merge sort)."""
    rows = format_rows(sites, stats, physical_pos)
    header = EVENT_COLS + "\t" + STAT_COLS
    with open(out_path, "w") as f:
        f.write(header + "\n")
        for r in rows:
            f.write(r + "\n")
    order = sorted(range(len(rows)), key=lambda i: _java_compare_key(float(stats["familywise_a2"][i])))
    with open(out_path + ".sorted_by_a2_maxT", "w") as f:
        f.write(header + "\n")
        for i in order:
            if stats["familywise_a2"][i] == stats["familywise_a2"][i]:
                f.write(rows[i] + "\n")
    order = sorted(order, key=lambda i: _java_compare_key(float(stats["familywise_a1"][i])))
    with open(out_path + ".sorted_by_a1_maxT", "w") as f:
        f.write(header + "\n")
        for i in order:
            if stats["familywise_a1"][i] == stats["familywise_a1"][i]:
                f.write(rows[i] + "\n")


# -------------------------------------------------------------------------------------------------
# call() with --use-precomputed-anc-recon
# -------------------------------------------------------------------------------------------------
def load_inputs(fasta, tree_file, pheno_file, map_file):
    """The part of call() in front of the two entry points (HC.java:230-296): loaders + cross-file checks, flattened to what
    the C ABI takes.  -> dict(tree, fasta (open FastaFile), physical_pos, n_sites, node_of_record, sample_node, label)"""
    with open(tree_file, newline="", encoding="utf-8", errors="replace") as f:
        text = f.read()
    newick = nexus_to_newick(tree_file) if text.lstrip().upper().startswith("#NEXUS") else text
    tree = parse_newick(newick)
    if any(n is None for n in tree.names):
        raise DataFormatError("every tree node needs a label (treetime writes NODE_xxxxxxx for internal nodes)")
    node_of_name = {}
    for v, n in enumerate(tree.names):
        if n in node_of_name:
            raise DataFormatError("duplicate node label %r" % n)
        node_of_name[n] = v
    fa = FastaFile(fasta)
    physical_pos = read_map(map_file)
    # The reference loops over the map rows (HC.java:1418) and reads snps[i] of EVERY record (HC.java:1383-1394): a shorter
    # record is an ArrayIndexOutOfBoundsException there, and check_num_seg_sites (HC.java:327-365) exits unless the first
    # record in its HashMap's iteration order has exactly as many sites as the map file has rows.  Here: every record must.
    S = len(physical_pos)
    for name, n in zip(fa.names, fa.lengths):
        if int(n) != S:
            raise DataFormatError("# sites of FASTA record %r (%d) != # rows in the map file (%d)" % (name, int(n), S))
    node_of_record = np.array([node_of_name.get(n, -1) for n in fa.names], dtype=np.int32)
    last = {}
    for i, n in enumerate(fa.names):                      # seg_sites.put(node_name, seq): a repeated name keeps its LAST record
        if n in last:
            node_of_record[last[n]] = -1
        last[n] = i
    missing = set(range(tree.n_nodes)) - set(int(v) for v in node_of_record if v >= 0)
    if missing:
        raise DataFormatError("tree node %r has no FASTA record" % tree.names[min(missing)])
    phenos = read_phenos(pheno_file)
    leaf_set = set(tree.names[v] for v in tree.leaf_nodes)
    for name in phenos:                                   # check_sample_names (HC.java:367-404)
        if name not in node_of_name or name not in leaf_set:
            raise DataFormatError("phenotype sample %r is not a leaf of the tree / not in the FASTA" % name)
    sample_node = np.array([node_of_name[n] for n in phenos], dtype=np.int32)
    label = np.array([1 if v == "1" else 0 if v == "0" else 2 for v in phenos.values()], dtype=np.uint8)
    return dict(tree=tree, fasta=fa, physical_pos=physical_pos, n_sites=S, node_of_record=node_of_record, sample_node=sample_node,
                label=label)


def run_homoplasy_counter(fasta, tree_file, pheno_file, map_file, out_path, num_replicates=100000, min_hcount=1, seed=0, device=0,
                          tile_sites=1 << 14, pack_threads=0, compact_upload=True, extras=False):
    """Files in, files out: same inputs as `poutine.sh -u -f -t -p -m`, same three result files.
    `tree_file` is a Newick file with internal node labels, or treetime's annotated_tree.nexus."""
    inp = load_inputs(fasta, tree_file, pheno_file, map_file)
    tree, fa, physical_pos, S = inp["tree"], inp["fasta"], inp["physical_pos"], inp["n_sites"]
    node_of_record, sample_node, label = inp["node_of_record"], inp["sample_node"], inp["label"]

    hc = HomoplasyCounter(num_replicates=num_replicates, min_hcount=min_hcount, device=device, seed=seed)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sample_node, label)
    hc.set_sites(S)
    hc.expect_site_output()      # records of tiles swept (and tailed) underneath the upload travel back underneath it
    # stream the alignment: pack tile t+1 on a worker thread (the C packer releases the GIL) while tile t uploads
    # 16384-site tiles: page-locked buffer set-up against per-tile overhead (profiles/r1_native_host_sweep.json)
    tile_sites = max(64, (min(tile_sites, S + 63) // 64) * 64)
    tw = tile_sites // 64
    # A tile whose sites all carry at most two bases (the usual SNP alignment) is compacted on the worker thread too and
    # goes up as one plane (+ V when some call is not a base) instead of three: buffer slot 3 holds X.
    bufs = []
    for _ in range(2):
        arrs = [_lib.pinned_empty((tree.n_nodes, tw), np.uint64) for _ in range(4)]
        bufs.append(arrs)
    starts = list(range(0, S, tile_sites))
    compacted = [None, None]
    colmasks = [np.empty((tw, 4), dtype=np.uint64) for _ in range(2)]
    n_direct = [0]

    def pack(k):
        s0 = starts[k]
        ns = min(tile_sites, S - s0)
        B0, B1, V, X = (a[0] for a in bufs[k & 1])
        # the usual alignment (every genotype a base, two bases per site) goes from the text straight to the compact format
        if compact_upload and fa.pack_tile_biallelic(node_of_record, tree.n_nodes, s0, ns, X, colmasks[k & 1], pack_threads):
            nw = (ns + 63) // 64
            compacted[k & 1] = (X[:, :nw], colmasks[k & 1][:nw], True)
            n_direct[0] += 1
            return
        fa.pack_tile(node_of_record, s0, ns, B0, B1, V, pack_threads)
        compacted[k & 1] = compact_planes(B0, B1, V, ns, X=X, n_threads=pack_threads) if compact_upload else None

    worker = None
    n_compact_tiles = 0
    if starts:
        pack(0)
    for k, s0 in enumerate(starts):
        if k + 1 < len(starts):
            worker = threading.Thread(target=pack, args=(k + 1,))
            worker.start()
        ns = min(tile_sites, S - s0)
        nw = (ns + 63) // 64
        B0, B1, V = (a[0][:, :nw] for a in bufs[k & 1][:3])
        c = compacted[k & 1]
        if c is not None:
            hc.put_planes_biallelic(c[0], None if c[2] else V, c[1], s0 // 64)
            n_compact_tiles += 1
        else:
            hc.put_planes(B0, B1, V, s0 // 64)
        if worker is not None:
            worker.join()
            worker = None
    all_events, counts = hc.count_all_homoplasy_events(compact=False)
    stats, maxT = hc.assoc()
    write_outputs(out_path, all_events, stats, physical_pos)
    if extras:
        # not a reference file: the exact point-wise null and two-sided Fisher per informative site (HomoplasyCounter.extras)
        ep, fp = hc.extras()
        with open(out_path + ".extras", "w") as f:
            f.write("segsite_ID\tpos\texact_pointwise_a1\texact_pointwise_a2\tfisher_two_sided\n")
            for i in range(len(stats)):
                f.write("%d\t%d\t%s\t%s\t%s\n" % (stats["segsite_id"][i], physical_pos[stats["site_index"][i]],
                                                  java_double_str(float(ep[i, 0])), java_double_str(float(ep[i, 1])), java_double_str(float(fp[i]))))
    result = dict(counts=counts, n_informative=len(stats), n_sites=S, n_nodes=tree.n_nodes, n_samples=len(label),
                  n_tiles=len(starts), n_compact_tiles=n_compact_tiles, n_direct_tiles=n_direct[0])
    fa.close()
    hc.close()
    for arrs in bufs:
        for _, p in arrs:
            _lib.pinned_free(p)
    return result
