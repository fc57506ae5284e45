"""Host-side loaders / packer / writer: semantics of the reference's own readers, no GPU needed.
The end-to-end files-in/files-out run is in the gpu section at the bottom."""
import os

import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import hostio, synth


@pytest.fixture(scope="module", autouse=True)
def _built():
    pb.build()


# ---- Newick (coevolution.jar state machine) ----------------------------------------------------------------
def _by_name(tree):
    return {tree.names[v]: (tree.names[tree.parent[v]] if tree.parent[v] >= 0 else None) for v in range(tree.n_nodes)}


@pytest.mark.parametrize("seed", range(4))
def test_newick_roundtrip(seed):
    t = synth.random_tree(40 + 10 * seed, seed, polytomy=0.3)
    u = hostio.parse_newick(t.to_newick())
    assert _by_name(u) == _by_name(t)
    # leaves come out in left-to-right DFS order, as NewickTree.buildCaches lists them
    assert [u.names[v] for v in u.leaf_nodes] == [t.names[v] for v in t.leaf_nodes]
    assert u.parent[u.root] == -1 and u.names[u.root] == t.names[t.root]


def test_newick_reader_quirks():
    t = hostio.parse_newick(" ( A:0.1 ,\n B :2e-3)R ; trailing garbage")   # whitespace splits tokens; parsing stops at ';'
    assert t.names == ["R", "A", "B"] and list(t.parent) == [-1, 0, 0]
    assert t.branch_len[1] == pytest.approx(0.1, rel=1e-6) and t.branch_len[2] == pytest.approx(2e-3, rel=1e-6)
    t = hostio.parse_newick("(A,,(B,C));")                # empty child between commas = unnamed node
    assert t.names == [None, "A", None, None, "B", "C"]
    for bad in ("(A,B", "(A,B));", "(A:x,B);", "((A,B);", "(A,B)"):   # the last one: "Unexpected end of data."
        with pytest.raises(hostio.DataFormatError):
            hostio.parse_newick(bad)


def test_nexus_to_newick(tmp_path):
    t = synth.yule_tree(12, 3)
    chars = synth.simulate_chars(t, 5, 1, nasty=False)
    sn, lab = synth.simulate_phenotypes(t, 2)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, nexus=True)
    nwk = hostio.nexus_to_newick(paths["tree"])
    assert nwk.startswith("(") and "[" not in nwk and "]" not in nwk
    assert _by_name(hostio.parse_newick(nwk)) == _by_name(t)


# ---- biallelic compaction (host helper of the compact upload) ---------------------------------------------
def _expand(X, colmask):
    A0, A1, D0, D1 = (colmask[:, i][None, :] for i in range(4))
    return A0 ^ (X & D0), A1 ^ (X & D1)


@pytest.mark.parametrize("n_sites,gaps", [(64 * 70 + 13, False), (64 * 3, True), (50, True), (64 * 130 + 1, False)])
def test_compact_planes_roundtrip(n_sites, gaps):
    t = synth.yule_tree(60, 5)
    B0, B1, V = synth.simulate_planes(t, n_sites, 9, leaf_gap=gaps)
    rng = np.random.default_rng(1)
    if gaps:   # garbage under V = 0 must not matter (lower-case / N calls carry arbitrary B bits)
        B0 = B0 ^ (rng.integers(0, 2**63, size=B0.shape, dtype=np.uint64) & ~V)
        B1 = B1 ^ (rng.integers(0, 2**63, size=B1.shape, dtype=np.uint64) & ~V)
    c = pb.compact_planes(B0, B1, V, n_sites)
    assert c is not None
    X, colmask, all_valid = c
    assert all_valid == (not gaps)
    e0, e1 = _expand(X, colmask)
    np.testing.assert_array_equal(e0 & V, B0 & V)
    np.testing.assert_array_equal(e1 & V, B1 & V)
    np.testing.assert_array_equal(X & ~V, np.zeros_like(X))            # X is canonical: 0 where the genotype is not a base
    # multi-threaded and single-threaded runs agree
    c1 = pb.compact_planes(B0, B1, V, n_sites, n_threads=1)
    np.testing.assert_array_equal(c1[0], X)
    np.testing.assert_array_equal(c1[1], colmask)


def test_compact_planes_property_random_alignments():
    """random small alignments over a nasty alphabet: compaction succeeds exactly when no site has three distinct valid
    bases over all nodes; X / colmask then reproduce the planes under V; all_valid exactly when every call is a base"""
    rng = np.random.default_rng(42)
    valid = np.frombuffer(b"ACGT", np.uint8)
    n_ok = n_bad = 0
    for trial in range(300):
        N = int(rng.integers(1, 9))
        S = int(rng.integers(1, 200))
        k = int(rng.integers(1, 4))                                   # bases available per site: 1..3
        junk = rng.random() < 0.5
        chars = np.empty((N, S), dtype=np.uint8)
        for s_ in range(S):
            pool = rng.choice(valid, size=min(k, 1 + int(rng.integers(0, 3))), replace=False)
            col = rng.choice(pool, size=N)
            if junk:
                bad = rng.random(N) < 0.2
                col = np.where(bad, rng.choice(np.frombuffer(b"Nn-acgtRY", np.uint8), size=N), col)
            chars[:, s_] = col
        B0, B1, V = pb.pack_chars(chars)
        isv = np.isin(chars, valid)
        n_distinct = np.array([len(set(chars[isv[:, s_], s_].tolist())) for s_ in range(S)])
        c = pb.compact_planes(B0, B1, V, S, n_threads=int(rng.integers(0, 3)))
        if n_distinct.max(initial=0) >= 3:
            assert c is None
            n_bad += 1
            continue
        n_ok += 1
        assert c is not None
        X, colmask, all_valid = c
        assert all_valid == bool(isv.all())
        e0, e1 = _expand(X, colmask)
        np.testing.assert_array_equal(e0 & V, B0 & V)
        np.testing.assert_array_equal(e1 & V, B1 & V)
        assert not np.any(X & ~V)
    assert n_ok > 50 and n_bad > 50


def test_compact_planes_rejects_third_base():
    t = synth.yule_tree(40, 2)
    n_sites = 64 * 5 + 9
    B0, B1, V = synth.simulate_planes(t, n_sites, 4)
    assert pb.compact_planes(B0, B1, V, n_sites) is not None
    chars = synth.planes_to_chars(B0, B1, V, 0, n_sites)
    col = chars[:, 200]
    third = [b for b in b"ACGT" if b not in set(col.tolist())][0]
    chars[17, 200] = third                                           # one internal or leaf node carries a third base
    b0, b1, v = pb.pack_chars(chars)
    assert pb.compact_planes(b0, b1, v, n_sites) is None
    chars[17, 200] = ord("n")                                        # ... an invalid call instead is fine, with V uploaded
    b0, b1, v = pb.pack_chars(chars)
    c = pb.compact_planes(b0, b1, v, n_sites)
    assert c is not None and c[2] is False


# ---- one pass from characters / FASTA text to the compact upload format ------------------------------------------------------
def test_pack_chars_biallelic_property_random_alignments():
    """random small alignments over a nasty alphabet: the one-pass packer qualifies a matrix exactly when every character is an
    upper-case ACGT and no site carries three of them; X / colmask then expand to the planes pack_chars gives"""
    rng = np.random.default_rng(7)
    valid = np.frombuffer(b"ACGT", np.uint8)
    n_ok = n_bad = 0
    for trial in range(400):
        N = int(rng.integers(1, 12))
        S = int(rng.integers(1, 300))
        k = int(rng.integers(1, 4))
        chars = np.empty((N, S), dtype=np.uint8)
        for s_ in range(S):
            pool = rng.choice(valid, size=min(k, 1 + int(rng.integers(0, 3))), replace=False)
            chars[:, s_] = rng.choice(pool, size=N)
        if rng.random() < 0.3:
            chars[int(rng.integers(0, N)), int(rng.integers(0, S))] = rng.choice(np.frombuffer(b"Nn-acgtRY", np.uint8))
        want_ok = bool(np.isin(chars, valid).all()) and max(len(set(chars[:, s_].tolist())) for s_ in range(S)) <= 2
        c = pb.pack_chars_biallelic(chars, n_threads=int(rng.integers(0, 3)))
        assert (c is not None) == want_ok, trial
        if c is None:
            n_bad += 1
            continue
        n_ok += 1
        X, colmask = c
        B0, B1, V = pb.pack_chars(chars)
        e0, e1 = _expand(X, colmask)
        np.testing.assert_array_equal(e0 & V, B0)
        np.testing.assert_array_equal(e1 & V, B1)
        np.testing.assert_array_equal(X & ~V, np.zeros_like(X))        # padding bits of the last word stay clear
    assert n_ok > 80 and n_bad > 80


@pytest.mark.parametrize("line_width,eol,S,tile", [(0, "\n", 64 * 9 + 17, 256), (60, "\r\n", 64 * 7, 192), (61, "\n", 1000, 1024), (0, "\r", 130, 64)])
def test_fasta_pack_tile_biallelic_matches_three_plane_path(tmp_path, line_width, eol, S, tile):
    t = synth.yule_tree(70, 3)
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 4), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 5)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, line_width=line_width, eol=eol)
    inp = hostio.load_inputs(paths["fasta"], paths["tree"], paths["pheno"], paths["map"])
    fa, nor, N = inp["fasta"], inp["node_of_record"], inp["tree"].n_nodes
    for s0 in range(0, S, tile):
        ns = min(tile, S - s0)
        W = (ns + 63) // 64
        B0, B1, V = (np.zeros((N, W), dtype=np.uint64) for _ in range(3))
        fa.pack_tile(nor, s0, ns, B0, B1, V)
        X = np.full((N, W + 3), 0xDEADBEEF, dtype=np.uint64)[:, :W]     # a pitched destination
        cm = np.empty((W, 4), dtype=np.uint64)
        assert fa.pack_tile_biallelic(nor, N, s0, ns, X, cm, n_threads=3)
        e0, e1 = _expand(X, cm)
        np.testing.assert_array_equal(e0 & V, B0)
        np.testing.assert_array_equal(e1 & V, B1)
        np.testing.assert_array_equal(X & ~V, np.zeros_like(X))
    # a lower-case call, a third base, an unmapped node: the tile does not qualify (the three-plane path takes it)
    W = (S + 63) // 64
    X = np.zeros((N, W), dtype=np.uint64)
    cm = np.empty((W, 4), dtype=np.uint64)
    nor2 = np.array(nor).copy()
    nor2[3] = -1
    assert not fa.pack_tile_biallelic(nor2, N, 0, S, X, cm)
    fa.close()
    for edit in ("n", None):
        ch2 = chars.copy()
        col = ch2[:, 77]
        ch2[11, 77] = ord(edit) if edit else [b for b in b"ACGT" if b not in set(col.tolist())][0]
        d2 = tmp_path / ("e_%s" % edit)
        d2.mkdir()
        p2 = synth.write_inputs(str(d2), t, ch2, sn, lab, line_width=line_width, eol=eol)
        inp2 = hostio.load_inputs(p2["fasta"], p2["tree"], p2["pheno"], p2["map"])
        assert not inp2["fasta"].pack_tile_biallelic(inp2["node_of_record"], N, 0, S, X, cm)
        assert inp2["fasta"].pack_tile_biallelic(inp2["node_of_record"], N, 128, min(64, S - 128), X, cm)     # other tiles still qualify
        inp2["fasta"].close()


# ---- pheno / map ---------------------------------------------------------------------------------------------
def test_pheno_and_map_readers(tmp_path):
    p = tmp_path / "ph.tsv"
    p.write_text("s1\t1\ns2\t0\ns3\tNA\ns1\t0\n")          # repeated id: HashMap.put keeps the last
    ph = hostio.read_phenos(str(p))
    assert ph == {"s1": "0", "s2": "0", "s3": "NA"}
    m = tmp_path / "a.map"
    m.write_text("1\tx\t0\t100\n1\ty\t250\n")               # last column, any number of columns
    assert list(hostio.read_map(str(m))) == [100, 250]


# ---- FASTA packer ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("line_width,eol", [(0, "\n"), (60, "\n"), (7, "\r\n"), (64, "\n"), (1, "\n")])
def test_fasta_pack_matches_char_packer(tmp_path, line_width, eol):
    t = synth.yule_tree(30, 5)
    S = 64 * 5 + 13
    chars = synth.simulate_chars(t, S, 6, nasty=True)
    sn, lab = synth.simulate_phenotypes(t, 7)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, line_width=line_width, eol=eol, seed=3)
    fa = hostio.FastaFile(paths["fasta"])
    assert sorted(fa.names) == sorted(t.names) and np.all(fa.lengths == S)
    node_of = {n: v for v, n in enumerate(t.names)}
    nor = np.array([node_of[n] for n in fa.names], dtype=np.int32)
    R0, R1, RV = pb.pack_chars(chars)
    W = (S + 63) // 64
    # whole alignment in one tile, then in 2-word tiles written into a wider buffer (pitch != tile width)
    B0, B1, V = (np.zeros((t.n_nodes, W), dtype=np.uint64) for _ in range(3))
    fa.pack_tile(nor, 0, S, B0, B1, V, 3)
    for a, b in ((B0, R0), (B1, R1), (V, RV)):
        np.testing.assert_array_equal(a, b)
    T0, T1, TV = (np.zeros((t.n_nodes, W), dtype=np.uint64) for _ in range(3))
    for w0 in range(0, W, 2):
        nw = min(2, W - w0)
        ns = min(S - 64 * w0, 64 * nw)
        fa.pack_tile(nor, 64 * w0, ns, T0[:, w0:w0 + nw], T1[:, w0:w0 + nw], TV[:, w0:w0 + nw], 1)
    for a, b in ((T0, R0), (T1, R1), (TV, RV)):
        np.testing.assert_array_equal(a, b)
    fa.close()


def test_fasta_reader_semantics(tmp_path):
    p = tmp_path / "x.fa"
    # header trimmed (String.trim), ragged line widths, bare CR as a line break, record without trailing newline
    p.write_bytes(b">  n1 extra words \nACG\nT\nAC\n>n2\rAC\rGTAC\n>n3\nacgtN-AC")
    fa = hostio.FastaFile(str(p))
    assert fa.names == ["n1 extra words", "n2", "n3"]
    assert list(fa.lengths) == [6, 6, 8]
    B0, B1, V = (np.zeros((3, 1), dtype=np.uint64) for _ in range(3))
    fa.pack_tile(np.array([2, 0, 1], dtype=np.int32), 0, 8, B0, B1, V, 1)
    back = synth.planes_to_chars(B0, B1, V, 0, 8)
    assert back[2].tobytes() == b"ACGTACNN"        # n1: "ACGTAC", the tail is missing
    assert back[0].tobytes() == b"ACGTACNN"        # n2
    assert back[1].tobytes() == b"NNNNNNAC"        # n3: lower case / N / '-' are missing (HC.java:1600)
    fa.close()


def test_non_ascii_sample_names_match_across_files(tmp_path):
    """tree, FASTA and pheno readers decode with one charset (the reference's FileReader / StringReader: the platform default,
    UTF-8): a sample called 'Sé_01' is the same key in all three"""
    nwk = tmp_path / "t.nwk"
    nwk.write_text("(Sé_01:1,S日本:2)R;\n", encoding="utf-8")
    fa = tmp_path / "a.fasta"
    fa.write_text(">Sé_01\nAC\n>S日本\nAG\n>R\nAC\n", encoding="utf-8")
    ph = tmp_path / "p.tsv"
    ph.write_text("Sé_01\t1\nS日本\t0\n", encoding="utf-8")
    t = hostio.parse_newick(open(nwk, encoding="utf-8").read())
    f = hostio.FastaFile(str(fa))
    assert set(t.names) == set(f.names) == {"Sé_01", "S日本", "R"}
    assert set(hostio.read_phenos(str(ph))) <= set(f.names)
    f.close()


# ---- writer -----------------------------------------------------------------------------------------------------
def test_java_double_to_string():
    j = hostio.java_double_str
    cases = {1.0: "1.0", 0.5: "0.5", 0.001: "0.001", 0.00099: "9.9E-4", 1e7: "1.0E7", 9999999.0: "9999999.0", 1e-5: "1.0E-5",
             9.99990000099999e-06: "9.99990000099999E-6", 1.0 / 3: "0.3333333333333333", 2.0 ** -53: "1.1102230246251565E-16",
             float("nan"): "NaN", 0.0: "0.0", 12345678.0: "1.2345678E7", 3.0 / 100001: "2.999970000299997E-5"}
    for x, s in cases.items():
        assert j(x) == s, (x, j(x), s)


def test_write_outputs_format_and_sorts(tmp_path):
    from poutine_b200._lib import SITE_DTYPE, STAT_DTYPE
    sites = np.zeros(6, dtype=SITE_DTYPE)
    sites["segsite_id"] = np.arange(1, 7)
    sites["site_class"] = 2
    sites["allele1"], sites["allele2"] = ord("A"), ord("G")
    sites["a1_count"], sites["a2_count"], sites["a1_extant"], sites["a2_extant"] = 0, 3, 0, 2
    st = np.zeros(4, dtype=STAT_DTYPE)
    st["site_index"] = [0, 2, 3, 5]
    st["segsite_id"] = [1, 3, 4, 6]
    nan = float("nan")
    st["r_a1"], st["r_maxT_a1"] = [-1, 5, -1, 7], [-1, 50, -1, 50]
    st["pointwise_a1"], st["obs_p_a1"] = [nan, 0.06, nan, 0.08], [nan, 0.01, nan, 0.02]
    st["familywise_a1"] = [nan, 0.5, nan, 0.5]                      # a tie: order must follow the a2-sorted array (stable)
    st["r_a2"], st["r_maxT_a2"] = [3, 0, 9, 1], [30, 0, 90, 10]
    st["pointwise_a2"], st["obs_p_a2"] = [0.04, 1e-5, 0.1, 0.02], [0.001, 1e-7, 0.05, 0.004]
    st["familywise_a2"] = [0.3, 9.99990000099999e-06, 0.9, 0.1]
    st["obs_counts"] = [[0, 0, 2, 0], [1, 1, 2, 1], [0, 0, 1, 1], [2, 0, 3, 0]]
    out = str(tmp_path / "res.out")
    hostio.write_outputs(out, sites, st, np.arange(6) * 10 + 5)
    lines = open(out).read().splitlines()
    assert lines[0] == hostio.EVENT_COLS + "\t" + hostio.STAT_COLS
    assert lines[1] == "1\t5\tA\tG\t0\t3\t0\t2\t[0, 0, 2, 0]\t-1\t3\tNaN\t0.04\tNaN\t0.001\t-1\t30\tNaN\t0.3"
    assert lines[2].endswith("\t5\t0\t0.06\t1.0E-5\t0.01\t1.0E-7\t50\t0\t0.5\t9.99990000099999E-6")
    a2 = [l.split("\t")[0] for l in open(out + ".sorted_by_a2_maxT").read().splitlines()[1:]]
    assert a2 == ["3", "6", "1", "4"]
    a1 = [l.split("\t")[0] for l in open(out + ".sorted_by_a1_maxT").read().splitlines()[1:]]
    assert a1 == ["3", "6"]                                        # NaN rows dropped; tie keeps the a2 order


# ---- files in, files out (GPU) ----------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("nexus,line_width,multi", [(False, 0, 0.05), (True, 70, 0.05), (False, 60, 0.002)])
def test_run_homoplasy_counter_matches_oracle(tmp_path, oracle_mod, nexus, line_width, multi):
    t = synth.yule_tree(200, 11)
    S = 64 * 40 + 9
    # multi = 0.002: a handful of multi-allelic sites, so most upload tiles are compact (one plane) and a few are full
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 12, leaf_gap=True, multiallelic_frac=multi), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 13, na_frac=0.03, no_row_frac=0.02)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, line_width=line_width, nexus=nexus, seed=5)
    out = str(tmp_path / "run.out")
    m = 800
    res = hostio.run_homoplasy_counter(paths["fasta"], paths["tree"], paths["pheno"], paths["map"], out, num_replicates=m, seed=21,
                                       tile_sites=640)
    # the oracle on the same inputs (node numbering of the parsed tree differs; results are per site)
    o = oracle_mod.Oracle(t.parent, t.leaf_nodes)
    ev, cc = o.count_all_homoplasy_events(chars, paths["physical_pos"].astype(np.int32))
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=21, n_threads=4)
    assert res["counts"]["bi"] == cc["bi"] and res["n_informative"] == len(st)
    if multi < 0.01:
        assert 0 < res["n_compact_tiles"] < res["n_tiles"]
    from poutine_b200._lib import SITE_DTYPE, STAT_DTYPE
    sites = np.zeros(S, dtype=SITE_DTYPE)
    for f in ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant"):
        sites[f][ev["segsite_id"] - 1] = ev[f]
    gst = np.zeros(len(st), dtype=STAT_DTYPE)
    for f in STAT_DTYPE.names:
        if f != "site_index":
            gst[f] = st[f]
    gst["site_index"] = st["segsite_id"] - 1
    ref_out = str(tmp_path / "oracle.out")
    hostio.write_outputs(ref_out, sites, gst, paths["physical_pos"])
    for suffix in ("", ".sorted_by_a2_maxT", ".sorted_by_a1_maxT"):
        assert open(out + suffix).read() == open(ref_out + suffix).read(), suffix


@pytest.mark.gpu
def test_cli_precomputed_mode_with_extras(tmp_path, capsys):
    """`python -m poutine_b200 -u -f -t -p -m -r -d -o -X --seed --extras`: the reference's flags for the precomputed
    reconstruction mode; same three files as run_homoplasy_counter, plus the .extras table; refuses to overwrite without -X."""
    from poutine_b200.__main__ import main
    t = synth.yule_tree(120, 3)
    S = 64 * 9 + 5
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 4), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 5)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, nexus=True, seed=2)
    argv = ["-u", "-f", paths["fasta"], "-t", paths["tree"], "-p", paths["pheno"], "-m", paths["map"], "-r", "300",
            "-d", str(tmp_path / "o"), "-o", "run.out", "--seed", "9", "--extras"]
    assert main(argv) == 0
    assert "homoplasically informative sites" in capsys.readouterr().out
    out = str(tmp_path / "o" / "run.out")
    ref = str(tmp_path / "direct.out")
    res = hostio.run_homoplasy_counter(paths["fasta"], paths["tree"], paths["pheno"], paths["map"], ref, num_replicates=300, seed=9)
    for suffix in ("", ".sorted_by_a2_maxT", ".sorted_by_a1_maxT"):
        assert open(out + suffix).read() == open(ref + suffix).read(), suffix
    rows = open(out + ".extras").read().splitlines()
    assert rows[0].split("\t") == ["segsite_ID", "pos", "exact_pointwise_a1", "exact_pointwise_a2", "fisher_two_sided"]
    assert len(rows) == 1 + res["n_informative"]
    vals = [r.split("\t") for r in rows[1:]]
    assert all(0.0 < float(v[4]) <= 1.0 for v in vals)
    assert all(v[2] == "NaN" or 0.0 <= float(v[2]) <= 1.0 for v in vals)
    assert main(argv) == 1                                            # exists, no -X
    assert main(argv + ["-X"]) == 0
    with pytest.raises(SystemExit):                                   # without -u: argparse error, as only that mode is offered
        main(argv[1:])
