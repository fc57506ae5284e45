"""The native (C++) host around the two entry points: poutine_b200/host/ -> poutine_b200/poutine_b200_host.

CPU part: the loaders and the writer of the C++ host against the Python host (poutine_b200/hostio.py) on the same files —
both restate the reference's readers (coevolution.jar Newick state machine, HC.java:866-914 / 1207-1235 / 1972-2019) and its
writer (HC.java:2137-2210), so they must agree token for token and byte for byte — plus the Java-specific corners
(String.split, readLine, Integer.parseInt, Double.toString) where the C++ host follows the JDK.
GPU part (bottom): files in -> files out through the binary, byte-compared with the Python pipeline and the oracle's rows."""
import os
import struct
import subprocess

import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import hostio, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "poutine_b200", "poutine_b200_host")


@pytest.fixture(scope="module", autouse=True)
def _built():
    pb.build()
    assert os.path.exists(EXE), "poutine_b200_host was not built"


def run(*args, stdin=None, ok=True):
    r = subprocess.run([EXE, *args], input=stdin, capture_output=True, text=True, timeout=180)
    if ok:
        assert r.returncode == 0, r.stderr
    return r


def _cpp_tree(path):
    out = run("--x-parse-tree", str(path)).stdout.split("\n")
    n = int(out[0])
    rows = [l.split("\t") for l in out[1:1 + n]]
    parent = [int(r[0]) for r in rows]
    is_leaf = [int(r[1]) for r in rows]
    dist = [struct.unpack("<f", struct.pack("<I", int(r[2], 16)))[0] for r in rows]
    names = [r[4] if r[3] == "1" else None for r in rows]
    leaves = [int(x) for x in out[1 + n].split()]
    return parent, is_leaf, dist, names, leaves


@pytest.mark.parametrize("seed", range(4))
def test_newick_matches_python_host(tmp_path, seed):
    t = synth.random_tree(60 + 25 * seed, seed, polytomy=0.3)
    p = tmp_path / "t.nwk"
    p.write_text(t.to_newick() + "\n")
    u = hostio.parse_newick(p.read_text())
    parent, is_leaf, dist, names, leaves = _cpp_tree(p)
    assert parent == list(u.parent) and is_leaf == list(u.is_leaf) and names == u.names and leaves == list(u.leaf_nodes)
    # bit patterns: a node without ':dist' keeps NewickTreeNode's Float.NaN (the root here)
    assert np.array_equal(np.asarray(dist, dtype=np.float32).view(np.uint32), u.branch_len.astype(np.float32).view(np.uint32))


def test_newick_reader_quirks(tmp_path):
    p = tmp_path / "q.nwk"
    p.write_text(" ( A:0.1 ,\n B :2e-3)R ; trailing garbage")     # whitespace splits tokens; parsing stops at ';'
    parent, is_leaf, dist, names, leaves = _cpp_tree(p)
    assert names == ["R", "A", "B"] and parent == [-1, 0, 0] and leaves == [1, 2]
    assert dist[1] == np.float32(0.1) and dist[2] == np.float32(2e-3)
    p.write_text("(A,,(B,C));")                                     # empty child between commas = unnamed node
    parent, is_leaf, dist, names, leaves = _cpp_tree(p)
    assert names == [None, "A", None, None, "B", "C"] and parent == [-1, 0, 0, 0, 3, 3] and is_leaf == [0, 1, 1, 0, 1, 1]
    for bad, msg in (("(A,B", "Unexpected end of data"), ("(A,B));", "NODE_END state"), ("(A:x,B);", 'For input string: "x"'),
                     ("((A,B);", "More open brackets"), ("(A,B)", "Unexpected end of data"), (");", "NODE_START state")):
        p.write_text(bad)
        r = run("--x-parse-tree", str(p), ok=False)
        assert r.returncode == 255 and msg in r.stderr, (bad, r.stderr)
        with pytest.raises(hostio.DataFormatError):
            hostio.parse_newick(bad)


def test_nexus_to_newick_matches_python_host(tmp_path):
    t = synth.yule_tree(30, 3)
    chars = synth.simulate_chars(t, 5, 1, nasty=False)
    sn, lab = synth.simulate_phenotypes(t, 2)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, nexus=True)
    got = run("--x-nexus-to-newick", paths["tree"]).stdout.rstrip("\n")
    assert got == hostio.nexus_to_newick(paths["tree"]) and got.startswith("(") and "[" not in got
    # 'begin trees;' in any case, CRLF line ends (BufferedReader.readLine)
    p = tmp_path / "crlf.nexus"
    p.write_bytes(b"#NEXUS\r\nbegin TREES;\r\n Tree t=[&U](A[&x=1]:1,B:2)R[&y];\r\nEnd;\r\n")
    assert run("--x-nexus-to-newick", str(p)).stdout.rstrip("\n") == "(A:1,B:2)R;"
    # where the reference throws (no tree line / no '(' / ']' before '['), the host reports an error instead of guessing
    for body in ("#NEXUS\nBegin Taxa;\nEnd;\n", "#NEXUS\nBegin Trees;\nTree t = A;\n", "#NEXUS\nBegin Trees;\n(A]:1,B[x:2);\n"):
        p.write_text(body)
        assert run("--x-nexus-to-newick", str(p), ok=False).returncode == 255


def test_pheno_and_map_readers_follow_the_jdk(tmp_path):
    ph = tmp_path / "p.tsv"
    ph.write_bytes(b"s1\t1\r\ns2\t0\textra\ns3\tNA\ns1\t0\ns4\t1\t\t\n")   # CRLF, extra column, repeated id, trailing tabs
    out = run("--x-read-phenos", str(ph)).stdout.splitlines()
    assert out[0] == "1\t2\t1"                                     # cases, controls, missing (HC.java:1999-2007)
    assert out[1:] == ["s1\t0", "s2\t0", "s3\tNA", "s4\t1"]        # HashMap.put: the repeated id keeps the last label
    py = hostio.read_phenos(str(ph))
    assert [k + "\t" + v for k, v in py.items()] == out[1:]
    for bad in (b"s1\t1\ns2\n", b"s1\t1\ns2\t\n", b"s1\t1\n\ns3\t0\n"):    # cols[1] out of bounds: "malformed phenotype file"
        ph.write_bytes(bad)
        r = run("--x-read-phenos", str(ph), ok=False)
        assert r.returncode == 255 and "malformed phenotype file" in r.stderr
    mp = tmp_path / "m.map"
    mp.write_text("1\tsnp1\t0\t17\n1\tsnp2\t0\t+23\n9\n-4\t\t\n")
    assert run("--x-read-map", str(mp)).stdout.split() == ["17", "23", "9", "-4"]
    for bad in ("1\tsnp\t0\t12x\n", "1\tsnp\t0\t 12\n", "1\t2147483648\n", "\n"):   # Integer.parseInt throws on each
        mp.write_text(bad)
        r = run("--x-read-map", str(mp), ok=False)
        assert r.returncode == 255 and "malformed map file" in r.stderr, bad


def test_java_double_to_string_matches_python_host():
    rng = np.random.default_rng(5)
    m = 100000
    xs = [1.0, 0.5, 0.001, 0.00099, 1e7, 9999999.0, 1e-5, 9.99990000099999e-06, 1.0 / 3, 2.0 ** -53, float("nan"), 0.0, -0.0,
          12345678.0, 3.0 / 100001, 1e-300, 5e-324, 1.7976931348623157e308, float("inf"), -2.5, 100.0, 1234567.125, 1e21, 123456789012.0]
    xs += [(r + 1) / (m + 1) for r in rng.integers(0, m, 300)]            # what the p-value columns hold
    xs += list(np.exp(rng.uniform(-40, 20, 300)))
    xs += list(rng.integers(0, 2 ** 63, 300, dtype=np.uint64).view(np.float64))   # arbitrary bit patterns
    bits = "".join("%016x\n" % struct.unpack("<Q", struct.pack("<d", float(x)))[0] for x in xs)
    got = run("--x-format-doubles", stdin=bits).stdout.splitlines()
    assert len(got) == len(xs)
    for x, g in zip(xs, got):
        assert g == hostio.java_double_str(float(x)), (x, g)
    known = {1.0: "1.0", 0.00099: "9.9E-4", 1e7: "1.0E7", 9999999.0: "9999999.0", 2.0 ** -53: "1.1102230246251565E-16",
             12345678.0: "1.2345678E7", 3.0 / 100001: "2.999970000299997E-5", 5e-324: "4.9E-324", 100.0: "100.0", 1e21: "1.0E21"}
    for x, s in known.items():
        assert got[xs.index(x)] == s


def _random_records(rng, S, n_inf):
    from poutine_b200._lib import SITE_DTYPE, STAT_DTYPE
    sites = np.zeros(S, dtype=SITE_DTYPE)
    sites["segsite_id"] = np.arange(1, S + 1)
    sites["site_class"] = 2
    sites["allele1"] = rng.choice([65, 67, 71, 84], S)
    sites["allele2"] = rng.choice([65, 67, 71, 84], S)
    for f in ("a1_count", "a2_count", "a1_extant", "a2_extant"):
        sites[f] = rng.integers(0, 40, S)
    st = np.zeros(n_inf, dtype=STAT_DTYPE)
    st["site_index"] = np.sort(rng.choice(S, n_inf, replace=False))
    st["segsite_id"] = st["site_index"] + 1
    st["obs_counts"] = rng.integers(0, 30, (n_inf, 4))
    m = 1000
    for a in ("a1", "a2"):
        use = rng.random(n_inf) < 0.7
        r = rng.integers(0, m, n_inf)
        rm = rng.integers(0, 12, n_inf) * 80                       # few distinct values: many ties in the sort keys
        st["r_" + a] = np.where(use, r, -1)
        st["r_maxT_" + a] = np.where(use, rm, -1)
        st["pointwise_" + a] = np.where(use, (r + 1) / (m + 1), np.nan)
        st["obs_p_" + a] = np.where(use, np.exp(rng.uniform(-30, 0, n_inf)), np.nan)
        st["familywise_" + a] = np.where(use, (rm + 1) / (m + 1), np.nan)
    return sites, st


@pytest.mark.parametrize("seed,S,n_inf", [(0, 400, 250), (1, 50, 50), (2, 10, 0)])
def test_writer_matches_python_host_byte_for_byte(tmp_path, seed, S, n_inf):
    """the three result files: column order, Arrays.toString / Double.toString formatting, the a2 sort followed by the
    STABLE a1 sort of the same array (ties abound here), NaN rows dropped"""
    rng = np.random.default_rng(seed)
    sites, st = _random_records(rng, S, n_inf)
    pos = np.sort(rng.choice(np.arange(1, 100 * S), S, replace=False))
    mp = tmp_path / "m.map"
    mp.write_text("".join("1\ts%d\t0\t%d\n" % (i, p) for i, p in enumerate(pos)))
    sites.tofile(tmp_path / "sites.bin")
    st.tofile(tmp_path / "stats.bin")
    run("--x-write-records", str(tmp_path / "sites.bin"), str(tmp_path / "stats.bin"), str(mp), str(tmp_path / "cpp.out"))
    hostio.write_outputs(str(tmp_path / "py.out"), sites, st, pos)
    for suffix in ("", ".sorted_by_a2_maxT", ".sorted_by_a1_maxT"):
        a = open(str(tmp_path / "cpp.out") + suffix).read()
        assert a == open(str(tmp_path / "py.out") + suffix).read(), suffix
        assert a.count("\n") >= 1


def test_cli_errors_without_gpu_work(tmp_path):
    """argument handling mirrors the reference's required options; input errors surface before any device work"""
    r = run(ok=False)
    assert r.returncode == 2 and "required" in r.stderr
    r = run("-f", "a", "-t", "b", "-p", "c", "-m", "d", ok=False)
    assert r.returncode == 2 and "--use-precomputed-anc-recon" in r.stderr
    r = run("-u", "-f", "a", "-t", str(tmp_path / "missing.nwk"), "-p", "c", "-m", "d", "-d", str(tmp_path / "o"), ok=False)
    assert r.returncode == 255 and "cannot open" in r.stderr


@pytest.mark.parametrize("nexus,line_width,eol", [(False, 0, "\n"), (True, 70, "\n"), (False, 33, "\r\n")])
def test_flattened_inputs_agree_between_hosts_and_with_the_oracle(tmp_path, oracle_mod, nexus, line_width, eol):
    """loaders + cross-file glue of both hosts (load_inputs): same parent / leaf / node_of_record / sample / label arrays; and
    the alignment they describe — FASTA records packed into the parsed tree's node order — gives the oracle the same
    Homoplasy_Events, site by site, as the generator's own tree and characters (node numbering differs, results do not).  A
    repeated record name keeps its LAST record, as seg_sites.put does (HC.java:1255-1266)."""
    t = synth.yule_tree(40, 21)
    S = 64 * 3 + 17
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 22, leaf_gap=True, multiallelic_frac=0.05), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 23, na_frac=0.05, no_row_frac=0.05)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, line_width=line_width, eol=eol, nexus=nexus, seed=7)
    body = open(paths["fasta"], newline="").read()
    with open(paths["fasta"], "w", newline="") as f:            # an earlier, wrong record of one leaf: must lose against the real one
        f.write(">" + t.names[int(t.leaf_nodes[3])] + eol + "T" * S + eol + body)
    r = run("--x-flatten", paths["fasta"], paths["tree"], paths["pheno"], paths["map"])
    out = dict(l.split(" ", 1) if " " in l else (l, "") for l in r.stdout.strip().split("\n"))
    cpp = {k: np.array(v.split(), dtype=np.int64) for k, v in out.items() if k not in ("n_nodes",)}
    inp = hostio.load_inputs(paths["fasta"], paths["tree"], paths["pheno"], paths["map"])
    tree = inp["tree"]
    assert out["n_nodes"] == "%d n_sites %d" % (tree.n_nodes, S) and inp["n_sites"] == S
    assert np.array_equal(cpp["parent"], tree.parent) and np.array_equal(cpp["is_leaf"], tree.is_leaf)
    assert np.array_equal(cpp["node_of_record"], inp["node_of_record"]) and cpp["node_of_record"][0] == -1
    assert np.array_equal(cpp["sample_node"], inp["sample_node"]) and np.array_equal(cpp["label"], inp["label"])
    assert np.array_equal(cpp["physical_pos"], paths["physical_pos"]) and np.array_equal(inp["physical_pos"], paths["physical_pos"])
    assert [tree.names[v] for v in inp["sample_node"]] == [t.names[int(v)] for v in sn]          # pheno-file order
    assert np.array_equal(inp["label"], np.where(lab == 1, 1, np.where(lab == 0, 0, 2)))
    # the alignment in the parsed tree's node order, through the library's FASTA packer
    W = (S + 63) // 64
    B0, B1, V = (np.zeros((tree.n_nodes, W), dtype=np.uint64) for _ in range(3))
    inp["fasta"].pack_tile(inp["node_of_record"], 0, S, B0, B1, V, 2)
    got_chars = synth.planes_to_chars(B0, B1, V, 0, S)
    ev_a, cc_a = oracle_mod.Oracle(tree.parent, tree.leaf_nodes).count_all_homoplasy_events(got_chars)
    ev_b, cc_b = oracle_mod.Oracle(t.parent, t.leaf_nodes).count_all_homoplasy_events(chars)
    assert cc_a == cc_b and len(ev_a) == len(ev_b) > 20
    for f in ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant"):
        assert np.array_equal(ev_a[f], ev_b[f]), f
    inp["fasta"].close()


def test_cross_file_checks_before_any_device_work(tmp_path):
    """check_num_seg_sites / check_sample_names (HC.java:327-404) and what the reference's site loop implies (HC.java:1383-1394,
    1418): every FASTA record needs exactly as many sites as the map file has rows; every tree node needs a record; every pheno
    sample must be a leaf with a record.  Both hosts refuse before touching the GPU."""
    t = synth.yule_tree(12, 1)
    S = 70
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 2), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 3)

    def both(paths, needle):
        r = run("-u", "-f", paths["fasta"], "-t", paths["tree"], "-p", paths["pheno"], "-m", paths["map"], "-r", "50", "-d", str(tmp_path / "o"),
                "-o", "x.out", ok=False)
        assert r.returncode == 255 and needle in r.stderr, r.stderr
        with pytest.raises(hostio.DataFormatError, match=needle):
            hostio.run_homoplasy_counter(paths["fasta"], paths["tree"], paths["pheno"], paths["map"], str(tmp_path / "py.out"), num_replicates=50)

    paths = synth.write_inputs(str(tmp_path / "a"), t, chars, sn, lab)
    with open(paths["map"], "a") as f:                                  # one map row too many
        f.write("1\tsnpX\t0\t999999\n")
    both(paths, "rows in the map file")
    paths = synth.write_inputs(str(tmp_path / "b"), t, chars, sn, lab)
    lines = open(paths["fasta"]).read().split("\n")
    lines[3] = lines[3][:-1]                                            # second record one site short: AIOOBE in the reference
    open(paths["fasta"], "w").write("\n".join(lines))
    both(paths, "sites of FASTA record")
    paths = synth.write_inputs(str(tmp_path / "c"), t, chars, sn, lab)
    lines = open(paths["fasta"]).read().split("\n")
    open(paths["fasta"], "w").write("\n".join(lines[2:]))               # first record gone
    both(paths, "has no FASTA record")
    paths = synth.write_inputs(str(tmp_path / "d"), t, chars, sn, lab)
    with open(paths["pheno"], "a") as f:
        f.write("not_a_sample\t1\n" + t.names[t.root] + "\t0\n")
    both(paths, "not_a_sample")


def test_no_cpu_fallback_in_the_native_host(tmp_path):
    """without a GPU the binary fails at pb_create (after the loaders accepted the inputs), it never computes on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is covered by the gpu test below")
    t = synth.yule_tree(20, 1)
    S = 70
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 2), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 3)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab)
    r = run("-u", "-f", paths["fasta"], "-t", paths["tree"], "-p", paths["pheno"], "-m", paths["map"], "-r", "50", "-d", str(tmp_path / "o"),
            "-o", "x.out", ok=False)
    assert r.returncode == 255 and "pb status -2" in r.stderr, r.stderr
    assert not os.path.exists(str(tmp_path / "o" / "x.out"))


# ---- files in, files out on the GPU ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("nexus,line_width,multi,tile", [(False, 0, 0.05, 640), (True, 70, 0.002, 640), (False, 60, 0.0, 1 << 17)])
def test_native_host_files_match_python_host_and_oracle(tmp_path, oracle_mod, nexus, line_width, multi, tile):
    t = synth.yule_tree(200, 11)
    S = 64 * 40 + 9
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 12, leaf_gap=multi > 0, multiallelic_frac=multi), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 13, na_frac=0.03, no_row_frac=0.02)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, line_width=line_width, nexus=nexus, seed=5)
    m = 800
    r = run("-u", "-f", paths["fasta"], "-t", paths["tree"], "-p", paths["pheno"], "-m", paths["map"], "-r", str(m), "-d", str(tmp_path / "o"),
            "-o", "run.out", "--seed", "21", "--tile-sites", str(tile), "--extras")
    assert "homoplasically informative sites" in r.stdout
    out = str(tmp_path / "o" / "run.out")
    py_out = str(tmp_path / "py.out")
    res = hostio.run_homoplasy_counter(paths["fasta"], paths["tree"], paths["pheno"], paths["map"], py_out, num_replicates=m, seed=21,
                                       tile_sites=tile, extras=True)
    assert ("# homoplasically informative sites = %d" % res["n_informative"]) in r.stdout
    assert ("%d tiles (%d compact)" % (res["n_tiles"], res["n_compact_tiles"])) in r.stdout
    for suffix in ("", ".sorted_by_a2_maxT", ".sorted_by_a1_maxT", ".extras"):
        assert open(out + suffix).read() == open(py_out + suffix).read(), suffix
    # and against the oracle's rows (node numbering of the parsed tree differs; results are per site)
    o = oracle_mod.Oracle(t.parent, t.leaf_nodes)
    ev, cc = o.count_all_homoplasy_events(chars, paths["physical_pos"].astype(np.int32))
    st, _, _ = o.assoc(sn, lab, m, 1, perms=None, seed=21, n_threads=4)
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
    # refuses to overwrite without -X (HC.java:507-615), like the reference
    args = ["-u", "-f", paths["fasta"], "-t", paths["tree"], "-p", paths["pheno"], "-m", paths["map"], "-r", "100", "-d", str(tmp_path / "o"),
            "-o", "run.out", "--seed", "21"]
    assert run(*args, ok=False).returncode == 1
    assert run(*args, "-X").returncode == 0
    # min_hcount too high: the reference's message and exit status (HC.java:3866-3869)
    r = run(*args, "-X", "-c", "100000", ok=False)
    assert r.returncode == 255 and "is set too high and has filtered out all segregating sites" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("real_gpus", [False, True])
def test_native_host_multi_gpu_in_one_process(tmp_path, monkeypatch, real_gpus):
    """`--devices d0,d1,...`: the single-process host (the shape of the reference's one JVM, HC.java:230-319) drives several
    site shards through pb_group and must write the same three files (+ extras) as the one-GPU run, byte for byte.
    real_gpus=False puts three shards on GPU 0 (test hook) so that a one-GPU box covers it too."""
    import torch
    n = torch.cuda.device_count()
    if real_gpus and n < 2:
        pytest.skip("needs 2 GPUs")
    devices = ",".join(str(d) for d in range(min(n, 8))) if real_gpus else "0,0,0"
    if not real_gpus:
        monkeypatch.setenv("PB_GROUP_ALLOW_SAME_DEVICE", "1")
    t = synth.yule_tree(300, 31)
    S = 64 * 90 + 17
    chars = synth.planes_to_chars(*synth.simulate_planes(t, S, 32, leaf_gap=True, multiallelic_frac=0.01), 0, S)
    sn, lab = synth.simulate_phenotypes(t, 33, na_frac=0.02, no_row_frac=0.02)
    paths = synth.write_inputs(str(tmp_path), t, chars, sn, lab, line_width=80, seed=6)
    common = ["-u", "-f", paths["fasta"], "-t", paths["tree"], "-p", paths["pheno"], "-m", paths["map"], "-r", "2000", "--seed", "77",
              "--tile-sites", "1024", "--extras", "-d", str(tmp_path / "o")]
    r1 = run(*common, "-o", "one.out")
    rn = run(*common, "-o", "multi.out", "--devices", devices)
    assert "homoplasically informative sites" in rn.stdout
    for suffix in ("", ".sorted_by_a2_maxT", ".sorted_by_a1_maxT", ".extras"):
        assert open(str(tmp_path / "o" / "multi.out") + suffix).read() == open(str(tmp_path / "o" / "one.out") + suffix).read(), suffix
