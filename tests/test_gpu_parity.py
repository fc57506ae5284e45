"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(libpoutine_b200.so via poutine_b200.HomoplasyCounter), against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): event counts, descendant tallies, observed statistics and — in
replay mode and in Philox mode alike, since the oracle restates the sampling spec — r / maxT counts are
BIT-EXACT; binomial p-values are compared bit-for-bit as well (the device chain is the same
commons-math restatement compiled without FMA), which is stricter than the stated 1e-12 relative.
"""
import json
import os

import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import synth
from tests.helpers import assert_stats_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    pb.build()


def run_gpu_events(tree, chars=None, planes=None, n_sites=None, **kw):
    hc = pb.HomoplasyCounter(**kw)
    hc.set_tree(tree.parent if hasattr(tree, "parent") else tree[0], tree.is_leaf if hasattr(tree, "is_leaf") else tree[1])
    if chars is not None:
        hc.set_seg_sites(chars)
    else:
        hc.set_planes(*planes, n_sites)
    ev, cc = hc.count_all_homoplasy_events()
    return hc, ev, cc


def assert_events_equal(gpu_ev, gpu_cc, or_ev, or_cc):
    assert gpu_cc["mono"] == or_cc["mono"] and gpu_cc["bi"] == or_cc["bi"] and gpu_cc["tri"] == or_cc["tri"]
    assert gpu_cc["quad"] == or_cc["quad"] and gpu_cc["no_allele"] == or_cc["other"]
    assert len(gpu_ev) == len(or_ev)
    for f in ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant"):
        np.testing.assert_array_equal(gpu_ev[f], or_ev[f], err_msg=f)


def gpu_mrca_sets(hc, tree, ev_all):
    """Rebuild the per-site MRCA buckets from the raw device event list (+ the root rule)."""
    site, node = hc.raw_events()
    return site, node


# -------------------------------------------------------------------------------------------------
# K1: events
# -------------------------------------------------------------------------------------------------
def test_hand_fixture_events(oracle_mod, golden_dir):
    with open(os.path.join(golden_dir, "events_hand.json")) as f:
        fx = json.load(f)
    N = len(fx["names"])
    chars = np.array([[ord(s["column"][v]) for s in fx["sites"]] for v in range(N)], dtype=np.uint8)
    parent = np.array(fx["parent"], dtype=np.int32)
    is_leaf = np.zeros(N, dtype=np.uint8)
    is_leaf[fx["leaves"]] = 1
    hc, ev, cc = run_gpu_events((parent, is_leaf), chars=chars, num_replicates=10)
    got = {int(e["segsite_id"]): e for e in ev}
    for i, s in enumerate(fx["sites"]):
        if s["expected"] is None:
            assert i + 1 not in got, s["name"]
        else:
            e = got[i + 1]
            tup = [chr(e["allele1"]), chr(e["allele2"]), int(e["a1_count"]), int(e["a2_count"]), int(e["a1_extant"]), int(e["a2_extant"])]
            assert tup == s["expected"], s["name"]
    want = fx["class_counts"]
    assert (cc["mono"], cc["bi"], cc["tri"], cc["quad"], cc["no_allele"]) == (want["mono"], want["bi"], want["tri"], want["quad"], want["other"])


@pytest.mark.parametrize("suffix", ["", "_extra"])
def test_events_match_reference_bytecode(golden_dir, suffix):
    """CUDA path against tests/golden/events_refbytecode.json (and the second set, `_extra`: star tree, ladder, two-leaf tree,
    few pheno rows, one-class phenotypes, min_hcount 4 — every case there also through the compacting upload, which keeps the
    tiles that qualify as one plane) — outputs of the reference's own compiled
    count_all_homoplasy_events (Homoplasy_Counter.class run by oracle/tools/refjvm.py): Homoplasy_Events fields, class
    counters and, where neither bucket was cleared, the exact MRCA node sets.  The all-valid biallelic case runs in both
    storage modes (k1x_events_kernel and k1_events_kernel)."""
    from tests.test_oracle_events import refbytecode_cases
    n_cases = 0
    for c, chars in refbytecode_cases(golden_dir, suffix):
        n_cases += 1
        parent = np.asarray(c["parent"], dtype=np.int32)
        is_leaf = np.zeros(len(parent), dtype=np.uint8)
        is_leaf[c["leaf_nodes"]] = 1
        names = np.asarray(c["names"])
        root = int(np.where(parent < 0)[0][0])
        S = chars.shape[1]
        modes = [dict(compact=False)] if c["name"] != "biallelic_all_valid" else [dict(compact=True), dict(compact=True, full_planes=True)]
        if suffix:
            modes = [dict(compact=False), dict(compact=True)]
        for mode in modes:
            hc = pb.HomoplasyCounter(num_replicates=10, full_planes=mode.get("full_planes", False))
            hc.set_tree(parent, is_leaf)
            hc.set_planes(*pb.pack_chars(chars), S, compact=mode["compact"])
            ev, cc = hc.count_all_homoplasy_events()
            if c["name"] == "biallelic_all_valid":
                assert hc.timings()["plane_mode"] == (2 if mode.get("full_planes") else 1)
            want_cc = c["class_counts"]
            assert (cc["bi"], cc["mono"], cc["tri"], cc["quad"], cc["no_allele"]) == tuple(want_cc[k] for k in ("bi", "mono", "tri", "quad", "other"))
            assert len(ev) == len(c["events"])
            site, node = hc.raw_events()
            sets = {}
            for s_, v in zip(site.tolist(), node.tolist()):
                sets.setdefault(s_ + 1, set()).add(v)
            for e, r in zip(ev, c["events"]):
                got = [int(e["segsite_id"]), chr(e["allele1"]), chr(e["allele2"]), int(e["a1_count"]), int(e["a2_count"]),
                       int(e["a1_extant"]), int(e["a2_extant"])]
                assert got == [r[k] for k in ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant")], (c["name"], r["segsite_id"])
                want = set(r["a1_mrcas"]) | set(r["a2_mrcas"])
                have = set(names[sorted(sets.get(r["segsite_id"], set()))].tolist())
                assert want - {names[root]} <= have, (c["name"], r["segsite_id"])
                if r["a1_count"] >= 2 and r["a2_count"] >= 2:
                    assert want - {names[root]} == have, (c["name"], r["segsite_id"])
            hc.close()
    assert n_cases >= 5


@pytest.mark.parametrize("suffix", ["", "_extra"])
def test_assoc_matches_reference_bytecode(golden_dir, suffix):
    """CUDA path in replay mode against tests/golden/assoc_refbytecode.json — what the reference's own compiled resampling
    driver produced (commons-math's BinomialTest from its jar included) with the recorded shuffles: obs_counts, r counts,
    per-replicate min p, sorted maxT, point-wise and family-wise estimates; every double bit for bit."""
    from tests.test_oracle_events import assoc_refbytecode_cases, assert_stats_equal_refbytecode
    for c, chars, a in assoc_refbytecode_cases(golden_dir, suffix):
        parent = np.asarray(c["parent"], dtype=np.int32)
        is_leaf = np.zeros(len(parent), dtype=np.uint8)
        is_leaf[c["leaf_nodes"]] = 1
        for compact in ((False, True) if (c["name"] == "biallelic_all_valid" or suffix) else (False,)):
            hc = pb.HomoplasyCounter(num_replicates=a["m"], min_hcount=a["min_hcount"], replay=True)
            hc.set_tree(parent, is_leaf)
            hc.set_phenos(np.asarray(a["sample_node"], dtype=np.int32), np.asarray(a["label"], dtype=np.uint8))
            hc.set_planes(*pb.pack_chars(chars), chars.shape[1], compact=compact)
            hc.count_all_homoplasy_events()
            st, mt = hc.assoc(replay_perms=np.asarray(a["perms"], dtype=np.uint32))
            p_success = hc.timings().get("p_success")
            if p_success is None:
                n1, n0 = sum(1 for x in a["label"] if x == 1), sum(1 for x in a["label"] if x == 0)
                p_success = np.float64(n1) / np.float64(n1 + n0)
            assert_stats_equal_refbytecode(st, mt, hc.maxT_unsorted(), p_success, a)
            hc.close()


@pytest.mark.parametrize("seed", range(10))
def test_events_nasty_random_trees(oracle_mod, seed):
    """polytomies, leaves on the root, shuffled node ids; gaps, lowercase, third alleles, ties, N at root"""
    tree = synth.random_tree(15 + 23 * seed, 300 + seed, polytomy=0.3, root_leaf_children=seed % 3)
    S = [1, 63, 64, 65, 200, 777, 1000, 129, 640, 2049][seed]
    chars = synth.simulate_chars(tree, S, 400 + seed, nasty=True)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    hc, ev, cc = run_gpu_events(tree, chars=chars, num_replicates=10)
    assert_events_equal(ev, cc, or_ev, or_cc)
    # MRCA node sets, not only their sizes: raw events (pre-clearing, root excluded) vs the oracle's buckets
    site, node = hc.raw_events()
    gpu_sets = {}
    for s, v in zip(site.tolist(), node.tolist()):
        gpu_sets.setdefault(s + 1, set()).add(v)
    assert len(site) == sum(len(v) for v in gpu_sets.values()), "duplicate raw events"
    root = tree.root
    for i, e in enumerate(or_ev):
        want = set(o.mrca(i, 0).tolist()) | set(o.mrca(i, 1).tolist())
        have = gpu_sets.get(int(e["segsite_id"]), set())
        # oracle buckets are post-clearing: every surviving member other than the root must be a raw event
        assert want - {root} <= have, (seed, i)
        if e["a1_count"] >= 2 and e["a2_count"] >= 2:
            assert want - {root} == have, (seed, i)


def test_events_c1_slice_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "c1_slice.npz"))
    S = int(z["n_sites"])
    hc, ev, cc = run_gpu_events((z["parent"], z["is_leaf"]), planes=(z["B0"], z["B1"], z["V"]), n_sites=S, num_replicates=10)
    gold = z["events"]
    assert_events_equal(ev, cc, gold, dict(zip(("mono", "bi", "tri", "quad", "other"), z["class_counts"].tolist())))


def test_events_tiled_upload_and_pitch(oracle_mod):
    """pb_put_planes column tiles with a foreign row pitch give the same result as one upload."""
    tree = synth.yule_tree(120, 5)
    S = 64 * 37 + 11
    B0, B1, V = synth.simulate_planes(tree, S, 6, leaf_gap=True, multiallelic_frac=0.05, third_allele_internal_frac=0.05)
    hc1, ev1, cc1 = run_gpu_events(tree, planes=(B0, B1, V), n_sites=S, num_replicates=10)
    hc2 = pb.HomoplasyCounter(num_replicates=10)
    hc2.set_tree(tree.parent, tree.is_leaf)
    hc2.set_planes(B0, B1, V, S, tile_words=5)   # slices of a wider array: pitch != n_words
    ev2, cc2 = hc2.count_all_homoplasy_events()
    assert ev1.tobytes() == ev2.tobytes() and cc1 == cc2
    # host rows already padded to the device pitch (W rounded up to 16 words): pb_put_planes takes the flat-copy path
    W = (S + 63) // 64
    Wp = (W + 15) & ~15
    padded = []
    for a in (B0, B1, V):
        h = np.full((tree.n_nodes, Wp), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)   # garbage in the padding words
        h[:, :W] = a
        padded.append(h[:, :W])
    hc3 = pb.HomoplasyCounter(num_replicates=10)
    hc3.set_tree(tree.parent, tree.is_leaf)
    hc3.set_planes(*padded, S)
    ev3, cc3 = hc3.count_all_homoplasy_events()
    assert ev1.tobytes() == ev3.tobytes() and cc1 == cc3
    chars = synth.planes_to_chars(B0, B1, V, 0, S)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    assert_events_equal(ev1, cc1, or_ev, or_cc)
    assert cc1["tri"] + cc1["quad"] > 0


@pytest.mark.parametrize("gaps,S", [(False, 64 * 41 + 11), (True, 64 * 41 + 11), (False, 64 * 32)])
def test_events_compact_biallelic_upload(oracle_mod, gaps, S):
    """pb_put_planes_biallelic (one plane over PCIe, expanded on the device) == pb_put_planes == oracle; tiles that do
    not qualify (a third base somewhere in the tile) go up as full planes in the same run; end-to-end through assoc."""
    tree = synth.yule_tree(150, 8)
    B0, B1, V = synth.simulate_planes(tree, S, 3, leaf_gap=gaps)
    if gaps:   # arbitrary B bits under V = 0, as a packer fed lower-case / N calls would leave them
        rng = np.random.default_rng(5)
        B0 = B0 ^ (rng.integers(0, 2**63, size=B0.shape, dtype=np.uint64) & ~V)
        B1 = B1 ^ (rng.integers(0, 2**63, size=B1.shape, dtype=np.uint64) & ~V)
    chars = synth.planes_to_chars(B0, B1, V, 0, S)
    site = 64 * 12 + 5                                              # one site of tile 2 (tile_words = 6) gets a third base
    third = [b for b in b"ACGT" if b not in set(chars[:, site].tolist())][0]
    chars[tree.leaf_nodes[3], site] = third
    B0, B1, V = pb.pack_chars(chars)
    sn, lab = synth.simulate_phenotypes(tree, 4)
    res = []
    for compact in (False, True):
        hc = pb.HomoplasyCounter(num_replicates=256, seed=3)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_phenos(sn, lab)
        hc.set_planes(B0, B1, V, S, tile_words=6, compact=compact)
        ev, cc = hc.count_all_homoplasy_events()
        assert hc.timings()["plane_mode"] == 2     # tile 2 is not compact: three planes (promoted from X when gaps is False)
        st, mt = hc.assoc()
        res.append((ev.tobytes(), cc, st.tobytes(), mt.tobytes()))
        if compact:
            o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
            or_ev, or_cc = o.count_all_homoplasy_events(chars)
            assert_events_equal(ev, cc, or_ev, or_cc)
        hc.close()
    assert res[0] == res[1]
    W = (S + 63) // 64
    n_compact = sum(pb.compact_planes(B0[:, w:w + 6], B1[:, w:w + 6], V[:, w:w + 6], min(S, (w + 6) * 64) - w * 64) is not None
                    for w in range(0, W, 6))
    assert n_compact == (W + 5) // 6 - 1                            # exactly the doctored tile was refused
    # whole-alignment compact upload with device-pitch rows (flat-copy path)
    chars[tree.leaf_nodes[3], site] = chars[tree.root, site]
    B0, B1, V = pb.pack_chars(chars)
    Wp = (W + 15) & ~15
    Xp = np.full((tree.n_nodes, Wp), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    X, colmask, all_valid = pb.compact_planes(B0, B1, V, S, X=Xp)
    assert all_valid == (not gaps)
    Vp = None
    if gaps:
        Vp = np.zeros((tree.n_nodes, Wp), dtype=np.uint64)
        Vp[:, :W] = V
        Vp = Vp[:, :W]
    hc = pb.HomoplasyCounter(num_replicates=16)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_sites(S)
    hc.put_planes_biallelic(X, Vp, colmask)
    ev, cc = hc.count_all_homoplasy_events()
    assert hc.timings()["plane_mode"] == (2 if gaps else 1)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    assert_events_equal(ev, cc, or_ev, or_cc)
    hc.close()


@pytest.mark.parametrize("n_leaves,S,chunks", [(150, 64 * 41 + 11, None), (700, 500, "1"), (700, 500, "21"), (1300, 64 * 300, None)])
def test_events_one_plane_mode(oracle_mod, monkeypatch, n_leaves, S, chunks):
    """Every tile compact with all genotypes valid: the alignment stays ONE plane in HBM and k1x_events_kernel runs.
    Events, raw MRCA list, CSR and the whole assoc output are identical to the three-plane path and to the oracle."""
    if chunks:
        monkeypatch.setenv("PB_K1_CHUNKS", chunks)
    tree = synth.yule_tree(n_leaves, 8) if n_leaves != 700 else synth.random_tree(700, 4, polytomy=0.3)
    B0, B1, V = synth.simulate_planes(tree, S, 3)
    sn, lab = synth.simulate_phenotypes(tree, 4, na_frac=0.02, no_row_frac=0.02)
    res = []
    for full in (False, True):
        hc = pb.HomoplasyCounter(num_replicates=256, seed=3, full_planes=full, site_offset=640)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_phenos(sn, lab)
        hc.set_planes(B0, B1, V, S, tile_words=7, compact=True)
        ev, cc = hc.count_all_homoplasy_events()
        assert hc.timings()["plane_mode"] == (2 if full else 1)
        rs, rn = hc.raw_events()
        order = np.lexsort((rn, rs))
        off, idx = hc.event_csr()
        csr = [sorted(idx[off[i]:off[i + 1]].tolist()) for i in range(len(off) - 1)]
        st, mt = hc.assoc()
        res.append((ev.tobytes(), cc, rs[order].tobytes(), rn[order].tobytes(), csr, st.tobytes(), mt.tobytes()))
        if not full:
            o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
            or_ev, or_cc = o.count_all_homoplasy_events(synth.planes_to_chars(B0, B1, V, 0, S))
            e2 = ev.copy()
            e2["segsite_id"] -= 640
            assert_events_equal(e2, cc, or_ev, or_cc)
        hc.close()
    assert res[0] == res[1]


@pytest.mark.parametrize("seed,density", [(0, "dense"), (1, "dense"), (2, "sparse"), (3, "sparse"), (4, "dense")])
def test_one_plane_mode_census_from_records(oracle_mod, seed, density):
    """k1x_events_kernel keeps no leaf counters: the census is L * X(root) + the signed sum, over the records, of
    leaves_below(child).  Topologies that stress it: leaves hanging off the root (their records are census-only and must
    never become MRCAs or extant events), polytomies, shuffled node ids; dense columns put flips of both directions into
    one (row, word) cell (two records) and make X(root) = 1 at half of the sites.  Oracle parity + identity with the
    three-plane path for events, raw MRCA list, CSR and the assoc output."""
    tree = synth.random_tree(40 + 37 * seed, 900 + seed, polytomy=0.3, root_leaf_children=1 + seed % 3)
    S = [64 * 3 + 5, 64 * 9, 700, 64 * 17 + 63, 129][seed]
    rng = np.random.default_rng(50 + seed)
    N = tree.n_nodes
    if density == "dense":
        X = rng.integers(0, 2, size=(N, S)).astype(np.uint8)
    else:                                     # a few flipped subtrees per site, plus single flipped leaves
        X = np.zeros((N, S), dtype=np.uint8)
        X[:, rng.random(S) < 0.5] = 1         # root state
        flip = rng.random((N, S)) < 2.5 / N
        par = np.asarray(tree.parent)
        kids = {}
        for v in range(N):
            if par[v] >= 0:
                kids.setdefault(int(par[v]), []).append(v)
        todo = [int(tree.root)]                # X(v) = X(parent) ^ flip(v), parents before children
        while todo:
            u = todo.pop()
            for c in kids.get(u, []):
                X[c] = X[u] ^ flip[c]
                todo.append(c)
    first = rng.integers(0, 4, size=S)
    second = (first + 1 + rng.integers(0, 3, size=S)) % 4
    pair = np.frombuffer(b"ACGT", dtype=np.uint8)[np.stack([first, second])]
    chars = np.where(X == 1, pair[1][None, :], pair[0][None, :]).astype(np.uint8)
    B0, B1, V = pb.pack_chars(chars)
    sn, lab = synth.simulate_phenotypes(tree, 60 + seed, na_frac=0.03, no_row_frac=0.05)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    res = []
    for full in (False, True):
        hc = pb.HomoplasyCounter(num_replicates=128, seed=9, full_planes=full, min_hcount=seed % 2)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_phenos(sn, lab)
        hc.set_planes(B0, B1, V, S, tile_words=5, compact=True)
        ev, cc = hc.count_all_homoplasy_events()
        assert hc.timings()["plane_mode"] == (2 if full else 1)
        assert_events_equal(ev, cc, or_ev, or_cc)
        rs, rn = hc.raw_events()
        assert not np.any(np.asarray(tree.parent)[rn] == tree.root), "a child of the root is never an MRCA (HC.java:1674)"
        order = np.lexsort((rn, rs))
        off, idx = hc.event_csr()
        csr = [sorted(idx[off[i]:off[i + 1]].tolist()) for i in range(len(off) - 1)]
        try:
            st, mt = hc.assoc()
            tail = (st.tobytes(), mt.tobytes())
        except pb.NoInformativeSites:
            tail = None
        res.append((ev.tobytes(), cc, rs[order].tobytes(), rn[order].tobytes(), csr, tail))
        hc.close()
    assert res[0] == res[1]


def test_one_plane_mode_partial_upload_and_reuse(oracle_mod):
    """Columns nobody uploaded hold no valid genotype in either storage mode; a context re-used for a second alignment of
    the same shape keeps working; a full-plane tile arriving late promotes the context without losing earlier tiles."""
    tree = synth.yule_tree(90, 2)
    S = 64 * 20 + 3
    B0, B1, V = synth.simulate_planes(tree, S, 5)
    out = []
    for full in (False, True):
        hc = pb.HomoplasyCounter(num_replicates=16, full_planes=full)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_sites(S)
        c = pb.compact_planes(B0[:, 4:15], B1[:, 4:15], V[:, 4:15], 64 * 11)
        hc.put_planes_biallelic(c[0], None, c[1], 4)                # only words 4..14
        ev, cc = hc.count_all_homoplasy_events()
        assert cc["no_allele"] == S - 64 * 11
        out.append((ev.tobytes(), cc))
        hc.set_planes(B0, B1, V, S, compact=True)                   # same shape again, now everything
        ev2, cc2 = hc.count_all_homoplasy_events()
        out.append((ev2.tobytes(), cc2))
        hc.close()
    assert out[0] == out[2] and out[1] == out[3]
    # late promotion: tiles 0..1 compact (X mode), then a three-plane tile
    hc = pb.HomoplasyCounter(num_replicates=16)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_sites(S)
    W = (S + 63) // 64
    for w0 in (0, 7):
        c = pb.compact_planes(B0[:, w0:w0 + 7], B1[:, w0:w0 + 7], V[:, w0:w0 + 7], 64 * 7)
        hc.put_planes_biallelic(c[0], None, c[1], w0)
    hc.put_planes(B0[:, 14:W], B1[:, 14:W], V[:, 14:W], 14)
    ev3, cc3 = hc.count_all_homoplasy_events()
    assert hc.timings()["plane_mode"] == 2
    assert (ev3.tobytes(), cc3) == out[1]
    hc.close()
    # ... and a late COMPACT tile that carries a V plane (invalid calls, still at most two bases) promotes as well
    V2 = V.copy()
    V2[tree.leaf_nodes[::7], 14:W] &= np.uint64(0xF0F0F0F0F0F0F0F0)
    res = []
    for compact in (True, False):
        hc = pb.HomoplasyCounter(num_replicates=16)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_sites(S)
        if compact:
            for w0 in (0, 7):
                c = pb.compact_planes(B0[:, w0:w0 + 7], B1[:, w0:w0 + 7], V2[:, w0:w0 + 7], 64 * 7)
                assert c[2]
                hc.put_planes_biallelic(c[0], None, c[1], w0)
            c = pb.compact_planes(B0[:, 14:W], B1[:, 14:W], V2[:, 14:W], S - 64 * 14)
            assert c is not None and not c[2]
            hc.put_planes_biallelic(c[0], V2[:, 14:W], c[1], 14)
        else:
            hc.put_planes(B0, B1, V2)
        ev4, cc4 = hc.count_all_homoplasy_events()
        assert hc.timings()["plane_mode"] == 2
        res.append((ev4.tobytes(), cc4))
        hc.close()
    assert res[0] == res[1]
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(synth.planes_to_chars(B0, B1, V2, 0, S))
    assert_events_equal(np.frombuffer(res[0][0], dtype=ev3.dtype), res[0][1], or_ev, or_cc)


def test_events_many_chunks_and_site_offset(oracle_mod, monkeypatch):
    """force the edge schedule into many chunks (census partials are summed across chunks)"""
    tree = synth.yule_tree(700, 11)
    S = 500
    B0, B1, V = synth.simulate_planes(tree, S, 12, leaf_gap=True)
    chars = synth.planes_to_chars(B0, B1, V, 0, S)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    for chunks in ("1", "7", "21"):
        monkeypatch.setenv("PB_K1_CHUNKS", chunks)
        hc, ev, cc = run_gpu_events(tree, planes=(B0, B1, V), n_sites=S, num_replicates=10, site_offset=6400)
        ev = ev.copy()
        ev["segsite_id"] -= 6400
        assert_events_equal(ev, cc, or_ev, or_cc)


def test_events_caterpillar_tree(oracle_mod):
    """ladder topology (depth = #leaves): every internal node is re-visited as the current parent"""
    tree = synth.caterpillar_tree(400, 5)
    S = 300
    chars = synth.simulate_chars(tree, S, 6, nasty=True)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    hc, ev, cc = run_gpu_events(tree, chars=chars, num_replicates=10)
    assert_events_equal(ev, cc, or_ev, or_cc)


def test_events_census_segments(oracle_mod, monkeypatch):
    """one chunk holding > 248 leaves: the census is flushed in several segments and re-summed"""
    monkeypatch.setenv("PB_K1_CHUNKS", "1")
    tree = synth.yule_tree(1100, 13)
    S = 130
    B0, B1, V = synth.simulate_planes(tree, S, 14, leaf_gap=True, multiallelic_frac=0.1)
    chars = synth.planes_to_chars(B0, B1, V, 0, S)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    hc, ev, cc = run_gpu_events(tree, planes=(B0, B1, V), n_sites=S, num_replicates=10)
    assert_events_equal(ev, cc, or_ev, or_cc)


def test_events_dense_overflowing_queue(oracle_mod):
    """every edge an event at every site: shared queue overflow path + event-buffer regrow"""
    tree = synth.yule_tree(300, 21)
    S = 64 * 40
    rng = np.random.default_rng(3)
    N = tree.n_nodes
    chars = np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, size=(N, S))]
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars)
    hc, ev, cc = run_gpu_events(tree, chars=chars, num_replicates=10)
    assert_events_equal(ev, cc, or_ev, or_cc)
    assert cc["raw_events"] > 4 * S


def test_events_dense_regrow_record_list_and_csr(oracle_mod):
    """more than 2^20 raw records (every node differs from its parent in half of the sites of every word) and more extant
    events than the CSR buffer was sized for: pb_run_events grows both and repeats the pass; one-plane and three-plane mode"""
    tree = synth.yule_tree(600, 22)
    S = 64 * 900
    rng = np.random.default_rng(4)
    N = tree.n_nodes
    chars = np.frombuffer(b"GT", dtype=np.uint8)[rng.integers(0, 2, size=(N, S), dtype=np.uint8)]
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars, n_threads=8)
    sn, lab = synth.simulate_phenotypes(tree, 5)
    B0, B1, V = pb.pack_chars(chars)
    out = []
    for full in (False, True):
        hc = pb.HomoplasyCounter(num_replicates=4, seed=1, full_planes=full)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_phenos(sn, lab)
        hc.set_planes(B0, B1, V, S, compact=True)
        ev, cc = hc.count_all_homoplasy_events()
        assert hc.timings()["plane_mode"] == (2 if full else 1)
        assert_events_equal(ev, cc, or_ev, or_cc)
        assert cc["raw_events"] > 30_000_000
        tm = hc.timings()
        assert tm["n_extant_events"] > (1 << 20)                       # beyond the initial CSR capacity
        off, idx = hc.event_csr()
        assert off[-1] == tm["n_extant_events"] and idx.min() >= 0 and idx.max() < len(lab)
        st, _ = hc.assoc()
        np.testing.assert_array_equal(st["obs_counts"].sum(axis=1), ev["a1_extant"] + ev["a2_extant"])   # all leaves have a 0/1 label
        out.append((ev.tobytes(), st["obs_counts"].tobytes(), st["r_a1"].tobytes(), st["r_a2"].tobytes()))
        hc.close()
    assert out[0] == out[1]


# -------------------------------------------------------------------------------------------------
# K2/K3/K4: association
# -------------------------------------------------------------------------------------------------
def _assoc_case(oracle_mod, n_leaves, S, m, tseed, na_frac, no_row_frac, min_hcount=1, nasty=False, **kw):
    tree = synth.yule_tree(n_leaves, tseed)
    if nasty:
        chars = synth.simulate_chars(tree, S, tseed + 1, nasty=True)
    else:
        chars = synth.planes_to_chars(*synth.simulate_planes(tree, S, tseed + 1), 0, S)
    sn, lab = synth.simulate_phenotypes(tree, tseed + 2, na_frac=na_frac, no_row_frac=no_row_frac)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, _ = o.count_all_homoplasy_events(chars)
    hc = pb.HomoplasyCounter(num_replicates=m, min_hcount=min_hcount, **kw)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_seg_sites(chars)
    hc.count_all_homoplasy_events()
    hc.set_phenos(sn, lab)
    hc.count_all_homoplasy_events()   # labels reset the CSR: events must be re-run (documented call order)
    return tree, o, hc, sn, lab


@pytest.mark.parametrize("na_frac,no_row_frac,force_general,min_hcount", [
    (0.0, 0.0, False, 1),     # all labelled: bit-sliced threshold path
    (0.0, 0.0, True, 1),      # same inputs through the general scalar path
    (0.05, 0.03, False, 1),   # missing labels + leaves without a pheno row
    (0.0, 0.1, False, 2),
    (0.05, 0.0, False, 0),    # min_hcount 0: every biallelic site informative, n may be 0
])
def test_assoc_replay_bit_exact(oracle_mod, na_frac, no_row_frac, force_general, min_hcount):
    m = 700
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 150, 900, m, 31, na_frac, no_row_frac, min_hcount=min_hcount,
                                       replay=True, force_general_null=force_general)
    perms = oracle_mod.java_shuffle_perms(77, m, len(lab))
    st, mt_s, mt_u = o.assoc(sn, lab, m, min_hcount, perms=perms)
    g_st, g_mt = hc.assoc(replay_perms=perms)
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), mt_u.view(np.uint64))
    tm = hc.timings()
    if force_general:
        assert tm["n_fast_alleles"] == 0
    elif na_frac == 0:
        assert tm["n_fast_alleles"] > 0


@pytest.mark.parametrize("na_frac", [0.0, 0.04])
def test_assoc_philox_bit_exact(oracle_mod, na_frac):
    """Philox mode: the oracle restates the sampling spec, so r/maxT are bit-exact here too."""
    m = 1500
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 200, 1200, m, 41, na_frac, 0.02, seed=0xC0FFEE1234)
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=0xC0FFEE1234, n_threads=4)
    g_st, g_mt = hc.assoc()
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), mt_u.view(np.uint64))


def test_assoc_philox_replayed_groups(oracle_mod, monkeypatch):
    """the label generator's fast groups skip the per-draw Lemire rejection test and replay a group through the
    exact path when a retry may be needed; forcing the replay for every group must not change a bit"""
    m = 1100
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 180, 700, m, 43, 0.0, 0.0, seed=0xABCDEF0123)
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=0xABCDEF0123, n_threads=4)
    g_st, g_mt = hc.assoc()
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))
    g_st, g_mt = g_st.copy(), g_mt.copy()
    monkeypatch.setenv("PB_GEN_FORCE_REPLAY", "1")
    hc.count_all_homoplasy_events()
    g_st2, g_mt2 = hc.assoc()
    assert g_st2.tobytes() == g_st.tobytes() and g_mt2.tobytes() == g_mt.tobytes()


@pytest.mark.parametrize("na_frac,replay,gen_kb", [(0.0, False, None), (0.03, False, "128"), (0.0, True, "128"), (0.0, False, "16")])
def test_assoc_multi_tile(oracle_mod, monkeypatch, na_frac, replay, gen_kb):
    """replicates split into several generation tiles (double-buffered on the second stream), each counted in
    several L2-sized column tiles"""
    monkeypatch.setenv("PB_K3_TILE_KB", "16")
    if gen_kb:
        monkeypatch.setenv("PB_K3_GEN_KB", gen_kb)
    m = 7000                                          # 2048-replicate tiles -> 4 tiles, the last one partial
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 150, 500, m, 91, na_frac, 0.0, seed=77, replay=replay)
    if replay:
        perms = oracle_mod.java_shuffle_perms(3, m, len(lab))
        st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=perms, n_threads=4)
        g_st, g_mt = hc.assoc(replay_perms=perms)
    else:
        st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=77, n_threads=4)
        g_st, g_mt = hc.assoc()
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), mt_u.view(np.uint64))
    # a second pass on the same context (pre-generated tile 0 from pb_run_events) gives the same answer
    g_st, g_mt = g_st.copy(), g_mt.copy()             # results are views of reused page-locked buffers
    hc.count_all_homoplasy_events()
    g_st2, g_mt2 = hc.assoc(replay_perms=perms) if replay else hc.assoc()
    assert g_st2.tobytes() == g_st.tobytes() and g_mt2.tobytes() == g_mt.tobytes()


@pytest.mark.parametrize("na_frac", [0.0, 0.05])
def test_assoc_long_event_lists(oracle_mod, na_frac):
    """heavily homoplasic sites: alleles with n from 1 to > 64 extant events (every list-length class of K3b)"""
    m = 400
    tree = synth.yule_tree(700, 101)
    S = 128
    chars = synth.planes_to_chars(*synth.simulate_planes(tree, S, 102, mean_extra_mut=700.0), 0, S)
    chars2 = synth.planes_to_chars(*synth.simulate_planes(tree, S, 103, mean_extra_mut=12.0), 0, S)
    chars = np.ascontiguousarray(np.concatenate([chars, chars2], axis=1))
    sn, lab = synth.simulate_phenotypes(tree, 104, na_frac=na_frac)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, _ = o.count_all_homoplasy_events(chars)
    assert max(or_ev["a1_extant"].max(), or_ev["a2_extant"].max()) > 64
    hc = pb.HomoplasyCounter(num_replicates=m, seed=5)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_seg_sites(chars)
    ev, cc = hc.count_all_homoplasy_events()
    for f in ("a1_count", "a2_count", "a1_extant", "a2_extant"):
        np.testing.assert_array_equal(ev[f], or_ev[f], err_msg=f)
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=5, n_threads=4)
    g_st, g_mt = hc.assoc()
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))


def test_assoc_nasty_inputs_replay(oracle_mod):
    m = 300
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 90, 700, m, 51, 0.05, 0.05, nasty=True, replay=True)
    perms = oracle_mod.java_shuffle_perms(5, m, len(lab))
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=perms)
    g_st, g_mt = hc.assoc(replay_perms=perms)
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))


@pytest.mark.parametrize("tau", [1e-300, 1.0])
def test_assoc_maxt_threshold_extremes(oracle_mod, tau):
    """tau tiny: every replicate goes through the maxT fallback kernel; tau = 1: every allele is a candidate."""
    m = 256
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 120, 600, m, 61, 0.0, 0.0, replay=True, maxt_tau=tau)
    perms = oracle_mod.java_shuffle_perms(9, m, len(lab))
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=perms)
    g_st, g_mt = hc.assoc(replay_perms=perms)
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), mt_u.view(np.uint64))
    if tau < 1e-100:
        assert hc.timings()["n_maxt_fallback_reps"] == m


def test_assoc_c1_slice_golden(oracle_mod, golden_dir):
    z = np.load(os.path.join(golden_dir, "c1_slice.npz"))
    S, m = int(z["n_sites"]), int(z["m"])
    hc = pb.HomoplasyCounter(num_replicates=m, replay=True)
    hc.set_tree(z["parent"], z["is_leaf"])
    hc.set_phenos(z["sample_node"], z["label"])
    hc.set_planes(z["B0"], z["B1"], z["V"], S)
    hc.count_all_homoplasy_events()
    perms = oracle_mod.java_shuffle_perms(int(z["perm_seed"]), m, len(z["label"]))
    g_st, g_mt = hc.assoc(replay_perms=perms)
    assert_stats_equal(g_st, z["stats"])
    np.testing.assert_array_equal(g_mt.view(np.uint64), z["maxT_sorted"].view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), z["maxT_unsorted"].view(np.uint64))


def test_ptable_bit_exact_vs_commons_math_port(oracle_mod, golden_dir):
    """device f[n][k] == the oracle's commons-math restatement (itself pinned to the jar bytecode)."""
    m = 64
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 400, 800, m, 71, 0.0, 0.0)
    hc.assoc()
    n_max, tab = hc.ptable()
    p = float((lab == 1).sum()) / float((lab == 1).sum() + (lab == 0).sum())
    ref = oracle_mod.binomial_table(n_max, p)
    np.testing.assert_array_equal(tab.view(np.uint64), ref.view(np.uint64))
    assert n_max >= 2


def test_ptable_golden_values_on_device(oracle_mod, golden_dir):
    """the minijvm-generated golden vectors (commons-math 3.6.1 bytecode) at p = 0.379, through the device table"""
    with open(os.path.join(golden_dir, "binom_golden.json")) as f:
        t = json.load(f)["table"]
    p, g_nmax = float.fromhex(t["p"]), t["n_max"]
    assert p == 0.379
    gold = np.array([float.fromhex(x) for x in t["f"]])
    # 379 cases of 1000 samples -> calc_background_p_success = 379/1000 = 0.379 exactly as a double
    tree = synth.yule_tree(1000, 81)
    chars = synth.planes_to_chars(*synth.simulate_planes(tree, 3000, 82, mean_extra_mut=6.0), 0, 3000)
    lab = np.zeros(1000, dtype=np.uint8)
    lab[:379] = 1
    hc = pb.HomoplasyCounter(num_replicates=32)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(tree.leaf_nodes, lab)
    hc.set_seg_sites(chars)
    hc.count_all_homoplasy_events()
    hc.assoc()
    n_max, tab = hc.ptable()
    n_cmp = min(n_max, g_nmax)
    assert n_cmp >= 4
    cnt = (n_cmp + 1) * (n_cmp + 2) // 2
    np.testing.assert_array_equal(tab[:cnt].view(np.uint64), gold[:cnt].view(np.uint64))


def test_degenerate_trees(oracle_mod):
    """star tree (every leaf hangs off the root: no edge is ever tested), a 3-node tree, one site"""
    for parent in ([-1] + [0] * 9, [-1, 0, 0]):
        parent = np.array(parent, dtype=np.int32)
        N = len(parent)
        is_leaf = np.ones(N, dtype=np.uint8)
        is_leaf[0] = 0
        leaves = np.arange(1, N, dtype=np.int32)
        rng = np.random.default_rng(N)
        for S in (1, 70):
            chars = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=(N, S))]
            o = oracle_mod.Oracle(parent, leaves)
            or_ev, or_cc = o.count_all_homoplasy_events(chars)
            hc, ev, cc = run_gpu_events((parent, is_leaf), chars=chars, num_replicates=10)
            assert_events_equal(ev, cc, or_ev, or_cc)
            assert cc["raw_events"] == 0
            hc.set_phenos(leaves, (np.arange(N - 1) % 2).astype(np.uint8))
            hc.count_all_homoplasy_events()
            with pytest.raises(pb.NoInformativeSites):       # nothing can be homoplasic on a star
                hc.assoc()


@pytest.mark.parametrize("labels", ["all_cases", "all_controls", "one_sample"])
def test_assoc_degenerate_phenotypes(oracle_mod, labels):
    """p_success = 1, 0, or a single phenotyped sample: the commons-math chain goes through log(0) / n = 0 rows;
    whatever it yields (NaN included), device and oracle must agree bit for bit and nothing may hang"""
    m = 200
    tree = synth.yule_tree(80, 7)
    chars = synth.planes_to_chars(*synth.simulate_planes(tree, 400, 8, mean_extra_mut=4.0), 0, 400)
    if labels == "one_sample":
        sn, lab = tree.leaf_nodes[:1].copy(), np.array([1], dtype=np.uint8)
    else:
        sn = tree.leaf_nodes.copy()
        lab = np.full(len(sn), 1 if labels == "all_cases" else 0, dtype=np.uint8)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    o.count_all_homoplasy_events(chars)
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=3, n_threads=2)
    hc = pb.HomoplasyCounter(num_replicates=m, seed=3)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_seg_sites(chars)
    hc.count_all_homoplasy_events()
    g_st, g_mt = hc.assoc()
    assert len(g_st) == len(st)
    for f in ("obs_counts", "a1_in_use", "a2_in_use"):
        np.testing.assert_array_equal(g_st[f], st[f], err_msg=f)
    for f in ("obs_p_a1", "obs_p_a2"):
        np.testing.assert_array_equal(g_st[f].view(np.uint64), st[f].view(np.uint64), err_msg=f)
    for f in ("r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
        np.testing.assert_array_equal(g_st[f], st[f], err_msg=f)


@pytest.mark.parametrize("na_frac,min_hcount", [(0.0, 1), (0.05, 1), (0.03, 0)])
def test_extras_exact_pointwise_and_fisher(oracle_mod, na_frac, min_hcount):
    """pb_run_extras (parity unpinned — no reference code can run): the exact label-permutation null of the point-wise
    statistic and R-style two-sided Fisher, against exact rational arithmetic, relative tolerance 1e-12; and the
    Monte-Carlo r/m of the live path inside a 6-sigma binomial band around the exact value."""
    from tests.helpers import exact_pointwise_null, fisher_two_sided
    tree = synth.yule_tree(260, 31)
    S, m = 1500, 6000
    B0, B1, V = synth.simulate_planes(tree, S, 32, leaf_gap=True)
    sn, lab = synth.simulate_phenotypes(tree, 33, na_frac=na_frac, no_row_frac=0.02)
    hc = pb.HomoplasyCounter(num_replicates=m, seed=5, min_hcount=min_hcount)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, S)
    hc.count_all_homoplasy_events()
    st, _ = hc.assoc()
    ep, fp = hc.extras()
    off, idx = hc.event_csr()
    nmax, tab = hc.ptable()
    f = lambda n, k: tab[n * (n + 1) // 2 + k]
    P, n_case, n_miss = len(lab), int((lab == 1).sum()), int((lab == 2).sum())
    assert (n_miss > 0) == (na_frac > 0)
    checked = 0
    for i in range(len(st)):
        oc = [int(x) for x in st["obs_counts"][i]]
        want = float(fisher_two_sided(*oc))
        assert abs(fp[i] - want) <= 1e-12 * want, (i, oc, fp[i], want)
        for a, name in enumerate(("a1", "a2")):
            if not st[name + "_in_use"][i]:
                assert np.isnan(ep[i, a])
                continue
            n = int(off[2 * i + a + 1] - off[2 * i + a])
            p_obs = st["obs_p_" + name][i]
            want = float(exact_pointwise_null(P, n_case, n_miss, n, p_obs, f))
            assert abs(ep[i, a] - want) <= 1e-12 * max(want, 1e-300), (i, a, n, ep[i, a], want)
            mc = st["r_" + name][i] / m
            sd = np.sqrt(max(want * (1 - want), 1e-12) / m)
            assert abs(mc - want) < 6 * sd + 2.0 / m, (i, a, mc, want)
            checked += 1
    assert checked > 200
    hc.close()


def test_no_informative_sites_error(oracle_mod):
    tree = synth.yule_tree(50, 3)
    chars = synth.simulate_chars(tree, 40, 4, nasty=False)
    sn, lab = synth.simulate_phenotypes(tree, 5)
    hc = pb.HomoplasyCounter(num_replicates=10, min_hcount=1000)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_seg_sites(chars)
    hc.count_all_homoplasy_events()
    with pytest.raises(pb.NoInformativeSites):      # HC.java:3866-3869
        hc.assoc()


def test_call_order_errors():
    hc = pb.HomoplasyCounter(num_replicates=10)
    with pytest.raises(pb.PoutineError):
        hc.set_sites(100)                            # no tree yet
    tree = synth.yule_tree(20, 1)
    hc.set_tree(tree.parent, tree.is_leaf)
    with pytest.raises(pb.PoutineError):
        hc.set_phenos(np.array([0], dtype=np.int32), np.array([1], dtype=np.uint8))   # node 0 is the root, not a leaf
    hc.set_sites(100)
    hc.count_all_homoplasy_events()
    with pytest.raises(pb.PoutineError):
        hc.assoc()                                   # no labels


# -------------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE config C2 shape, scaled in S to keep the test short)
# -------------------------------------------------------------------------------------------------
def test_large_properties_and_slice_parity(oracle_mod):
    tree = synth.yule_tree(5000, 20262)
    S = 200_000
    B0, B1, V = synth.simulate_planes(tree, S, 20262 + 2000)
    sn, lab = synth.simulate_phenotypes(tree, 20262 + 1000)
    m = 20_000
    hc = pb.HomoplasyCounter(num_replicates=m, seed=7)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, S)
    ev, cc = hc.count_all_homoplasy_events()
    assert cc["mono"] + cc["bi"] + cc["tri"] + cc["quad"] + cc["no_allele"] == S
    # invariants of the reference code (SURVEY §4 i-iii)
    assert np.all((ev["a1_count"] == 0) | (ev["a1_count"] >= 2)) and np.all((ev["a2_count"] == 0) | (ev["a2_count"] >= 2))
    assert np.all(ev["a1_extant"] <= ev["a1_count"]) and np.all(ev["a2_extant"] <= ev["a2_count"])
    st, mt = hc.assoc()
    evs = hc.sites[st["site_index"]]
    expect = np.where(st["a1_in_use"] == 1, evs["a1_extant"], 0) + np.where(st["a2_in_use"] == 1, evs["a2_extant"], 0)
    np.testing.assert_array_equal(st["obs_counts"].sum(axis=1), expect)     # check_extant_nodes_only_code (HC.java:4345-4369)
    for a in ("a1", "a2"):
        use = st[a + "_in_use"] == 1
        r = st["r_" + a]
        assert np.all((r[use] >= 0) & (r[use] <= m)) and np.all(r[~use] == -1)
        np.testing.assert_array_equal(st["pointwise_" + a][use], (r[use] + 1.0) / (m + 1.0))
        np.testing.assert_array_equal(st["familywise_" + a][use], (st["r_maxT_" + a][use] + 1.0) / (m + 1.0))
        # pooled min over all alleles <= this allele's own p  =>  r_maxT >= r
        assert np.all(st["r_maxT_" + a][use] >= r[use])
    assert np.all(np.diff(mt) >= 0) and mt[0] > 0 and mt[-1] <= 1
    # a 1,536-site slice through the literal oracle: events bit-exact
    s0, s1 = 64 * 1000, 64 * 1024
    chars = synth.planes_to_chars(B0, B1, V, s0, s1)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, _ = o.count_all_homoplasy_events(chars)
    sl = hc.sites[s0:s1]
    sl = sl[sl["site_class"] == 2].copy()
    sl["segsite_id"] -= s0
    for f in ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant"):
        np.testing.assert_array_equal(sl[f], or_ev[f], err_msg=f)
    # Monte-Carlo acceptance of the point-wise p-values against the exact hypergeometric tail
    # (all labels present: k_r ~ Hypergeom(P, cases, n)); 6-sigma binomial band at m replicates
    from scipy.stats import hypergeom
    P, K = len(lab), int((lab == 1).sum())
    use = st["a1_in_use"] == 1
    n = st["obs_counts"][use][:, 0] + st["obs_counts"][use][:, 1]
    k = st["obs_counts"][use][:, 0]
    exact = hypergeom.sf(k - 1, P, K, n)
    mc = st["r_a1"][use] / m
    sd = np.sqrt(np.maximum(exact * (1 - exact), 1e-12) / m)
    # the statistic is f(n,k) <= f(n,k_obs) <=> k >= k_obs when f(n,.) is strictly decreasing (true for 0<p<1)
    assert np.all(np.abs(mc - exact) < 6 * sd + 2.0 / m)


@pytest.mark.parametrize("na_frac", [0.0, 0.03])
def test_assoc_philox_small_generator_blocks(oracle_mod, monkeypatch, na_frac):
    """the 64-thread instantiation of the label generator (what a sharded run uses when a rank is left with few replicates),
    forced here on one GPU: same matrix, bit for bit"""
    monkeypatch.setenv("PB_GEN_SMALL_BLOCKS", "1")
    m = 1500
    tree, o, hc, sn, lab = _assoc_case(oracle_mod, 200, 1200, m, 43, na_frac, 0.02, seed=0xABCDEF)
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=0xABCDEF, n_threads=4)
    g_st, g_mt = hc.assoc()
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), mt_u.view(np.uint64))


def test_assoc_philox_many_samples_uses_spec_v1(oracle_mod):
    """P >= 65536 pheno rows: four draw ranges no longer fit one 64-bit product, so the label generator falls back to one
    32-bit word per draw (sampling spec v1); the oracle restates both specs and switches at the same P."""
    tree = synth.yule_tree(66000, 3)
    S, m = 64 * 3 + 7, 192
    chars = synth.planes_to_chars(*synth.simulate_planes(tree, S, 4), 0, S)
    sn, lab = synth.simulate_phenotypes(tree, 5, na_frac=0.01)
    assert len(lab) >= 65536
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    o.count_all_homoplasy_events(chars, n_threads=8)
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=12345, n_threads=8)
    hc = pb.HomoplasyCounter(num_replicates=m, seed=12345)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_seg_sites(chars)
    hc.count_all_homoplasy_events()
    g_st, g_mt = hc.assoc()
    assert_stats_equal(g_st, st)
    np.testing.assert_array_equal(g_mt.view(np.uint64), mt_s.view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), mt_u.view(np.uint64))
    hc.close()
