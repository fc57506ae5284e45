"""The oracle's event detection / association: hand-derived fixture, independent closed form,
invariants of the reference code, JDK RNG known answers, and the frozen C1 slice."""
import json
import os

import numpy as np
import pytest

from poutine_b200 import synth
from tests.helpers import closed_form_events


def test_hand_derived_fixture(oracle_mod, golden_dir):
    with open(os.path.join(golden_dir, "events_hand.json")) as f:
        fx = json.load(f)
    chars = np.array([[ord(s["column"][v]) for s in fx["sites"]] for v in range(len(fx["names"]))], dtype=np.uint8)
    o = oracle_mod.Oracle(fx["parent"], fx["leaves"])
    ev, cc = o.count_all_homoplasy_events(chars)
    got = {int(e["segsite_id"]): e for e in ev}
    for i, s in enumerate(fx["sites"]):
        if s["expected"] is None:
            assert i + 1 not in got, s["name"]
        else:
            e = got[i + 1]
            tup = [chr(e["allele1"]), chr(e["allele2"]), int(e["a1_count"]), int(e["a2_count"]), int(e["a1_extant"]), int(e["a2_extant"])]
            assert tup == s["expected"], s["name"]
    assert cc == fx["class_counts"]


def refbytecode_cases(golden_dir, suffix=""):
    """tests/golden/events_refbytecode.json: what the reference's own compiled count_all_homoplasy_events produced
    (executed from Homoplasy_Counter.class by oracle/tools/refjvm.py, see gen_events_refbytecode_golden.py).
    suffix="_extra": the second set (degenerate topologies and phenotypes), used by the oracle tests only."""
    with open(os.path.join(golden_dir, "events_refbytecode%s.json" % suffix)) as f:
        fx = json.load(f)
    for c in fx["cases"]:
        chars = np.array([[ord(ch) for ch in row] for row in c["rows"]], dtype=np.uint8)
        yield c, chars


@pytest.mark.parametrize("suffix,min_sites", [("", 400), ("_extra", 500)])
def test_oracle_matches_reference_bytecode(oracle_mod, golden_dir, suffix, min_sites):
    """THE pin of rows a1-a5: every Homoplasy_Events field, both MRCA node sets and the class counters the reference logs,
    for 5 seeded alignments on nasty topologies, as computed by the reference's bytecode; the second set adds a star tree
    (no testable edge), a ladder, a two-leaf tree and four more."""
    n_sites = 0
    for c, chars in refbytecode_cases(golden_dir, suffix):
        o = oracle_mod.Oracle(c["parent"], c["leaf_nodes"])
        ev, cc = o.count_all_homoplasy_events(chars, np.asarray(c["physical_pos"], dtype=np.int32))
        assert cc == c["class_counts"], c["name"]
        assert len(ev) == len(c["events"]), c["name"]
        names = np.asarray(c["names"])
        for i, (e, r) in enumerate(zip(ev, c["events"])):
            got = [int(e["segsite_id"]), int(e["physical_pos"]), chr(e["allele1"]), chr(e["allele2"]), int(e["a1_count"]),
                   int(e["a2_count"]), int(e["a1_extant"]), int(e["a2_extant"])]
            want = [r[k] for k in ("segsite_id", "physical_pos", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant")]
            assert got == want, (c["name"], i)
            assert sorted(names[o.mrca(i, 0)].tolist()) == r["a1_mrcas"], (c["name"], i)
            assert sorted(names[o.mrca(i, 1)].tolist()) == r["a2_mrcas"], (c["name"], i)
        n_sites += len(ev)
    assert n_sites > min_sites


def _f64(hexbits):
    return np.array([int(hexbits, 16)], dtype=np.uint64).view(np.float64)[0]


def assert_stats_equal_refbytecode(st, mt_sorted, mt_unsorted, p_success, a):
    """st / maxT arrays (oracle or GPU) against one case of tests/golden/assoc_refbytecode.json, doubles bit for bit"""
    bits = lambda x: "%016x" % int(np.float64(x).view(np.uint64))
    assert bits(p_success) == a["p_success"]
    assert len(st) == len(a["sites"])
    assert [bits(x) for x in mt_sorted] == a["maxT_sorted"]
    if mt_unsorted is not None:
        assert [bits(x) for x in mt_unsorted] == a["maxT_unsorted"]
    for row, r in zip(st, a["sites"]):
        tag = (a["name"], r["segsite_id"])
        assert int(row["segsite_id"]) == r["segsite_id"], tag
        assert (int(row["a1_in_use"]), int(row["a2_in_use"])) == (r["a1_in_use"], r["a2_in_use"]), tag
        assert row["obs_counts"].tolist() == r["obs_counts"], tag
        for k in ("r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
            assert int(row[k]) == r[k], tag + (k,)
        for k in ("obs_p_a1", "obs_p_a2", "pointwise_a1", "pointwise_a2", "familywise_a1", "familywise_a2"):
            want = _f64(r[k])
            assert (np.isnan(row[k]) and np.isnan(want)) or bits(row[k]) == r[k], tag + (k, float(row[k]), float(want))


def assoc_refbytecode_cases(golden_dir, suffix=""):
    """(events case, chars, assoc case) triples: assoc_refbytecode.json holds what the reference's compiled
    subset_by_min_hcount + resampling driver (+ commons-math's own bytecode) produced on the events of the same-named case."""
    with open(os.path.join(golden_dir, "assoc_refbytecode%s.json" % suffix)) as f:
        fa = {c["name"]: c for c in json.load(f)["cases"]}
    for c, chars in refbytecode_cases(golden_dir, suffix):
        yield c, chars, fa[c["name"]]


@pytest.mark.parametrize("suffix,min_sites", [("", 400), ("_extra", 350)])
def test_oracle_assoc_matches_reference_bytecode(oracle_mod, golden_dir, suffix, min_sites):
    """THE pin of rows a6-a11: observed tallies, binomial p-values, r counts, per-replicate min p, point-wise and
    family-wise estimates of the reference's own bytecode, replayed through the oracle with the recorded shuffles.  The second
    set: alleles with n = 0 (min_hcount 0 on a star tree), 17 pheno rows for 50 leaves, 30 % missing labels, only cases / only
    controls (background p = 1 / 0: the boundary branches of the binomial chain), min_hcount = 4."""
    n = 0
    for c, chars, a in assoc_refbytecode_cases(golden_dir, suffix):
        o = oracle_mod.Oracle(c["parent"], c["leaf_nodes"])
        o.count_all_homoplasy_events(chars, np.asarray(c["physical_pos"], dtype=np.int32))
        st, mt_s, mt_u = o.assoc(np.asarray(a["sample_node"]), np.asarray(a["label"], dtype=np.uint8), a["m"], a["min_hcount"],
                                 perms=np.asarray(a["perms"], dtype=np.uint32))
        assert_stats_equal_refbytecode(st, mt_s, mt_u, o.p_success, a)
        n += len(st)
    assert n > min_sites


@pytest.mark.parametrize("seed", range(12))
def test_literal_walk_equals_closed_form(oracle_mod, seed):
    tree = synth.random_tree(12 + 5 * seed, 100 + seed, polytomy=0.3, root_leaf_children=seed % 3)
    chars = synth.simulate_chars(tree, 60, 200 + seed, nasty=True)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    ev, cc = o.count_all_homoplasy_events(chars)
    ref, rcc = closed_form_events(tree.parent, tree.leaf_nodes, chars)
    assert cc == rcc
    assert len(ev) == len(ref)
    for i, (e, r) in enumerate(zip(ev, ref)):
        for f in ("segsite_id", "a1_count", "a2_count", "a1_extant", "a2_extant"):
            assert int(e[f]) == r[f], (seed, i, f)
        assert (int(e["allele1"]), int(e["allele2"])) == (r["allele1"], r["allele2"])
        assert list(o.mrca(i, 0)) == r["mrca"][0] and list(o.mrca(i, 1)) == r["mrca"][1]


def test_reference_invariants(oracle_mod):
    tree = synth.yule_tree(200, 7)
    chars = synth.simulate_chars(tree, 300, 8, nasty=True)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    ev, _ = o.count_all_homoplasy_events(chars)
    assert np.all((ev["a1_count"] == 0) | (ev["a1_count"] >= 2))          # (i)
    assert np.all((ev["a2_count"] == 0) | (ev["a2_count"] >= 2))
    assert np.all(ev["a1_extant"] <= ev["a1_count"]) and np.all(ev["a2_extant"] <= ev["a2_count"])  # (ii)
    sn, lab = synth.simulate_phenotypes(tree, 9)                         # every leaf labelled 0/1
    m = 50
    perms = oracle_mod.java_shuffle_perms(1, m, len(lab))
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=perms)
    evs = ev[st["event_index"]]
    tot = st["obs_counts"].sum(axis=1)
    expect = np.where(st["a1_in_use"] == 1, evs["a1_extant"], 0) + np.where(st["a2_in_use"] == 1, evs["a2_extant"], 0)
    np.testing.assert_array_equal(tot, expect)                           # (iii) check_extant_nodes_only_code, HC.java:4345-4369
    for a in ("a1", "a2"):
        r = st["r_" + a]
        use = st[a + "_in_use"] == 1
        assert np.all((r[use] >= 0) & (r[use] <= m)) and np.all(r[~use] == -1)   # (iv)
        np.testing.assert_array_equal(st["pointwise_" + a][use], (r[use] + 1.0) / (m + 1.0))  # (v)
        assert np.all(np.isnan(st["pointwise_" + a][~use]))
    assert np.all(np.diff(mt_s) >= 0) and sorted(mt_u) == list(mt_s)


def test_min_hcount_filters_everything(oracle_mod):
    tree = synth.yule_tree(50, 3)
    chars = synth.simulate_chars(tree, 40, 4, nasty=False)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    o.count_all_homoplasy_events(chars)
    sn, lab = synth.simulate_phenotypes(tree, 5)
    st, _, _ = o.assoc(sn, lab, 10, 1000, perms=oracle_mod.java_shuffle_perms(1, 10, len(lab)))
    assert st is None   # HC.java:3866-3869


def test_java_util_random_known_answers(oracle_mod):
    # JDK-documented LCG: values every Java programmer can reproduce
    assert list(oracle_mod.java_random_ints(0, 2)) == [-1155484576, -723955400]
    assert list(oracle_mod.java_random_ints(42, 2)) == [-1170105035, 234785527]
    assert list(oracle_mod.java_random_ints(42, 5, 10)) == [0, 3, 8, 4, 0]
    p = oracle_mod.java_shuffle_perms(123, 4, 37)
    assert all(sorted(row) == list(range(37)) for row in p.tolist())
    assert len({tuple(r) for r in p.tolist()}) == 4


def test_philox_labels_are_exact_arrangements(oracle_mod):
    P, nc, nt = 203, 70, 120
    seen = set()
    for rep in range(200):
        lab = oracle_mod.philox_labels(99, rep, P, nc, nt)
        assert (lab == 1).sum() == nc and (lab == 0).sum() == nt and (lab == 2).sum() == P - nc - nt
        seen.add(lab.tobytes())
    assert len(seen) == 200
    # marginal: each sample is a case with probability nc/P
    tot = sum(oracle_mod.philox_labels(5, rep, P, nc, nt).astype(np.int64) == 1 for rep in range(4000))
    assert abs(tot.mean() / 4000 - nc / P) < 0.01
    assert tot.std() / 4000 < 0.02


def _philox4x32_10(c, k):
    """Philox4x32-10 (Salmon et al. 2011) in plain Python ints — second opinion for the oracle's restatement."""
    c0, c1, c2, c3 = c
    k0, k1 = k
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & M, (p0 >> 32) ^ c3 ^ k1, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return [c0, c1, c2, c3]


def _labels_spec_v3(seed, rep, P, n_case, n_ctrl):
    """Sampling spec v3 written from its text with Python big ints: xoshiro128++ stream seeded by
    Philox(ctr=(0, rep, 2)), four draws per 64-bit word by mixed-radix multiply-shift, exact rejection from the Philox retry
    stream (ctr word 3 = 1)."""
    M = 0xFFFFFFFF
    key = (seed & M, seed >> 32)
    r0, r1 = rep & M, rep >> 32
    s = _philox4x32_10((0, r0, r1, 2), key)
    if not any(s):
        s[0] = 1
    rotl = lambda v, k: ((v << k) | (v >> (32 - k))) & M

    def nxt():
        res = (rotl((s[0] + s[3]) & M, 7) + s[0]) & M
        t = (s[1] << 9) & M
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
        s[2] ^= t
        s[3] = rotl(s[3], 11)
        return res

    out, q, retries = [], 0, 0
    for g in range((P + 3) // 4):
        lo = nxt(); hi = nxt()
        x = (hi << 32) | lo
        c = min(4, P - 4 * g)
        while True:
            l, prod, us = x, 1, []
            for j in range(c):
                R = P - 4 * g - j
                pr = l * R
                us.append(pr >> 64); l = pr & (2 ** 64 - 1); prod *= R
            if l >= (2 ** 64) % prod:
                break
            rw = _philox4x32_10((q >> 1, r0, r1, 1), key)
            x = (rw[2 * (q & 1) + 1] << 32) | rw[2 * (q & 1)]
            q += 1; retries += 1
        for u in us:
            if u < n_case:
                out.append(1); n_case -= 1
            elif u < n_case + n_ctrl:
                out.append(0); n_ctrl -= 1
            else:
                out.append(2)
    return np.array(out, dtype=np.uint8), retries


def test_philox_labels_match_spec_v3_text(oracle_mod):
    """The oracle's C++ restatement of the sampling spec against the plain-Python one above, label for label; the huge-P
    cases make retries likely enough to be seen (prod(R_j) / 2^64 per group)."""
    assert _philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]     # Random123 known answer
    assert _philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    seen_retry = 0
    for seed, rep, P, nc, nt in [(1, 0, 1, 1, 0), (2, 5, 2, 1, 1), (3, 7, 3, 1, 1), (4, 9, 4, 2, 2), (5, 11, 5, 2, 2), (99, 3, 203, 70, 120),
                                 (2 ** 63 + 12345, 2 ** 40 + 17, 1301, 400, 880), (20260, 99999, 5000, 1900, 3100),
                                 (7, 1, 65535, 30000, 35000), (7, 2, 65535, 20000, 30000), (8, 3, 60001, 30000, 30001)]:
        want, retries = _labels_spec_v3(seed, rep, P, nc, nt)
        got = oracle_mod.philox_labels(seed, rep, P, nc, nt)
        assert np.array_equal(got, want), (seed, rep, P)
        seen_retry += retries
    assert seen_retry > 0


def test_philox_labels_uniform_over_arrangements(oracle_mod):
    """Sampling spec v3 draws every arrangement of the labels equally often: chi-square over all 30 arrangements of
    (2 cases, 2 controls, 1 missing) on 5 samples, and over the 20 case subsets of (3 cases, 3 controls); consecutive
    replicates of one seed and the same replicate under consecutive seeds (the Philox-seeded streams must not be correlated
    across either)."""
    import itertools
    for P, nc, nt in ((5, 2, 2), (6, 3, 3)):
        n_arr = len(set(itertools.permutations([1] * nc + [0] * nt + [2] * (P - nc - nt))))
        for vary in ("rep", "seed"):
            n = 3000 * n_arr
            counts = {}
            for i in range(n):
                lab = oracle_mod.philox_labels(77, i, P, nc, nt) if vary == "rep" else oracle_mod.philox_labels(1000 + i, 3, P, nc, nt)
                counts[lab.tobytes()] = counts.get(lab.tobytes(), 0) + 1
            assert len(counts) == n_arr
            exp = n / n_arr
            chi2 = sum((c - exp) ** 2 / exp for c in counts.values())
            # chi-square with n_arr - 1 degrees of freedom: mean df, sd sqrt(2 df); 5 sd is generous and still catches a bias of ~2 %
            df = n_arr - 1
            assert chi2 < df + 5 * (2 * df) ** 0.5, (P, vary, chi2, df)
    # pairs of samples are exchangeable as well: P(sample i and sample j both cases) = nc (nc-1) / (P (P-1)) for a larger P
    P, nc, nt, n = 41, 15, 20, 20000
    both = np.zeros((P, P))
    for rep in range(n):
        c = (oracle_mod.philox_labels(5, rep, P, nc, nt) == 1).astype(np.float64)
        both += np.outer(c, c)
    want = nc * (nc - 1) / (P * (P - 1))
    off = both[~np.eye(P, dtype=bool)] / n
    sd = (want * (1 - want) / n) ** 0.5
    assert np.abs(off - want).max() < 5 * sd, (np.abs(off - want).max(), sd)


def test_c1_slice_regression(oracle_mod, golden_dir):
    z = np.load(os.path.join(golden_dir, "c1_slice.npz"))
    S, m = int(z["n_sites"]), int(z["m"])
    chars = synth.planes_to_chars(z["B0"], z["B1"], z["V"], 0, S)
    o = oracle_mod.Oracle(z["parent"], z["leaf_nodes"])
    ev, cc = o.count_all_homoplasy_events(chars, np.arange(S, dtype=np.int32) * 7 + 11)
    assert ev.tobytes() == z["events"].tobytes()
    perms = oracle_mod.java_shuffle_perms(int(z["perm_seed"]), m, len(z["label"]))
    st, mt_s, mt_u = o.assoc(z["sample_node"], z["label"], m, 1, perms=perms, n_threads=3)  # threads must not change results
    assert st.tobytes() == z["stats"].tobytes()
    np.testing.assert_array_equal(mt_u, z["maxT_unsorted"])
