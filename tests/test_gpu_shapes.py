"""GPU parity at the BASELINE tree shapes (run with -m gpu on a B200): both event sweeps — k1x_events_kernel (one plane)
and k1_events_kernel (three planes) — and the whole association path against the CPU oracle at the C2 (5,000 leaves),
C3 (20,000 leaves: census segments, 15-bit leaf counters, leaves_below sums) and C5 (gaps / multi-allelic / lower case)
tree sizes, on site slices the literal oracle (leaf -> root walks, HC.java:1647-1698) finishes in seconds.

Events: every Homoplasy_Events field and the class counters, bit-exact (HC.java:1402-1511).
Association: replay mode with the oracle's java.util.Random shuffles and Philox mode (the oracle restates the sampling
spec), every integer and every double bit for bit (HC.java:3013-3186).
"""
import numpy as np
import pytest

import poutine_b200 as pb
from poutine_b200 import synth
from tests.helpers import assert_stats_equal

pytestmark = pytest.mark.gpu

EV_FIELDS = ("segsite_id", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant")


@pytest.fixture(scope="module", autouse=True)
def _built():
    pb.build()


def _case(config, S, seed_off=0):
    cfg = synth.CONFIGS[config]
    tree = synth.yule_tree(cfg["n_leaves"], cfg["seed"])
    planes = synth.simulate_planes(tree, S, cfg["seed"] + 2000 + seed_off, leaf_gap=cfg.get("leaf_gap", False),
                                   multiallelic_frac=cfg.get("multiallelic_frac", 0.0),
                                   third_allele_internal_frac=cfg.get("third_allele_internal_frac", 0.0))
    sn, lab = synth.simulate_phenotypes(tree, cfg["seed"] + 1000, na_frac=cfg.get("na_frac", 0.0))
    return tree, planes, sn, lab


@pytest.mark.parametrize("config,S,one_plane", [
    ("C2", 64 * 257 + 13, True),      # the headline kernel at the headline tree (BASELINE configs[1])
    ("C2", 64 * 257 + 13, False),     # same input held as three planes
    ("C3", 64 * 129 + 7, True),       # 20,000 leaves / 39,999 nodes: more than one wave of edge chunks per column block
    ("C3", 64 * 129 + 7, False),      # 81 census segments of 248 leaves, 15-bit counters
    ("C5", 64 * 257 + 13, False),     # leaf gaps, tri/quad-allelic sites, third alleles on internal nodes (three planes by necessity)
    ("C1", 64 * 300 + 1, True),
])
def test_events_and_assoc_at_baseline_shapes(oracle_mod, config, S, one_plane):
    tree, (B0, B1, V), sn, lab = _case(config, S)
    chars = synth.planes_to_chars(B0, B1, V, 0, S)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(chars, n_threads=8)
    m = 320
    perms = oracle_mod.java_shuffle_perms(4711, m, len(lab))
    or_st, or_mt_s, or_mt_u = o.assoc(sn, lab, m, 1, perms=perms, n_threads=8)

    hc = pb.HomoplasyCounter(num_replicates=m, replay=True, full_planes=not one_plane)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, S, tile_words=97, compact=True)
    ev, cc = hc.count_all_homoplasy_events()
    assert hc.timings()["plane_mode"] == (1 if one_plane else 2)
    assert (cc["mono"], cc["bi"], cc["tri"], cc["quad"], cc["no_allele"]) == tuple(or_cc[k] for k in ("mono", "bi", "tri", "quad", "other"))
    assert len(ev) == len(or_ev) and len(ev) > S // 2
    for f in EV_FIELDS:
        np.testing.assert_array_equal(ev[f], or_ev[f], err_msg=f)
    st, mt = hc.assoc(replay_perms=perms)
    assert_stats_equal(st, or_st)
    np.testing.assert_array_equal(mt.view(np.uint64), or_mt_s.view(np.uint64))
    np.testing.assert_array_equal(hc.maxT_unsorted().view(np.uint64), or_mt_u.view(np.uint64))
    hc.close()

    # Philox mode on the same events (product-defined sampling, restated by the oracle): bit-exact as well
    m2 = 1024
    or_st, or_mt_s, or_mt_u = o.assoc(sn, lab, m2, 1, perms=None, seed=99, n_threads=8)
    hp = pb.HomoplasyCounter(num_replicates=m2, seed=99, full_planes=not one_plane)
    hp.set_tree(tree.parent, tree.is_leaf)
    hp.set_phenos(sn, lab)
    hp.set_planes(B0, B1, V, S, compact=True)
    hp.count_all_homoplasy_events()
    st, mt = hp.assoc()
    assert_stats_equal(st, or_st)
    np.testing.assert_array_equal(mt.view(np.uint64), or_mt_s.view(np.uint64))
    hp.close()


def test_c3_shard_slice_parity_inside_a_large_run(oracle_mod):
    """A 20,000-leaf context at a size where the sweep runs many column blocks (64k sites = 1/78 of C3): a 2,048-site slice
    out of its middle, taken from the context's own per-site output, must equal the oracle on the same columns — in both
    storage modes."""
    S = 64 * 1024
    tree, (B0, B1, V), sn, lab = _case("C3", S, seed_off=5)
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    s0, s1 = 64 * 500, 64 * 532
    or_ev, _ = o.count_all_homoplasy_events(synth.planes_to_chars(B0, B1, V, s0, s1), n_threads=8)
    for full in (False, True):
        hc = pb.HomoplasyCounter(num_replicates=256, seed=3, full_planes=full)
        hc.set_tree(tree.parent, tree.is_leaf)
        hc.set_phenos(sn, lab)
        hc.set_planes(B0, B1, V, S, compact=True)
        hc.count_all_homoplasy_events()
        assert hc.timings()["plane_mode"] == (2 if full else 1)
        sl = hc.sites[s0:s1]
        sl = sl[sl["site_class"] == 2].copy()
        sl["segsite_id"] -= s0
        assert len(sl) == len(or_ev)
        for f in EV_FIELDS:
            np.testing.assert_array_equal(sl[f], or_ev[f], err_msg=f)
        hc.close()


@pytest.mark.parametrize("expect_output", [False, True])
@pytest.mark.parametrize("mode", ["whole", "tiles256", "tiles512_async", "reverse", "unaligned", "reupload_first", "labels_late"])
def test_sweep_under_upload_equals_plain_run(oracle_mod, mode, expect_output):
    """One-plane tiles that arrive in column order, cut at the sweep's block width (256 site words), are swept as they land
    (timings()["eager_words"]) and — with the phenotype rows already set — have their tail run as well ("tailed_words"); with
    pb_expect_site_output their per-site records are on their way back before pb_run_events is called ("early_sites").
    Anything else falls back to the full pass inside pb_run_events.  Same results either way, equal to the oracle; a second
    pass over the resident planes (no upload in between) must do everything again."""
    tree = synth.yule_tree(300, 21)
    S = 64 * (256 * 3 + 100) + 5
    W = (S + 63) // 64
    B0, B1, V = synth.simulate_planes(tree, S, 22)
    sn, lab = synth.simulate_phenotypes(tree, 23)
    X, colmask, all_valid = pb.compact_planes(B0, B1, V, S)
    assert all_valid
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    or_ev, or_cc = o.count_all_homoplasy_events(synth.planes_to_chars(B0, B1, V, 0, S), n_threads=8)
    m = 512
    or_st, or_mt, _ = o.assoc(sn, lab, m, 1, perms=None, seed=5, n_threads=8)

    hc = pb.HomoplasyCounter(num_replicates=m, seed=5)
    hc.set_tree(tree.parent, tree.is_leaf)
    if mode != "labels_late":
        hc.set_phenos(sn, lab)
    hc.set_sites(S)
    if expect_output:
        hc.expect_site_output()

    def put(w0, w1, asynchronous=False):
        hc.put_planes_biallelic(np.ascontiguousarray(X[:, w0:w1]), None, np.ascontiguousarray(colmask[w0:w1]), w0, asynchronous=asynchronous)

    want_tailed = None                   # default: everything swept ahead is tailed ahead
    if mode == "whole":
        put(0, W); want_eager = W
    elif mode == "labels_late":
        for w0 in range(0, W, 256):
            put(w0, min(W, w0 + 256))
        hc.set_phenos(sn, lab)           # which leaves are sampled decides tallies and CSR: everything done ahead is dropped
        want_eager = 0
    elif mode == "tiles256":
        for w0 in range(0, W, 256):
            put(w0, min(W, w0 + 256))
        want_eager = W
    elif mode == "tiles512_async":
        for w0 in range(0, W, 512):
            put(w0, min(W, w0 + 512), asynchronous=True)
        hc.upload_sync()
        want_eager = W
    elif mode == "reverse":
        for w0 in reversed(range(0, W, 256)):
            put(w0, min(W, w0 + 256))
        want_eager = 256                 # only the tile at word 0 (uploaded last) could start a pass
    elif mode == "unaligned":
        for w0 in range(0, W, 200):
            put(w0, min(W, w0 + 200))
        want_eager = 0
    else:
        for w0 in range(0, W, 256):
            put(w0, min(W, w0 + 256))
        put(0, 256)                      # the first tile again: the head start is dropped ...
        want_eager = 256                 # ... and a new in-order run starts with it; the rest is swept by pb_run_events
    if want_tailed is None:
        want_tailed = want_eager
    for pass_ in range(2):
        if expect_output:
            ev, cc = hc.count_all_homoplasy_events(compact=False)
            assert len(ev) == S and np.array_equal(ev["segsite_id"], np.arange(1, S + 1))
            ev = ev[ev["site_class"] == 2]
        else:
            ev, cc = hc.count_all_homoplasy_events()
        tm = hc.timings()
        assert tm["plane_mode"] == 1
        assert tm["eager_words"] == (want_eager if pass_ == 0 else 0)
        assert tm["tailed_words"] == (want_tailed if pass_ == 0 else 0)
        assert tm["early_sites"] == (min(S, 64 * want_tailed) if (pass_ == 0 and expect_output) else 0)
        assert cc["bi"] == or_cc["bi"] and len(ev) == len(or_ev)
        for f in EV_FIELDS:
            np.testing.assert_array_equal(ev[f], or_ev[f], err_msg=f)
        st, mt = hc.assoc()
        assert_stats_equal(st, or_st)
        np.testing.assert_array_equal(mt.view(np.uint64), or_mt.view(np.uint64))
    hc.close()


def test_fused_entry_point_equals_the_two_calls(oracle_mod):
    """pb_run_events_assoc (both reference entry points in one C call, per-site records copied out underneath the association)
    must return exactly what pb_run_events + pb_run_assoc return"""
    tree = synth.yule_tree(400, 31)
    S = 64 * 300 + 9
    B0, B1, V = synth.simulate_planes(tree, S, 32, leaf_gap=True, multiallelic_frac=0.02)
    sn, lab = synth.simulate_phenotypes(tree, 33, na_frac=0.02)
    m = 1000
    hc = pb.HomoplasyCounter(num_replicates=m, seed=8)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, S, compact=True)
    ev, cc = hc.count_all_homoplasy_events(compact=False)
    ev = ev.copy()
    st, mt = hc.assoc()
    st, mt = st.copy(), mt.copy()
    for _ in range(2):
        ev2, cc2, st2, mt2 = hc.count_all_homoplasy_events_and_assoc()
        assert cc2 == cc
        assert ev2.tobytes() == ev.tobytes()
        np.testing.assert_array_equal(mt2.view(np.uint64), mt.view(np.uint64))
        assert len(st2) == len(st)
        for f in st.dtype.names:
            a, b = st2[f], st[f]
            if a.dtype.kind == "f":
                np.testing.assert_array_equal(a.view(np.uint64), b.view(np.uint64), err_msg=f)
            else:
                np.testing.assert_array_equal(a, b, err_msg=f)
    hc.close()
