"""Multi-GPU parity: site-sharded runs must reproduce the single-GPU result bit for bit — every Homoplasy_Events field, the
class counters, every Binomial_Test_Stat column (the family-wise ones depend on the minimum of maxT over ALL shards: the
path's one exchange) and the pooled maxT itself.

  * in-process: pb_group (one host thread, one context + worker thread per GPU, exchange over peer memory) — the drop-in
    shape for the single-JVM reference host (HC.java:230-319);
  * one process per GPU: peer-memory exchange through cudaIpc handles (pb_comm_export / pb_comm_connect) and the NCCL
    all-reduce (pb_comm_init).

Needs >= 2 B200s for the real thing (`gpurun --gpus 2` / 8); the same group code also runs with several shards on ONE
GPU (PB_GROUP_ALLOW_SAME_DEVICE test hook), so a one-GPU box still covers the protocol and the merge.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_GPUS = torch.cuda.device_count() if torch.cuda.is_available() else 0
EV_FIELDS = ("segsite_id", "site_class", "allele1", "allele2", "a1_count", "a2_count", "a1_extant", "a2_extant")


def _inputs(S, leaf_gap, empty_tail_from=None, n_leaves=300):
    from poutine_b200 import synth
    tree = synth.yule_tree(n_leaves, 5)
    B0, B1, V = synth.simulate_planes(tree, S, 6, leaf_gap=leaf_gap)
    sn, lab = synth.simulate_phenotypes(tree, 7, na_frac=0.02 if leaf_gap else 0.0)
    if empty_tail_from is not None:          # everything from this word on is monomorphic 'A': shards with no informative site
        B0[:, empty_tail_from:] = 0
        B1[:, empty_tail_from:] = 0
    return tree, B0, B1, V, sn, lab


def _single(tree, B0, B1, V, sn, lab, S, m, compact, replay_perms=None, seed=4242):
    import poutine_b200 as pb
    hc = pb.HomoplasyCounter(num_replicates=m, seed=seed, replay=replay_perms is not None)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, S, compact=compact)
    ev, cc = hc.count_all_homoplasy_events(compact=False)
    ev = ev.copy()
    st, mt = hc.assoc(replay_perms=replay_perms)
    out = (ev, cc, st.copy(), mt.copy(), hc.maxT_unsorted())
    hc.close()
    return out


def _assert_same(got, want):
    ev, cc, st, mt, mtu = got
    ev0, cc0, st0, mt0, mtu0 = want
    assert cc == cc0
    for f in EV_FIELDS:
        np.testing.assert_array_equal(ev[f], ev0[f], err_msg=f)
    assert len(st) == len(st0)
    for f in ("site_index", "segsite_id", "a1_in_use", "a2_in_use", "obs_counts", "r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2"):
        np.testing.assert_array_equal(st[f], st0[f], err_msg=f)
    for f in ("obs_p_a1", "obs_p_a2", "pointwise_a1", "pointwise_a2", "familywise_a1", "familywise_a2"):
        a, b = st[f], st0[f]
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f)
        ok = ~np.isnan(b)
        np.testing.assert_array_equal(a[ok].view(np.uint64), b[ok].view(np.uint64), err_msg=f)
    np.testing.assert_array_equal(mt.view(np.uint64), mt0.view(np.uint64))
    np.testing.assert_array_equal(mtu.view(np.uint64), mtu0.view(np.uint64))


def _group_run(devices, tree, B0, B1, V, sn, lab, S, m, compact, tile_words, replay_perms=None, seed=4242, passes=1):
    import poutine_b200 as pb
    hg = pb.HomoplasyCounterGroup(devices, num_replicates=m, seed=seed, replay=replay_perms is not None)
    hg.set_tree(tree.parent, tree.is_leaf)
    hg.set_phenos(sn, lab)
    hg.set_planes(B0, B1, V, S, tile_words=tile_words, compact=compact)     # tiles straddle shard boundaries
    for _ in range(passes):                                                  # several epochs of the exchange
        ev, cc = hg.count_all_homoplasy_events(compact=False)
        st, mt = hg.assoc(replay_perms=replay_perms)
    out = (ev.copy(), cc, st.copy(), mt.copy(), hg.maxT_unsorted())
    tm = hg.timings()
    hg.close()
    return out, tm


@pytest.mark.parametrize("n_shards,compact,empty_tail", [(1, False, False), (2, True, False), (3, False, False), (4, True, True)])
def test_group_on_one_gpu_equals_single_context(monkeypatch, n_shards, compact, empty_tail):
    """pb_group with several shards on ONE device (test hook): sharding, tile routing, the peer-memory exchange protocol
    (flags, epochs, fused min in the sort kernel) and the output merge against a single context."""
    monkeypatch.setenv("PB_GROUP_ALLOW_SAME_DEVICE", "1")
    S, m = 64 * 301 + 5, 4096
    tree, B0, B1, V, sn, lab = _inputs(S, leaf_gap=not compact, empty_tail_from=(150 if empty_tail else None))
    want = _single(tree, B0, B1, V, sn, lab, S, m, compact)
    got, tm = _group_run([0] * n_shards, tree, B0, B1, V, sn, lab, S, m, compact, tile_words=37, passes=3)
    _assert_same(got, want)
    assert tm["n_informative"] == len(want[2])


@pytest.mark.parametrize("n_shards", [1, 2, 3])
def test_group_streamed_upload_with_expected_output(monkeypatch, n_shards):
    """Streamed one-plane tiles through a group with pb_group_expect_site_output: each shard sweeps and tails what arrives in
    order at the block width underneath the upload and sends the records back early; the merged result is a single context's."""
    import poutine_b200 as pb
    monkeypatch.setenv("PB_GROUP_ALLOW_SAME_DEVICE", "1")
    S, m = 64 * (256 * 4 + 40) + 5, 2048
    W = (S + 63) // 64
    tree, B0, B1, V, sn, lab = _inputs(S, leaf_gap=False)
    want = _single(tree, B0, B1, V, sn, lab, S, m, True)
    X, colmask, all_valid = pb.compact_planes(B0, B1, V, S)
    assert all_valid
    hg = pb.HomoplasyCounterGroup([0] * n_shards, num_replicates=m, seed=4242)
    hg.set_tree(tree.parent, tree.is_leaf)
    hg.set_phenos(sn, lab)
    hg.set_sites(S)
    hg.expect_site_output()
    for p in range(2):
        for w0 in range(0, W, 256):
            w1 = min(W, w0 + 256)
            hg.put_planes_biallelic(np.ascontiguousarray(X[:, w0:w1]), None, np.ascontiguousarray(colmask[w0:w1]), w0, asynchronous=True)
        hg.upload_sync()
        ev, cc = hg.count_all_homoplasy_events(compact=False)
        tm = hg.timings()
        assert tm["tailed_words"] > 0 and tm["early_sites"] > 0 and tm["tailed_words"] <= tm["eager_words"]
        if n_shards == 1:
            assert tm["tailed_words"] == W and tm["early_sites"] == S
        st, mt = hg.assoc()
        _assert_same((ev.copy(), cc, st.copy(), mt.copy(), hg.maxT_unsorted()), want)
    hg.close()


def test_group_replay_mode_and_no_informative(monkeypatch, oracle_mod):
    import poutine_b200 as pb
    monkeypatch.setenv("PB_GROUP_ALLOW_SAME_DEVICE", "1")
    S, m = 64 * 40 + 3, 300
    tree, B0, B1, V, sn, lab = _inputs(S, leaf_gap=True, n_leaves=120)
    perms = oracle_mod.java_shuffle_perms(5, m, len(lab))
    want = _single(tree, B0, B1, V, sn, lab, S, m, False, replay_perms=perms)
    got, _ = _group_run([0, 0], tree, B0, B1, V, sn, lab, S, m, False, tile_words=11, replay_perms=perms)
    _assert_same(got, want)
    # and against the oracle directly (replay mode is the reference-pinned one)
    from poutine_b200 import synth
    from tests.helpers import assert_stats_equal
    o = oracle_mod.Oracle(tree.parent, tree.leaf_nodes)
    o.count_all_homoplasy_events(synth.planes_to_chars(B0, B1, V, 0, S))
    or_st, or_mt, _ = o.assoc(sn, lab, m, 1, perms=perms)
    assert_stats_equal(got[2], or_st)
    np.testing.assert_array_equal(got[3].view(np.uint64), or_mt.view(np.uint64))
    # min_hcount filters out every site of every shard: the group reports the reference's exit (HC.java:3866-3869)
    hg = pb.HomoplasyCounterGroup([0, 0], num_replicates=64, min_hcount=10 ** 6)
    hg.set_tree(tree.parent, tree.is_leaf)
    hg.set_phenos(sn, lab)
    hg.set_planes(B0, B1, V, S)
    hg.count_all_homoplasy_events()
    with pytest.raises(pb.NoInformativeSites):
        hg.assoc()
    hg.close()


@pytest.mark.skipif(N_GPUS < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("compact,empty_tail,gen_kb", [(False, False, None), (True, False, None), (True, True, None), (False, False, "96"), (True, False, "64")])
def test_group_all_gpus_equals_single_gpu(monkeypatch, compact, empty_tail, gen_kb):
    """the real thing: one process, one host thread, every GPU of the box (2 on a 2-GPU lease, 8 on a full box).  The shards
    share the label matrix (each draws 1/G of its replicate columns and pulls the rest over NVLink); gen_kb forces several
    generation tiles per pass, so the double-buffered matrices and the "slice pulled" acknowledgements are exercised too."""
    S, m = 64 * 1201 + 5, 8192
    devices = list(range(min(N_GPUS, 8)))
    tree, B0, B1, V, sn, lab = _inputs(S, leaf_gap=not compact, empty_tail_from=(700 if empty_tail else None), n_leaves=700)
    if gen_kb:
        monkeypatch.setenv("PB_K3_GEN_KB", gen_kb)
        monkeypatch.setenv("PB_K3_TILE_KB", "32")
    want = _single(tree, B0, B1, V, sn, lab, S, m, compact)
    got, tm = _group_run(devices, tree, B0, B1, V, sn, lab, S, m, compact, tile_words=97, passes=3)
    _assert_same(got, want)
    assert abs(tm["gen_shard_frac"] - 1.0 / len(devices)) < 1e-6          # the label matrix was drawn cooperatively


# ---- one process per GPU ------------------------------------------------------------------------------------------
def _worker(rank, world, port, S, m, q, transport, compact, empty_tail):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)   # host channel for the ids / handles only
    import poutine_b200 as pb
    from poutine_b200 import shard
    tree, B0, B1, V, sn, lab = _inputs(S, leaf_gap=not compact, empty_tail_from=(150 if empty_tail else None))
    s0, s1 = shard.site_range(S, rank, world)
    hc = pb.HomoplasyCounter(num_replicates=m, device=rank, seed=4242, site_offset=s0)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    w0, w1 = s0 // 64, (s1 + 63) // 64
    hc.set_planes(np.ascontiguousarray(B0[:, w0:w1]), np.ascontiguousarray(B1[:, w0:w1]), np.ascontiguousarray(V[:, w0:w1]), s1 - s0,
                  compact=compact)
    if transport == "nccl":
        hc.comm_init(shard.share_unique_id(dist, pb.comm_unique_id, rank), rank, world)
    else:
        handles = [None] * world
        dist.all_gather_object(handles, hc.comm_export())
        hc.comm_connect(handles, rank, world)
        if transport == "p2p_shared_labels":       # the ranks draw the label matrix cooperatively
            blobs = [None] * world
            dist.all_gather_object(blobs, hc.comm_export_labels())
            hc.comm_connect_labels(blobs)
    for _ in range(3):
        ev, cc = hc.count_all_homoplasy_events(compact=False)
        st, mt = hc.assoc()
    frac = hc.timings()["gen_shard_frac"]
    assert abs(frac - (1.0 / world if transport == "p2p_shared_labels" else 1.0)) < 1e-6, frac
    q.put((rank, ev.tobytes(), cc, st.tobytes(), mt.tobytes(), hc.maxT_unsorted().tobytes()))
    dist.barrier()          # no rank frees its exchange block while a peer may still be reading it
    hc.close()
    dist.destroy_process_group()


@pytest.mark.skipif(N_GPUS < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport,compact,empty_tail", [("p2p", False, False), ("p2p", True, True), ("nccl", False, True), ("nccl", True, False),
                                                         ("p2p_shared_labels", False, False), ("p2p_shared_labels", True, True)])
def test_process_per_gpu_sharded_equals_single_gpu(transport, compact, empty_tail):
    import torch.multiprocessing as mp
    import poutine_b200 as pb
    from poutine_b200 import _lib, shard
    pb.build()
    world = min(N_GPUS, 8)
    S, m = 64 * 301 + 5, 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, m, q, transport, compact, empty_tail)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    tree, B0, B1, V, sn, lab = _inputs(S, leaf_gap=not compact, empty_tail_from=(150 if empty_tail else None))
    want = _single(tree, B0, B1, V, sn, lab, S, m, compact)
    evs = [np.frombuffer(r[1], dtype=_lib.SITE_DTYPE) for r in res]
    sts = [np.frombuffer(r[3], dtype=_lib.STAT_DTYPE).copy() for r in res]
    for rk, p in enumerate(sts):
        p["site_index"] += shard.site_range(S, rk, world)[0]     # shard-local -> global
    cc = {k: sum(r[2][k] for r in res) for k in res[0][2]}
    for r in res:                                                 # every rank holds the pooled maxT, sorted and unsorted
        assert r[4] == res[0][4] and r[5] == res[0][5]
    got = (np.concatenate(evs), cc, np.concatenate(sts), np.frombuffer(res[0][4], dtype=np.float64), np.frombuffer(res[0][5], dtype=np.float64))
    _assert_same(got, want)
