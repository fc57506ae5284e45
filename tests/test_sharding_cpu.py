"""N > 1 host logic on CPU (gloo, world_size 2): the site-shard decomposition the multi-GPU path uses —
r local per shard, all-reduce(min) of maxT — reproduces the unsharded result exactly.  The per-shard
compute here is the CPU oracle (this is a test of the decomposition + plumbing, not of the kernels)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from poutine_b200 import shard, synth


def test_ranges_tile_the_sites():
    for S in (1, 63, 64, 65, 1000, 64 * 15625):
        for world in (1, 2, 3, 8):
            ends = 0
            for r in range(world):
                s0, s1 = shard.site_range(S, r, world)
                assert s0 == min(ends, s0) or s0 == ends or s1 == s0
                assert s0 % 64 == 0 and s0 <= s1 <= S
                if s1 > s0:
                    assert s0 == ends
                    ends = s1
            assert ends == S


def _worker(rank, world, port, S, m, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    tree = synth.yule_tree(120, 5)
    chars = synth.simulate_chars(tree, S, 6, nasty=True)
    sn, lab = synth.simulate_phenotypes(tree, 7, na_frac=0.03)
    ident = shard.share_unique_id(dist, lambda: bytes(range(128)), rank)
    assert ident == bytes(range(128))
    s0, s1 = shard.site_range(S, rank, world)
    o = oracle.Oracle(tree.parent, tree.leaf_nodes)
    o.count_all_homoplasy_events(np.ascontiguousarray(chars[:, s0:s1]))
    st, mt_s, mt_u = o.assoc(sn, lab, m, 1, perms=None, seed=99)   # same seed on every rank -> same permutations
    t = torch.from_numpy(mt_u.copy())
    dist.all_reduce(t, op=dist.ReduceOp.MIN)                        # the one collective of the path
    pooled = np.sort(t.numpy())
    # family-wise counts against the pooled maxT
    fw = {}
    for a in ("a1", "a2"):
        use = st[a + "_in_use"] == 1
        fw[a] = np.where(use, np.searchsorted(pooled, st["obs_p_" + a], side="right"), -1)
    q.put((rank, s0, st["segsite_id"] + s0, st["r_a1"], st["r_a2"], fw["a1"], fw["a2"], pooled))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo():
    import oracle
    oracle.build()
    S, m, world = 700, 300, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tree = synth.yule_tree(120, 5)
    chars = synth.simulate_chars(tree, S, 6, nasty=True)
    sn, lab = synth.simulate_phenotypes(tree, 7, na_frac=0.03)
    o = oracle.Oracle(tree.parent, tree.leaf_nodes)
    o.count_all_homoplasy_events(chars)
    st, mt_s, _ = o.assoc(sn, lab, m, 1, perms=None, seed=99)
    np.testing.assert_array_equal(np.concatenate([r[2] for r in res]), st["segsite_id"])
    np.testing.assert_array_equal(np.concatenate([r[3] for r in res]), st["r_a1"])
    np.testing.assert_array_equal(np.concatenate([r[4] for r in res]), st["r_a2"])
    np.testing.assert_array_equal(np.concatenate([r[5] for r in res]), st["r_maxT_a1"])
    np.testing.assert_array_equal(np.concatenate([r[6] for r in res]), st["r_maxT_a2"])
    for r in res:
        np.testing.assert_array_equal(r[7], mt_s)
