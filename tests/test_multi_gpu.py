"""Two-GPU NCCL path (needs >= 2 B200s: `gpurun --gpus 2`): site-sharded contexts whose only exchange is the
all-reduce(min) of maxT inside pb_run_assoc must reproduce the single-GPU result bit for bit."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mono_tail(B0, B1, V, S, world):
    """make rank 1's shard monomorphic ('A' everywhere): it has no informative site but still joins the all-reduce"""
    from poutine_b200 import shard
    s0, s1 = shard.site_range(S, 1, world)
    B0[:, s0 // 64:] = 0
    B1[:, s0 // 64:] = 0


def _worker(rank, world, port, S, m, q, empty_rank1=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)   # host channel for the 128-byte id only
    import poutine_b200 as pb
    from poutine_b200 import shard, synth
    tree = synth.yule_tree(300, 5)
    B0, B1, V = synth.simulate_planes(tree, S, 6, leaf_gap=True)
    sn, lab = synth.simulate_phenotypes(tree, 7, na_frac=0.02)
    s0, s1 = shard.site_range(S, rank, world)
    if empty_rank1:
        _mono_tail(B0, B1, V, S, world)
    hc = pb.HomoplasyCounter(num_replicates=m, device=rank, seed=4242, site_offset=s0)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    w0, w1 = s0 // 64, (s1 + 63) // 64
    hc.set_planes(np.ascontiguousarray(B0[:, w0:w1]), np.ascontiguousarray(B1[:, w0:w1]), np.ascontiguousarray(V[:, w0:w1]), s1 - s0)
    hc.comm_init(shard.share_unique_id(dist, pb.comm_unique_id, rank), rank, world)
    hc.count_all_homoplasy_events()
    st, mt = hc.assoc()
    q.put((rank, st.tobytes(), mt.tobytes(), hc.timings()["ms_allreduce"]))
    dist.barrier()
    hc.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("empty_rank1", [False, True])
def test_two_gpu_sharded_equals_single_gpu(empty_rank1):
    import torch.multiprocessing as mp
    import poutine_b200 as pb
    from poutine_b200 import _lib, synth
    pb.build()
    S, m, world = 64 * 301 + 5, 4096, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, m, q, empty_rank1)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    tree = synth.yule_tree(300, 5)
    B0, B1, V = synth.simulate_planes(tree, S, 6, leaf_gap=True)
    sn, lab = synth.simulate_phenotypes(tree, 7, na_frac=0.02)
    if empty_rank1:
        _mono_tail(B0, B1, V, S, world)
    hc = pb.HomoplasyCounter(num_replicates=m, seed=4242)
    hc.set_tree(tree.parent, tree.is_leaf)
    hc.set_phenos(sn, lab)
    hc.set_planes(B0, B1, V, S)
    hc.count_all_homoplasy_events()
    st, mt = hc.assoc()
    parts = [np.frombuffer(r[1], dtype=_lib.STAT_DTYPE).copy() for r in res]
    from poutine_b200 import shard
    for rk, p in enumerate(parts):
        p["site_index"] += shard.site_range(S, rk, world)[0]     # shard-local -> global
    got = np.concatenate(parts)
    for f in ("segsite_id", "obs_counts", "r_a1", "r_a2", "r_maxT_a1", "r_maxT_a2", "a1_in_use", "a2_in_use"):
        np.testing.assert_array_equal(got[f], st[f], err_msg=f)
    for f in ("obs_p_a1", "pointwise_a2", "familywise_a1", "familywise_a2"):
        a, b = got[f], st[f]
        ok = ~np.isnan(b)
        np.testing.assert_array_equal(a[ok].view(np.uint64), b[ok].view(np.uint64), err_msg=f)
    for r in res:
        assert r[2] == mt.tobytes()                 # every rank holds the pooled, sorted maxT
