"""Site sharding for multi-GPU runs: contiguous 64-site-word ranges, one per rank.

Sites are independent in count_all_homoplasy_events (the loop at HC.java:1418 carries only the class
counters) and in the point-wise statistic; the family-wise statistic needs the per-replicate minimum
over ALL sites (HC.java:3239-3243), which is the path's one exchange: all-reduce(min) of maxT[m].
"""
from __future__ import annotations


def word_range(n_words: int, rank: int, world: int) -> tuple[int, int]:
    """[w0, w1) of the 64-site words owned by `rank`."""
    assert 0 <= rank < world
    return n_words * rank // world, n_words * (rank + 1) // world


def site_range(n_sites: int, rank: int, world: int) -> tuple[int, int]:
    """[s0, s1) of the sites owned by `rank`; s0 is 64-aligned (pb_config.site_offset)."""
    W = (n_sites + 63) // 64
    w0, w1 = word_range(W, rank, world)
    return 64 * w0, min(n_sites, 64 * w1)


def share_unique_id(dist, make_id, rank: int, src: int = 0) -> bytes:
    """Rank `src` creates the 128-byte NCCL id (pb_comm_unique_id); everyone receives it over the
    host's own channel (torch.distributed here; any transport works)."""
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])
