"""Data-parallel host logic (device-agnostic so it can be tested on CPU with gloo).

Clips shard across ranks (contiguous ranges of the global batch); the only collectives on the data
path are one scalar all-reduce (the global visible count of train.py:111-113) and the bucketed
sum-all-reduce of the flat gradient buffer.  Inference has no collective at all.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_clips: int, world: int, rank: int):
    """[lo, hi) of the global batch owned by ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(num_clips, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_denominator(local_visible_count: torch.Tensor, group=None) -> float:
    """max(sum over the GLOBAL batch of query_tracks_visible, 1)."""
    c = local_visible_count.clone().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c, group=group)
    return max(float(c.item()), 1.0)


def bucketed_allreduce(flat: torch.Tensor, bucket_elems: int, group=None, async_op=False):
    """Sum-all-reduce ``flat`` in contiguous buckets (the buffer is laid out in backward-completion
    order, so bucket i is complete before bucket i+1).  Returns the work handles when async."""
    works = []
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return works
    for s in range(0, flat.numel(), bucket_elems):
        w = dist.all_reduce(flat[s : s + bucket_elems], op=dist.ReduceOp.SUM, group=group, async_op=True)
        works.append(w)
    if not async_op:
        for w in works:
            w.wait()
        return []
    return works
