"""Multi-GPU sharding of a batch: by stream, byte-balanced, NO data-path collective.

Every zlib stream is independent, so a batch is partitioned across the ranks of one
node and each rank runs the ordinary single-GPU batch call on its shard.  The only inter-rank
traffic is the gather of the per-stream results (status, out_len) -- and of the payload when the
caller wants it in one place -- over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
A single stream is never split across GPUs.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def partition_lpt(costs: Sequence[int], world: int) -> list[np.ndarray]:
    """Longest-processing-time greedy: sort by cost descending, give each item to the least loaded
    rank.  Deterministic (ties broken by index), so every rank computes the same partition without
    talking.  Returns, per rank, the sorted stream indices it owns."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.lexsort((np.arange(costs.size), -costs))
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(costs.size, dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))  # first minimum: deterministic
        owner[i] = r
        load[r] += max(int(costs[i]), 1)
    return [np.nonzero(owner == r)[0] for r in range(world)]


def shard_inflate(ctx, streams: Sequence[bytes], out_caps: Sequence[int], rank: int, world: int, group=None,
                  flags: int = 0, gather_payload: bool = True):
    """Inflate `streams` with the work split over `world` ranks.  Every rank passes the same arguments;
    rank 0 returns (status int32[n], outputs list[bytes] | None, out_len int64[n]); other ranks return
    their local results only.  cost = compressed + expected uncompressed bytes."""
    import torch
    import torch.distributed as dist

    n = len(streams)
    costs = [len(s) + int(c) for s, c in zip(streams, out_caps)]
    parts = partition_lpt(costs, world)
    mine = parts[rank]
    st, outs, _ = ctx.inflate_batch([streams[i] for i in mine], [out_caps[i] for i in mine], flags)
    status = torch.full((n,), -99, dtype=torch.int32)
    out_len = torch.zeros(n, dtype=torch.int64)
    status[torch.from_numpy(mine)] = torch.from_numpy(np.asarray(st, dtype=np.int32))
    out_len[torch.from_numpy(mine)] = torch.tensor([len(o) for o in outs], dtype=torch.int64)
    if world > 1:
        # the only collectives: results to everyone (max works because unowned entries are the minimum)
        dist.all_reduce(status, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(out_len, op=dist.ReduceOp.MAX, group=group)
    payload = None
    if gather_payload:
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object((mine.tolist(), outs), gathered, dst=0, group=group)
            if rank == 0:
                payload = [b""] * n
                for idx, o in gathered:
                    for i, b in zip(idx, o):
                        payload[i] = b
        else:
            payload = [b""] * n
            for i, b in zip(mine.tolist(), outs):
                payload[i] = b
    return status.numpy(), payload, out_len.numpy()


def shard_deflate_ultrafast(ctx, inputs: Sequence[bytes], rank: int, world: int, group=None):
    """Ultra-fast deflate of `inputs` split over `world` ranks; rank 0 gets every stream."""
    import torch.distributed as dist

    parts = partition_lpt([len(b) for b in inputs], world)
    mine = parts[rank]
    outs = ctx.deflate_ultrafast_batch([inputs[i] for i in mine])
    if world == 1:
        res = [b""] * len(inputs)
        for i, b in zip(mine.tolist(), outs):
            res[i] = b
        return res
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((mine.tolist(), outs), gathered, dst=0, group=group)
    if rank != 0:
        return None
    res = [b""] * len(inputs)
    for idx, o in gathered:
        for i, b in zip(idx, o):
            res[i] = b
    return res
