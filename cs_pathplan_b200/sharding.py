"""Trajectory-index sharding of a batch across the GPUs of one node.

Trajectories are independent, so the data path has NO collective: every rank solves a contiguous range of
trajectory indices.  Uniform batches split evenly; ragged (CSR) batches split where the running segment count
crosses k/world of the total, so ranks get equal work (sum of ns), not equal trajectory counts.  The only cross-rank
artefact is the exclusive scan of per-rank sample counts needed to place each rank's samples in one global CSR
output; that is host-side bookkeeping off the timed path (``global_sample_base``).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_bounds(B: int, world: int, seg_offset: Optional[np.ndarray] = None) -> List[Tuple[int, int]]:
    """Contiguous trajectory ranges [b0, b1) for ranks 0..world-1; together they cover [0, B) exactly once."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if B < 0:
        raise ValueError("B must be >= 0")
    if seg_offset is None:
        cuts = [(B * r) // world for r in range(world + 1)]
    else:
        so = np.asarray(seg_offset, dtype=np.int64)
        if so.shape[0] != B + 1:
            raise ValueError("seg_offset must have B + 1 entries")
        total = int(so[-1])
        cuts = [0]
        for r in range(1, world):
            target = (total * r) // world
            # first trajectory boundary at or after the target, never moving backwards
            b = int(np.searchsorted(so, target, side="left"))
            cuts.append(min(max(b, cuts[-1]), B))
        cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_batch(waypoints: np.ndarray, rank: int, world: int, ns: Optional[int] = None,
                seg_offset: Optional[np.ndarray] = None):
    """This rank's slice of a batch.  Returns (waypoints_local, ns or None, seg_offset_local or None, (b0, b1))."""
    wp = np.asarray(waypoints)
    if seg_offset is None:
        if ns is None or ns < 1 or wp.shape[0] % (ns + 1):
            raise ValueError("uniform batch needs ns >= 1 and rows divisible by ns + 1")
        B = wp.shape[0] // (ns + 1)
        b0, b1 = shard_bounds(B, world)[rank]
        return wp[b0 * (ns + 1): b1 * (ns + 1)], ns, None, (b0, b1)
    so = np.asarray(seg_offset, dtype=np.int64)
    B = so.shape[0] - 1
    b0, b1 = shard_bounds(B, world, so)[rank]
    p0, p1 = int(so[b0]) + b0, int(so[b1]) + b1
    return wp[p0:p1], None, so[b0: b1 + 1] - so[b0], (b0, b1)


def shard_rows(rows: np.ndarray, row_offset: np.ndarray, rank: int, world: int):
    """This rank's slice of a batch of sampled trajectories (rows [n, 3] with CSR row_offset [B+1]) -- the layout the
    WGS84 <-> ENU and altitude-optimisation stages work on.  Balanced on the number of rows.  Returns
    (rows_local, row_offset_local, (b0, b1)); rows_local is a view of the caller's array."""
    off = np.asarray(row_offset, dtype=np.int64)
    B = off.shape[0] - 1
    b0, b1 = shard_bounds(B, world, off)[rank]
    return np.asarray(rows)[int(off[b0]): int(off[b1])], off[b0: b1 + 1] - off[b0], (b0, b1)


def global_sample_base(local_rows: int, group=None) -> Tuple[int, int]:
    """Exclusive scan of per-rank sample-row counts over the process group (host-side, off the timed path).
    Returns (first global row of this rank, total rows).  Works with the gloo and nccl backends."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return 0, int(local_rows)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([int(local_rows)], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    counts = [int(t.item()) for t in allc]
    return sum(counts[:rank]), sum(counts)
