"""Chunk sharding for multi-GPU inference: backbone passes over chunks / test
fragments / `_chunked_forward` sub-chunks are independent (pointcept/models/default.py:134-176,
pointcept/engines/test.py:315-349), so ranks take disjoint chunk sets and never exchange activations.
Pure host logic (no CUDA); covered by world_size-2 gloo tests."""
from __future__ import annotations

from typing import List, Sequence


def assign_chunks(sizes: Sequence[int], world_size: int, policy: str = "lpt") -> List[List[int]]:
    """Partition chunk indices over ranks.  "round_robin": chunk i -> rank i % world (the DistributedSampler
    layout of pointcept/engines/test.py:94); "lpt": longest-processing-time-first greedy on the voxel counts
    (imbalance comes only from unequal chunk sizes).  Deterministic; every rank computes the same table."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    out: List[List[int]] = [[] for _ in range(world_size)]
    if policy == "round_robin":
        for i in range(len(sizes)):
            out[i % world_size].append(i)
        return out
    if policy != "lpt":
        raise ValueError(policy)
    load = [0] * world_size
    for i in sorted(range(len(sizes)), key=lambda j: (-int(sizes[j]), j)):
        r = min(range(world_size), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += int(sizes[i])
    for r in range(world_size):
        out[r].sort()
    return out


def chunk_ranges(n: int, chunk_size: int):
    """Contiguous index-range chunks of `_chunked_forward` (default.py:134-139)."""
    return [(s, min(s + chunk_size, n)) for s in range(0, n, chunk_size)]


def job_throughput(units_per_rank: Sequence[float], ms_per_rank: Sequence[float]) -> float:
    """Whole-job units/s = all ranks' units / the slowest rank's device time."""
    return float(sum(units_per_rank)) / (max(ms_per_rank) * 1e-3)
