"""Partitioning a collection of independent archives over ranks / GPUs: archives are the unit, nothing is
exchanged between ranks, so this is all the "parallelism strategy" the path has."""
import heapq
from typing import List, Sequence


def partition(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy: indices of the archives each rank decodes, balanced by compressed size."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + sizes[i], r))
    for lst in out:
        lst.sort()
    return out


def decode_collection(archives: Sequence[bytes], rank: int, world_size: int, device: int = 0, _library=None, **fields):
    """Decodes this rank's share of `archives`.  Returns (indices, results); no communication happens here."""
    from .decoder import decode_batch
    mine = partition([len(a) for a in archives], world_size)[rank]
    return mine, decode_batch([archives[i] for i in mine], device=device, _library=_library, **fields)
