"""Multi-GPU host logic: reads are independent given the index, so they are sharded across ranks (one process per GPU)
with the index replicated; the only exchange is the final hit/score reduction.

Backend-agnostic torch.distributed code: NCCL with CUDA tensors on the GPU box, gloo with CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .api import LOCATE_DTYPE


def shard_ranges(lens: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous read-index ranges, balanced by total bases (not read count): DP work grows with length."""
    lens = np.asarray(lens, dtype=np.int64)
    n = len(lens)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    cum = np.cumsum(lens)
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        cuts.append(int(np.searchsorted(cum, target, side="left")) + (1 if target > 0 else 0))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def counters_of(recs: np.ndarray) -> np.ndarray:
    """[mapped reads, sum of costs over mapped reads, reference-equivalent DP cells, kept reads]"""
    f = recs["found"] == 1
    return np.array([int(f.sum()), int(recs["cost"][f].sum()), int(recs["cells"].sum()), len(recs)], dtype=np.int64)


def reduce_counters(recs: np.ndarray, device: torch.device | str = "cpu") -> np.ndarray:
    """The path's one all-reduce: hit/score counters summed over ranks."""
    t = torch.from_numpy(counters_of(recs)).to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


def gather_records(recs: np.ndarray, device: torch.device | str = "cpu", dst: int = 0):
    """Gather every rank's records on `dst`, renumbering nseq (rank among kept reads, locator.cpp:72,84) into the
    global order of the contiguous shards.  Returns the concatenated array on dst, None elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return recs.copy()
    rank = dist.get_rank()
    n = torch.tensor([len(recs)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros(cap * LOCATE_DTYPE.itemsize, dtype=torch.uint8, device=device)
    raw = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).reshape(-1).copy())
    buf[: raw.numel()].copy_(raw)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    parts, base = [], 0
    for r in range(world):
        a = out[r].cpu().numpy()[: counts[r] * LOCATE_DTYPE.itemsize].view(LOCATE_DTYPE).copy()
        a["nseq"] += base  # prefix sum of the kept counts of the earlier shards (SURVEY Q-L1)
        base += counts[r]
        parts.append(a)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=LOCATE_DTYPE)
