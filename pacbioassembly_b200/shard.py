"""Multi-GPU host logic: reads are independent given the index, so they are sharded across ranks (one process per GPU)
with the index replicated; the only exchange is the final hit/score reduction.

Backend-agnostic torch.distributed code: NCCL with CUDA tensors on the GPU box, gloo with CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .api import LOCATE_DTYPE, PAIR_DTYPE, PAIR_STATS


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


# ---- all-vs-all (config 5): the QUERY reads are sharded, the set and its index are replicated -----------------------

def reduce_pair_stats(stats: dict, device: torch.device | str = "cpu") -> dict:
    """all-reduce (sum) of pb_pairs_job_stats over the ranks"""
    t = torch.tensor([int(stats.get(k, 0)) for k in PAIR_STATS], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return dict(zip(PAIR_STATS, (int(x) for x in t.cpu().tolist())))


def gather_pair_records(recs: np.ndarray, device: torch.device | str = "cpu", dst: int = 0):
    """Gather every rank's pair records on `dst`.  Shards are contiguous query ranges and each rank's records come in ascending
    (read_id, ref_id), so the concatenation in rank order is globally sorted; ids are set-wide already (no renumbering)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return recs.copy()
    rank = dist.get_rank()
    n = torch.tensor([len(recs)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1) * PAIR_DTYPE.itemsize
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    raw = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).reshape(-1).copy())
    buf[: raw.numel()].copy_(raw)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([out[r].cpu().numpy()[: counts[r] * PAIR_DTYPE.itemsize].view(PAIR_DTYPE).copy() for r in range(world)])


class FinalReduction:
    """The path's one exchange, with its buffers allocated once: an all-reduce of the four counters and a gather of every
    rank's 56-byte records on rank 0 (fixed-size slots of `cap` records behind an 8-byte count).  Both are queued on a side
    stream and nothing blocks the host: a step's exchange runs under the NEXT step's kernels (a collective's kernel needs a
    free SM slot, which the aligner's persistent kernels only give up when their band class drains -- waiting for it on the
    host would stall the pipeline for most of a step).  `wait()` joins the exchange in flight; the next `__call__`, `totals()`,
    `records()` and `finish()` call it.  sync=True is the blocking form.  NCCL (device="cuda") or gloo (device="cpu")."""

    def __init__(self, cap: int, device="cpu", dst: int = 0):
        self.cap, self.device, self.dst = int(cap), device, dst
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        nbytes = self.cap * LOCATE_DTYPE.itemsize + 8  # 8-byte header: number of valid records
        self.cuda = str(device).startswith("cuda")
        self.h_send = torch.empty(nbytes, dtype=torch.uint8, pin_memory=self.cuda)
        self.d_send = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.counters = torch.zeros(4, dtype=torch.int64, device=device)
        self.h_counters = torch.zeros(4, dtype=torch.int64, pin_memory=self.cuda)
        self.h_totals = torch.zeros(4, dtype=torch.int64, pin_memory=self.cuda)
        self.side = torch.cuda.Stream() if self.cuda else None
        self.pending = None  # (work handles, event or None) of the exchange in flight
        self.local = None
        if self.rank == dst and self.world > 1:
            self.d_recv = [torch.empty(nbytes, dtype=torch.uint8, device=device) for _ in range(self.world)]
            self.h_recv = torch.empty(self.world * nbytes, dtype=torch.uint8, pin_memory=self.cuda)
        else:
            self.d_recv = None

    def send_view(self, n: int) -> np.ndarray:
        """n records of the pinned send buffer: results fetched straight into it need no staging copy in __call__
        (join the exchange in flight first: wait())"""
        assert n <= self.cap
        return self.h_send.numpy()[8: 8 + n * LOCATE_DTYPE.itemsize].view(LOCATE_DTYPE)

    def wait(self):
        """join the exchange started by the last call (a no-op when none is in flight)"""
        if self.pending is None:
            return
        works, ev = self.pending
        self.pending = None
        if ev is not None:
            ev.synchronize()
            return
        for w in works:
            w.wait()
        self.h_totals.copy_(self.counters)
        if self.rank == self.dst:
            nbytes = self.d_send.numel()
            for r in range(self.world):
                self.h_recv[r * nbytes: (r + 1) * nbytes].copy_(self.d_recv[r])

    def finish(self):
        self.wait()
        return self.totals()

    def totals(self) -> np.ndarray:
        """counters of the last call summed over the ranks: [mapped reads, sum of costs, reference-equivalent cells, kept reads]"""
        self.wait()
        return self.h_totals.numpy().copy()

    def __call__(self, recs: np.ndarray, want_records: bool = True, sync: bool = True):
        """sync: returns (summed counters as numpy int64[4], concatenated records on dst or None); else (None, None) at once"""
        self.wait()  # the send buffers are reused
        self.h_counters.numpy()[:] = counters_of(recs)
        if self.world == 1:
            self.local = recs
            self.h_totals.numpy()[:] = self.h_counters.numpy()
            return self.h_totals.numpy().copy(), (recs if want_records else None)
        n = len(recs)
        assert n <= self.cap
        hs = self.h_send.numpy()
        hs[:8].view(np.int64)[0] = n
        if n and recs.ctypes.data != hs[8:].ctypes.data:  # not already fetched into the send buffer
            hs[8: 8 + n * LOCATE_DTYPE.itemsize] = recs.view(np.uint8).reshape(-1)
        if self.cuda:
            with torch.cuda.stream(self.side):
                self.counters.copy_(self.h_counters, non_blocking=True)
                w1 = dist.all_reduce(self.counters, async_op=True)
                w1.wait()  # orders the side stream behind the collective, does not block the host
                self.h_totals.copy_(self.counters, non_blocking=True)
                self.d_send.copy_(self.h_send, non_blocking=True)
                w2 = dist.gather(self.d_send, self.d_recv, dst=self.dst, async_op=True)
                w2.wait()
                if self.rank == self.dst:
                    nbytes = self.d_send.numel()
                    for r in range(self.world):
                        self.h_recv[r * nbytes: (r + 1) * nbytes].copy_(self.d_recv[r], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.side)
            self.pending = ((w1, w2), ev)
        else:
            self.counters.copy_(self.h_counters)
            self.d_send.copy_(self.h_send)
            w1 = dist.all_reduce(self.counters, async_op=True)
            w2 = dist.gather(self.d_send, self.d_recv, dst=self.dst, async_op=True)
            self.pending = ((w1, w2), None)
        if not sync:
            return None, None
        tot = self.totals()
        if not want_records or self.rank != self.dst:
            return tot, None
        return tot, self.records()

    def records(self) -> np.ndarray:
        """dst only, after a call: every rank's records as one array, nseq renumbered into the global kept order"""
        self.wait()
        if self.world == 1:
            return self.local.copy()
        nbytes = self.d_send.numel()
        hr = self.h_recv.numpy()
        parts, base = [], 0
        for r in range(self.world):
            blk = hr[r * nbytes: (r + 1) * nbytes]
            cnt = int(blk[:8].view(np.int64)[0])
            a = blk[8: 8 + cnt * LOCATE_DTYPE.itemsize].view(LOCATE_DTYPE).copy()
            a["nseq"] += base  # prefix sum of the kept counts of the earlier shards (SURVEY Q-L1)
            base += cnt
            parts.append(a)
        return np.concatenate(parts)
