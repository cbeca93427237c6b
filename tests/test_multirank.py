"""N>1 host logic on CPU: world_size 2 over gloo (no GPU): sharding, the counter all-reduce, the record gather and the
global nseq renumbering.  Records are produced by the CPU oracle here; on the GPU box bench.py feeds the same functions
with the CUDA path's records over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_libs
    import workload
    from pacbioassembly_b200 import shard
    from pacbioassembly_b200.api import LOCATE_DTYPE

    mask = 0xff3c3ffc
    ref = workload.reference(71, 60000)
    lens = workload.read_lengths(72, 24, mean=900.0, sigma_log=0.5, lo=300, hi=2500)
    txt, offs, lens, _ = workload.reads(73, ref, lens, 0.03, 0.02, 0.01, nthreads=1)
    o = cpu_libs.oracle()
    ix = o.index_build(ref, mask, 0)
    lo, hi = shard.shard_ranges(lens, world)[rank]
    mine = o.locate(ix, ref, txt, offs[lo:hi], lens[lo:hi], mask, R=0.3)
    recs = np.zeros(len(mine), dtype=LOCATE_DTYPE)
    for n in LOCATE_DTYPE.names:
        if n != "_pad":
            recs[n] = mine[n]
    tot = shard.reduce_counters(recs)
    allrecs = shard.gather_records(recs)
    tot2, allrecs2 = shard.FinalReduction(len(lens))(recs)  # the buffered variant bench.py uses
    assert tot2.tolist() == tot.tolist()
    if rank == 0:
        assert all((allrecs2[n] == allrecs[n]).all() for n in allrecs.dtype.names)
    if rank == 0:
        whole = o.locate(ix, ref, txt, offs, lens, mask, R=0.3)
        ok = len(allrecs) == len(whole)
        for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
            ok = ok and bool((allrecs[n] == whole[n]).all())
        f = whole["found"] == 1
        ok = ok and tot.tolist() == [int(f.sum()), int(whole["cost"][f].sum()), int(whole["cells"].sum()), len(whole)]
        q.put(ok)
    o.index_free(ix)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_balance_by_bases():
    sys.path.insert(0, ROOT)
    from pacbioassembly_b200 import shard
    lens = np.array([100] * 50 + [5000] * 10, dtype=np.int32)
    rs = shard.shard_ranges(lens, 4)
    assert rs[0][0] == 0 and rs[-1][1] == len(lens)
    assert all(rs[i][1] == rs[i + 1][0] for i in range(3))
    bases = [int(lens[a:b].sum()) for a, b in rs]
    assert max(bases) - min(bases) <= 5000
    assert shard.shard_ranges(lens, 1) == [(0, 60)]
    assert shard.shard_ranges(np.zeros(0, np.int32), 2) == [(0, 0), (0, 0)]


def test_world2_gloo_reduction_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


# ---- all-vs-all (config 5): query reads sharded, set replicated ------------------------------------------------------

def _pairs_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_libs
    from allpairs_util import PAIR_FIELDS, allpairs_workload, expected_pairs, kept_texts, oracle_overlap_fn
    from pacbioassembly_b200 import shard
    from pacbioassembly_b200.api import PAIR_DTYPE

    texts, image = allpairs_workload(361, 5000, 24)
    kt = kept_texts(texts)
    pairs, tot_ncand, tot_cells = expected_pairs(oracle_overlap_fn(cpu_libs.oracle(), 0xff3c3ffc), texts, image)

    def records(lo, hi):  # what a rank's pb_overlap_all_run(q_first=lo, q_count=hi-lo) returns with found_only
        rows = sorted((Q, T) for (T, Q), r in pairs.items() if lo <= Q < hi and r["found"])
        out = np.zeros(len(rows), dtype=PAIR_DTYPE)
        for i, (Q, T) in enumerate(rows):
            r = pairs[(T, Q)]
            out[i]["read_id"], out[i]["ref_id"], out[i]["found"] = Q, T, 1
            for n in PAIR_FIELDS + ("ncand", "cells"):
                out[i][n] = r[n]
        return out

    def stats(lo, hi):
        mine = {k: r for k, r in pairs.items() if lo <= k[1] < hi}
        return {"pairs": len(mine), "pairs_found": sum(int(r["found"]) for r in mine.values()),
                "try_align_calls": sum(int(r["ncand"]) for r in mine.values()), "ref_cells": sum(int(r["cells"]) for r in mine.values())}

    lens = np.array([len(t) for t in kt], dtype=np.int32)
    lo, hi = shard.shard_ranges(lens, world)[rank]
    tot = shard.reduce_pair_stats(stats(lo, hi))
    allrecs = shard.gather_pair_records(records(lo, hi))
    if rank == 0:
        whole = records(0, len(kt))
        ok = len(allrecs) == len(whole) > 20 and all((allrecs[n] == whole[n]).all() for n in PAIR_DTYPE.names)
        ok = ok and tot["pairs"] == len(pairs) and tot["try_align_calls"] == tot_ncand and tot["ref_cells"] == tot_cells
        ok = ok and tot["pairs_found"] == len(whole)
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_allpairs_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_pairs_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
