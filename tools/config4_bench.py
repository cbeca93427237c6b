#!/usr/bin/env python
"""BASELINE config 4: a 64 Mbp (chr20-size) reference, 1 M CLR reads sharded across the GPUs of one box.

    python tools/config4_bench.py [--ref 64000000] [--reads 1000000] [--steps 2] [--sample 48]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config4_bench.py ...

Strong scaling: the read count is fixed and rank r maps reads/N of them (its own deterministic shard of the generator:
seed 5 + 1000 r), in batches of at most --batch reads streamed through the pipelined entry points; the index is replicated (built on every rank); the only exchange is the final
reduction (one all-reduce of the counters; records stay on their rank here, bench.py shows the gather).  Rank 0 checks an evenly
spaced sample of its shard bit-exactly against the CPU oracle (incl. candidate and cell counts) and all ranks check that
located reads sit on their true locus.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", type=int, default=64_000_000)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=125_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--sample", type=int, default=48)
    ap.add_argument("--mask", default="fff0ccfc")  # weight 11: ~15 random candidates per probe at 64 Mbp
    ap.add_argument("--serial", action="store_true", help="blocking pb_locate_batch calls instead of the pipelined entry points")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    mask, R = int(a.mask, 16), 0.3

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import workload
    from pacbioassembly_b200 import Context

    t0 = time.time()
    ref = workload.reference(4, a.ref)
    mine = a.reads // world + (1 if rank < a.reads % world else 0)
    lens = workload.read_lengths(5 + 1000 * rank, mine, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
    pinned = torch.empty(int(lens.astype(np.int64).sum()), dtype=torch.uint8, pin_memory=True)  # reads wait in pinned host memory
    txt, offs, lens, starts = workload.reads(5 + 1000 * rank, ref, lens, nthreads=max(1, (os.cpu_count() or 8) // world), out=pinned.numpy())
    if rank == 0:
        print(f"workload {time.time() - t0:.1f}s: ref {a.ref}, {a.reads} reads over {world} GPU(s), {mine} on rank 0 ({len(txt)} bases)", flush=True)
    ctx = Context(local)
    t0 = time.time()
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, mask)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"index: {ix.nentries} entries, {ix.nkeys} keys, {time.time() - t0:.2f}s wall", flush=True)

    # batches of a rank: at most --batch reads, and at least four per rank so that copies and planning of one batch run under the
    # alignment of the one before (pb_locate_submit / pb_locate_collect); --serial: one blocking pb_locate_batch per batch
    bsz = a.batch if a.serial else min(a.batch, max(20_000, (mine + 3) // 4))

    def one_pass():
        parts, cand, k3 = [], 0, 0.0
        prev = None
        for b0 in range(0, mine, bsz):
            b1 = min(mine, b0 + bsz)
            base = int(offs[b0])
            end = int(offs[b1 - 1] + lens[b1 - 1])
            if a.serial:
                recs = ctx.locate(ix, txt[base:end], offs[b0:b1] - base, lens[b0:b1], R=R)  # pb_locate_batch: host text in, records out
                parts.append(recs)
                k3 += ctx.timings()["align"]
                continue
            cur = ctx.locate_submit(ix, txt[base:end], offs[b0:b1] - base, lens[b0:b1], R=R)
            if prev is not None:
                parts.append(prev.collect())
                k3 += ctx.timings()["align"]
            prev = cur
        if prev is not None:
            parts.append(prev.collect())
            k3 += ctx.timings()["align"]
        recs = np.concatenate(parts)
        return recs, int(recs["ncand"].sum()), k3

    best = None
    for it in range(a.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        recs, cand, k3 = one_pass()
        torch.cuda.synchronize()
        dt = time.time() - t0
        f = recs["found"] == 1
        c = torch.tensor([int(f.sum()), int(recs["cost"][f].sum()), int(recs["cells"].sum()), len(recs), cand], dtype=torch.int64, device="cuda")
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(c)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = [int(x) for x in c.tolist()]
        dt_all = float(t.item())
        if rank == 0:
            print(f"step {it}: {dt_all:.3f}s (max over ranks): {tot[0]}/{tot[3]} located, {tot[4]} align() calls of the reference, "
                  f"{tot[2]:.3e} reference DP cells; rank-0 K3 {k3:.1f} ms", flush=True)
        if best is None or dt_all < best[0]:
            best = (dt_all, tot)
    near = np.abs(recs["pos"][f].astype(np.int64) - recs["j"][f] - starts[f]) < 0.35 * lens[f]
    assert near.mean() > 0.99, near.mean()
    checked = 0
    if rank == 0 and a.sample > 0:
        import cpu_libs
        o = cpu_libs.oracle()
        ids = np.arange(0, mine, max(1, mine // a.sample))[: a.sample]
        s_lens = lens[ids]
        s_offs = np.zeros(len(ids), dtype=np.int64)
        np.cumsum(s_lens[:-1], out=s_offs[1:])
        s_txt = np.concatenate([txt[offs[i]: offs[i] + lens[i]] for i in ids])
        t0 = time.time()
        oix = o.index_build(ref, mask, 0)
        t_ix = time.time() - t0
        t0 = time.time()
        want = o.locate(oix, ref, s_txt, s_offs, s_lens, mask, R=R, nthreads=min(os.cpu_count() or 1, 32))
        t_cpu = time.time() - t0
        got = recs[ids]  # every read is >= 500 long here, so kept rank == read index
        for n in ("found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
            assert (got[n] == want[n]).all(), n
        checked = len(ids)
        print(f"sample of {checked} reads bit-exact vs oracle (incl. ncand and cells); located reads on their true locus: {near.mean():.4f}; "
              f"oracle: index {t_ix:.0f}s, {checked} reads in {t_cpu:.1f}s on {min(os.cpu_count() or 1, 32)} threads", flush=True)
        cpu = {"kind": "port", "cores": min(os.cpu_count() or 1, 32), "sample": f"{checked} evenly spaced reads of rank 0's shard",
               "value": checked / t_cpu, "unit": "reads/s"}
    if rank == 0:
        dt, tot = best
        line = {"config": f"config4: {a.ref} bp iid reference, {a.reads} CLR reads (mean 5 kbp, ins 9/del 4/sub 2 %) sharded over {world} GPU(s), "
                          f"mask {a.mask}, R={R}, locator.cpp semantics, host text in / records out ({'pb_locate_batch' if a.serial else 'pb_locate_submit / pb_locate_collect, batches of ' + str(bsz)})",
                "n_gpus": world, "seconds": dt, "reads_per_s": a.reads / dt, "scaling": "strong", "located": tot[0], "reads": tot[3],
                "reference_align_calls": tot[4], "reference_dp_cells": tot[2], "oracle_checked_reads": checked,
                "cpu_baseline": cpu if checked else None}
        print(json.dumps(line), flush=True)
        if a.out:
            with open(a.out, "w") as fh:
                fh.write(json.dumps(line) + "\n")
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
