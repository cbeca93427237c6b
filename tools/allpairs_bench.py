#!/usr/bin/env python
"""BASELINE config 5: all-vs-all read overlap detection (spaced-seed hits + banded verify) on 50 k reads.

    python tools/allpairs_bench.py [--genome 4600000] [--reads 50000] [--batch 50000] [--targets 3] [--steps 2]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allpairs_bench.py ...

Every rank holds the whole read set and its seed index (replicated); the QUERY reads are sharded over the ranks in
contiguous ranges balanced by bases, each range processed in batches of --batch reads (bounds the candidate arrays).  The only
exchange is the final reduction: an all-reduce of the counters and a gather of the found 56-byte records on rank 0.
Parity: for --targets sampled reads T, every pair (T, Q) is recomputed by the CPU oracle (the reference's trial loop with T
as the locked reference) and compared field by field with the GPU's records for that T; plus size-independent checks on the
whole result (found pairs really overlap on the genome).
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
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--reads", type=int, default=50_000)
    ap.add_argument("--mean", type=float, default=5000.0)
    ap.add_argument("--batch", type=int, default=50_000)
    ap.add_argument("--queries", type=int, default=0, help="only the first N query reads (0 = all)")
    ap.add_argument("--targets", type=int, default=3, help="target reads checked against the CPU oracle")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--mask", default="ff3c3ffc")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    mask = int(a.mask, 16)
    R = 0.3

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
    from pacbioassembly_b200.api import PAIR_DTYPE
    from pacbioassembly_b200.shard import gather_pair_records, reduce_pair_stats, shard_ranges

    t0 = time.time()
    g = workload.reference(2, a.genome)
    lens = workload.read_lengths(7, a.reads, mean=a.mean, sigma_log=0.5, lo=501, hi=19999)
    txt, offs, lens, starts = workload.reads(8, g, lens)
    if rank == 0:
        print(f"workload {time.time() - t0:.1f}s: genome {a.genome}, {a.reads} reads, {len(txt)} bases "
              f"({len(txt) / a.genome:.1f}x)", flush=True)
    ctx = Context(local)
    t0 = time.time()
    rs = ctx.seqset(txt, offs, lens)
    ix = ctx.index_set(rs, mask)
    torch.cuda.synchronize()
    t_index = time.time() - t0
    if rank == 0:
        print(f"set + index: {ix.nentries} entries, {ix.nkeys} keys, {t_index:.2f}s wall, {ctx.timings()}", flush=True)
    nq = a.queries or a.reads
    q0, q1 = shard_ranges(lens[:nq], world)[rank]

    def one_pass():
        parts, stats = [], None
        tm = {}
        for b0 in range(q0, q1, a.batch):
            recs, st = ctx.overlap_all(ix, b0, min(a.batch, q1 - b0), found_only=True, R=R)
            parts.append(recs)
            stats = st if stats is None else {k: stats[k] + st[k] for k in st}
            for k, v in ctx.timings().items():
                tm[k] = tm.get(k, 0.0) + v
        recs = np.concatenate(parts) if parts else np.zeros(0, dtype=PAIR_DTYPE)
        return recs, stats or {}, tm

    results = []
    for it in range(a.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches
        t0 = time.time()
        recs, stats, tm = one_pass()
        torch.cuda.synchronize()
        dt = time.time() - t0
        launches = ctx.launches - l0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_all = float(t.item())
            tot = reduce_pair_stats(stats, "cuda")
        else:
            dt_all, tot = dt, stats
        results.append((dt_all, tot, tm, launches))
        if rank == 0:
            print(f"step {it}: {dt_all:.3f}s (max over ranks), queries {nq}: {tot}", flush=True)
            print(f"   rank-0 stage ms: { {k: round(v, 2) for k, v in tm.items()} }", flush=True)

    # final reduction: gather the found records on rank 0
    if world > 1:
        recs = gather_pair_records(recs, "cuda")
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    dt, tot, tm, launches = min(results, key=lambda x: x[0])
    # size-independent check: a found pair's reads must overlap on the genome where the alignment says
    T, Q = recs["ref_id"], recs["read_id"]
    assert ((T != Q) & (recs["found"] == 1)).all()
    key = Q.astype(np.int64) * a.reads + T
    assert (np.diff(key) > 0).all(), "records are not in ascending (read, reference) order"
    # forward: read position read_pos aligns to reference position ref_pos; backward: the seed's last base (pos+15)
    gq = starts[Q] + recs["read_pos"] / 1.05  # read coordinates run ~5 % ahead of genome coordinates (ins 9 %, del 4 %)
    gt = starts[T] + recs["ref_pos"] / 1.05
    near = np.abs(gq - gt) < 0.25 * np.maximum(recs["read_pos"], recs["ref_pos"]) + 200
    print(f"found pairs whose seed anchors agree on the genome: {near.mean():.4f} of {len(recs)}")
    assert near.mean() > 0.98

    # sampled targets against the CPU oracle
    checked = 0
    cpu_s, cpu_targets = 0.0, 0
    if a.targets > 0:
        import cpu_libs
        o = cpu_libs.oracle()
        t0 = time.time()
        image = b"".join(o.text2bin(txt[offs[k]: offs[k] + lens[k]].tobytes()) for k in range(nq))
        print(f"oracle image of the {nq} query reads: {time.time() - t0:.1f}s", flush=True)
        tids = np.linspace(0, a.reads - 1, a.targets + 2).astype(int)[1:-1]
        for t in tids:
            tt = txt[offs[t]: offs[t] + lens[t]]
            t0 = time.time()
            oix = o.index_build(tt, mask, policy=1)
            want = o.overlap(oix, tt, image, mask, R=R, quirk=False, nthreads=os.cpu_count() or 8)
            o.index_free(oix)
            assert len(want) == nq
            wf = {int(q): want[q] for q in np.nonzero(want["found"] == 1)[0] if q != t}
            got = {int(r["read_id"]): r for r in recs[recs["ref_id"] == t]}
            assert sorted(got) == sorted(wf), (t, sorted(set(got) ^ set(wf))[:10])
            for q, w in wf.items():
                for f in ("j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
                    assert int(got[q][f]) == int(w[f]), (t, q, f, int(got[q][f]), int(w[f]))
            checked += len(wf)
            cpu_s += time.time() - t0
            cpu_targets += 1
            print(f"target {t} (len {lens[t]}): {len(wf)} overlapping reads, bit-exact vs oracle ({time.time() - t0:.1f}s CPU)", flush=True)
    line = {"config": f"config5: all-vs-all, {a.reads} CLR reads (mean {a.mean:.0f}, ins 9/del 4/sub 2 %) from a {a.genome} bp genome, "
                      f"mask {a.mask}, R={R}, max_trial 32, queries {nq}",
            "n_gpus": world, "seconds": dt, "query_reads_per_s": nq / dt, "pairs_found": tot.get("pairs_found"),
            "stats": tot, "rank0_stage_ms": tm, "gpu_launches": launches, "index_build_s": t_index,
            "tcups_k3": tot.get("k3_cells", 0) / max(sum(v for k, v in tm.items() if k == "align"), 1e-9) / 1e9 if world == 1 else None,
            "oracle_checked_pairs": checked,
            # the reference's way to do this job: one seed map + one trial loop over all reads per target read (C port of the
            # reference, all host threads); a reported baseline next to the GPU figure, extrapolated from the sampled targets
            "cpu_baseline": None if not cpu_targets else {
                "kind": "port", "cores": os.cpu_count(), "sample": f"{cpu_targets} target reads x {nq} query reads",
                "seconds_per_target": cpu_s / cpu_targets, "est_seconds_all_targets": cpu_s / cpu_targets * a.reads}}
    print(json.dumps(line), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            f.write(json.dumps(line) + "\n")
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
