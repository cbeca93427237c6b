#!/usr/bin/env python
"""BASELINE config 4 at single-GPU scale: 64 Mbp reference, CLR reads; checks a subsample bit-exactly against the CPU
oracle (C port; the compiled reference's fixed 1.9 GB matrix and 4 s/Mbp map build make it impractical here) and the
whole batch against size-independent properties.

    python tools/scale_check.py [ref_len] [nreads] [nsample] [mask_hex]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_libs  # noqa: E402
import workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402


def main():
    ref_len = int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000
    nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    nsample = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    mask = int(sys.argv[4], 16) if len(sys.argv) > 4 else 0xfff0ccfc  # weight 11: ~15 random candidates per probe
    R = 0.3
    t0 = time.time()
    ref = workload.reference(4, ref_len)
    lens = workload.read_lengths(5, nreads, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
    txt, offs, lens, starts = workload.reads(5, ref, lens)
    print(f"workload {time.time() - t0:.1f}s: ref {ref_len}, {nreads} reads, {len(txt)} bases", flush=True)
    ctx = Context(0)
    t0 = time.time()
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, mask)
    print(f"index: {ix.nentries} entries, {ix.nkeys} keys, {time.time() - t0:.2f}s, {ctx.timings()}", flush=True)
    for it in range(2):
        t0 = time.time()
        s = ctx.seqset(txt, offs, lens)
        job = ctx.locate_run(ix, s, R=R)
        recs = job.fetch()
        st = job.stats()
        print(f"step {it}: {time.time() - t0:.3f}s wall, {int(recs['found'].sum())}/{len(recs)} located, "
              f"{st['ncand']} candidates, {st['dp_alignments']} DP runs, {st['dp_cells']:.3e} DP cells, {ctx.timings()}", flush=True)
        job.free()
        s.free()
    f = recs["found"] == 1
    near = np.abs(recs["pos"][f].astype(np.int64) - recs["j"][f] - starts[f]) < 0.35 * lens[f]
    print(f"located reads whose position is their true locus: {near.mean():.4f}")
    assert near.mean() > 0.99
    # subsample against the oracle
    ids = np.arange(0, nreads, max(1, nreads // nsample))[:nsample]
    s_lens = lens[ids]
    s_offs = np.zeros(len(ids), dtype=np.int64)
    np.cumsum(s_lens[:-1], out=s_offs[1:])
    s_txt = np.concatenate([txt[offs[i]: offs[i] + lens[i]] for i in ids])
    o = cpu_libs.oracle()
    t0 = time.time()
    oix = o.index_build(ref, mask, 0)
    print(f"oracle index {time.time() - t0:.1f}s; nkeys/nentries oracle {o.index_stats(oix)} gpu {(ix.nkeys, ix.nentries)}", flush=True)
    assert o.index_stats(oix) == (ix.nkeys, ix.nentries)
    t0 = time.time()
    want = o.locate(oix, ref, s_txt, s_offs, s_lens, mask, R=R, nthreads=min(os.cpu_count() or 1, 32))
    print(f"oracle locate of {len(ids)} reads {time.time() - t0:.1f}s", flush=True)
    got = recs[ids]  # every read is >= 500 long here, so kept rank == read index
    for n in ("found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
        assert (got[n] == want[n]).all(), n
    print(f"subsample of {len(ids)} reads bit-exact vs oracle (incl. ncand and cells); mean candidates/read {want['ncand'].mean():.1f}")
    ctx.close()


if __name__ == "__main__":
    main()
