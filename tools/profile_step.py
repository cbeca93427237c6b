#!/usr/bin/env python
"""One locate step on a slice of the config-2 workload, for ncu (B200_PROFILING.md): small, no torch, no CPU legs.

    python tools/profile_step.py [nreads] [steps] [lo hi]     (lo hi: uniform read lengths instead of the CLR log-normal)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402

MASK, R = 0xff3c3ffc, 0.3


def main():
    nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ref = workload.reference(2, 4_600_000)
    if len(sys.argv) > 4:
        lens = workload.read_lengths(3, nreads, sigma_log=0.0, lo=int(sys.argv[3]), hi=int(sys.argv[4]))
    else:
        lens = workload.read_lengths(3, nreads, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
    txt, offs, lens, _ = workload.reads(3, ref, lens)
    ctx = Context(0)
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASK)
    s = ctx.seqset(txt, offs, lens)
    for it in range(steps):
        t0 = time.time()
        job = ctx.locate_run(ix, s, R=R)
        recs = job.fetch()
        t, st = ctx.timings(), job.stats()
        job.free()
        print(f"step {it}: {len(recs)} reads, {int(recs['found'].sum())} located, {int(recs['cells'].sum())} ref-equivalent cells, "
              f"{time.time() - t0:.3f}s wall, align {t['align']:.2f} ms, launches so far {ctx.launches}, K3 stats {st}")
    print("seed bulk:", s.seeds_device(MASK))
    ctx.close()


if __name__ == "__main__":
    main()
