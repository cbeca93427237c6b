#!/usr/bin/env python
"""Host-side timeline of the pipelined locate (pb_locate_submit / pb_locate_collect): how long each call blocks.

    python tools/pipe_probe.py [nreads] [steps] [text|bin]
"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402

MASK, R = 0xff3c3ffc, 0.3


def main():
    nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    mode = sys.argv[3] if len(sys.argv) > 3 else "text"
    ref = workload.reference(2, 4_600_000)
    lens = workload.read_lengths(3, nreads)
    txt, offs, lens, _ = workload.reads(3, ref, lens)
    src = workload.pack_bin(txt, offs, lens) if mode == "bin" else txt
    pin = torch.empty(len(src), dtype=torch.uint8, pin_memory=True)
    pin.numpy()[:] = src
    buf = pin.numpy()
    ctx = Context(0)
    ix = ctx.index(ctx.seqset_one(ref), MASK)
    prev, t_start = None, time.perf_counter()
    for k in range(steps + 1):
        t0 = time.perf_counter()
        cur = None
        if k < steps:
            cur = ctx.locate_submit_bin(ix, buf, R=R) if mode == "bin" else ctx.locate_submit(ix, buf, offs, lens, R=R)
        t1 = time.perf_counter()
        n = 0
        if prev is not None:
            recs = prev.collect()
            n = int(recs["found"].sum())
        t2 = time.perf_counter()
        print(f"step {k}: submit {1e3 * (t1 - t0):7.1f} ms, collect {1e3 * (t2 - t1):7.1f} ms, located {n}, stages {ctx.timings() if prev is not None else ''}",
              flush=True)
        prev = cur
    print(f"{steps} steps in {1e3 * (time.perf_counter() - t_start):.1f} ms wall")
    ctx.close()


if __name__ == "__main__":
    main()
