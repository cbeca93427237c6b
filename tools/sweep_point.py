#!/usr/bin/env python
"""One point of the config-3 sweep (for ncu): python tools/sweep_point.py len band [npairs] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402

alen, band = int(sys.argv[1]), int(sys.argv[2])
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
P = [workload.sweep_pair(1000 * band + alen, k, alen, band) for k in range(npairs)]
A, B, R = [x[0] for x in P], [x[1] for x in P], P[0][2]
a_len = np.array([len(x) for x in A], dtype=np.int32)
b_len = np.array([len(x) for x in B], dtype=np.int32)
a_off = np.zeros(npairs, dtype=np.int64); np.cumsum(a_len[:-1], out=a_off[1:])
b_off = np.zeros(npairs, dtype=np.int64); np.cumsum(b_len[:-1], out=b_off[1:])
ctx = Context(0)
for rep in range(reps):
    recs, _ = ctx.align_batch(b"".join(A), a_off, a_len, b"".join(B), b_off, b_len, R, 26000, 6000, want_ops=False)
    t = ctx.timings()
    cells = int(recs["cells"].sum())
    print(f"len {alen} band {band} pairs {npairs}: aligned {int((recs['ret'] >= 0).sum())}, K3 {t['align']:.3f} ms, {cells / t['align'] / 1e6:.0f} GCUPS")
ctx.close()
