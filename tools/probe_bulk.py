#!/usr/bin/env python
"""K2 bulk measurement on its own (A/B knob: PB_PROBE_V = uint4 groups per thread in the count pass, PB_PROBE1 = the
one-query-per-thread kernels): every position of the config-2 read set probed against the config-2 index.

    python tools/probe_bulk.py [nreads]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402

nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
ref = workload.reference(2, 4_600_000)
lens = workload.read_lengths(3, nreads, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
txt, offs, lens, _ = workload.reads(3, ref, lens)
ctx = Context(0)
rs = ctx.seqset_one(ref)
ix = ctx.index(rs, 0xff3c3ffc)
s = ctx.seqset(txt, offs, lens)
for it in range(3):
    p = ix.probe_bulk(s)
q, c = p["queries"], p["candidates"]
print(f"PB_PROBE_V={os.environ.get('PB_PROBE_V', '-')} PB_PROBE1={os.environ.get('PB_PROBE1', '-')}: {q} queries, {c} candidates; "
      f"count {p['count_ms']:.3f} ms = {12.0 * q / p['count_ms'] / 1e6:.0f} GB/s algorithmic, {q / p['count_ms'] / 1e6:.1f} G queries/s; "
      f"gather {p['gather_ms']:.3f} ms = {(12.0 * q + 12.0 * c) / p['gather_ms'] / 1e6:.0f} GB/s algorithmic; seed {p['seed_ms']:.3f} ms")
for mb in (2, 16, 64, 256, 1024):
    r = ctx.random_gather_peak(mb << 20)
    print(f"random 4-byte reads of a {mb} MB table: {r / 1e9:.1f} G reads/s ({r * 32 / 1e9:.0f} GB/s of sectors); "
          f"count pass = {q / (p['count_ms'] / 1e3) / r:.2f} of it")
ctx.close()
