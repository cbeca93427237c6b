#!/usr/bin/env python
"""A small pass through every pipeline (locate, all-vs-all, consensus rounds, bulk probe): a quick whole-library smoke,
sized for a run under a checker (compute-sanitizer is closed on this pool: it answered exit 86 without running)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workload  # noqa: E402
from allpairs_util import allpairs_workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402
from pacbioassembly_b200.assemble import assemble_rounds  # noqa: E402

MASK = 0xff3c3ffc
ctx = Context(0)
ref = workload.reference(2, 60_000)
lens = workload.read_lengths(3, 48, mean=1500.0, sigma_log=0.5, lo=500, hi=5000)
txt, offs, lens, _ = workload.reads(3, ref, lens, 0.05, 0.03, 0.02, nthreads=1)
rs = ctx.seqset_one(ref)
ix = ctx.index(rs, MASK)
recs, ops = ctx.locate(ix, txt, offs, lens, want_ops=True, R=0.3)
print("locate:", int(recs["found"].sum()), "of", len(recs))
texts, image = allpairs_workload(301, 6000, 40)
reads = ctx.seqset_from_bin(image)
aix = ctx.index_set(reads, MASK)
pr, st = ctx.overlap_all(aix, found_only=False)
print("all-vs-all:", st)
ref0 = np.frombuffer(texts[3], dtype=np.uint8)
cons, fr, _, passes = assemble_rounds(ctx, ref0, image, [MASK, 0x3fcfccf3], seed_at_quirk=1)
print("assemble:", [len(c) for c in cons], int((fr > 0).sum()), passes)
p = ix.probe_bulk(ctx.seqset(txt, offs, lens))
print("probe bulk:", p["queries"], p["candidates"])
ctx.close()
print("ok")
