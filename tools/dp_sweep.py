#!/usr/bin/env python
"""BASELINE config 3: banded-DP sweep, band 32..512 x read length 1..20 kbp, unit costs (the reference's scoring;
`quality.cpp` is not a scoring function).  Pairs: a = mutated copy of b's prefix with edits spaced so the
early-failure line cost(i,i) <= i*R is respected (otherwise the pair aborts after ~11 rows and "GCUPS" means nothing),
R = (band - 0.5) / len so that max_dst == band.

    python tools/dp_sweep.py [npairs] [--check N] [--weighted]
        -> JSON lines + a summary table; --check compares N pairs per point with the CPU oracle (bit-exact, transcripts
           included); --weighted runs the quality-weighted EXTENSION (pb_align_weighted_batch: per-base weights 1..4 from the
           PRNG, early-failure line scaled by the largest weight) against the extended oracle instead
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workload  # noqa: E402
from pacbioassembly_b200 import Context  # noqa: E402


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 4096
    ncheck = int(sys.argv[sys.argv.index("--check") + 1]) if "--check" in sys.argv else 0
    weighted = "--weighted" in sys.argv
    rng = np.random.default_rng(2024)
    ctx = Context(0)
    int_peak = ctx.int_pipe_peak()  # warp instructions per second of a register-only LOP3 / SHF kernel, measured now
    oracle = None
    if ncheck:
        import cpu_libs
        oracle = cpu_libs.oracle()
    rows = []
    for alen in (1000, 2000, 5000, 10000, 19999):
        for band in (32, 64, 128, 256, 512):
            A, B = [], []
            R = (band - 0.5) / alen
            t0 = time.time()
            for k in range(npairs):
                a, b, R = workload.sweep_pair(1000 * band + alen, k, alen, band)
                A.append(a)
                B.append(b)
            a_blob, b_blob = b"".join(A), b"".join(B)
            a_len = np.array([len(x) for x in A], dtype=np.int32)
            b_len = np.array([len(x) for x in B], dtype=np.int32)
            a_off = np.zeros(npairs, dtype=np.int64); np.cumsum(a_len[:-1], out=a_off[1:])
            b_off = np.zeros(npairs, dtype=np.int64); np.cumsum(b_len[:-1], out=b_off[1:])
            best = None
            if weighted:
                wa = rng.integers(1, 5, size=len(a_blob)).astype(np.uint8)
                wb = rng.integers(1, 5, size=len(b_blob)).astype(np.uint8)
            for rep in range(3):
                if weighted:
                    recs, ops = ctx.align_weighted_batch(a_blob, wa, a_off, a_len, b_blob, wb, b_off, b_len, R, 4.0, 26000, 6000,
                                                         want_ops=(rep == 0))
                else:
                    recs, ops = ctx.align_batch(a_blob, a_off, a_len, b_blob, b_off, b_len, R, 26000, 6000, want_ops=(rep == 0))
                t = ctx.timings()
                if rep == 0:
                    keep_ops = ops
                best = t["align"] if best is None else min(best, t["align"])
            cells = int(recs["cells"].sum())
            ok = int((recs["ret"] >= 0).sum())
            assert (recs["max_dst"] == band).all(), (alen, band, recs["max_dst"][:4])
            assert ok == npairs, (alen, band, ok)  # every pair of the generator aligns: the GCUPS include the traceback
            checked = 0
            for k in range(min(ncheck, npairs)):
                if weighted:
                    w = oracle.align_weighted(A[k], wa[a_off[k]: a_off[k] + a_len[k]], B[k], wb[b_off[k]: b_off[k] + b_len[k]], R, 4.0)
                else:
                    w = oracle.align(A[k], B[k], R)
                for f in ("ret", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit", "fail_row", "cells"):
                    assert int(recs[f][k]) == w[f], (alen, band, k, f)
                if w["ret"] >= 0:
                    assert (keep_ops[k] == w["ops"]).all()
                checked += 1
            # integer-ALU warp instructions of the kernels that ran (counted in their SASS, row loops and -- for the
            # one-alignment-per-thread kernels -- traceback steps) over K3 time, against the measured peak of the pipe
            alu = None
            if not weighted:
                nw = (2 * band + 1 + 31) // 32
                if nw <= 9:    # align_pairs_thread_kernel<3|5|9>: 32 alignments per warp; 63 / 90 / 175 per row, 18 per traceback step
                    per_row = {3: 63, 5: 90, 9: 175}[3 if nw <= 3 else (5 if nw <= 5 else 9)]
                    alu = ((npairs + 31) // 32) * float(a_len.mean()) * (per_row + 18)
                else:          # strip pass, one warp per alignment: 11 per band word and lane + 12 per row (band 256 and 512: S = 1)
                    alu = npairs * float(a_len.mean()) * (11 * 1 + 12)
            row = {"len": alen, "band": band, "pairs": npairs, "aligned": ok, "cells": cells, "k3_ms": best,
                   "int_pipe_frac": None if alu is None else round(alu / (best / 1e3) / int_peak, 4),
                   "gcups": cells / (best / 1e3) / 1e9, "mean_cost": float(recs["cost"][recs["ret"] >= 0].mean()) if ok else None,
                   "checked_vs_oracle": checked, "gen_s": round(time.time() - t0, 2)}
            rows.append(row)
            print(json.dumps(row), flush=True)
    print("\nlen \\ band " + "".join(f"{b:>9d}" for b in (32, 64, 128, 256, 512)) + "    (GCUPS, K3 kernel time, "
          + ("quality-weighted extension" if weighted else "unit costs") + ")")
    for alen in (1000, 2000, 5000, 10000, 19999):
        print(f"{alen:>10d} " + "".join(f"{r['gcups']:9.0f}" for r in rows if r["len"] == alen))
    if not weighted:
        print(f"\nlen \\ band " + "".join(f"{b:>9d}" for b in (32, 64, 128, 256, 512)) + f"    (integer-pipe fraction: ALU warp instructions "
              f"of the row loops / K3 time / measured peak {int_peak / 1e9:.0f} G warp-instr/s)")
        for alen in (1000, 2000, 5000, 10000, 19999):
            print(f"{alen:>10d} " + "".join(f"{r['int_pipe_frac']:9.3f}" for r in rows if r["len"] == alen))
    ctx.close()


if __name__ == "__main__":
    main()
