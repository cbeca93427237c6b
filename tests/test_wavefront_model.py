"""The CPU model of the register-resident weighted wavefront (tools/wavefront_model.py = alignw_reg_kernel<CPL> lane by lane)
against the extended oracle: every field and the transcript, over band classes, both goal sides, failures and short inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from wavefront_model import align_weighted_model  # noqa: E402

FIELDS = ("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit", "fail_row")


def test_wavefront_model_matches_extended_oracle(oracle):
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    naligned = nfailed = 0
    for case in range(260):
        n = int(rng.integers(1, 90))
        a = acgt[rng.integers(0, 4, n)]
        rate = float(rng.choice((0.0, 0.05, 0.15, 0.4)))
        b = []
        for ch in a.tolist():
            u = rng.random()
            if u < rate * 0.4:
                b += [int(acgt[rng.integers(0, 4)]), ch]
            elif u < rate * 0.7:
                continue
            elif u < rate:
                b.append(int(acgt[rng.integers(0, 4)]))
            else:
                b.append(ch)
        b += acgt[rng.integers(0, 4, int(rng.integers(0, 12)))].tolist() if case % 3 else []
        b = np.array(b or [65], dtype=np.uint8)
        if case % 2:
            a, b = b, a
        wa = rng.integers(1, 5, size=len(a)).astype(np.uint8)
        wb = rng.integers(1, 5, size=len(b)).astype(np.uint8)
        if case % 5 == 0:
            wa[:] = 1
            wb[:] = 1
        R, fs = ((0.3, 4.0), (0.15, 2.5), (0.3, 1.0), (0.45, 4.0))[case % 4]
        want = oracle.align_weighted(a.tobytes(), wa, b.tobytes(), wb, R, fs)
        cells = want["max_dst"] + 1
        for CPL, lanes in ((1, 32), (2, 32), (3, 32), (5, 8)):  # the last: an 8-lane group
            if cells > CPL * lanes:
                continue
            got = align_weighted_model(a.tolist(), wa.tolist(), b.tolist(), wb.tolist(), R, fs, CPL, lanes)
            for f in FIELDS:
                assert got[f] == want[f], (case, CPL, lanes, f, got[f], want[f])
            if want["ret"] >= 0:
                assert got["ops"] == want["ops"].tolist(), (case, CPL, lanes)
        naligned += want["ret"] >= 0
        nfailed += want["fail_row"] > 0
    assert naligned > 80 and nfailed > 20, (naligned, nfailed)
