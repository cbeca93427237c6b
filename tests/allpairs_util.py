"""All-vs-all expected results (test infrastructure): the reference's trial loop (spaced_seed.cpp:424-436) run once per
target read T with T as the locked reference -- through the CPU oracle, or through the compiled reference itself."""
from __future__ import annotations

import numpy as np

import workload

PAIR_FIELDS = ("j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit")


def allpairs_workload(seed=301, genome_len=9000, nreads=48, mean=1300.0, lo=520, hi=3500, err=(0.04, 0.02, 0.01)):
    """reads drawn from a small genome at ~7x coverage, as a .bin image (binary_test.cpp:55-63 layout)"""
    import cpu_libs
    o = cpu_libs.oracle()
    g = workload.reference(seed, genome_len)
    lens = workload.read_lengths(seed + 1, nreads, mean=mean, sigma_log=0.4, lo=lo, hi=hi)
    txt, offs, lens, _ = workload.reads(seed + 2, g, lens, *err, nthreads=1)
    texts = [txt[offs[k]: offs[k] + lens[k]].tobytes() for k in range(nreads)]
    image = b"".join(o.text2bin(t) for t in texts)
    return texts, image


def kept_texts(texts, min_excl=500, max_excl=20000):
    return [t for t in texts if min_excl < len(t) < max_excl]


def expected_pairs(overlap_fn, texts, image):
    """overlap_fn(T_text: np.ndarray, image) -> OVERLAP_DTYPE records of every kept read against reference T.
    Returns {(T, Q): record} for the pairs with at least one try_align call, and the totals."""
    kt = kept_texts(texts)
    pairs = {}
    tot_ncand = tot_cells = 0
    for T, t in enumerate(kt):
        recs = overlap_fn(np.frombuffer(t, dtype=np.uint8), image)
        assert len(recs) == len(kt)
        for Q in range(len(kt)):
            if Q == T or recs["ncand"][Q] == 0:
                continue
            pairs[(T, Q)] = recs[Q].copy()
            tot_ncand += int(recs["ncand"][Q])
            tot_cells += int(recs["cells"][Q])
    return pairs, tot_ncand, tot_cells


def oracle_overlap_fn(oracle, mask, R=0.3, quirk=False, **kw):
    def fn(t, image):
        ix = oracle.index_build(t, mask, policy=1)
        try:
            return oracle.overlap(ix, t, image, mask, R=R, quirk=quirk, nthreads=4, **kw)
        finally:
            oracle.index_free(ix)
    return fn
