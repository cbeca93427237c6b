"""Pins the CPU oracle (oracle/pb_oracle.c) -- CPU only.

Three layers of evidence, strongest first:
1. the assertions of the reference's own unit tests (test/dna_test.cpp, test/aligner_test.cpp,
   test/ref_test.cpp:119-128), restated as known answers;
2. tests/golden/ref_vectors.json -- outputs of the unmodified reference (tests/golden/make_golden.py);
3. live differential runs against oracle/_ref/libpbref.so when it is present.
"""
import hashlib

import numpy as np
import pytest

import workload
from cpu_libs import DELETE, INSERT, MATCH


def ops_str(ops):
    return "".join(chr(48 + int(x)) for x in ops)


# ---------------------------------------------------------------------------
# 1. the reference's own tests
# ---------------------------------------------------------------------------

def test_dna_test_binary(oracle):
    """test/dna_test.cpp:20-30"""
    s = b"ACGTGTCATCGGATCAACCGGTT"
    rec = oracle.text2bin(s)
    assert len(rec) == 10
    assert oracle.bin2text(rec) == s
    assert oracle.seed_at(rec, 0) == 0x34DAB41B
    assert oracle.seed_at(rec, 1) == 0xD068D36E
    assert oracle.seed_at(rec, 2) == 0x41A34DBB
    assert oracle.seed_at(rec, 7) == 0xAF058D36


def edit_tester(ref_elems: bytes, d):
    """aligner_test.cpp:29-41: every MATCH/INSERT val equals the next element of seg_b."""
    j = 0
    for op, val in zip(d["ops"], d["vals"]):
        if op in (MATCH, INSERT):
            assert ref_elems[j] == val
            j += 1


def test_aligner_test_forward(oracle):
    """aligner_test.cpp:44-64"""
    dna_ref, seg1, seg3 = b"ACGTAACCGGTT", b"CGTAAGC", b"TCGTAAC"
    d = oracle.align(seg1[:6], dna_ref[:7])
    assert 6 <= d["ret"] <= 7 and d["cost"] == 2
    edit_tester(dna_ref[:7], d)
    d = oracle.align(seg1[:7], dna_ref[:8])
    assert d["ret"] == 7 and d["cost"] == 2
    edit_tester(dna_ref[:8], d)
    d = oracle.align(seg3[:7], dna_ref[:8])
    assert d["ret"] == 7 and d["cost"] == 1
    edit_tester(dna_ref[:8], d)


def test_aligner_test_backward(oracle):
    """aligner_test.cpp:66-72: accessors (dna_ref+7, false, 7) and (dna_seg1+6, false, 7)"""
    dna_ref, seg1 = b"ACGTAACCGGTT", b"CGTAAGC"
    d = oracle.align(seg1[:7], dna_ref[1:8], a_fwd=False, b_fwd=False)
    assert d["ret"] == 7 and d["cost"] == 1
    edit_tester(dna_ref[1:8][::-1], d)


def test_aligner_test_overlay(oracle):
    """aligner_test.cpp:74-80"""
    dna_ref, seg2 = b"ACGTAACCGGTT", b"GTAACGGGTTAA"
    d = oracle.align(seg2, dna_ref[2:12])
    assert d["ret"] == 10 and d["cost"] == 1
    edit_tester(dna_ref[2:12], d)


def test_aligner_test_remove(oracle):
    """aligner_test.cpp:82-98"""
    dna_ref = b"ACGTAACCGGTT"
    d = oracle.align(dna_ref[1:10], dna_ref[:10])
    assert d["ret"] == 10 and d["nedit"] == 10 and d["ops"][0] == INSERT and d["cost"] == 1
    edit_tester(dna_ref[:10], d)
    d = oracle.align(dna_ref[:10], dna_ref[1:10])
    assert d["ret"] == 9 and d["nedit"] == 10 and d["ops"][0] == DELETE and d["cost"] == 1
    edit_tester(dna_ref[1:10], d)


def test_aligner_test_sample(oracle, golden):
    """aligner_test.cpp:100-117: pair 1 aligns backward, pair 2 fails forward (inputs from the golden file)."""
    real = {(x["pair"], x["order"], x["a_fwd"]): x for x in golden["real_align"]}
    x = real[(1, "seg,ref", False)]
    d = oracle.align(x["a"].encode(), x["b"].encode(), a_fwd=False, b_fwd=False)
    assert d["ret"] > 0
    edit_tester(x["b"].encode()[::-1], d)
    x = real[(2, "seg,ref", True)]
    assert oracle.align(x["a"].encode(), x["b"].encode())["ret"] == -1


def test_ref_test_basic_seedmap(oracle):
    """test/ref_test.cpp:119-128 with that file's fixture string dna_txt (:70; 43 bases, mask 0xFFFFFFFF)."""
    txt = np.frombuffer(b"ACGTAACCGGTTAAACCCGGGTTTTGCAAAAAAAAAAAAAAAA", dtype=np.uint8)
    sz = len(txt)
    ix = oracle.index_build(txt, 0xFFFFFFFF, policy=1)
    nkeys, _ = oracle.index_stats(ix)
    assert nkeys == sz - 15 - 1
    for i in range(sz - 16):
        assert oracle.index_find(ix, oracle.encode(txt[i:i + 16].tobytes()))
    assert not oracle.index_find(ix, oracle.encode(txt[sz - 15:].tobytes()))  # reads the NUL terminator -> code 3
    oracle.index_free(ix)


# ---------------------------------------------------------------------------
# 2. golden vectors produced by the unmodified reference
# ---------------------------------------------------------------------------

def test_golden_masks_and_encode(oracle, golden):
    for m in golden["masks"]:
        assert oracle.parse_pattern(m["pattern"].encode()) == m["mask"], m
    seeds_txt = [m["mask"] for m in golden["masks"][:8]]
    assert seeds_txt == [0xff3c3ffc, 0xff33f3fc, 0xfff0ccfc, 0x3fcfccf3, 0xffccc3f3, 0xffccf3fc, 0x3fcff3fc, 0x3fcfc3fc]
    for e in golden["encode"]:
        assert oracle.encode(e["text"].encode("latin1")) == e["code"]
        assert oracle.decode(e["code"]) == e["decoded"].encode("latin1")


def test_golden_packed(oracle, golden):
    for p in golden["packed"]:
        t = p["text"].encode()
        rec = oracle.text2bin(t)
        assert rec.hex() == p["bin"]
        assert oracle.bin2text(rec).decode() == p["roundtrip"] == p["text"]
        for pos, want in enumerate(p["seed_at"]):
            # the reference's pos%4==0 branch reads byte offset pos (Q-S1): reproduced by quirk=True;
            # canonical value = encode(text+pos) for every pos
            assert oracle.seed_at(rec, pos, quirk=True) == want, (p["text"], pos)
            if pos % 4 or pos == 0:
                assert oracle.seed_at(rec, pos) == want
            assert oracle.seed_at(rec, pos) == oracle.encode(t[pos:pos + 16])


def check_align(oracle, x):
    maxn, maxm = (26000, 6000) if x["which"] == 0 else (40000, 6000)
    d = oracle.align(x["a"].encode("latin1"), x["b"].encode("latin1"), x["R"], x["a_fwd"], x["b_fwd"], maxn, maxm)
    assert d["ret"] == x["ret"], (x["a"], x["b"], x["R"])
    if x["ret"] >= 0:
        for k in ("len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit"):
            assert d[k] == x[k], (k, d[k], x[k])
        assert ops_str(d["ops"]) == x["ops"]
        assert bytes(d["vals"]).decode("latin1") == x["vals"]


def test_golden_real_align(oracle, golden):
    ok = 0
    for x in golden["real_align"]:
        check_align(oracle, x)
        ok += x["ret"] >= 0
    assert ok == 8  # pairs 1(bwd),3,5,6 in both argument orders
    costs = {(x["pair"], x["order"]): x["cost"] for x in golden["real_align"] if x["ret"] >= 0}
    assert costs[(1, "seg,ref")] == 7 and costs[(3, "seg,ref")] == 157
    assert costs[(5, "seg,ref")] == 284 and costs[(6, "seg,ref")] == 346


def test_golden_random_align(oracle, golden):
    for x in golden["random_align"]:
        check_align(oracle, x)


def _digest(find, keys):
    h = hashlib.sha256()
    for k in sorted(keys):
        lst = find(k)
        h.update(np.array([k, len(lst)] + lst, dtype=np.int64).tobytes())
    return h.hexdigest()


def golden_index_ref(g):
    ref = workload.reference(g["seed"], g["length"])
    if g["patched"]:
        ref[100:140] = ord("A")
        ref[2990:] = ord("T")
    return ref


def test_golden_index(oracle, golden):
    for g in golden["index"]:
        ref = golden_index_ref(g)
        ix = oracle.index_build(ref, g["mask"], g["policy"])
        nkeys, _ = oracle.index_stats(ix)
        assert nkeys == g["nkeys"]
        for s in g["sample"]:
            assert oracle.index_find(ix, s["key"]) == s["pos"]
        keys = {oracle.encode(ref[i:i + 16].tobytes()) & g["mask"] for i in range(len(ref))}
        present = [k for k in keys if k and oracle.index_find(ix, k)]
        assert _digest(lambda k: oracle.index_find(ix, k), present) == g["digest"]
        assert not oracle.index_find(ix, 0)
        oracle.index_free(ix)


def golden_locate_inputs(g):
    ref = workload.reference(g["ref_seed"], g["ref_len"])
    lens = workload.read_lengths(g["lens_seed"], g["nreads"], mean=900.0, sigma_log=0.5, lo=300, hi=2500)
    txt, offs, lens, _ = workload.reads(g["reads_seed"], ref, lens, *g["perr"], nthreads=1)
    return ref, txt, offs, lens


def test_golden_locate(oracle, golden):
    for g in golden["locate"]:
        ref, txt, offs, lens = golden_locate_inputs(g)
        ix = oracle.index_build(ref, g["mask"], 0)
        recs, ops = oracle.locate(ix, ref, txt, offs, lens, g["mask"], R=g["R"], want_ops=True, nthreads=2)
        assert len(recs) == len(g["records"])
        for k, row in enumerate(g["records"]):
            for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit"):
                assert int(recs[n][k]) == row[n], (k, n)
            assert hashlib.sha256(ops[k].tobytes()).hexdigest()[:16] == row["ops_sha"]
        oracle.index_free(ix)


# ---------------------------------------------------------------------------
# 3. live differential runs against the compiled reference
# ---------------------------------------------------------------------------

def mutate(rng, a, rate):
    out = []
    for ch in a:
        u = rng.random()
        if u < rate * 0.5:
            out.append(int(rng.integers(0, 4)))
            out.append(ch)
        elif u < rate * 0.8:
            continue
        elif u < rate:
            out.append((ch + 1 + int(rng.integers(0, 3))) & 3)
        else:
            out.append(ch)
    return out


def test_live_align_differential(oracle, ref):
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    nsucc = 0
    for case in range(160):
        n = int(rng.integers(20, 1500))
        a = rng.integers(0, 4, n).tolist()
        b = mutate(rng, a, float(rng.choice([0.0, 0.03, 0.1, 0.2]))) + rng.integers(0, 4, int(rng.integers(0, 400))).tolist()
        a_t, b_t = acgt[a].tobytes(), acgt[np.array(b, dtype=np.int64)].tobytes()
        if case % 2:
            a_t, b_t = b_t, a_t
        R = float(rng.choice([0.1, 0.15, 0.3]))
        fwd = bool(case % 4 != 3)
        d0 = ref.align(a_t, b_t, R, fwd, fwd, which=case % 2)
        d1 = oracle.align(a_t, b_t, R, fwd, fwd, *((26000, 6000) if case % 2 == 0 else (40000, 6000)))
        assert d0["ret"] == d1["ret"]
        if d0["ret"] >= 0:
            nsucc += 1
            for k in ("len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit"):
                assert d0[k] == d1[k], k
            assert (d0["ops"] == d1["ops"]).all() and (d0["vals"] == d1["vals"]).all()
    assert nsucc > 40


def test_live_align_long(oracle, ref):
    """5 kbp CLR-like read against its true locus at R=0.3 (band 3003) and an early failure."""
    g = workload.reference(31, 60000)
    lens = np.array([5200, 5200, 3000], dtype=np.int32)
    txt, offs, lens, starts = workload.reads(32, g, lens, nthreads=1)
    seen = set()
    for k in range(3):
        a = txt[offs[k]: offs[k] + lens[k]].tobytes()
        b = g[starts[k]:].tobytes()
        d0, d1 = ref.align(a, b, 0.3), oracle.align(a, b, 0.3)
        assert d0["ret"] == d1["ret"]
        seen.add(d0["ret"] >= 0)
        if d0["ret"] >= 0:
            for kk in ("matlen_a", "matlen_b", "cost", "diag_cost", "nedit"):
                assert d0[kk] == d1[kk]
            assert (d0["ops"] == d1["ops"]).all()
    b = g[1000:9000].tobytes()
    assert ref.align(txt[:5200].tobytes(), b, 0.3)["ret"] == oracle.align(txt[:5200].tobytes(), b, 0.3)["ret"] == -1


def test_live_index_and_locate(oracle, ref):
    g = workload.reference(41, 50000)
    for policy, mask in ((0, 0xff3c3ffc), (1, 0xffccc3f3)):
        nkeys = ref.index_build(g, mask, policy)
        ix = oracle.index_build(g, mask, policy)
        assert oracle.index_stats(ix)[0] == nkeys
        rng = np.random.default_rng(policy)
        for i in rng.integers(0, len(g) - 16, 300).tolist() + list(range(len(g) - 40, len(g))):
            key = oracle.encode(g[i:i + 16].tobytes()) & mask
            assert ref.index_find(key) == oracle.index_find(ix, key), (policy, i)
        oracle.index_free(ix)
    lens = workload.read_lengths(42, 40, mean=1200.0, sigma_log=0.4, lo=400, hi=3000)
    txt, offs, lens, _ = workload.reads(43, g, lens, 0.05, 0.03, 0.02, nthreads=1)
    ix = oracle.index_build(g, 0xff3c3ffc, 0)
    r0 = ref.locate(g, txt, offs, lens, 0xff3c3ffc, R=0.3, nthreads=2)
    r1 = oracle.locate(ix, g, txt, offs, lens, 0xff3c3ffc, R=0.3, nthreads=2)
    for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand"):
        assert (r0[n] == r1[n]).all(), n
    assert r0["found"].sum() > 10
    oracle.index_free(ix)


def test_live_encode_c2i(oracle, ref):
    rng = np.random.default_rng(9)
    for ch in range(256):
        assert ref.c2i(ch if ch < 128 else ch - 256) == oracle.lib.pbo_c2i(ch if ch < 128 else ch - 256)
    for _ in range(300):
        t = bytes(rng.integers(1, 256, 16).tolist())
        assert ref.encode(t) == oracle.encode(t)


def overlap_workload(seed=81, ref_len=30000, nreads=40):
    """a contig and reads drawn from it, as a .bin image (binary_test.cpp:55-63 layout: records back to back)"""
    import cpu_libs
    o = cpu_libs.oracle()
    ref = workload.reference(seed, ref_len)
    lens = workload.read_lengths(seed + 1, nreads, mean=1500.0, sigma_log=0.4, lo=300, hi=4000)
    txt, offs, lens, _ = workload.reads(seed + 2, ref, lens, 0.04, 0.02, 0.01, nthreads=1)
    image = b"".join(o.text2bin(txt[offs[k]: offs[k] + lens[k]].tobytes()) for k in range(nreads))
    return ref, image


def test_live_overlap_trial_loop(oracle, ref):
    """spaced_seed.cpp:424-436 / try_align :261-299 through the reference's own ref_seq::try_align, seed_at (with its
    pos%4==0 branch, Q-S1) and get_seedmap, against the restatement"""
    g, image = overlap_workload()
    for mask in (0xff3c3ffc, 0x3fcfccf3):
        want = ref.overlap(g, image, mask, R=0.3)
        ix = oracle.index_build(g, mask, policy=1)
        got = oracle.overlap(ix, g, image, mask, R=0.3, quirk=True, nthreads=2)
        oracle.index_free(ix)
        assert len(got) == len(want) > 20
        for n in ("id", "found", "j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit", "ncand"):
            assert (got[n] == want[n]).all(), n
        assert want["found"].sum() > 5
        assert (want["dir"][want["found"] == 1] == -1).any() and (want["dir"][want["found"] == 1] == 1).any()


def test_golden_overlap(oracle, golden):
    for g in golden["overlap"]:
        ref, image = overlap_workload(g["seed"], g["ref_len"], g["nreads"])
        ix = oracle.index_build(ref, g["mask"], policy=1)
        got = oracle.overlap(ix, ref, image, g["mask"], R=g["R"], quirk=True)
        oracle.index_free(ix)
        assert len(got) == len(g["records"])
        for k, row in enumerate(g["records"]):
            for n, v in row.items():
                assert int(got[n][k]) == v, (k, n)


# ---------------------------------------------------------------------------------------------
# all-vs-all (BASELINE config 5): the trial loop once per target read, the target as the locked reference
# ---------------------------------------------------------------------------------------------

def test_golden_allpairs(oracle, golden):
    """the restatement against the compiled reference's output (shipped seed_at behaviour): successful pairs field by
    field, number of pairs with a try_align call, number of calls"""
    from allpairs_util import PAIR_FIELDS, allpairs_workload, expected_pairs, oracle_overlap_fn
    for g in golden["allpairs"]:
        texts, image = allpairs_workload(g["seed"], g["genome_len"], g["nreads"])
        pairs, tot_ncand, _ = expected_pairs(oracle_overlap_fn(oracle, g["mask"], R=g["R"], quirk=True), texts, image)
        assert len(pairs) == g["pairs_with_calls"] and tot_ncand == g["try_align_calls"]
        found = {k: v for k, v in pairs.items() if v["found"]}
        assert len(found) == len(g["found"]) > 100
        for row in g["found"]:
            rec = found[(row["ref_id"], row["read_id"])]
            for n in PAIR_FIELDS:
                assert int(rec[n]) == row[n], (row, n)


def test_live_allpairs(oracle, ref):
    """same, live against the compiled reference, on a different workload and with transposed roles checked:
    overlap(T, Q) and overlap(Q, T) are separate results (T is indexed whole, Q only probes its head / tail trials)"""
    from allpairs_util import PAIR_FIELDS, allpairs_workload, expected_pairs, oracle_overlap_fn
    texts, image = allpairs_workload(331, 6000, 30)
    mask = 0xff3c3ffc
    want, wn, _ = expected_pairs(lambda t, img: ref.overlap(t, img, mask, R=0.3), texts, image)
    got, gn, _ = expected_pairs(oracle_overlap_fn(oracle, mask, quirk=True), texts, image)
    assert sorted(want) == sorted(got) and wn == gn
    for k in want:
        for n in ("found", "ncand") + PAIR_FIELDS:
            assert int(want[k][n]) == int(got[k][n]), (k, n)
    assert any((q, t) not in want for (t, q) in want)


# ---------------------------------------------------------------------------------------------
# consensus voting and the unlocked assembler rounds (ref_seq.h:25-41,47-183,207-276,317-362; spaced_seed.cpp:408-453)
# ---------------------------------------------------------------------------------------------

def assemble_workload(seed=501, genome_len=12000, nreads=80, ref_read=5):
    from allpairs_util import allpairs_workload
    texts, image = allpairs_workload(seed, genome_len, nreads, mean=1500.0, lo=520, hi=4000)
    return np.frombuffer(texts[ref_read], dtype=np.uint8), image


ASM_FIELDS = ("found", "j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit", "ncand")


def test_live_assemble_unlocked(oracle, ref):
    """the reference's own ref_seq (try_align votes and grows, evolve rewrites the text) over four rounds that grow a 1.5 kbp
    read into an ~11 kbp contig, against the restatement: consensus text of every round, who was found when, the records"""
    ref0, image = assemble_workload()
    masks = [0xff3c3ffc, 0x3fcfccf3, 0xff3c3ffc, 0xfff0ccfc]
    for weight in (1, 3):
        cw, fw, rw = ref.assemble(ref0, image, masks, weight=weight)
        cg, fg, rg = oracle.assemble(ref0, image, masks, weight=weight, quirk=True)
        assert cw == cg and (fw == fg).all()
        for n in ASM_FIELDS:
            assert (rw[n] == rg[n]).all(), n
        assert len(cw[-1]) > 4 * len(ref0) and (fw > 0).sum() > 60 and len(set(fw.tolist())) >= 4


def test_golden_assemble(oracle, golden):
    for g in golden["assemble"]:
        ref0, image = assemble_workload(g["seed"], g["genome_len"], g["nreads"], g["ref_read"])
        cons, fr, recs = oracle.assemble(ref0, image, g["masks"], weight=g["weight"], quirk=True)
        assert [hashlib.sha1(c).hexdigest() for c in cons] == g["consensus_sha1"] and [len(c) for c in cons] == g["consensus_len"]
        assert fr.tolist() == g["found_round"]


def test_weighted_extension_is_pinned_by_unit_weights(oracle):
    """pbo_align_weighted (the quality-weighted EXTENSION of config 3) with all weights 1 and fail_scale 1 is pbo_align --
    the function pinned against the compiled reference -- field for field and transcript for transcript; and a weighted
    cost is what its own transcript adds up to"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    nal = 0
    for case in range(120):
        n = int(rng.integers(12, 500))
        a = acgt[rng.integers(0, 4, size=n)]
        keep = rng.random(n) > 0.05
        b = a[keep].copy()
        sub = rng.random(len(b)) < 0.05
        b[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
        b = np.concatenate([b, acgt[rng.integers(0, 4, size=int(rng.integers(0, 200)))]])
        if case % 3 == 0:
            a, b = b, a
        a, b = a.tobytes(), b.tobytes()
        R = float(rng.choice([0.1, 0.3, 0.45]))
        u = oracle.align(a, b, R)
        w = oracle.align_weighted(a, np.ones(len(a), np.uint8), b, np.ones(len(b), np.uint8), R, 1.0)
        for k in ("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit", "fail_row", "cells"):
            assert u[k] == w[k], (case, k)
        if u["ret"] >= 0:
            assert (u["ops"] == w["ops"]).all()
            nal += 1
        wa, wb = rng.integers(1, 5, size=len(a)).astype(np.uint8), rng.integers(1, 5, size=len(b)).astype(np.uint8)
        q = oracle.align_weighted(a, wa, b, wb, R, 4.0)
        if q["ret"] >= 0:
            i = j = cost = 0
            for op in q["ops"].tolist():
                if op == 1:
                    cost += int(wa[i]) if a[i] != b[j] else 0; i += 1; j += 1
                elif op == 2:
                    cost += int(wb[j]); j += 1
                else:
                    cost += int(wa[i]); i += 1
            assert (i, j, cost) == (q["matlen_a"], q["matlen_b"], q["cost"])
    assert nal > 30
