"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the golden vectors
(outputs of the unmodified reference).  Bit-exact: these are integer / byte / index results."""
import hashlib
import os

import numpy as np
import pytest

import workload
from cpu_libs import DELETE, INSERT, MATCH

pytestmark = pytest.mark.gpu

MASKS = [0xff3c3ffc, 0xff33f3fc, 0xfff0ccfc, 0x3fcfccf3, 0xffccc3f3, 0xffccf3fc, 0x3fcff3fc, 0x3fcfc3fc]


@pytest.fixture(scope="module")
def ctx():
    from pacbioassembly_b200 import Context
    c = Context(0)
    yield c
    c.close()


def ops_str(ops):
    return "".join(chr(48 + int(x)) for x in ops)


# ---------------------------------------------------------------------------------------------
# L0
# ---------------------------------------------------------------------------------------------

def test_codec_golden(ctx, golden, oracle):
    for e in golden["encode"]:
        assert int(ctx.encode(e["text"].encode("latin1"))[0]) == e["code"]
    codes = [e["code"] for e in golden["encode"]]
    for e, txt in zip(golden["encode"], ctx.decode(codes)):
        assert txt == e["decoded"].encode("latin1")
    for p in golden["packed"]:
        t = p["text"].encode()
        rec = ctx.text2bin(t)
        assert rec.hex() == p["bin"]
        assert ctx.bin2text(rec) == t
        if p["seed_at"]:
            pos = list(range(len(p["seed_at"])))
            assert ctx.seed_at(rec, pos, quirk=True).tolist() == p["seed_at"]
            assert ctx.seed_at(rec, pos).tolist() == [oracle.encode(t[q:q + 16]) for q in pos]
    # dna_test.cpp:20-30
    rec = ctx.text2bin(b"ACGTGTCATCGGATCAACCGGTT")
    assert len(rec) == 10
    assert ctx.seed_at(rec, [0, 1, 2, 7]).tolist() == [0x34DAB41B, 0xD068D36E, 0x41A34DBB, 0xAF058D36]
    with pytest.raises(Exception):
        ctx.text2bin(b"ACGTACGT", cap=5)  # the reference asserts (dna_seq.h:118)


def test_encode_tail_and_nonacgt(ctx, oracle):
    t = b"ACGTNNacgtXGATTACAGATTACA"
    offs = list(range(len(t)))
    got = ctx.encode(t, offs).tolist()
    assert got == [oracle.encode(t[o:]) for o in offs]


def test_seqset_roundtrip_and_views(ctx, oracle):
    g = workload.reference(5, 1000)
    txt = g.tobytes()
    offs = [0, 10, 999, 500, 64]
    lens = [1000, 33, 1000, 0, 1]
    strides = [1, 1, -1, 1, 1]
    s = ctx.seqset(g, offs, lens, strides)
    assert len(s) == 5
    assert s.text(0) == txt
    assert s.text(1) == txt[10:43]
    assert s.text(2) == txt[::-1]
    assert s.text(3) == b""
    assert s.text(4) == txt[64:65]
    assert s.packed(1) == oracle.text2bin(txt[10:43])[4:]
    image = b"".join(oracle.text2bin(workload.reference(100 + k, n).tobytes()) for k, n in enumerate((600, 20, 501, 20000, 777)))
    sb = ctx.seqset_from_bin(image)  # keeps 500 < len < 20000 (spaced_seed.cpp:336)
    assert len(sb) == 3 and [sb.length(i) for i in range(3)] == [600, 501, 777]
    assert sb.text(2) == workload.reference(104, 777).tobytes()


# ---------------------------------------------------------------------------------------------
# L1
# ---------------------------------------------------------------------------------------------

def test_seed_extract_all_positions(ctx, oracle):
    g = workload.reference(6, 5000)
    g[100:140] = ord("A")
    g[777] = ord("N")
    s = ctx.seqset_one(g)
    txt = g.tobytes()
    for mask in (MASKS[0], MASKS[3], 0xFFFFFFFF):
        got = s.seeds(0, mask)
        want = np.array([oracle.encode(txt[p:p + 16]) & mask for p in range(len(txt))], dtype=np.uint32)
        assert (got == want).all()  # the last 15 positions read code 3 past the end (Q-S3)
    # several sequences in one set: windows never leak into a neighbour
    parts = [workload.reference(50 + k, n) for k, n in enumerate((17, 512, 31, 1000))]
    blob = np.concatenate(parts)
    offs = np.cumsum([0] + [len(p) for p in parts[:-1]])
    ms = ctx.seqset(blob, offs, [len(p) for p in parts])
    for k, p in enumerate(parts):
        t = p.tobytes()
        want = [oracle.encode(t[q:q + 16]) & MASKS[1] for q in range(len(t))]
        assert ms.seeds(k, MASKS[1]).tolist() == want


def _digest(find_batch, keys):
    keys = sorted(keys)
    lists = find_batch(keys)
    h = hashlib.sha256()
    for k, lst in zip(keys, lists):
        h.update(np.array([k, len(lst)] + lst, dtype=np.int64).tobytes())
    return h.hexdigest()


def test_index_golden(ctx, golden, oracle):
    from test_oracle import golden_index_ref
    for g in golden["index"]:
        ref = golden_index_ref(g)
        s = ctx.seqset_one(ref)
        ix = ctx.index(s, g["mask"], g["policy"])
        assert ix.nkeys == g["nkeys"]
        got = ix.find_batch([x["key"] for x in g["sample"]])
        assert got == [x["pos"] for x in g["sample"]]
        txt = ref.tobytes()
        keys = {oracle.encode(txt[i:i + 16]) & g["mask"] for i in range(len(txt))}
        keys.discard(0)
        lists = ix.find_batch(sorted(keys))
        present = [k for k, l in zip(sorted(keys), lists) if l]
        assert _digest(ix.find_batch, present) == g["digest"]
        assert ix.find(0) == []
        ix.free()


def test_index_ref_test_basic(ctx):
    """test/ref_test.cpp:119-128 through the device index."""
    txt = b"ACGTAACCGGTTAAACCCGGGTTTTGCAAAAAAAAAAAAAAAA"
    s = ctx.seqset_one(txt)
    ix = ctx.index(s, 0xFFFFFFFF, policy=1)
    sz = len(txt)
    assert ix.nkeys == sz - 15 - 1
    keys = ctx.encode(txt, list(range(sz - 16)) + [sz - 15]).tolist()
    lists = ix.find_batch(keys)
    assert all(lists[i] for i in range(sz - 16))
    assert lists[-1] == []


def test_index_from_pairs_keeps_insertion_order(ctx):
    """pb_index_build_pairs: seedmap[key].push_back(pos) in insertion order (locator.cpp:62-66), duplicates kept."""
    rng = np.random.default_rng(77)
    keys = rng.integers(1, 5000, 60000).astype(np.uint32) * np.uint32(0x9E3779B1)
    keys[1000:1400] = keys[0]  # one long list
    pos = rng.integers(0, 1 << 30, len(keys)).astype(np.int32)
    ix = ctx.index_from_pairs(keys, pos)
    want: dict[int, list[int]] = {}
    for k, p in zip(keys.tolist(), pos.tolist()):
        want.setdefault(k, []).append(p)
    assert ix.nkeys == len(want) and ix.nentries == len(keys)
    probe = list(want)[:700] + [int(keys[0]), 12345]
    got = ix.find_batch(probe)
    for k, lst in zip(probe, got):
        assert lst == want.get(k, []), hex(k)
    ix.free()


def test_index_vs_oracle_large(ctx, oracle):
    g = workload.reference(41, 300000)
    g[5000:5600] = ord("T")  # a repeat: one bucket far above the small-bucket sort threshold
    g[9000:9100] = np.frombuffer(b"AC" * 50, dtype=np.uint8)
    s = ctx.seqset_one(g)
    for policy, mask in ((0, MASKS[0]), (0, MASKS[2]), (1, MASKS[4]), (0, 0xFFFFFFFF)):
        ix = ctx.index(s, mask, policy)
        oix = oracle.index_build(g, mask, policy)
        nk, ne = oracle.index_stats(oix)
        assert (ix.nkeys, ix.nentries) == (nk, ne)
        rng = np.random.default_rng(policy + mask % 7)
        probe = rng.integers(0, len(g) - 16, 400).tolist() + [5100, 9010, len(g) - 3, len(g) - 16, 0]
        txt = g.tobytes()
        keys = [oracle.encode(txt[i:i + 16]) & mask for i in probe] + [0x12345678 & mask, 0]
        got = ix.find_batch(keys)
        for k, lst in zip(keys, got):
            assert lst == oracle.index_find(oix, k), (policy, hex(mask), hex(k))
        oracle.index_free(oix)
        ix.free()


# ---------------------------------------------------------------------------------------------
# L2
# ---------------------------------------------------------------------------------------------

def check_against(d, want, keys=("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit")):
    assert d["ret"] == want["ret"]
    if want["ret"] >= 0:
        for k in keys:
            assert d[k] == want[k], (k, d[k], want[k])


def test_aligner_test_cases(ctx):
    """test/aligner_test.cpp:44-98 through the CUDA aligner."""
    dna_ref, seg1, seg2, seg3 = b"ACGTAACCGGTT", b"CGTAAGC", b"GTAACGGGTTAA", b"TCGTAAC"

    def edit_tester(ref_elems, d):
        j = 0
        for op, val in zip(d["ops"], d["vals"]):
            if op in (MATCH, INSERT):
                assert ref_elems[j] == val
                j += 1

    d = ctx.align(seg1[:6], dna_ref[:7]); assert 6 <= d["ret"] <= 7 and d["cost"] == 2; edit_tester(dna_ref[:7], d)
    d = ctx.align(seg1[:7], dna_ref[:8]); assert d["ret"] == 7 and d["cost"] == 2; edit_tester(dna_ref[:8], d)
    d = ctx.align(seg3[:7], dna_ref[:8]); assert d["ret"] == 7 and d["cost"] == 1; edit_tester(dna_ref[:8], d)
    d = ctx.align(seg1[:7], dna_ref[1:8], a_fwd=False, b_fwd=False)
    assert d["ret"] == 7 and d["cost"] == 1; edit_tester(dna_ref[1:8][::-1], d)
    d = ctx.align(seg2, dna_ref[2:12]); assert d["ret"] == 10 and d["cost"] == 1; edit_tester(dna_ref[2:12], d)
    d = ctx.align(dna_ref[1:10], dna_ref[:10])
    assert d["ret"] == 10 and d["nedit"] == 10 and d["ops"][0] == INSERT and d["cost"] == 1
    d = ctx.align(dna_ref[:10], dna_ref[1:10])
    assert d["ret"] == 9 and d["nedit"] == 10 and d["ops"][0] == DELETE and d["cost"] == 1


def test_align_golden(ctx, golden):
    n_ok = n_irr = 0
    for x in golden["real_align"] + golden["random_align"]:
        a, b = x["a"].encode("latin1"), x["b"].encode("latin1")
        maxn, maxm = (26000, 6000) if x["which"] == 0 else (40000, 6000)
        n_irr += bool(set(a + b) - set(b"ACGT"))  # raw-byte compare (seq_aligner.h:136): the byte-exact aligner variant
        d = ctx.align(a, b, x["R"], x["a_fwd"], x["b_fwd"], maxn, maxm)
        check_against(d, x)
        if x["ret"] >= 0:
            assert ops_str(d["ops"]) == x["ops"]
            assert bytes(d["vals"]).decode("latin1") == x["vals"]
            n_ok += 1
    assert n_ok > 80 and n_irr >= 10


def make_pairs(rng, n, maxlen, rates=(0.0, 0.03, 0.1, 0.2)):
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    A, B = [], []
    for case in range(n):
        m = int(rng.integers(1, maxlen))
        a = rng.integers(0, 4, m)
        rate = float(rng.choice(rates))
        u = rng.random(m)
        out = []
        for ch, uu in zip(a.tolist(), u.tolist()):
            if uu < rate * 0.5:
                out += [int(rng.integers(0, 4)), ch]
            elif uu < rate * 0.8:
                continue
            elif uu < rate:
                out.append((ch + 1 + int(rng.integers(0, 3))) & 3)
            else:
                out.append(ch)
        extra = int(rng.integers(0, maxlen // 3 + 1)) if case % 3 else 0
        b = np.array(out + rng.integers(0, 4, extra).tolist(), dtype=np.int64)
        if case % 5 == 0 and len(b) > 4:
            b = b[: len(b) - int(rng.integers(0, len(b) // 3))]
        if len(b) == 0:
            b = np.array([0])
        at, bt = acgt[a].tobytes(), acgt[b].tobytes()
        if case % 2:
            at, bt = bt, at
        A.append(at)
        B.append(bt)
    return A, B


def run_batch_vs_oracle(ctx, oracle, A, B, R, maxn=26000, maxm=6000, fwd=True):
    a_blob, b_blob = b"".join(A), b"".join(B)
    a_len, b_len = [len(x) for x in A], [len(x) for x in B]
    a_off = np.cumsum([0] + a_len[:-1])
    b_off = np.cumsum([0] + b_len[:-1])
    if fwd:
        recs, ops = ctx.align_batch(a_blob, a_off, a_len, b_blob, b_off, b_len, R, maxn, maxm)
    else:
        recs, ops = ctx.align_batch(a_blob, a_off + np.array(a_len) - 1, a_len, b_blob, b_off + np.array(b_len) - 1, b_len, R,
                                    maxn, maxm, a_stride=[-1] * len(A), b_stride=[-1] * len(B))
    nsucc = 0
    for i, (a, b) in enumerate(zip(A, B)):
        w = oracle.align(a, b, R, fwd, fwd, maxn, maxm)
        for k in ("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit", "fail_row", "cells"):
            assert int(recs[k][i]) == w[k], (i, k, int(recs[k][i]), w[k], len(a), len(b))
        if w["ret"] >= 0:
            assert (ops[i] == w["ops"]).all(), i
            nsucc += 1
    return nsucc


def test_align_random_small(ctx, oracle):
    rng = np.random.default_rng(11)
    A, B = make_pairs(rng, 400, 300)
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.3) > 100
    assert run_batch_vs_oracle(ctx, oracle, A[:150], B[:150], 0.15, 40000, 6000) > 20
    assert run_batch_vs_oracle(ctx, oracle, A[:150], B[:150], 0.3, fwd=False) > 20
    assert run_batch_vs_oracle(ctx, oracle, A[:100], B[:100], 0.05) >= 0


def test_align_random_multiword(ctx, oracle):
    """bands of 1k..4k bits: several words per lane, carries crossing lanes"""
    rng = np.random.default_rng(12)
    A, B = make_pairs(rng, 60, 6000, rates=(0.0, 0.05, 0.12))
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.3) > 15


def test_align_non_acgt_bytes(ctx, oracle):
    """N, lower case and other bytes compare by raw value in the DP (seq_aligner.h:136) while seeding maps them to
    code 3 (dna_seq.h:21): pairs that hold them take the byte-exact aligner variant"""
    from pacbioassembly_b200 import PbError
    rng = np.random.default_rng(21)
    A, B = make_pairs(rng, 120, 900)
    A2, B2 = [], []
    for k, (a, b) in enumerate(zip(A, B)):
        a, b = bytearray(a), bytearray(b)
        for seq in (a, b):
            for _ in range(int(rng.integers(0, 6))):
                seq[int(rng.integers(0, len(seq)))] = int(rng.choice(list(b"NNNnx-")))  # four distinct values at most
        if k % 7 == 0:  # the same unusual byte on both sides at aligned places must MATCH
            m = min(len(a), len(b))
            for q in range(0, m, 37):
                a[q] = b[q] = ord("N")
        A2.append(bytes(a)); B2.append(bytes(b))
    for R in (0.3, 0.15):
        assert run_batch_vs_oracle(ctx, oracle, A2, B2, R) > 20
    assert run_batch_vs_oracle(ctx, oracle, A2[:40], B2[:40], 0.3, fwd=False) > 5
    # long pair with N runs (multi-word band, class S=6/16 of the byte-exact variant)
    g = bytearray(workload.reference(91, 9000).tobytes())
    a = bytearray(g[:6000]); a[1000:1010] = b"N" * 10; a[3000] = ord("n")
    b = bytearray(g[:8000]); b[1003:1007] = b"NNNN"; b[5000:5003] = b"nnn"
    assert run_batch_vs_oracle(ctx, oracle, [bytes(a), bytes(b[:5000])], [bytes(b), bytes(a)], 0.3) >= 1
    # more than four distinct unusual byte values in seg_a: outside the bit-parallel byte-exact variant (four extra Eq planes);
    # pb_align_batch hands such pairs to the wavefront aligner, which compares raw bytes -- forward and backward views, mixed
    # with ordinary pairs in one batch, transcripts of the neighbours untouched
    wide_a = [b"ACGTNXYZWACGT" * 9, bytes(A2[3]), b"acgtnACGTNRYKMSW" * 20 + b"ACGT" * 30, bytes(A2[5])]
    wide_b = [b"ACGTACGTACGTA" * 9 + b"GG", bytes(B2[3]), b"acgtnACGTNRYKMSW" * 20 + b"ACGT" * 40, bytes(B2[5])]
    assert run_batch_vs_oracle(ctx, oracle, wide_a, wide_b, 0.3) >= 2
    assert run_batch_vs_oracle(ctx, oracle, wide_a, wide_b, 0.45, fwd=False) >= 2
    # the read-set pipelines keep the four-value limit and say so on every entry point
    g = workload.reference(92, 30000)
    weird = bytearray(g[1000:1700].tobytes()); weird[100:105] = b"NXYZW"
    rs = ctx.seqset_one(g)
    ix = ctx.index(rs, MASKS[0])
    with pytest.raises(PbError):
        ctx.locate(ix, np.frombuffer(bytes(weird), np.uint8), [0], [len(weird)], R=0.3)
    with pytest.raises(PbError):
        ctx.locate_submit(ix, np.frombuffer(bytes(weird), np.uint8), [0], [len(weird)], R=0.3)
    ix2 = ctx.index(rs, MASKS[0], policy=1)
    with pytest.raises(PbError):
        ctx.overlap(ix2, ctx.seqset(np.frombuffer(bytes(weird), np.uint8), [0], [len(weird)]), R=0.3)


def test_align_domain_limits(ctx, oracle):
    """seq_aligner.h:104-107 with the Q-D3 domain: len_a >= maxn or max_dst >= maxm -> -1"""
    g = workload.reference(77, 3000).tobytes()
    A, B = [g[:2000], g[:1500], g[:100]], [g[:2100], g[:1500], g[:100]]
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.3, maxn=1800, maxm=6000) == 2
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.3, maxn=26000, maxm=500) == 2
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.3, maxn=50, maxm=10) == 0


def test_align_clr_reads_all_band_classes(ctx, oracle):
    """CLR-like reads against their true locus, 0.5-20 kbp, R=0.3: every band class up to 12 words per lane"""
    g = workload.reference(31, 400000)
    lens = np.array([600, 1500, 2500, 3400, 4800, 5200, 6500, 8000, 9900, 12000, 15000, 19999], dtype=np.int32)
    txt, offs, lens, starts = workload.reads(32, g, lens, 0.05, 0.03, 0.02, nthreads=1)
    A = [txt[offs[k]: offs[k] + lens[k]].tobytes() for k in range(len(lens))]
    B = [g[starts[k]:].tobytes()[: int(lens[k] * 1.4) + 50] for k in range(len(lens))]
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.3) >= 9
    assert run_batch_vs_oracle(ctx, oracle, B[:6], A[:6], 0.3) >= 4  # seg_a longer than seg_b: last-column goal


# ---------------------------------------------------------------------------------------------
# locate
# ---------------------------------------------------------------------------------------------

def test_locate_golden(ctx, golden):
    from test_oracle import golden_locate_inputs
    for g in golden["locate"]:
        ref, txt, offs, lens = golden_locate_inputs(g)
        rs = ctx.seqset_one(ref)
        ix = ctx.index(rs, g["mask"])
        recs, ops = ctx.locate(ix, txt, offs, lens, want_ops=True, R=g["R"])
        assert len(recs) == len(g["records"])
        for k, row in enumerate(g["records"]):
            for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit"):
                assert int(recs[n][k]) == row[n], (k, n)
            assert hashlib.sha256(ops[k].tobytes()).hexdigest()[:16] == row["ops_sha"]


def locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, mask, R, nthreads=8, **kw):
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, mask)
    recs, ops = ctx.locate(ix, txt, offs, lens, want_ops=True, R=R, **kw)
    oix = oracle.index_build(ref, mask, 0)
    want, wops = oracle.locate(oix, ref, txt, offs, lens, mask, R=R, nthreads=nthreads, want_ops=True, **kw)
    oracle.index_free(oix)
    assert len(recs) == len(want)
    for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
        bad = np.nonzero(recs[n] != want[n])[0]
        assert len(bad) == 0, (n, bad[:5], recs[n][bad[:5]], want[n][bad[:5]])
    for k in range(len(recs)):
        assert (ops[k] == wops[k]).all(), k
    return recs


def test_locate_config1(ctx, oracle):
    """BASELINE config 1 (scaled to CPU-oracle seconds): 1 Mbp reference, reads 500-3000 at 5 % error, two masks, both R"""
    ref = workload.reference(1, 1_000_000)
    lens = workload.read_lengths(7, 240, mean=1500.0, sigma_log=0.5, lo=300, hi=3000)
    txt, offs, lens, _ = workload.reads(8, ref, lens, 0.025, 0.015, 0.01)
    found = 0
    for mask, R in ((MASKS[0], 0.15), (MASKS[3], 0.3)):
        recs = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, mask, R)
        found += int(recs["found"].sum())
    assert found > 200


def test_locate_clr_and_edges(ctx, oracle):
    """CLR error model (ins 9 / del 4 / sub 2 %), R=0.3; plus empty batch, all-short batch, reads at the contig end"""
    ref = workload.reference(2, 400_000)
    lens = workload.read_lengths(3, 96, mean=3000.0, sigma_log=0.5, lo=500, hi=9000)
    txt, offs, lens, _ = workload.reads(3, ref, lens)
    recs = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, MASKS[0], 0.3)
    assert 10 < recs["found"].sum() < len(recs)
    # reads cut from the very end of the contig: seg_a longer than what is left of the reference
    tail = ref[-2600:]
    t2 = np.concatenate([tail[:2000], tail[300:2600], tail[1000:2600], ref[:400]])
    o2 = np.array([0, 2000, 4300, 5900])
    l2 = np.array([2000, 2300, 1600, 400], dtype=np.int32)
    recs = locate_vs_oracle(ctx, oracle, ref, t2, o2, l2, MASKS[0], 0.15)
    assert len(recs) == 3 and recs["found"].sum() >= 2
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASKS[0])
    assert len(ctx.locate(ix, np.zeros(0, np.uint8), [], [])) == 0
    assert len(ctx.locate(ix, ref[:900], [0, 400], [400, 499])) == 0
    # custom trial count / minimum length
    recs = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, MASKS[5], 0.3, ntrial=8, minlen=1000)


def test_locate_with_n_in_reads_and_contig(ctx, oracle):
    """real contigs carry N runs and locator.cpp's N->A loop never runs (Q-S4): seeds see code 3, the DP sees raw bytes"""
    ref = workload.reference(33, 120_000)
    ref[40_000:40_050] = ord("N")
    ref[77_777] = ord("n")
    lens = workload.read_lengths(34, 60, mean=1200.0, sigma_log=0.4, lo=500, hi=3000)
    txt, offs, lens, starts = workload.reads(35, ref, lens, 0.02, 0.01, 0.01)
    txt = txt.copy()
    for k in range(0, 60, 3):
        txt[offs[k] + 100 + k] = ord("N")
    recs = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, MASKS[0], 0.15)
    assert recs["found"].sum() > 30
    clean = workload.reference(33, 120_000)
    recs = locate_vs_oracle(ctx, oracle, clean, txt, offs, lens, MASKS[3], 0.3)  # N only in the reads


def test_locate_properties_at_scale(ctx):
    """Size-independent properties on a batch too large for the CPU oracle: transcripts replay to the right lengths and
    costs, error-free reads map to their true position with cost 0, and the run is deterministic."""
    ref = workload.reference(2, 4_600_000)
    lens = workload.read_lengths(3, 3000, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
    txt, offs, lens, starts = workload.reads(3, ref, lens)
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASKS[0])
    recs, ops = ctx.locate(ix, txt, offs, lens, want_ops=True, R=0.3)
    recs2 = ctx.locate(ix, txt, offs, lens, R=0.3)
    for n in recs.dtype.names:
        assert (recs[n] == recs2[n]).all()
    f = recs["found"] == 1
    assert 0.3 < f.mean() < 0.9
    g = ref.tobytes()
    for k in np.nonzero(f)[0][:400]:
        o = ops[k]
        n_m, n_i, n_d = int((o == MATCH).sum()), int((o == INSERT).sum()), int((o == DELETE).sum())
        assert n_m + n_d == recs["matlen_a"][k] and n_m + n_i == recs["matlen_b"][k]
        a = txt[offs[k] + recs["j"][k]: offs[k] + lens[k]]
        b = ref[recs["pos"][k]: recs["pos"][k] + recs["matlen_b"][k]]
        i = j = cost = 0
        for op in o.tolist():
            if op == MATCH:
                cost += int(a[i] != b[j]); i += 1; j += 1
            elif op == INSERT:
                cost += 1; j += 1
            else:
                cost += 1; i += 1
        assert cost == recs["cost"][k]
        assert abs(int(recs["pos"][k]) - int(recs["j"][k]) - int(starts[k])) < 0.35 * lens[k]
    # diagonal-bin tally (K2 voting, diagnostic): located reads' fullest bin sits on their true diagonal
    s_all = ctx.seqset(txt, offs, lens)
    job = ctx.locate_run(ix, s_all, R=0.3)
    r3 = job.fetch()
    votes, bdiag = job.votes()
    for n in recs.dtype.names:
        assert (r3[n] == recs[n]).all()
    ok = np.nonzero(f & (votes >= 3))[0]
    assert len(ok) > 500
    assert (np.abs(bdiag[ok].astype(np.int64) - starts[ok]) < 0.1 * lens[ok] + 512).mean() > 0.95
    job.free(); s_all.free()
    # error-free reads
    l0 = np.full(64, 2000, dtype=np.int32)
    t0, o0, l0, s0 = workload.reads(9, ref, l0, 0.0, 0.0, 0.0)
    r0 = ctx.locate(ix, t0, o0, l0, R=0.15)
    assert (r0["found"] == 1).all() and (r0["cost"] == 0).all() and (r0["j"] == 0).all()
    assert (r0["pos"] == s0).all()


def test_locate_config1_full(ctx, oracle):
    """BASELINE config 1 at its stated size: 1 Mbp reference (seed 1), 2 000 reads of 500-3000 bases at 5 % error, every
    mask of seeds.txt, locator semantics at R = 0.15 and R = 0.3 (locator.cpp:51-54, 62-66, 70-92) -- all against the oracle"""
    ref = workload.reference(1, 1_000_000)
    rng = np.random.default_rng(11)
    lens = rng.integers(500, 3001, size=2000).astype(np.int32)
    txt, offs, lens, _ = workload.reads(8, ref, lens, 0.025, 0.015, 0.01)
    nthreads = min(os.cpu_count() or 1, 32)
    for mask in MASKS:
        for R in (0.15, 0.3):
            recs = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, mask, R, nthreads=nthreads)
            assert recs["found"].sum() > 1500, (hex(mask), R)


def test_locate_strip_redo_paths(ctx, oracle):
    """K3's first pass runs a certified strip of the band and hands what it cannot certify to the full band.  Reads built to
    land on both sides of the certificate: a block deletion / insertion late in the read drives the final cost towards R*len and
    the path far off the diagonal; plus reads at the contig's end (goal on the last column)."""
    ref = workload.reference(2, 300_000)
    rng = np.random.default_rng(5)
    L = 4000
    lens = np.full(120, L, dtype=np.int32)
    base, offs, lens, starts = workload.reads(21, ref, lens, 0.01, 0.01, 0.01)
    parts = []
    for k in range(len(lens)):
        r = base[offs[k]: offs[k] + lens[k]]
        frac = 0.10 + 0.0025 * k  # block size as a share of the read: 10 % .. 40 %
        blk = int(frac * L)
        at = int(0.93 * L) - blk // 2
        if k % 2 == 0:  # deletion from the read: the path jumps right of the diagonal
            r = np.concatenate([r[:at], r[at + blk:]])
        else:           # insertion of random bases into the read: the path jumps left
            r = np.concatenate([r[:at], np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=blk)], r[at:]])
        parts.append(r)
    # reads reaching past the contig's end
    tail = ref[-3000:]
    parts += [tail[:2500].copy(), tail[200:].copy(), np.concatenate([tail[500:], np.frombuffer(b"ACGT" * 100, np.uint8)])]
    lens2 = np.array([len(x) for x in parts], dtype=np.int32)
    offs2 = np.zeros(len(parts), dtype=np.int64)
    np.cumsum(lens2[:-1], out=offs2[1:])
    txt2 = np.concatenate(parts)
    recs = locate_vs_oracle(ctx, oracle, ref, txt2, offs2, lens2, MASKS[0], 0.3)
    assert recs["found"].sum() > 20
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASKS[0])
    s = ctx.seqset(txt2, offs2, lens2)
    job = ctx.locate_run(ix, s, R=0.3)
    r2 = job.fetch()
    st = job.stats()
    for n in recs.dtype.names:
        assert (r2[n] == recs[n]).all()
    if os.environ.get("PB_NARROW", "1") != "0":
        assert st["redone"] > 0 and st["band_cells"] > 0  # both passes ran
    job.free(); s.free()


def test_locate_traceback_rounds_all_classes(ctx, oracle):
    """The strip pass stores no parents: its traceback recomputes (block, lane) tiles from checkpoints, 32 per round, around the
    path's predicted diagonal.  Reads of every strip class (600 .. 19 000 bases), insertion-heavy, deletion-heavy and balanced --
    the path drifts one way, the other, or wanders -- with transcripts compared op for op against the oracle."""
    ref = workload.reference(2, 600_000)
    parts_t, parts_o, parts_l = [], [], []
    pos = 0
    for seed, (pi, pd, ps) in enumerate(((0.14, 0.02, 0.01), (0.02, 0.14, 0.01), (0.07, 0.07, 0.02), (0.01, 0.01, 0.01))):
        lens = np.array([600, 1500, 2500, 2600, 3000, 5200, 5300, 6000, 9000, 12000, 15000, 19000], dtype=np.int32)
        t, o, l, _ = workload.reads(300 + seed, ref, lens, pi, pd, ps)
        parts_t.append(t); parts_o.append(o + pos); parts_l.append(l)
        pos += len(t)
    txt, offs, lens = np.concatenate(parts_t), np.concatenate(parts_o), np.concatenate(parts_l)
    recs = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, MASKS[0], 0.3, nthreads=16)
    assert recs["found"].sum() >= 24
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASKS[0])
    s = ctx.seqset(txt, offs, lens)
    job = ctx.locate_run(ix, s, R=0.3)
    job.fetch()
    st = job.stats()
    if os.environ.get("PB_NARROW", "1") != "0":
        assert st["tb_rounds"] >= recs["found"].sum() and st["tb_cold"] >= st["tb_rounds"]  # the checkpointed traceback ran
    job.free(); s.free()


def test_traceback_prefetch_stress(ctx, oracle):
    """The traceback's asynchronous window ring (cp.async into shared memory) once faulted on zero-fill copies with a dummy
    source -- only where prefetched windows reach above row 1, i.e. on SHORT alignments.  Short and long
    alignments, every band class, both aligner passes, 40 repetitions: no CUDA error and the same records every time."""
    ref = workload.reference(2, 200_000)
    lens = np.concatenate([np.full(96, 500), np.full(64, 523), np.arange(500, 2500, 25), np.array([4000, 7000, 12000, 19000])]).astype(np.int32)
    txt, offs, lens, _ = workload.reads(77, ref, lens, 0.04, 0.02, 0.02)
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASKS[0])
    first = locate_vs_oracle(ctx, oracle, ref, txt, offs, lens, MASKS[0], 0.3)
    for rep in range(40):
        recs, ops = ctx.locate(ix, txt, offs, lens, want_ops=True, R=0.3)
        for n in first.dtype.names:
            assert (recs[n] == first[n]).all(), (rep, n)
    from allpairs_util import allpairs_workload
    texts, image = allpairs_workload(305, 7000, 44)
    reads = ctx.seqset_from_bin(image)
    aix = ctx.index_set(reads, MASKS[0])
    r0, s0 = ctx.overlap_all(aix, found_only=False, R=0.3)
    for rep in range(40):
        r1, s1 = ctx.overlap_all(aix, found_only=False, R=0.3)
        assert s1 == s0 and all((r1[n] == r0[n]).all() for n in r0.dtype.names), rep


def test_locate_pipelined_submit_collect(ctx, oracle):
    """pb_locate_submit / pb_locate_collect (two batches in flight, host text and .bin image) return what pb_locate_batch does"""
    ref = workload.reference(2, 400_000)
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, MASKS[0])
    batches = []
    for b in range(3):
        lens = workload.read_lengths(40 + b, 80 + 16 * b, mean=2500.0, sigma_log=0.5, lo=300, hi=8000)
        batches.append(workload.reads(50 + b, ref, lens))
    want = [ctx.locate(ix, t, o, l, R=0.3) for t, o, l, _ in batches]
    steps, got = [], []
    for b, (t, o, l, _) in enumerate(batches):  # submit k+1 before collecting k
        if b % 2 == 0:
            steps.append(ctx.locate_submit(ix, t, o, l, R=0.3))
        else:
            steps.append(ctx.locate_submit_bin(ix, workload.pack_bin(t, o, l), R=0.3))
        if b > 0:
            got.append(steps[b - 1].collect())
    got.append(steps[-1].collect())
    for w, g, st in zip(want, got, steps):
        assert len(w) == len(g) > 0
        for n in w.dtype.names:
            assert (w[n] == g[n]).all(), n
        assert st.stats["dp_alignments"] > 0
    # the same batches with their text already in device memory (pb_locate_submit_device), two in flight
    import torch
    dev = [torch.from_numpy(np.ascontiguousarray(t)).cuda() for t, _, _, _ in batches]
    dsteps = [ctx.locate_submit_device(ix, d.data_ptr(), d.numel(), o, l, R=0.3) for d, (_, o, l, _) in zip(dev[:2], batches[:2])]
    dgot = [dsteps[0].collect()]
    dsteps.append(ctx.locate_submit_device(ix, dev[2].data_ptr(), dev[2].numel(), batches[2][1], batches[2][2], R=0.3))
    dgot += [dsteps[1].collect(), dsteps[2].collect()]
    for w, g in zip(want, dgot):
        assert len(w) == len(g) and all((w[n] == g[n]).all() for n in w.dtype.names)
    with pytest.raises(Exception):  # a view past the end of the blob
        ctx.locate_submit_device(ix, dev[0].data_ptr(), 100, batches[0][1], batches[0][2], R=0.3)
    # an empty batch and a step that is dropped without being collected
    assert len(ctx.locate_submit(ix, np.zeros(0, np.uint8), [], [], R=0.3).collect()) == 0
    ctx.locate_submit(ix, *batches[0][:3], R=0.3).free()
    assert (ctx.locate(ix, *batches[0][:3], R=0.3)["pos"] == want[0]["pos"]).all()


def test_locate_config4_reference_size(ctx, oracle):
    """BASELINE config 4 at its reference size, fewer reads: 64 Mbp reference (seed 4), weight-11 mask (~15 random candidates
    per probe, hundreds of failing alignments per read before the true locus), CLR reads (seed 5), R = 0.3 -- index and
    every record field incl. candidate and cell counts against the oracle"""
    ref = workload.reference(4, 64_000_000)
    lens = workload.read_lengths(5, 320, mean=5000.0, sigma_log=0.5, lo=500, hi=19999)
    txt, offs, lens, starts = workload.reads(5, ref, lens)
    mask = MASKS[2]
    rs = ctx.seqset_one(ref)
    ix = ctx.index(rs, mask)
    oix = oracle.index_build(ref, mask, 0)
    assert oracle.index_stats(oix) == (ix.nkeys, ix.nentries)
    recs = ctx.locate(ix, txt, offs, lens, R=0.3)
    want = oracle.locate(oix, ref, txt, offs, lens, mask, R=0.3, nthreads=min(os.cpu_count() or 1, 32))
    oracle.index_free(oix)
    for n in ("found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
        assert (recs[n] == want[n]).all(), n
    f = recs["found"] == 1
    assert f.sum() > 150 and want["ncand"].mean() > 100
    assert (np.abs(recs["pos"][f].astype(np.int64) - recs["j"][f] - starts[f]) < 0.35 * lens[f]).mean() > 0.99
    ix.free(); rs.free()


# ---------------------------------------------------------------------------------------------
# assembler-side trial loop (spaced_seed.cpp:424-436, try_align :261-299)
# ---------------------------------------------------------------------------------------------

OV_FIELDS = ("id", "found", "j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit", "ncand")


def test_overlap_golden_and_oracle(ctx, golden, oracle):
    from test_oracle import overlap_workload
    for g in golden["overlap"]:
        ref, image = overlap_workload(g["seed"], g["ref_len"], g["nreads"])
        rs = ctx.seqset_one(ref)
        ix = ctx.index(rs, g["mask"], policy=1)
        reads = ctx.seqset_from_bin(image)
        # the shipped behaviour (seed_at reads byte offset pos when pos%4==0): against the reference's own output
        got = ctx.overlap(ix, reads, R=g["R"], seed_at_quirk=1)
        assert len(got) == len(g["records"])
        for k, row in enumerate(g["records"]):
            for n, v in row.items():
                assert int(got[n][k]) == v, (k, n, int(got[n][k]), v)
        # the intended behaviour (seed at base pos): against the oracle, transcripts included
        got, ops = ctx.overlap(ix, reads, R=g["R"], want_ops=True)
        oix = oracle.index_build(ref, g["mask"], policy=1)
        want = oracle.overlap(oix, ref, image, g["mask"], R=g["R"], quirk=False, nthreads=4)
        oracle.index_free(oix)
        for n in OV_FIELDS + ("cells",):
            assert (got[n] == want[n]).all(), n
        f = got["found"] == 1
        assert f.sum() > 5 and (got["dir"][f] == -1).any() and (got["dir"][f] == 1).any()
        for k in np.nonzero(f)[0]:
            o = ops[k]
            assert len(o) == got["nedit"][k]
            assert int((o != INSERT).sum()) == got["matlen_a"][k] and int((o != DELETE).sum()) == got["matlen_b"][k]


def test_overlap_longer_reads(ctx, oracle):
    """5 kbp-class CLR reads against a 45 kbp reference (head and tail windows of get_seedmap both populated)"""
    from test_oracle import overlap_workload
    ref, image = overlap_workload(91, 45000, 30)
    import cpu_libs
    o = cpu_libs.oracle()
    lens = workload.read_lengths(92, 30, mean=4000.0, sigma_log=0.4, lo=600, hi=12000)
    txt, offs, lens, _ = workload.reads(93, ref, lens, 0.05, 0.03, 0.02, nthreads=1)
    image = b"".join(o.text2bin(txt[offs[k]: offs[k] + lens[k]].tobytes()) for k in range(len(lens)))
    rs = ctx.seqset_one(ref)
    for quirk in (0, 1):
        ix = ctx.index(rs, MASKS[0], policy=1)
        reads = ctx.seqset_from_bin(image)
        got = ctx.overlap(ix, reads, R=0.3, seed_at_quirk=quirk)
        oix = oracle.index_build(ref, MASKS[0], policy=1)
        want = oracle.overlap(oix, ref, image, MASKS[0], R=0.3, quirk=bool(quirk), nthreads=8)
        oracle.index_free(oix)
        for n in OV_FIELDS + ("cells",):
            assert (got[n] == want[n]).all(), (quirk, n)
    assert got["found"].sum() > 3


def test_align_narrow_bands_packed(ctx, oracle):
    """bands of 65..513 bits: several alignments per warp (packed kernels, 4/8/16 lanes each) and the one-word-per-lane
    class; mixed lengths in one batch, failures and seg_a longer than seg_b included"""
    rng = np.random.default_rng(31)
    for R, maxlen in ((0.03, 900), (0.06, 1100), (0.03, 4000), (0.12, 2000)):
        A, B = make_pairs(rng, 160, maxlen, rates=(0.0, 0.01, 0.03, 0.08))
        assert run_batch_vs_oracle(ctx, oracle, A, B, R) > 30
    A, B = make_pairs(rng, 64, 700, rates=(0.0, 0.02))
    assert run_batch_vs_oracle(ctx, oracle, A, B, 0.04, fwd=False) >= 5
    # the config-3 generator itself: band == max_dst exactly, edits spaced to survive the early-failure line
    for alen, band in ((1000, 32), (2000, 64), (5000, 128), (1000, 256)):
        P = [workload.sweep_pair(7 * band + alen, k, alen, band) for k in range(24)]
        Rr = P[0][2]
        assert run_batch_vs_oracle(ctx, oracle, [x[0] for x in P], [x[1] for x in P], Rr) == 24


def test_align_thread_kernel_blocks_and_rounds(ctx, oracle):
    """one alignment per thread (bands of at most 96 / 160 / 288 bits): whole 32-row blocks take the unrolled loop, the rest the
    general one, the traceback moves in rounds -- batches built around those seams: every alignment of a warp failing, equal
    sequences, lengths of 32k-1 / 32k / 32k+1 / 32k+2 rows, fewer than 32 rows, seg_a longer than seg_b by up to the band,
    seg_b's tail of insertions past the last row, a lone item in a second warp"""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, n)].tobytes()

    def mutate(a, rate):
        out = bytearray()
        for ch in a:
            u = rng.random()
            if u < rate * 0.4:
                out += bytes([int(acgt[rng.integers(0, 4)]), ch])
            elif u < rate * 0.7:
                continue
            elif u < rate:
                out.append(int(acgt[rng.integers(0, 4)]))
            else:
                out.append(ch)
        return bytes(out)

    for R, base in ((0.02, 1500), (0.04, 1500), (0.045, 3000)):  # max_dst ~ 30 / 60 / 135: W = 3 / 5 / 9
        A, B = [], []
        for k in range(33):  # 33: the second warp holds one item
            n = base + (-1, 0, 1, 2, 31, 33)[k % 6] - 32 * (k % 5)
            a = rnd(n)
            b = mutate(a[:64], 0.0) + mutate(a[64:], R / 4 if k % 4 else 0.0)  # a clean start: the failure line is tight in the first rows
            if k % 7 == 3:
                b = b + rnd(int(rng.integers(1, int(R * n))))       # insertions after seg_a's end (goal on the last row)
            if k % 7 == 5:
                a = a + rnd(int(rng.integers(1, int(R * n))))       # seg_a longer: goal on the last column
            A.append(a)
            B.append(b)
        assert run_batch_vs_oracle(ctx, oracle, A, B, R) >= 10
        assert run_batch_vs_oracle(ctx, oracle, [rnd(n) for n in range(600, 640)], [rnd(n) for n in range(600, 640)], R) == 0  # all fail early
        short_a = [rnd(n) for n in range(12, 44)]
        assert run_batch_vs_oracle(ctx, oracle, short_a, [mutate(a, 0.02) + b"A" for a in short_a], 0.3) > 5  # under one block
        # the same long pairs as one batch with early failures and short ones mixed in: lanes leave the unrolled loop's
        # company at different rows
        mix_a = A[:20] + [rnd(900) for _ in range(6)] + short_a[:6]
        mix_b = B[:20] + [rnd(900) for _ in range(6)] + [a + b"C" for a in short_a[:6]]
        assert run_batch_vs_oracle(ctx, oracle, mix_a, mix_b, R) >= 5


def test_align_batch_strip_pass_and_redo(ctx, oracle):
    """wide bands of pb_align_batch take the certified strip first (align_pairs_nb_kernel) and the full band redoes what it could
    not certify: pairs whose cost runs from well inside the strip's goal side (0.75 max_dst) to beyond it, goal on the last row and
    on the last column, failures and rejected pairs mixed in"""
    rng = np.random.default_rng(401)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, n)].tobytes()

    def mutate(a, rate):
        out = bytearray()
        u = rng.random(len(a))
        for ch, x in zip(a, u):
            if x < rate * 0.45:
                out += bytes([int(acgt[rng.integers(0, 4)]), ch])
            elif x < rate * 0.75:
                continue
            elif x < rate:
                out.append(int(acgt[rng.integers(0, 4)]))
            else:
                out.append(ch)
        return bytes(out)

    A, B = [], []
    for k in range(72):
        n = int(rng.integers(1200, 5200))
        a = rnd(n)
        b = mutate(a, (0.05, 0.16, 0.2, 0.23, 0.26, 0.29)[k % 6])
        if k % 4 == 1:
            b = b + rnd(int(rng.integers(1, n // 3)))  # goal on the last row, seg_b cut to len_a + max_dst
        if k % 4 == 2:
            a = a + rnd(int(rng.integers(1, n // 3)))  # goal on the last column
        if k % 9 == 8:
            b = rnd(len(b))                            # fails early
        A.append(a)
        B.append(b)
    recs_ok = run_batch_vs_oracle(ctx, oracle, A, B, 0.3)
    assert recs_ok >= 30
    # costs on both sides of the strip's goal width: the redo pass must have had work
    costs = [oracle.align(a, b, 0.3, True, True, 26000, 6000) for a, b in zip(A, B)]
    ratio = [w["cost"] / w["max_dst"] for w in costs if w["ret"] >= 0]
    assert min(ratio) < 0.5 and max(ratio) > 0.8, (min(ratio), max(ratio))
    assert run_batch_vs_oracle(ctx, oracle, A[:24], B[:24], 0.3, maxn=4000, maxm=1000) >= 5  # some pairs rejected by the domain check


def test_config3_sweep_every_point(ctx, oracle):
    """BASELINE config 3, unit costs: every (length, band) point of the sweep, 32 pairs each, against the oracle with
    transcripts; every pair of the generator aligns (band == max_dst, edits spaced under the early-failure line)"""
    for alen in (1000, 2000, 5000, 10000, 19999):
        for band in (32, 64, 128, 256, 512):
            P = [workload.sweep_pair(1000 * band + alen, k, alen, band) for k in range(32)]
            assert run_batch_vs_oracle(ctx, oracle, [x[0] for x in P], [x[1] for x in P], P[0][2]) == 32, (alen, band)


def run_weighted_vs_oracle(ctx, oracle, A, WA, B, WB, R, fail_scale, maxn=26000, maxm=6000):
    a_blob, b_blob = b"".join(A), b"".join(B)
    wa_blob, wb_blob = np.concatenate(WA), np.concatenate(WB)
    a_len, b_len = [len(x) for x in A], [len(x) for x in B]
    a_off, b_off = np.cumsum([0] + a_len[:-1]), np.cumsum([0] + b_len[:-1])
    recs, ops = ctx.align_weighted_batch(a_blob, wa_blob, a_off, a_len, b_blob, wb_blob, b_off, b_len, R, fail_scale, maxn, maxm)
    nsucc = 0
    for i, (a, b) in enumerate(zip(A, B)):
        w = oracle.align_weighted(a, WA[i], b, WB[i], R, fail_scale, maxn, maxm)
        for k in ("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit", "fail_row", "cells"):
            assert int(recs[k][i]) == w[k], (i, k, int(recs[k][i]), w[k], len(a), len(b))
        if w["ret"] >= 0:
            assert (ops[i] == w["ops"]).all(), i
            nsucc += 1
    return nsucc, recs, ops


def test_align_quality_weighted(ctx, oracle):
    """the quality-weighted extension (config 3): per-element weights 1..4 through the two scoring hooks, against the
    extended oracle; with all weights 1 it must reproduce the unit-cost aligner (pb_align_batch) bit for bit"""
    rng = np.random.default_rng(77)
    A, B = make_pairs(rng, 160, 700, rates=(0.0, 0.02, 0.08, 0.15))
    A, B = [a for a, b in zip(A, B) if len(a) and len(b)], [b for a, b in zip(A, B) if len(a) and len(b)]
    ones_a, ones_b = [np.ones(len(a), np.uint8) for a in A], [np.ones(len(b), np.uint8) for b in B]
    n1, r1, o1 = run_weighted_vs_oracle(ctx, oracle, A, ones_a, B, ones_b, 0.3, 1.0)
    assert n1 > 40
    a_len, b_len = [len(x) for x in A], [len(x) for x in B]
    ru, ou = ctx.align_batch(b"".join(A), np.cumsum([0] + a_len[:-1]), a_len, b"".join(B), np.cumsum([0] + b_len[:-1]), b_len, 0.3)
    for k in ru.dtype.names:
        assert (ru[k] == r1[k]).all(), k
    for x, y in zip(ou, o1):
        assert (x is None and y is None) or (x == y).all()
    WA = [rng.integers(1, 5, size=len(a)).astype(np.uint8) for a in A]
    WB = [rng.integers(1, 5, size=len(b)).astype(np.uint8) for b in B]
    for R, fs in ((0.3, 1.0), (0.3, 4.0), (0.1, 2.5), (0.45, 4.0)):
        n, _, _ = run_weighted_vs_oracle(ctx, oracle, A, WA, B, WB, R, fs)
        assert n >= 10, (R, fs, n)
    # multi-word bands and the sweep generator's pairs
    A, B = make_pairs(rng, 24, 5000, rates=(0.02, 0.1))
    WA = [rng.integers(1, 5, size=len(a)).astype(np.uint8) for a in A]
    WB = [rng.integers(1, 5, size=len(b)).astype(np.uint8) for b in B]
    assert run_weighted_vs_oracle(ctx, oracle, A, WA, B, WB, 0.3, 4.0)[0] > 5
    for alen, band in ((1000, 32), (5000, 128), (19999, 512)):
        P = [workload.sweep_pair(1000 * band + alen, k, alen, band) for k in range(8)]
        A, B = [x[0] for x in P], [x[1] for x in P]
        WA = [rng.integers(1, 5, size=len(a)).astype(np.uint8) for a in A]
        WB = [rng.integers(1, 5, size=len(b)).astype(np.uint8) for b in B]
        assert run_weighted_vs_oracle(ctx, oracle, A, WA, B, WB, P[0][2], 4.0)[0] == 8


# ---------------------------------------------------------------------------------------------
# all-vs-all overlap detection (BASELINE config 5): pb_index_build_set + pb_overlap_all_run
# ---------------------------------------------------------------------------------------------

def check_allpairs(ctx, oracle, texts, image, mask, R=0.3, quirk=0, golden_row=None, splits=1):
    from allpairs_util import PAIR_FIELDS, expected_pairs, oracle_overlap_fn
    reads = ctx.seqset_from_bin(image)
    ix = ctx.index_set(reads, mask)
    n = len(reads)
    recs_l, stats_l = [], []
    for s in range(splits):  # query sub-ranges: how callers bound memory and shard over GPUs
        q0, q1 = n * s // splits, n * (s + 1) // splits
        recs, stats = ctx.overlap_all(ix, q0, q1 - q0, found_only=False, R=R, seed_at_quirk=quirk)
        assert ((recs["read_id"] >= q0) & (recs["read_id"] < q1)).all()
        recs_l.append(recs)
        stats_l.append(stats)
    recs = np.concatenate(recs_l)
    stats = {k: sum(s[k] for s in stats_l) for k in stats_l[0]}
    want, tot_ncand, tot_cells = expected_pairs(oracle_overlap_fn(oracle, mask, R=R, quirk=bool(quirk)), texts, image)
    # totals over every (T, Q) pair with a seed hit, aligned or not
    assert stats["pairs"] == len(want) and stats["try_align_calls"] == tot_ncand and stats["ref_cells"] == tot_cells
    assert stats["pairs_aligned"] == len(recs)
    key = list(zip(recs["read_id"].tolist(), recs["ref_id"].tolist()))
    assert key == sorted(key) and len(set(key)) == len(key)
    for rec in recs:  # every pair that reached the aligner, found or not
        w = want[(int(rec["ref_id"]), int(rec["read_id"]))]
        for f in ("found", "ncand", "cells") + PAIR_FIELDS:
            assert int(rec[f]) == int(w[f]), (int(rec["ref_id"]), int(rec["read_id"]), f, int(rec[f]), int(w[f]))
    got_found = {(int(r["ref_id"]), int(r["read_id"])) for r in recs if r["found"]}
    assert got_found == {k for k, v in want.items() if v["found"]}
    assert stats["pairs_found"] == len(got_found)
    f2, st2 = ctx.overlap_all(ix, found_only=True, R=R, seed_at_quirk=quirk)
    assert len(f2) == len(got_found) == st2["pairs_found"] and (f2["found"] == 1).all()
    if golden_row is not None:  # the compiled reference's own output
        assert len(f2) == len(golden_row["found"]) and st2["try_align_calls"] == golden_row["try_align_calls"]
        assert st2["pairs"] == golden_row["pairs_with_calls"]
        for rec, row in zip(f2, golden_row["found"]):
            for f in ("ref_id", "read_id") + PAIR_FIELDS:
                assert int(rec[f]) == row[f], (row, f)
    ix.free()
    reads.free()
    return len(got_found)


def test_allpairs_golden_and_oracle(ctx, golden, oracle):
    from allpairs_util import allpairs_workload
    for g in golden["allpairs"]:
        texts, image = allpairs_workload(g["seed"], g["genome_len"], g["nreads"])
        # shipped seed_at behaviour: against the compiled reference's output and the oracle
        assert check_allpairs(ctx, oracle, texts, image, g["mask"], R=g["R"], quirk=1, golden_row=g) > 100
        # intended behaviour (seed at base pos): against the oracle
        assert check_allpairs(ctx, oracle, texts, image, g["mask"], R=g["R"], quirk=0, splits=3) > 100


def test_allpairs_clr_reads(ctx, oracle):
    """CLR-like error (ins 9 / del 4 / sub 2 %), longer reads, high coverage: many seed hits per pair, dir = -1 and +1"""
    from allpairs_util import allpairs_workload
    texts, image = allpairs_workload(341, 20000, 60, mean=3000.0, lo=520, hi=9000, err=(0.09, 0.04, 0.02))
    assert check_allpairs(ctx, oracle, texts, image, MASKS[0], splits=2) > 30
    # a read set with a short read (len <= 500 is dropped by open_binary) and a duplicate read
    texts2 = texts[:20] + [texts[3], texts[5][:400]]
    import cpu_libs
    o = cpu_libs.oracle()
    image2 = b"".join(o.text2bin(t) for t in texts2)
    assert check_allpairs(ctx, oracle, texts2, image2, MASKS[1]) > 5


def test_allpairs_config5_set_size(ctx, oracle):
    """BASELINE config 5 at its set size: 50 000 CLR reads of a 4.6 Mbp genome (~54x), whole-set index, all-vs-all over a
    slice of the query reads; sampled target reads are recomputed by the oracle the reference's way (one seed map and one
    trial loop over the queries per target, spaced_seed.cpp:424-436) and every pair's fields must agree"""
    nreads, nq = 50_000, 4000
    g = workload.reference(2, 4_600_000)
    lens = workload.read_lengths(3, nreads, mean=5000.0, sigma_log=0.5, lo=520, hi=19999)
    txt, offs, lens, starts = workload.reads(3, g, lens)
    mask = MASKS[0]
    rs = ctx.seqset(txt, offs, lens)
    ix = ctx.index_set(rs, mask)
    recs, stats = ctx.overlap_all(ix, 0, nq, found_only=True, R=0.3)
    assert stats["pairs_found"] == len(recs) > 10_000
    image = b"".join(oracle.text2bin(txt[offs[k]: offs[k] + lens[k]].tobytes()) for k in range(nq))
    by_target = {}
    for r in recs:
        by_target.setdefault(int(r["ref_id"]), {})[int(r["read_id"])] = r
    targets = [t for t in sorted(by_target, key=lambda t: -len(by_target[t]))[:2]] + [17, 25_000, 49_999]
    checked = 0
    for t in targets:
        tt = txt[offs[t]: offs[t] + lens[t]]
        oix = oracle.index_build(tt, mask, policy=1)
        want = oracle.overlap(oix, tt, image, mask, R=0.3, quirk=False, nthreads=min(os.cpu_count() or 1, 32))
        oracle.index_free(oix)
        wf = {int(q): want[q] for q in np.nonzero(want["found"] == 1)[0] if q != t}
        got = by_target.get(t, {})
        assert sorted(got) == sorted(wf), (t, sorted(set(got) ^ set(wf))[:10])
        for q, w in wf.items():
            for f in ("j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
                assert int(got[q][f]) == int(w[f]), (t, q, f, int(got[q][f]), int(w[f]))
        checked += len(wf)
    assert checked > 20
    ix.free(); rs.free()


def test_allpairs_errors(ctx):
    from allpairs_util import allpairs_workload
    from pacbioassembly_b200 import PbError
    texts, image = allpairs_workload(351, 3000, 8)
    reads = ctx.seqset_from_bin(image)
    single = ctx.index(reads, MASKS[0], policy=1, seq=0)
    with pytest.raises(PbError):
        ctx.overlap_all(single)  # not a whole-set index
    ix = ctx.index_set(reads, MASKS[0])
    with pytest.raises(PbError):
        ctx.overlap(ix, reads)  # and a whole-set index is not a single-reference one
    with pytest.raises(PbError):
        ctx.overlap_all(ix, 0, len(reads) + 1)
    recs, st = ctx.overlap_all(ix, 0, 0)
    assert len(recs) == 0 and st["pairs"] == 0
    long_ref = ctx.seqset_one(workload.reference(5, 20017))
    with pytest.raises(PbError):
        ctx.index_set(long_ref, MASKS[0])


def test_allpairs_long_runs(ctx, oracle):
    """a tiny genome at very high coverage: every probe hits nearly every other read, so a query read's run of hits exceeds
    the 8192 keys the regrouping sort holds in shared memory (in-place global-memory network), and every pair has hits"""
    from allpairs_util import allpairs_workload
    texts, image = allpairs_workload(371, 800, 150, mean=540.0, lo=520, hi=600, err=(0.01, 0.005, 0.005))
    assert check_allpairs(ctx, oracle, texts, image, MASKS[0]) > 20000


def test_packed_bucket_headers_escape(ctx, oracle):
    """probe kernels read one packed word per query (start | capped count); with PB_PK_SHIFT=30 the cap is 3, so the
    repeats of this reference take the exact-count escape path; results must not change"""
    rng = np.random.default_rng(5)
    unit = workload.reference(77, 3000)
    ref = np.concatenate([unit] * 6 + [workload.reference(78, 20000)])  # every seed of `unit` has 6 positions
    lens = workload.read_lengths(79, 40, mean=1200.0, sigma_log=0.4, lo=520, hi=3000)
    txt, offs, lens, _ = workload.reads(80, ref, lens, 0.03, 0.02, 0.01, nthreads=1)
    oix = oracle.index_build(ref, MASKS[0], 0)
    want = oracle.locate(oix, ref, txt, offs, lens, MASKS[0], R=0.3, nthreads=4)
    oracle.index_free(oix)
    rs = ctx.seqset_one(ref)
    for shift in (None, "30"):
        if shift:
            os.environ["PB_PK_SHIFT"] = shift
        try:
            ix = ctx.index(rs, MASKS[0])
        finally:
            os.environ.pop("PB_PK_SHIFT", None)
        got = ctx.locate(ix, txt, offs, lens, R=0.3)
        for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand", "cells"):
            assert (got[n] == want[n]).all(), (shift, n)
        ix.free()
    assert want["found"].sum() > 10 and want["ncand"].max() >= 6


# ---------------------------------------------------------------------------------------------
# consensus voting: pb_consensus_* against the oracle's ref_seq restatement and the reference's own output
# ---------------------------------------------------------------------------------------------

def test_consensus_primitives(ctx, oracle):
    """create / elect / append / prepend / evolve, box by box: one batch of matches voted on the GPU in one launch against the
    same matches voted one by one through the oracle's elect (transcripts from the oracle's own align)"""
    from test_oracle import overlap_workload
    ref, image = overlap_workload(601, 9000, 50)
    mask = MASKS[0]
    cons = ctx.consensus(ref, weight=2)
    oc = oracle.cons_create(ref.tobytes(), 2)
    assert (cons.votes() == oracle.cons_votes(oc)).all() and len(cons) == len(ref)
    reads = ctx.seqset_from_bin(image)
    cur = cons.seqset()
    ix = ctx.index(cur, mask, policy=1)
    recs, ops, ops_off = ctx.overlap(ix, reads, want_ops="raw", R=0.3)
    assert recs["found"].sum() > 10 and (recs["dir"][recs["found"] == 1] == -1).any()
    cons.elect(reads, recs, ops, ops_off)
    from pacbioassembly_b200.assemble import kept_records
    rec_bytes = kept_records(image)
    for k in np.nonzero(recs["found"] == 1)[0]:
        r = recs[k]
        fwd = r["dir"] == 1
        text = oracle.bin2text(rec_bytes[k])
        r_off = int(r["ref_pos"]) + (0 if fwd else 15)
        s_off = int(r["read_pos"]) + (0 if fwd else 15)
        a = ref.tobytes()[r_off:] if fwd else ref.tobytes()[: r_off + 1]
        b = text[s_off:] if fwd else text[: s_off + 1]
        al = oracle.align(a, b, R=0.3, a_fwd=fwd, b_fwd=fwd)
        assert al["nedit"] == r["nedit"] and (al["ops"] == ops[ops_off[k]: ops_off[k] + r["nedit"]]).all()
        oracle.cons_elect(oc, r_off, al["ops"], al["vals"], fwd)
    assert (cons.votes() == oracle.cons_votes(oc)).all()
    # growth at both ends, then evolve; twice (the second evolve sees the suppliments the first one left behind)
    for rnd in range(2):
        cons.append(b"ACGTTGCAAC" * (rnd + 1))
        cons.prepend(b"TTGACCA")
        oracle.lib.pbo_cons_append(oc, b"ACGTTGCAAC" * (rnd + 1), 10 * (rnd + 1))
        oracle.lib.pbo_cons_prepend(oc, b"TTGACCA", 7)
        full, before = oracle.cons_full_text(oc)
        assert cons.text(full=True) == full and cons.extent() == (before, len(full))
        assert (cons.votes() == oracle.cons_votes(oc)).all()
        cons.evolve()
        oracle.lib.pbo_cons_evolve(oc)
        assert cons.text() == oracle.cons_text(oc) and cons.extent() == (0, len(cons))
        assert (cons.votes() == oracle.cons_votes(oc)).all()
    assert cons.text() != ref.tobytes()  # votes changed the text (insertions / deletions took effect)
    oracle.lib.pbo_cons_free(oc)


def test_assemble_unlocked_rounds(ctx, golden, oracle):
    """the assembler's rounds with voting and growth (pacbioassembly_b200/assemble.py over pb_overlap_batch with ref_shift,
    pb_consensus_elect_batch / append / prepend / evolve): consensus of every round, who was found when and the records, against
    the compiled reference's own output (golden) and the oracle"""
    import hashlib
    from pacbioassembly_b200.assemble import assemble_rounds
    from test_oracle import ASM_FIELDS, assemble_workload
    for g in golden["assemble"]:
        ref0, image = assemble_workload(g["seed"], g["genome_len"], g["nreads"], g["ref_read"])
        cons, fr, recs, passes = assemble_rounds(ctx, ref0, image, g["masks"], weight=g["weight"], seed_at_quirk=1)
        assert [len(c) for c in cons] == g["consensus_len"]
        assert [hashlib.sha1(c).hexdigest() for c in cons] == g["consensus_sha1"]
        assert fr.tolist() == g["found_round"]
        wc, wf, wr = oracle.assemble(ref0, image, g["masks"], weight=g["weight"], quirk=True)
        hit = fr > 0  # a record describes the round in which its read was found; reads never found have none
        for n in ASM_FIELDS:
            assert (recs[n][hit] == wr[n][hit]).all(), n
        assert (recs["found"][~hit] == 0).all() and (wr["found"][~hit] == 0).all()
        assert max(passes) > 3  # growth really happened inside rounds
    # intended seed_at behaviour (no golden: the compiled reference only has the shipped one)
    ref0, image = assemble_workload(541, 8000, 50, 7)
    masks = [MASKS[0], MASKS[3], MASKS[1]]
    cons, fr, recs, _ = assemble_rounds(ctx, ref0, image, masks, seed_at_quirk=0)
    wc, wf, wr = oracle.assemble(ref0, image, masks, quirk=False)
    assert cons == wc and (fr == wf).all()
    for n in ASM_FIELDS:
        assert (recs[n][fr > 0] == wr[n][fr > 0]).all(), n
