"""ctypes bindings for the CPU checkers used by tests/, smoke() and bench.py's cpu legs.

* ``oracle``  -> oracle/liboracle.so      (plain-C restatement, oracle/pb_oracle.c)
* ``ref``     -> oracle/_ref/libpbref.so  (the unmodified reference behind oracle/ref_shim.cpp);
                 None when it has not been built (needs /root/reference at build time).
These are TEST INFRASTRUCTURE: nothing under pacbioassembly_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REFERENCE = os.environ.get("PB_REFERENCE", "/root/reference")

MATCH, INSERT, DELETE = 1, 2, 3


class AlignOut(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost", "diag_cost", "nedit", "fail_row")] \
        + [("cells", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class LocateRec(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b", "nedit", "ncand")] \
        + [("cells", C.c_int64)]


LOCATE_DTYPE = np.dtype([(n, np.int32) for n in
                         ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a", "matlen_b",
                          "nedit", "ncand")] + [("_pad", np.int32), ("cells", np.int64)])
assert LOCATE_DTYPE.itemsize == C.sizeof(LocateRec) == 56

OVERLAP_DTYPE = np.dtype([(n, np.int32) for n in
                          ("id", "found", "j", "ref_pos", "cost", "read_pos", "dir", "matlen_a", "matlen_b", "nedit",
                           "ncand", "_pad")] + [("cells", np.int64)])
assert OVERLAP_DTYPE.itemsize == 56


def build_oracle(with_ref: bool = True) -> None:
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
    if with_ref and os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref", f"REFERENCE={REFERENCE}"])


class Oracle:
    def __init__(self):
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle(with_ref=False)
        L = self.lib = C.CDLL(so)
        L.pbo_encode.restype = C.c_uint32
        L.pbo_encode.argtypes = [C.c_char_p, C.c_size_t]
        L.pbo_decode.argtypes = [C.c_uint32, C.c_char_p]
        L.pbo_text2bin.restype = C.c_size_t
        L.pbo_text2bin.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.pbo_bin2text.restype = C.c_size_t
        L.pbo_bin2text.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.pbo_seed_at.restype = C.c_uint32
        L.pbo_seed_at.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.pbo_parse_pattern.restype = C.c_uint32
        L.pbo_parse_pattern.argtypes = [C.c_char_p]
        L.pbo_index_build.restype = C.c_void_p
        L.pbo_index_build.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int]
        L.pbo_index_free.argtypes = [C.c_void_p]
        L.pbo_index_nkeys.restype = C.c_size_t
        L.pbo_index_nkeys.argtypes = [C.c_void_p]
        L.pbo_index_nentries.restype = C.c_size_t
        L.pbo_index_nentries.argtypes = [C.c_void_p]
        L.pbo_index_find.restype = C.c_size_t
        L.pbo_index_find.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(C.c_int32))]
        L.pbo_align.restype = C.c_int
        L.pbo_align.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                C.POINTER(AlignOut), C.c_void_p, C.c_void_p, C.c_size_t]
        L.pbo_align_weighted.restype = C.c_int
        L.pbo_align_weighted.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double,
                                         C.c_double, C.c_int, C.c_int, C.POINTER(AlignOut), C.c_void_p, C.c_size_t]
        L.pbo_locate.restype = C.c_int64
        L.pbo_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p]

        L.pbo_overlap.restype = C.c_int64
        L.pbo_overlap.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32,
                                  C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def overlap(self, ix, ref: np.ndarray, image: bytes, mask: int, R: float = 0.3, max_trial: int = 32, min_overlap: int = 64,
                maxn: int = 26000, maxm: int = 6000, quirk: bool = False, nthreads: int = 1, min_excl: int = 500,
                max_excl: int = 20000):
        """spaced_seed.cpp:424-436 + try_align :261-299 over a .bin image against a REFSEQ-policy index"""
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        img = np.frombuffer(image, dtype=np.uint8)
        args = (ix, ref.ctypes.data, len(ref), img.ctypes.data, len(img), min_excl, max_excl, mask, R, max_trial, min_overlap,
                maxn, maxm, int(quirk), nthreads)
        nk = self.lib.pbo_overlap(*args, None)
        recs = np.zeros(nk, dtype=OVERLAP_DTYPE)
        self.lib.pbo_overlap(*args, recs.ctypes.data)
        return recs

    def assemble(self, ref: np.ndarray, image: bytes, round_masks, weight: int = 1, R: float = 0.3, max_trial: int = 32,
                 min_overlap: int = 64, maxn: int = 26000, maxm: int = 6000, quirk: bool = False, min_excl: int = 500,
                 max_excl: int = 20000):
        """unlocked assembler rounds (spaced_seed.cpp:408-453 with ref_seq voting / growth / evolve):
        returns (list of consensus bytes per round, found_round int32[nkept], OVERLAP_DTYPE records)"""
        L = self.lib
        L.pbo_assemble.restype = C.c_int64
        L.pbo_assemble.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        img = np.frombuffer(image, dtype=np.uint8)
        masks = np.ascontiguousarray(round_masks, dtype=np.uint32)
        head = (ref.ctypes.data, len(ref), weight, img.ctypes.data, len(img), min_excl, max_excl, masks.ctypes.data, len(masks), R,
                max_trial, min_overlap, maxn, maxm, int(quirk))
        nk = L.pbo_assemble(*head, None, 0, None, None, None)
        stride = 800000
        cons = np.zeros(len(masks) * stride, dtype=np.uint8)
        clen = np.zeros(len(masks), dtype=np.int32)
        fr = np.zeros(nk, dtype=np.int32)
        recs = np.zeros(nk, dtype=OVERLAP_DTYPE)
        L.pbo_assemble(*head, cons.ctypes.data, stride, clen.ctypes.data, fr.ctypes.data, recs.ctypes.data)
        return [cons[r * stride: r * stride + clen[r]].tobytes() for r in range(len(masks))], fr, recs

    # ref_seq's voting state, piece by piece (the GPU consensus primitives are checked against these)
    def cons_create(self, text: bytes, weight: int = 1):
        L = self.lib
        L.pbo_cons_create.restype = C.c_void_p
        L.pbo_cons_create.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_size_t]
        L.pbo_cons_free.argtypes = [C.c_void_p]
        L.pbo_cons_length.restype = C.c_size_t
        L.pbo_cons_length.argtypes = [C.c_void_p]
        L.pbo_cons_extent.restype = C.c_size_t
        L.pbo_cons_extent.argtypes = [C.c_void_p, C.c_void_p]
        L.pbo_cons_text.restype = C.c_void_p
        L.pbo_cons_text.argtypes = [C.c_void_p]
        L.pbo_cons_append.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.pbo_cons_prepend.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.pbo_cons_elect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.pbo_cons_evolve.argtypes = [C.c_void_p]
        L.pbo_cons_votes.restype = C.c_int64
        L.pbo_cons_votes.argtypes = [C.c_void_p, C.c_void_p]
        return L.pbo_cons_create(text, len(text), weight, 800000)

    def cons_text(self, c) -> bytes:
        return C.string_at(self.lib.pbo_cons_text(c), self.lib.pbo_cons_length(c))

    def cons_full_text(self, c):
        """(text of [pre, post), beg - pre)"""
        before = C.c_long(0)
        n = self.lib.pbo_cons_extent(c, C.byref(before))
        return C.string_at(self.lib.pbo_cons_text(c) - before.value, n), before.value

    def cons_votes(self, c) -> np.ndarray:
        n = self.lib.pbo_cons_votes(c, None)
        out = np.zeros((n, 9), dtype=np.int32)
        self.lib.pbo_cons_votes(c, out.ctypes.data)
        return out

    def cons_elect(self, c, pos: int, ops: np.ndarray, vals: np.ndarray, forward: bool):
        ops = np.ascontiguousarray(ops, dtype=np.uint8)
        vals = np.ascontiguousarray(vals, dtype=np.uint8)
        self.lib.pbo_cons_elect(c, pos, ops.ctypes.data, vals.ctypes.data, len(ops), int(forward))

    # -- L0 --
    def encode(self, text: bytes) -> int:
        return self.lib.pbo_encode(text, len(text))

    def decode(self, code: int) -> bytes:
        buf = C.create_string_buffer(16)
        self.lib.pbo_decode(code, buf)
        return buf.raw

    def text2bin(self, text: bytes) -> bytes:
        cap = 4 + (len(text) + 3) // 4
        buf = (C.c_uint8 * cap)()
        n = self.lib.pbo_text2bin(text, len(text), buf, cap)
        return bytes(buf[:n])

    def bin2text(self, rec: bytes) -> bytes:
        n = int.from_bytes(rec[:4], "little")
        out = C.create_string_buffer(n + 1)
        src = (C.c_uint8 * len(rec)).from_buffer_copy(rec)
        self.lib.pbo_bin2text(src, out, n + 1)
        return out.raw[:n]

    def seed_at(self, rec: bytes, pos: int, quirk: bool = False) -> int:
        src = (C.c_uint8 * len(rec)).from_buffer_copy(rec)
        return self.lib.pbo_seed_at(src, len(rec), pos, 1 if quirk else 0)

    def parse_pattern(self, pat: bytes) -> int:
        return self.lib.pbo_parse_pattern(pat)

    # -- L1 --
    def index_build(self, ref: np.ndarray, mask: int, policy: int = 0):
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        return self.lib.pbo_index_build(ref.ctypes.data, len(ref), mask, policy)

    def index_free(self, ix):
        self.lib.pbo_index_free(ix)

    def index_stats(self, ix):
        return self.lib.pbo_index_nkeys(ix), self.lib.pbo_index_nentries(ix)

    def index_find(self, ix, key: int) -> list[int]:
        p = C.POINTER(C.c_int32)()
        n = self.lib.pbo_index_find(ix, key, C.byref(p))
        return [p[i] for i in range(n)]

    # -- L2 --
    def align(self, a: bytes, b: bytes, R: float = 0.3, a_fwd: bool = True, b_fwd: bool = True,
              maxn: int = 26000, maxm: int = 6000, want_ops: bool = True):
        """a/b are the element sequences in accessor order is NOT assumed: pass the underlying text and direction;
        a backward accessor starts at the LAST byte of ``a`` and walks down (seq_accessor semantics)."""
        abuf = np.frombuffer(a, dtype=np.uint8) if len(a) else np.zeros(1, np.uint8)
        bbuf = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
        ap = abuf.ctypes.data + (0 if a_fwd else max(len(a) - 1, 0))
        bp = bbuf.ctypes.data + (0 if b_fwd else max(len(b) - 1, 0))
        out = AlignOut()
        cap = len(a) + len(b) + 8
        ops = np.zeros(cap, dtype=np.uint8)
        vals = np.zeros(cap, dtype=np.uint8)
        self.lib.pbo_align(ap, len(a), 1 if a_fwd else -1, bp, len(b), 1 if b_fwd else -1, R, maxn, maxm,
                           C.byref(out), ops.ctypes.data if want_ops else None,
                           vals.ctypes.data if want_ops else None, cap)
        d = out.as_dict()
        if want_ops and out.ret >= 0:
            d["ops"] = ops[: out.nedit].copy()
            d["vals"] = vals[: out.nedit].copy()
        return d

    def align_weighted(self, a: bytes, wa, b: bytes, wb, R: float = 0.3, fail_scale: float = 1.0, maxn: int = 26000,
                       maxm: int = 6000):
        """the quality-weighted EXTENSION (pb_oracle.h): forward views, weights 1..4 per element of a and of b"""
        abuf = np.frombuffer(a, dtype=np.uint8) if len(a) else np.zeros(1, np.uint8)
        bbuf = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
        wa = np.ascontiguousarray(wa, dtype=np.uint8) if len(a) else np.ones(1, np.uint8)
        wb = np.ascontiguousarray(wb, dtype=np.uint8) if len(b) else np.ones(1, np.uint8)
        assert len(wa) >= len(a) and len(wb) >= len(b)
        out = AlignOut()
        cap = len(a) + len(b) + 8
        ops = np.zeros(cap, dtype=np.uint8)
        self.lib.pbo_align_weighted(abuf.ctypes.data, len(a), wa.ctypes.data, bbuf.ctypes.data, len(b), wb.ctypes.data, R,
                                    fail_scale, maxn, maxm, C.byref(out), ops.ctypes.data, cap)
        d = out.as_dict()
        if out.ret >= 0:
            d["ops"] = ops[: out.nedit].copy()
        return d

    def locate(self, ix, ref: np.ndarray, reads: np.ndarray, offs: np.ndarray, lens: np.ndarray, mask: int,
               R: float = 0.15, ntrial: int = 50, minlen: int = 500, maxn: int = 40000, maxm: int = 6000,
               nthreads: int = 1, want_ops: bool = False):
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        kept = lens >= minlen
        nk = int(kept.sum())
        recs = np.zeros(nk, dtype=LOCATE_DTYPE)
        ops = ops_off = None
        if want_ops:
            caps = (lens[kept].astype(np.int64) * 2 + maxm + 16)
            ops_off = np.zeros(nk, dtype=np.int64)
            if nk:
                np.cumsum(caps[:-1], out=ops_off[1:])
            ops = np.zeros(int(caps.sum()) + 1, dtype=np.uint8)
        n = self.lib.pbo_locate(ix, ref.ctypes.data, len(ref), reads.ctypes.data, offs.ctypes.data, lens.ctypes.data,
                                len(lens), mask, R, ntrial, minlen, maxn, maxm, nthreads, recs.ctypes.data,
                                ops.ctypes.data if want_ops else None, ops_off.ctypes.data if want_ops else None)
        assert n == nk
        if want_ops:
            return recs, [ops[ops_off[k]: ops_off[k] + recs["nedit"][k]].copy() for k in range(nk)]
        return recs


class Ref:
    """The unmodified reference behind oracle/ref_shim.cpp (oracle/_ref/libpbref.so)."""

    def __init__(self, so: str):
        L = self.lib = C.CDLL(so)
        L.pbref_encode.restype = C.c_uint32
        L.pbref_encode.argtypes = [C.c_char_p]
        L.pbref_decode.argtypes = [C.c_uint32, C.c_char_p]
        L.pbref_text2bin.restype = C.c_uint32
        L.pbref_text2bin.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32]
        L.pbref_bin2text.restype = C.c_uint32
        L.pbref_bin2text.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.pbref_seed_at.restype = C.c_uint32
        L.pbref_seed_at.argtypes = [C.c_void_p, C.c_int]
        L.pbref_parse_pattern.restype = C.c_uint32
        L.pbref_parse_pattern.argtypes = [C.c_char_p]
        L.pbref_c2i.argtypes = [C.c_int]
        L.pbref_align.restype = C.c_int
        L.pbref_align.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                  C.POINTER(AlignOut), C.c_void_p, C.c_void_p, C.c_size_t]
        L.pbref_index_build.restype = C.c_long
        L.pbref_index_build.argtypes = [C.c_void_p, C.c_long, C.c_uint32, C.c_int]
        L.pbref_index_find.restype = C.c_long
        L.pbref_index_find.argtypes = [C.c_uint32, C.c_void_p, C.c_long]
        L.pbref_locate.restype = C.c_int64
        L.pbref_locate.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32,
                                   C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

        L.pbref_locator_open.restype = C.c_void_p
        L.pbref_locator_open.argtypes = [C.c_void_p, C.c_long, C.c_uint32]
        L.pbref_locator_close.argtypes = [C.c_void_p]
        L.pbref_locator_run.restype = C.c_int64
        L.pbref_locator_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    def overlap(self, ref: np.ndarray, image: bytes, mask: int, R: float = 0.3, max_trial: int = 32, min_excl: int = 500,
                max_excl: int = 20000):
        """the reference's own try_align machinery (locked ref_seq, get_seedmap, seed_at incl. its pos%4==0 branch)"""
        self.lib.pbref_overlap.restype = C.c_int64
        self.lib.pbref_overlap.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_uint32, C.c_double,
                                           C.c_int, C.c_void_p]
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        img = np.frombuffer(image + b"\0" * 65536, dtype=np.uint8).copy()  # the over-reads of seed_at land in zeros
        nrec = 0
        p = 0
        while p + 4 <= len(image):
            l = int.from_bytes(image[p:p + 4], "little")
            nrec += min_excl < l < max_excl
            p += 4 + (l + 3) // 4
        recs = np.zeros(nrec, dtype=OVERLAP_DTYPE)
        n = self.lib.pbref_overlap(ref.ctypes.data, len(ref), img.ctypes.data, len(image), min_excl, max_excl, mask, R, max_trial,
                                   recs.ctypes.data)
        assert n == nrec
        return recs

    def assemble(self, ref: np.ndarray, image: bytes, round_masks, weight: int = 1, R: float = 0.3, max_trial: int = 32,
                 min_excl: int = 500, max_excl: int = 20000):
        """unlocked rounds through the reference's own ref_seq (try_align votes / grows, evolve); same outputs as Oracle.assemble"""
        L = self.lib
        L.pbref_assemble.restype = C.c_int64
        L.pbref_assemble.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_double, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        img = np.frombuffer(image + b"\0" * 65536, dtype=np.uint8).copy()
        masks = np.ascontiguousarray(round_masks, dtype=np.uint32)
        head = (ref.ctypes.data, len(ref), weight, img.ctypes.data, len(image), min_excl, max_excl, masks.ctypes.data, len(masks), R,
                max_trial)
        nk = L.pbref_assemble(*head, None, 0, None, None, None)
        stride = 800000
        cons = np.zeros(len(masks) * stride, dtype=np.uint8)
        clen = np.zeros(len(masks), dtype=np.int32)
        fr = np.zeros(nk, dtype=np.int32)
        recs = np.zeros(nk, dtype=OVERLAP_DTYPE)
        L.pbref_assemble(*head, cons.ctypes.data, stride, clen.ctypes.data, fr.ctypes.data, recs.ctypes.data)
        return [cons[r * stride: r * stride + clen[r]].tobytes() for r in range(len(masks))], fr, recs

    def locator_open(self, ref: np.ndarray, mask: int):
        """locator.cpp:57-66 (contig + seed map), built once"""
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        return self.lib.pbref_locator_open(ref.ctypes.data, len(ref), mask)

    def locator_close(self, h):
        self.lib.pbref_locator_close(h)

    def locator_run(self, h, reads, offs, lens, R=0.15, ntrial=50, minlen=500, nthreads=1):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        nk = int((lens >= minlen).sum())
        recs = np.zeros(nk, dtype=LOCATE_DTYPE)
        n = self.lib.pbref_locator_run(h, reads.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(lens), R, ntrial,
                                       minlen, nthreads, recs.ctypes.data, None, None)
        assert n == nk
        return recs

    def encode(self, text: bytes) -> int:
        return self.lib.pbref_encode(text + b"\0" * 17)

    def decode(self, code: int) -> bytes:
        buf = C.create_string_buffer(17)
        self.lib.pbref_decode(code, buf)
        return buf.raw[:16]

    def text2bin(self, text: bytes) -> bytes:
        cap = 4 + (len(text) + 3) // 4
        buf = (C.c_uint8 * (cap + 8))()
        n = self.lib.pbref_text2bin(text + b"\0", buf, cap)
        return bytes(buf[:n])

    def bin2text(self, rec: bytes) -> bytes:
        n = int.from_bytes(rec[:4], "little")
        out = C.create_string_buffer(n + 8)
        src = (C.c_uint8 * (len(rec) + 8)).from_buffer_copy(rec + b"\0" * 8)
        self.lib.pbref_bin2text(src, out, n + 1)
        return out.raw[:n]

    def seed_at(self, rec: bytes, pos: int) -> int:
        """rec is zero-padded here so the reference's one-byte over-read is defined."""
        pad = rec + b"\0" * (pos + 16)
        src = (C.c_uint8 * len(pad)).from_buffer_copy(pad)
        return self.lib.pbref_seed_at(src, pos)

    def parse_pattern(self, pat: bytes) -> int:
        return self.lib.pbref_parse_pattern(pat)

    def c2i(self, ch: int) -> int:
        return self.lib.pbref_c2i(ch)

    def align(self, a: bytes, b: bytes, R: float = 0.3, a_fwd: bool = True, b_fwd: bool = True, which: int = 0,
              want_ops: bool = True):
        abuf = np.frombuffer(a, dtype=np.uint8).copy() if len(a) else np.zeros(1, np.uint8)
        bbuf = np.frombuffer(b, dtype=np.uint8).copy() if len(b) else np.zeros(1, np.uint8)
        ap = abuf.ctypes.data + (0 if a_fwd else max(len(a) - 1, 0))
        bp = bbuf.ctypes.data + (0 if b_fwd else max(len(b) - 1, 0))
        out = AlignOut()
        cap = len(a) + len(b) + 8
        ops = np.zeros(cap, dtype=np.uint8)
        vals = np.zeros(cap, dtype=np.uint8)
        self.lib.pbref_align(which, ap, len(a), int(a_fwd), bp, len(b), int(b_fwd), R, C.byref(out),
                             ops.ctypes.data if want_ops else None, vals.ctypes.data if want_ops else None, cap)
        d = out.as_dict()
        if want_ops and out.ret >= 0:
            d["ops"] = ops[: out.nedit].copy()
            d["vals"] = vals[: out.nedit].copy()
        return d

    def index_build(self, ref: np.ndarray, mask: int, policy: int = 0) -> int:
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        return self.lib.pbref_index_build(ref.ctypes.data, len(ref), mask, policy)

    def index_find(self, key: int, cap: int = 1 << 16) -> list[int]:
        buf = np.zeros(cap, dtype=np.int32)
        n = self.lib.pbref_index_find(key, buf.ctypes.data, cap)
        assert n <= cap
        return buf[:n].tolist()

    def locate(self, ref: np.ndarray, reads: np.ndarray, offs: np.ndarray, lens: np.ndarray, mask: int,
               R: float = 0.15, ntrial: int = 50, minlen: int = 500, nthreads: int = 1, want_ops: bool = False):
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        kept = lens >= minlen
        nk = int(kept.sum())
        recs = np.zeros(nk, dtype=LOCATE_DTYPE)
        ops = ops_off = None
        if want_ops:
            caps = (lens[kept].astype(np.int64) * 2 + 6000 + 16)
            ops_off = np.zeros(nk, dtype=np.int64)
            if nk:
                np.cumsum(caps[:-1], out=ops_off[1:])
            ops = np.zeros(int(caps.sum()) + 1, dtype=np.uint8)
        n = self.lib.pbref_locate(ref.ctypes.data, len(ref), reads.ctypes.data, offs.ctypes.data, lens.ctypes.data,
                                  len(lens), mask, R, ntrial, minlen, nthreads, recs.ctypes.data,
                                  ops.ctypes.data if want_ops else None, ops_off.ctypes.data if want_ops else None)
        assert n == nk
        if want_ops:
            return recs, [ops[ops_off[k]: ops_off[k] + recs["nedit"][k]].copy() for k in range(nk)]
        return recs


_oracle = None
_ref = False


def oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def ref() -> Ref | None:
    """The compiled reference, or None if oracle/_ref/libpbref.so is absent and cannot be built here."""
    global _ref
    if _ref is False:
        so = os.path.join(ORACLE_DIR, "_ref", "libpbref.so")
        if not os.path.exists(so) and os.path.isdir(os.path.join(REFERENCE, "src")):
            try:
                build_oracle(with_ref=True)
            except Exception:
                pass
        _ref = Ref(so) if os.path.exists(so) else None
    return _ref
