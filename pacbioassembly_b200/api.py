"""ctypes binding of the C ABI in include/pacbio_b200.h (Python host side of the drop-in boundary).

Every method maps 1:1 onto an entry point of libpacbio_b200.so; there is no Python or CPU implementation
behind any of them.  If the shared library is missing, or no sm_100a device is visible, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB") or os.path.join(_HERE, "libpacbio_b200.so")  # PB_LIB: experiment builds

MATCH, INSERT, DELETE = 1, 2, 3
POLICY_LOCATOR, POLICY_REFSEQ = 0, 1
T_NAMES = ("h2d", "ingest", "seed", "index", "probe", "prefilter", "align", "d2h", "total")

ALIGN_DTYPE = np.dtype([(n, np.int32) for n in ("ret", "len_a", "len_b", "max_dst", "matlen_a", "matlen_b", "cost",
                                                 "diag_cost", "nedit", "fail_row")] + [("cells", np.int64)])
LOCATE_DTYPE = np.dtype([(n, np.int32) for n in ("nseq", "found", "j", "pos", "cost", "seg_len", "diag_cost", "matlen_a",
                                                  "matlen_b", "nedit", "ncand", "_pad")] + [("cells", np.int64)])
OVERLAP_DTYPE = np.dtype([(n, np.int32) for n in ("id", "found", "j", "ref_pos", "cost", "read_pos", "dir", "matlen_a",
                                                   "matlen_b", "nedit", "ncand", "_pad")] + [("cells", np.int64)])
PAIR_DTYPE = np.dtype([(n, np.int32) for n in ("read_id", "found", "j", "ref_pos", "cost", "read_pos", "dir", "matlen_a",
                                               "matlen_b", "nedit", "ncand", "ref_id")] + [("cells", np.int64)])
PAIR_STATS = ("seed_hits", "pairs", "pairs_aligned", "pairs_found", "try_align_calls", "ref_cells", "k3_alignments", "k3_cells")
assert ALIGN_DTYPE.itemsize == 48 and LOCATE_DTYPE.itemsize == 56 and OVERLAP_DTYPE.itemsize == 56 and PAIR_DTYPE.itemsize == 56


class LocateParams(C.Structure):
    _fields_ = [("R", C.c_double), ("ntrial", C.c_int32), ("minlen", C.c_int32), ("maxn", C.c_int32),
                ("maxm", C.c_int32), ("want_ops", C.c_int32), ("reserved", C.c_int32)]


class OverlapParams(C.Structure):
    _fields_ = [("R", C.c_double), ("max_trial", C.c_int32), ("min_overlap", C.c_int32), ("maxn", C.c_int32),
                ("maxm", C.c_int32), ("seed_at_quirk", C.c_int32), ("want_ops", C.c_int32), ("ref_shift", C.c_int32)]


class PbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pacbio_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load libpacbio_b200.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m pacbioassembly_b200.build` "
                          "(the product path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_size_t
    P = C.POINTER
    sig = {
        "pb_abi_version": (C.c_int, []),
        "pb_device_count": (C.c_int, []),
        "pb_ctx_create": (C.c_int, [C.c_int, P(vp)]),
        "pb_ctx_destroy": (None, [vp]),
        "pb_last_error": (C.c_char_p, [vp]),
        "pb_ctx_stream": (vp, [vp]),
        "pb_ctx_launch_count": (i64, [vp]),
        "pb_ctx_set_scratch_limit": (C.c_int, [vp, sz]),
        "pb_ctx_timings": (C.c_int, [vp, vp]),
        "pb_int_pipe_peak": (C.c_int, [vp, P(C.c_double)]),
        "pb_random_gather_peak": (C.c_int, [vp, sz, P(C.c_double)]),
        "pb_encode_batch": (C.c_int, [vp, vp, sz, vp, i64, vp]),
        "pb_decode_batch": (C.c_int, [vp, vp, i64, vp]),
        "pb_text2bin": (C.c_int, [vp, vp, sz, vp, sz, P(sz)]),
        "pb_bin2text": (C.c_int, [vp, vp, vp, sz, P(sz)]),
        "pb_seed_at_batch": (C.c_int, [vp, vp, sz, vp, i64, C.c_int, vp]),
        "pb_parse_pattern": (u32, [C.c_char_p]),
        "pb_seqset_from_text": (C.c_int, [vp, vp, vp, vp, vp, i64, P(vp)]),
        "pb_seqset_from_device_text": (C.c_int, [vp, vp, sz, vp, vp, vp, i64, P(vp)]),
        "pb_seqset_from_bin": (C.c_int, [vp, vp, sz, C.c_int, C.c_int, P(vp)]),
        "pb_seqset_free": (None, [vp]),
        "pb_seqset_count": (i64, [vp]),
        "pb_seqset_length": (i32, [vp, i64]),
        "pb_seqset_text": (C.c_int, [vp, vp, i64, vp, sz]),
        "pb_seqset_packed": (C.c_int, [vp, vp, i64, vp, sz]),
        "pb_seed_extract": (C.c_int, [vp, vp, i64, u32, vp]),
        "pb_seed_extract_all_device": (C.c_int, [vp, vp, u32, P(i64), P(C.c_float)]),
        "pb_probe_bulk_device": (C.c_int, [vp, vp, vp, P(i64), P(i64), P(C.c_float), P(C.c_float), P(C.c_float)]),
        "pb_index_build": (C.c_int, [vp, vp, i64, u32, C.c_int, P(vp)]),
        "pb_index_build_pairs": (C.c_int, [vp, vp, vp, i64, P(vp)]),
        "pb_index_free": (None, [vp]),
        "pb_index_nkeys": (i64, [vp]),
        "pb_index_nentries": (i64, [vp]),
        "pb_index_nscanned": (i64, [vp]),
        "pb_index_mask": (u32, [vp]),
        "pb_index_find_batch": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, i64]),
        "pb_align_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, C.c_double, C.c_int, C.c_int, vp, vp, vp]),
        "pb_align_weighted_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, C.c_double, C.c_double, C.c_int, C.c_int,
                                              vp, vp, vp]),
        "pb_locate_default_params": (None, [P(LocateParams)]),
        "pb_locate_batch": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, i64, P(LocateParams), vp, P(i64), vp, vp]),
        "pb_locate_run": (C.c_int, [vp, vp, vp, i64, vp, P(LocateParams), vp, P(vp)]),
        "pb_locate_job_nkept": (i64, [vp]),
        "pb_locate_job_ncand": (i64, [vp]),
        "pb_locate_job_ops_layout": (C.c_int, [vp, vp, P(i64)]),
        "pb_locate_job_stats": (C.c_int, [vp, vp]),
        "pb_locate_job_votes": (C.c_int, [vp, vp, vp, vp]),
        "pb_locate_fetch": (C.c_int, [vp, vp, vp, vp]),
        "pb_locate_job_free": (None, [vp]),
        "pb_locate_submit": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, i64, P(LocateParams), P(vp)]),
        "pb_locate_submit_bin": (C.c_int, [vp, vp, vp, i64, vp, sz, C.c_int, C.c_int, P(LocateParams), P(vp)]),
        "pb_locate_submit_device": (C.c_int, [vp, vp, vp, i64, vp, sz, vp, vp, i64, P(LocateParams), P(vp)]),
        "pb_locate_step_nkept": (i64, [vp]),
        "pb_locate_step_ops_extent": (i64, [vp]),
        "pb_locate_collect": (C.c_int, [vp, vp, vp, P(i64), vp, vp]),
        "pb_locate_step_free": (None, [vp]),
        "pb_overlap_default_params": (None, [P(OverlapParams)]),
        "pb_overlap_batch": (C.c_int, [vp, vp, vp, i64, vp, P(OverlapParams), vp, vp, vp]),
        "pb_overlap_subset": (C.c_int, [vp, vp, vp, i64, vp, vp, i64, P(OverlapParams), vp, vp, vp]),
        "pb_consensus_create": (C.c_int, [vp, vp, i64, C.c_int, P(vp)]),
        "pb_consensus_free": (None, [vp]),
        "pb_consensus_length": (i64, [vp]),
        "pb_consensus_extent": (C.c_int, [vp, P(i64), P(i64)]),
        "pb_consensus_text": (C.c_int, [vp, vp, C.c_int, vp, sz]),
        "pb_consensus_seqset": (C.c_int, [vp, vp, C.c_int, P(vp)]),
        "pb_consensus_append": (C.c_int, [vp, vp, vp, i32]),
        "pb_consensus_prepend": (C.c_int, [vp, vp, vp, i32]),
        "pb_consensus_elect_batch": (C.c_int, [vp, vp, vp, vp, i64, vp, vp]),
        "pb_consensus_evolve": (C.c_int, [vp, vp]),
        "pb_consensus_votes": (C.c_int, [vp, vp, vp, i64]),
        "pb_index_build_set": (C.c_int, [vp, vp, u32, P(vp)]),
        "pb_overlap_all_run": (C.c_int, [vp, vp, vp, i64, i64, P(OverlapParams), P(vp)]),
        "pb_pairs_job_stats": (C.c_int, [vp, vp]),
        "pb_pairs_job_fetch": (C.c_int, [vp, vp, C.c_int, vp, i64, P(i64)]),
        "pb_pairs_job_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            if os.environ.get("PB_LIB"):
                continue  # an older experiment build may lack newer entry points
            raise
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def _u8(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def parse_pattern(pattern: str | bytes) -> int:
    """parse_pattern, spaced_seed.cpp:166-180 (pure string -> mask)."""
    if isinstance(pattern, str):
        pattern = pattern.encode()
    return int(lib().pb_parse_pattern(pattern))


def default_locate_params(**kw) -> LocateParams:
    p = LocateParams()
    lib().pb_locate_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Context:
    """One per process per GPU (pb_ctx)."""

    def __init__(self, device: int = 0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.pb_ctx_create(device, C.byref(h))
        if rc != 0:
            raise PbError(rc, (self._L.pb_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self._children = weakref.WeakSet()  # handles must be released before the context they live in

    def close(self):
        if getattr(self, "h", None):
            for child in list(self._children):
                child.free()
            self._L.pb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise PbError(rc, (self._L.pb_last_error(self.h) or b"").decode())

    @property
    def stream(self) -> int:
        return int(self._L.pb_ctx_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self._L.pb_ctx_launch_count(self.h))

    def set_scratch_limit(self, nbytes: int):
        self.check(self._L.pb_ctx_set_scratch_limit(self.h, nbytes))

    def int_pipe_peak(self) -> float:
        """measured peak of the integer ALU pipe, warp instructions per second (pb_int_pipe_peak)"""
        v = C.c_double(0.0)
        self.check(self._L.pb_int_pipe_peak(self.h, C.byref(v)))
        return float(v.value)

    def random_gather_peak(self, table_bytes: int = 0) -> float:
        """measured peak of independent random 4-byte reads of a table (default 64 MB), reads per second"""
        v = C.c_double(0.0)
        self.check(self._L.pb_random_gather_peak(self.h, table_bytes, C.byref(v)))
        return float(v.value)

    def timings(self) -> dict:
        ms = (C.c_float * len(T_NAMES))()
        self.check(self._L.pb_ctx_timings(self.h, ms))
        return {n: float(ms[i]) for i, n in enumerate(T_NAMES)}

    # ---- dna_seq statics -------------------------------------------------------------------
    def encode(self, text, offsets=None) -> np.ndarray:
        t = _u8(text)
        off = np.ascontiguousarray([0] if offsets is None else offsets, dtype=np.int64)
        out = np.zeros(len(off), dtype=np.uint32)
        self.check(self._L.pb_encode_batch(self.h, _ptr(t), len(t), _ptr(off), len(off), _ptr(out)))
        return out

    def decode(self, codes) -> list[bytes]:
        c = np.ascontiguousarray(codes, dtype=np.uint32).reshape(-1)
        out = np.zeros(16 * len(c), dtype=np.uint8)
        self.check(self._L.pb_decode_batch(self.h, _ptr(c), len(c), _ptr(out)))
        return [out[16 * i:16 * i + 16].tobytes() for i in range(len(c))]

    def text2bin(self, text: bytes, cap: int | None = None) -> bytes:
        t = _u8(text)
        need = 4 + (len(t) + 3) // 4
        cap = need if cap is None else cap
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        w = C.c_size_t(0)
        self.check(self._L.pb_text2bin(self.h, _ptr(t) if len(t) else None, len(t), _ptr(out), cap, C.byref(w)))
        return out[: w.value].tobytes()

    def bin2text(self, rec: bytes, cap: int | None = None) -> bytes:
        r = _u8(rec)
        n = int.from_bytes(bytes(r[:4]), "little")
        cap = n + 1 if cap is None else cap
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        tl = C.c_size_t(0)
        self.check(self._L.pb_bin2text(self.h, _ptr(r), _ptr(out), cap, C.byref(tl)))
        return out[: tl.value].tobytes()

    def seed_at(self, rec: bytes, positions, quirk: bool = False) -> np.ndarray:
        r = _u8(rec)
        pos = np.ascontiguousarray(positions, dtype=np.int32).reshape(-1)
        out = np.zeros(len(pos), dtype=np.uint32)
        self.check(self._L.pb_seed_at_batch(self.h, _ptr(r), len(r), _ptr(pos), len(pos), int(quirk), _ptr(out)))
        return out

    # ---- sequences ------------------------------------------------------------------------------
    def seqset(self, text, offs, lens, strides=None) -> "SeqSet":
        t = _u8(text)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        st = None if strides is None else np.ascontiguousarray(strides, dtype=np.int32)
        h = C.c_void_p()
        self.check(self._L.pb_seqset_from_text(self.h, _ptr(t), _ptr(offs), _ptr(lens), _ptr(st), len(lens), C.byref(h)))
        return SeqSet(self, h)

    def seqset_one(self, text) -> "SeqSet":
        t = _u8(text)
        return self.seqset(t, [0], [len(t)])

    def seqset_from_device(self, dptr: int, nbytes: int, offs, lens, strides=None) -> "SeqSet":
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        st = None if strides is None else np.ascontiguousarray(strides, dtype=np.int32)
        h = C.c_void_p()
        self.check(self._L.pb_seqset_from_device_text(self.h, dptr, nbytes, _ptr(offs), _ptr(lens), _ptr(st), len(lens),
                                                      C.byref(h)))
        return SeqSet(self, h)

    def seqset_from_bin(self, image: bytes, min_excl: int = 500, max_excl: int = 20000) -> "SeqSet":
        b = _u8(image)
        h = C.c_void_p()
        self.check(self._L.pb_seqset_from_bin(self.h, _ptr(b), len(b), min_excl, max_excl, C.byref(h)))
        return SeqSet(self, h)

    # ---- index ---------------------------------------------------------------------------------
    def index(self, ref: "SeqSet", mask: int, policy: int = POLICY_LOCATOR, seq: int = 0) -> "Index":
        h = C.c_void_p()
        self.check(self._L.pb_index_build(self.h, ref.h, seq, mask, policy, C.byref(h)))
        return Index(self, h, ref, seq)

    def index_from_pairs(self, keys, pos) -> "Index":
        """the seed map as seedmap[key].push_back(pos) fills it (locator.cpp:65): explicit (key, position) pairs in
        insertion order; serves find() only (what hash_table::operator[] hands over)"""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        assert len(keys) == len(pos)
        h = C.c_void_p()
        self.check(self._L.pb_index_build_pairs(self.h, _ptr(keys), _ptr(pos), len(keys), C.byref(h)))
        return Index(self, h, None, 0)

    # ---- aligner -------------------------------------------------------------------------------
    def align_batch(self, a_text, a_off, a_len, b_text, b_off, b_len, R: float = 0.3, maxn: int = 26000, maxm: int = 6000,
                    a_stride=None, b_stride=None, want_ops: bool = True):
        """n seq_aligner::align calls; returns (records[ALIGN_DTYPE], list of op arrays or None)."""
        at, bt = _u8(a_text), _u8(b_text)
        a_off = np.ascontiguousarray(a_off, dtype=np.int64)
        b_off = np.ascontiguousarray(b_off, dtype=np.int64)
        a_len = np.ascontiguousarray(a_len, dtype=np.int32)
        b_len = np.ascontiguousarray(b_len, dtype=np.int32)
        ast = None if a_stride is None else np.ascontiguousarray(a_stride, dtype=np.int32)
        bst = None if b_stride is None else np.ascontiguousarray(b_stride, dtype=np.int32)
        n = len(a_len)
        out = np.zeros(n, dtype=ALIGN_DTYPE)
        ops = ops_off = None
        if want_ops:
            slots = (a_len.astype(np.int64) + b_len + 1 + 15) & ~15
            ops_off = np.zeros(n, dtype=np.int64)
            if n:
                np.cumsum(slots[:-1], out=ops_off[1:])
            ops = np.zeros(int(slots.sum()) + 16, dtype=np.uint8)
        self.check(self._L.pb_align_batch(self.h, _ptr(at), _ptr(a_off), _ptr(a_len), _ptr(ast), _ptr(bt), _ptr(b_off),
                                          _ptr(b_len), _ptr(bst), n, R, maxn, maxm, _ptr(out), _ptr(ops), _ptr(ops_off)))
        if want_ops:
            return out, [ops[ops_off[i]: ops_off[i] + max(int(out["nedit"][i]), 0)] if out["ret"][i] >= 0 else None
                         for i in range(n)]
        return out, None

    def align_weighted_batch(self, a_text, a_w, a_off, a_len, b_text, b_w, b_off, b_len, R: float = 0.3,
                             fail_scale: float = 1.0, maxn: int = 26000, maxm: int = 6000, want_ops: bool = True):
        """The quality-weighted EXTENSION (pb_align_weighted_batch): per-element weights 1..4 for a and b, forward views."""
        at, bt, aw, bw = _u8(a_text), _u8(b_text), _u8(a_w), _u8(b_w)
        assert len(aw) >= len(at) and len(bw) >= len(bt)
        a_off = np.ascontiguousarray(a_off, dtype=np.int64)
        b_off = np.ascontiguousarray(b_off, dtype=np.int64)
        a_len = np.ascontiguousarray(a_len, dtype=np.int32)
        b_len = np.ascontiguousarray(b_len, dtype=np.int32)
        n = len(a_len)
        out = np.zeros(n, dtype=ALIGN_DTYPE)
        ops = ops_off = None
        if want_ops:
            slots = (a_len.astype(np.int64) + b_len + 1 + 15) & ~15
            ops_off = np.zeros(n, dtype=np.int64)
            if n:
                np.cumsum(slots[:-1], out=ops_off[1:])
            ops = np.zeros(int(slots.sum()) + 16, dtype=np.uint8)
        self.check(self._L.pb_align_weighted_batch(self.h, _ptr(at), _ptr(aw), _ptr(a_off), _ptr(a_len), _ptr(bt), _ptr(bw),
                                                   _ptr(b_off), _ptr(b_len), n, R, fail_scale, maxn, maxm, _ptr(out), _ptr(ops),
                                                   _ptr(ops_off)))
        if want_ops:
            return out, [ops[ops_off[i]: ops_off[i] + max(int(out["nedit"][i]), 0)] if out["ret"][i] >= 0 else None
                         for i in range(n)]
        return out, None

    def align(self, a: bytes, b: bytes, R: float = 0.3, a_fwd: bool = True, b_fwd: bool = True, maxn: int = 26000,
              maxm: int = 6000):
        """One seq_aligner::align(seg_a, seg_b) (batch of one).  a/b are the underlying texts; a backward accessor
        starts at the last byte and walks down."""
        a, b = _u8(a), _u8(b)
        rec, ops = self.align_batch(a, [0 if a_fwd else max(len(a) - 1, 0)], [len(a)], b,
                                    [0 if b_fwd else max(len(b) - 1, 0)], [len(b)], R, maxn, maxm,
                                    [1 if a_fwd else -1], [1 if b_fwd else -1])
        d = {k: int(rec[k][0]) for k in ALIGN_DTYPE.names}
        if d["ret"] >= 0:
            d["ops"] = ops[0].copy()
            bb = b if b_fwd else b[::-1]
            vals, j = np.zeros(len(d["ops"]), dtype=np.uint8), 0
            for k, op in enumerate(d["ops"]):  # edit.val = seg_b->at(j-1) under MATCH/INSERT (seq_aligner.h:219,225)
                if op != DELETE:
                    vals[k] = bb[j]
                    j += 1
            d["vals"] = vals
        return d

    # ---- locate ----------------------------------------------------------------------------------
    def locate(self, index: "Index", reads_text, offs, lens, want_ops: bool = False, **params):
        """locator.cpp:70-92 for a batch of reads given as host text.  Returns records (and transcripts)."""
        prm = default_locate_params(want_ops=int(want_ops), **params)
        t = _u8(reads_text)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        kept = lens >= prm.minlen
        nk = int(kept.sum())
        recs = np.zeros(nk, dtype=LOCATE_DTYPE)
        ops = ops_off = None
        if want_ops:
            slots = (2 * lens[kept].astype(np.int64) + prm.maxm + 16 + 15) & ~15
            ops_off = np.zeros(nk, dtype=np.int64)
            if nk:
                np.cumsum(slots[:-1], out=ops_off[1:])
            ops = np.zeros(int(slots.sum()) + 16, dtype=np.uint8)
        got = C.c_int64(0)
        self.check(self._L.pb_locate_batch(self.h, index.h, index.ref.h, index.seq, _ptr(t), _ptr(offs), _ptr(lens),
                                           len(lens), C.byref(prm), _ptr(recs), C.byref(got), _ptr(ops), _ptr(ops_off)))
        assert got.value == nk
        if want_ops:
            return recs, [ops[ops_off[k]: ops_off[k] + int(recs["nedit"][k])] for k in range(nk)]
        return recs

    def locate_submit(self, index: "Index", reads_text, offs, lens, **params) -> "LocateStep":
        """Pipelined locate (pb_locate_submit): queue one batch of host text; collect() it after submitting the next one."""
        prm = default_locate_params(**params)
        t = _u8(reads_text)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        h = C.c_void_p()
        self.check(self._L.pb_locate_submit(self.h, index.h, index.ref.h, index.seq, _ptr(t), _ptr(offs), _ptr(lens), len(lens),
                                            C.byref(prm), C.byref(h)))
        return LocateStep(self, h, keep=(t, offs, lens))

    def locate_submit_device(self, index: "Index", dptr: int, nbytes: int, offs, lens, **params) -> "LocateStep":
        """Pipelined locate of a batch whose text already lies in device memory (dptr, e.g. a torch tensor's data_ptr):
        no staging copy; the caller keeps the buffer alive until the step is collected."""
        prm = default_locate_params(**params)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        h = C.c_void_p()
        self.check(self._L.pb_locate_submit_device(self.h, index.h, index.ref.h, index.seq, C.c_void_p(dptr), nbytes, _ptr(offs),
                                                   _ptr(lens), len(lens), C.byref(prm), C.byref(h)))
        return LocateStep(self, h, keep=(offs, lens))

    def locate_submit_bin(self, index: "Index", image, **params) -> "LocateStep":
        """Pipelined locate of a batch given as a .bin image (binary_test.cpp:55-63); records of len >= minlen are kept."""
        prm = default_locate_params(**params)
        b = _u8(image)
        h = C.c_void_p()
        self.check(self._L.pb_locate_submit_bin(self.h, index.h, index.ref.h, index.seq, _ptr(b), b.size, prm.minlen - 1,
                                                2 ** 31 - 1, C.byref(prm), C.byref(h)))
        return LocateStep(self, h, keep=(b,))

    def overlap(self, index: "Index", reads: "SeqSet", want_ops: bool = False, ref: "SeqSet | None" = None, ids=None, **params):
        """spaced_seed.cpp:424-436 / try_align: head and tail trials of every read against a REFSEQ-policy index.
        ref (with ref_shift=beg-pre): the whole text of a reference that has grown beyond the indexed [beg, end).
        ids: only these reads of the set, in this order (pb_overlap_subset)."""
        prm = OverlapParams()
        self._L.pb_overlap_default_params(C.byref(prm))
        prm.want_ops = 1 if want_ops else 0
        for k, v in params.items():
            setattr(prm, k, v)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
        n = len(reads) if ids is None else len(ids)
        recs = np.zeros(n, dtype=OVERLAP_DTYPE)
        ops = ops_off = None
        if want_ops:
            lens = np.array([reads.length(int(i)) for i in (range(n) if ids is None else ids)], dtype=np.int64)
            slots = (3 * lens + 2 * prm.maxm + 16 + 15) & ~15
            ops_off = np.zeros(n, dtype=np.int64)
            if n:
                np.cumsum(slots[:-1], out=ops_off[1:])
            ops = np.zeros(int(slots.sum()) + 16, dtype=np.uint8)
        self.check(self._L.pb_overlap_subset(self.h, index.h, (index.ref if ref is None else ref).h, index.seq if ref is None else 0,
                                             reads.h, _ptr(ids), n, C.byref(prm), _ptr(recs), _ptr(ops), _ptr(ops_off)))
        if want_ops == "raw":  # the buffers as pb_consensus_elect_batch takes them
            return recs, ops, ops_off
        if want_ops:
            return recs, [ops[ops_off[k]: ops_off[k] + int(recs["nedit"][k])] for k in range(n)]
        return recs

    # ---- consensus voting ---------------------------------------------------------------------------
    def consensus(self, text, weight: int = 1) -> "Consensus":
        t = _u8(text)
        h = C.c_void_p()
        self.check(self._L.pb_consensus_create(self.h, _ptr(t) if len(t) else None, len(t), weight, C.byref(h)))
        return Consensus(self, h)

    # ---- all-vs-all ------------------------------------------------------------------------------
    def index_set(self, seqs: "SeqSet", mask: int) -> "Index":
        """one seed index over every sequence of the set (get_seedmap's head pass per sequence, ref_seq.h:291-300)"""
        h = C.c_void_p()
        self.check(self._L.pb_index_build_set(self.h, seqs.h, mask, C.byref(h)))
        return Index(self, h, seqs, -1)

    def overlap_all(self, index: "Index", q_first: int = 0, q_count: int | None = None, found_only: bool = True, **params):
        """the trial loop of spaced_seed.cpp:424-436 with every read of the set as the locked reference in turn;
        returns (records[PAIR_DTYPE], stats dict)"""
        prm = OverlapParams()
        self._L.pb_overlap_default_params(C.byref(prm))
        for k, v in params.items():
            setattr(prm, k, v)
        seqs = index.ref
        if q_count is None:
            q_count = len(seqs) - q_first
        h = C.c_void_p()
        self.check(self._L.pb_overlap_all_run(self.h, index.h, seqs.h, q_first, q_count, C.byref(prm), C.byref(h)))
        try:
            st = (C.c_int64 * 8)()
            self.check(self._L.pb_pairs_job_stats(h, st))
            stats = dict(zip(PAIR_STATS, (int(x) for x in st)))
            n = C.c_int64(0)
            self.check(self._L.pb_pairs_job_fetch(self.h, h, int(found_only), None, 0, C.byref(n)))
            recs = np.zeros(n.value, dtype=PAIR_DTYPE)
            self.check(self._L.pb_pairs_job_fetch(self.h, h, int(found_only), _ptr(recs), n.value, C.byref(n)))
        finally:
            self._L.pb_pairs_job_free(h)
        return recs, stats

    def locate_run(self, index: "Index", reads: "SeqSet", want_ops: bool = False, **params) -> "LocateJob":
        prm = default_locate_params(want_ops=int(want_ops), **params)
        h = C.c_void_p()
        self.check(self._L.pb_locate_run(self.h, index.h, index.ref.h, index.seq, reads.h, C.byref(prm), None, C.byref(h)))
        return LocateJob(self, h)


class SeqSet:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        ctx._children.add(self)

    def free(self):
        if self.h and self.ctx.h:
            self.ctx._L.pb_seqset_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return int(self.ctx._L.pb_seqset_count(self.h))

    def length(self, i: int) -> int:
        return int(self.ctx._L.pb_seqset_length(self.h, i))

    def text(self, i: int) -> bytes:
        n = self.length(i)
        out = np.zeros(n + 1, dtype=np.uint8)
        self.ctx.check(self.ctx._L.pb_seqset_text(self.ctx.h, self.h, i, _ptr(out), n + 1))
        return out[:n].tobytes()

    def packed(self, i: int) -> bytes:
        n = (self.length(i) + 3) // 4
        out = np.zeros(max(n, 1), dtype=np.uint8)
        self.ctx.check(self.ctx._L.pb_seqset_packed(self.ctx.h, self.h, i, _ptr(out), n))
        return out[:n].tobytes()

    def seeds(self, i: int, mask: int) -> np.ndarray:
        """K1 bulk: encode(text+p) & mask for every p of sequence i."""
        out = np.zeros(self.length(i), dtype=np.uint32)
        self.ctx.check(self.ctx._L.pb_seed_extract(self.ctx.h, self.h, i, mask, _ptr(out)))
        return out

    def seeds_device(self, mask: int) -> tuple[int, float]:
        n, ms = C.c_int64(0), C.c_float(0)
        self.ctx.check(self.ctx._L.pb_seed_extract_all_device(self.ctx.h, self.h, mask, C.byref(n), C.byref(ms)))
        return n.value, ms.value


class Consensus:
    """ref_seq's voting state on the device (ref_seq.h:47-183, 317-362): text [pre, post) + one vote box per base."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        ctx._children.add(self)

    def free(self):
        if self.h and self.ctx.h:
            self.ctx._L.pb_consensus_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.ctx._L.pb_consensus_length(self.h))

    def extent(self) -> tuple[int, int]:
        """(beg - pre, post - pre)"""
        a, b = C.c_int64(0), C.c_int64(0)
        self.ctx.check(self.ctx._L.pb_consensus_extent(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def text(self, full: bool = False) -> bytes:
        n = self.extent()[1] if full else len(self)
        out = np.zeros(n + 1, dtype=np.uint8)
        self.ctx.check(self.ctx._L.pb_consensus_text(self.ctx.h, self.h, int(full), _ptr(out), n + 1))
        return out[:n].tobytes()

    def seqset(self, full: bool = False) -> "SeqSet":
        h = C.c_void_p()
        self.ctx.check(self.ctx._L.pb_consensus_seqset(self.ctx.h, self.h, int(full), C.byref(h)))
        return SeqSet(self.ctx, h)

    def append(self, seg):
        t = _u8(seg)
        self.ctx.check(self.ctx._L.pb_consensus_append(self.ctx.h, self.h, _ptr(t) if len(t) else None, len(t)))

    def prepend(self, seg):
        t = _u8(seg)
        self.ctx.check(self.ctx._L.pb_consensus_prepend(self.ctx.h, self.h, _ptr(t) if len(t) else None, len(t)))

    def elect(self, reads: "SeqSet", recs: np.ndarray, ops: np.ndarray, ops_off: np.ndarray):
        recs = np.ascontiguousarray(recs)
        ops_off = np.ascontiguousarray(ops_off, dtype=np.int64)
        assert recs.dtype == OVERLAP_DTYPE and len(ops_off) == len(recs)
        self.ctx.check(self.ctx._L.pb_consensus_elect_batch(self.ctx.h, self.h, reads.h, _ptr(recs), len(recs), _ptr(ops), _ptr(ops_off)))

    def evolve(self):
        self.ctx.check(self.ctx._L.pb_consensus_evolve(self.ctx.h, self.h))

    def votes(self) -> np.ndarray:
        n = self.extent()[1]
        out = np.zeros((max(n, 1), 9), dtype=np.int32)
        self.ctx.check(self.ctx._L.pb_consensus_votes(self.ctx.h, self.h, _ptr(out), max(n, 1)))
        return out[:n]


class Index:
    """Device seed index; find() is hash_table::find (common.h:54)."""

    def __init__(self, ctx: Context, h, ref: SeqSet, seq: int):
        self.ctx, self.h, self.ref, self.seq = ctx, h, ref, seq
        ctx._children.add(self)

    def free(self):
        if self.h and self.ctx.h:
            self.ctx._L.pb_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def nkeys(self) -> int:
        return int(self.ctx._L.pb_index_nkeys(self.h))

    @property
    def nentries(self) -> int:
        return int(self.ctx._L.pb_index_nentries(self.h))

    @property
    def nscanned(self) -> int:
        return int(self.ctx._L.pb_index_nscanned(self.h))

    @property
    def mask(self) -> int:
        return int(self.ctx._L.pb_index_mask(self.h))

    def probe_bulk(self, s: "SeqSet") -> dict:
        """every position of `s` probed against this index on the device (K1 bulk + K2 count/gather); sizes and kernel times"""
        nq, nc = C.c_int64(0), C.c_int64(0)
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self.ctx.check(self.ctx._L.pb_probe_bulk_device(self.ctx.h, self.h, s.h, C.byref(nq), C.byref(nc), C.byref(a), C.byref(b),
                                                        C.byref(c)))
        return {"queries": nq.value, "candidates": nc.value, "seed_ms": a.value, "count_ms": b.value, "gather_ms": c.value}

    def find_batch(self, keys) -> list[list[int]]:
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1)
        n = len(keys)
        cnt = np.zeros(n, dtype=np.int64)
        self.ctx.check(self.ctx._L.pb_index_find_batch(self.ctx.h, self.h, _ptr(keys), n, _ptr(cnt), None, None, 0))
        off = np.zeros(n, dtype=np.int64)
        if n:
            np.cumsum(cnt[:-1], out=off[1:])
        pos = np.zeros(int(cnt.sum()) + 1, dtype=np.int32)
        cap = int(cnt.max()) if n else 0
        self.ctx.check(self.ctx._L.pb_index_find_batch(self.ctx.h, self.h, _ptr(keys), n, _ptr(cnt), _ptr(pos), _ptr(off), cap))
        return [pos[off[i]: off[i] + cnt[i]].tolist() for i in range(n)]

    def find(self, key: int) -> list[int]:
        return self.find_batch([key])[0]


class LocateStep:
    """One batch in flight through pb_locate_submit; collect() waits for it, returns its records and frees it."""

    def __init__(self, ctx: Context, h, keep=()):
        self.ctx, self.h, self._keep = ctx, h, keep  # the host buffers must outlive the asynchronous copy
        self.stats = None

    @property
    def nkept(self) -> int:
        return int(self.ctx._L.pb_locate_step_nkept(self.h))

    def collect(self, recs: np.ndarray | None = None) -> np.ndarray:
        nk = self.nkept
        if recs is None:
            recs = np.zeros(nk, dtype=LOCATE_DTYPE)
        out = np.zeros(8, dtype=np.int64)
        got = C.c_int64(0)
        h, self.h = self.h, None
        self.ctx.check(self.ctx._L.pb_locate_collect(self.ctx.h, h, _ptr(recs), C.byref(got), None, _ptr(out)))
        self.stats = {"ncand": int(out[0]), "dp_alignments": int(out[1]), "dp_cells": int(out[2]), "band_cells": int(out[3]),
                      "redone": int(out[4]), "alu_instr": int(out[5]), "tb_rounds": int(out[6]), "tb_cold": int(out[7])}
        self._keep = ()
        return recs[: got.value]

    def free(self):
        if self.h and self.ctx.h:
            self.ctx._L.pb_locate_step_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class LocateJob:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        ctx._children.add(self)

    def free(self):
        if self.h and self.ctx.h:
            self.ctx._L.pb_locate_job_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def nkept(self) -> int:
        return int(self.ctx._L.pb_locate_job_nkept(self.h))

    @property
    def ncand(self) -> int:
        return int(self.ctx._L.pb_locate_job_ncand(self.h))

    def stats(self) -> dict:
        """valid after fetch(): candidates gathered, alignments run by K3, DP cells computed by K3"""
        out = np.zeros(8, dtype=np.int64)
        self.ctx.check(self.ctx._L.pb_locate_job_stats(self.h, _ptr(out)))
        return {"ncand": int(out[0]), "dp_alignments": int(out[1]), "dp_cells": int(out[2]), "band_cells": int(out[3]),
                "redone": int(out[4]), "alu_instr": int(out[5]), "tb_rounds": int(out[6]), "tb_cold": int(out[7])}

    def votes(self):
        """(votes, best_diag) per kept read: the diagonal-bin tally of its seed hits (diagnostic only)"""
        v = np.zeros(max(self.nkept, 1), dtype=np.int32)
        d = np.zeros(max(self.nkept, 1), dtype=np.int32)
        self.ctx.check(self.ctx._L.pb_locate_job_votes(self.ctx.h, self.h, _ptr(v), _ptr(d)))
        return v[: self.nkept], d[: self.nkept]

    def ops_layout(self):
        off = np.zeros(max(self.nkept, 1), dtype=np.int64)
        ext = C.c_int64(0)
        self.ctx.check(self.ctx._L.pb_locate_job_ops_layout(self.h, _ptr(off), C.byref(ext)))
        return off[: self.nkept], ext.value

    def fetch(self, want_ops: bool = False, recs: np.ndarray | None = None, ops: np.ndarray | None = None):
        nk = self.nkept
        if recs is None:
            recs = np.zeros(nk, dtype=LOCATE_DTYPE)
        if want_ops:
            off, ext = self.ops_layout()
            if ops is None:
                ops = np.zeros(ext + 16, dtype=np.uint8)
            self.ctx.check(self.ctx._L.pb_locate_fetch(self.ctx.h, self.h, _ptr(recs), _ptr(ops)))
            return recs, [ops[off[k]: off[k] + int(recs["nedit"][k])] for k in range(nk)]
        self.ctx.check(self.ctx._L.pb_locate_fetch(self.ctx.h, self.h, _ptr(recs), None))
        return recs
