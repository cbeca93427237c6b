"""Deterministic synthetic workload generator (bench/test INPUTS only; see workload/synth.c).

Not on the product path and not part of the oracle.  The C source is compiled on demand with
gcc (``build()``); ``__graft_entry__.build()`` prebuilds it so the .so travels to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpbsynth.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm", "-lpthread"])
    return _SO


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.pbs_reference.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
        _lib.pbs_read_lengths.argtypes = [C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]
        _lib.pbs_reads.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        _lib.pbs_pack_bin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.pbs_pack_bin.restype = C.c_int64
        _lib.pbs_sweep_pair.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    return _lib


def reference(seed: int, length: int) -> np.ndarray:
    """iid-uniform ACGT reference as a uint8 array of ASCII codes."""
    out = np.empty(length, dtype=np.uint8)
    _L().pbs_reference(seed, length, out.ctypes.data)
    return out


def read_lengths(seed: int, nreads: int, mean: float = 5000.0, sigma_log: float = 0.5,
                 lo: int = 500, hi: int = 19999) -> np.ndarray:
    lens = np.empty(nreads, dtype=np.int32)
    _L().pbs_read_lengths(seed, nreads, mean, sigma_log, lo, hi, lens.ctypes.data)
    return lens


def reads(seed: int, ref: np.ndarray, lens: np.ndarray, p_ins: float = 0.09, p_del: float = 0.04,
          p_sub: float = 0.02, nthreads: int | None = None, out: np.ndarray | None = None):
    """Simulated CLR-like reads.  Returns (text uint8 blob, offs int64[n], lens int32[n], starts int64[n])."""
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n = len(lens)
    offs = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(lens[:-1], out=offs[1:])
    total = int(lens.sum())
    if out is None:
        out = np.empty(total, dtype=np.uint8)
    starts = np.empty(n, dtype=np.int64)
    if nthreads is None:
        nthreads = min(os.cpu_count() or 1, 32)
    _L().pbs_reads(seed, ref.ctypes.data, len(ref), n, lens.ctypes.data, offs.ctypes.data,
                   p_ins, p_del, p_sub, nthreads, out.ctypes.data, starts.ctypes.data)
    return out, offs, lens, starts


def pack_bin(text: np.ndarray, offs: np.ndarray, lens: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """The batch as a .bin image (binary_test.cpp:55-63): u32 length + ceil(len/4) packed bytes per read."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n = _L().pbs_pack_bin(text.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(lens), None)
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.size >= n
    _L().pbs_pack_bin(text.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(lens), out.ctypes.data)
    return out[:n]


def sweep_pair(seed: int, idx: int, alen: int, band: int, tail: int | None = None, nedits: int | None = None):
    """One DP-sweep pair (config 3): returns (a bytes, b bytes, R) with max_dst == band for len(a).
    tail = bases of b beyond a's length; the default band // 2 keeps the pair inside align's coverage test
    matlen_b >= len_b * (1 - R) (seq_aligner.h:114) at every point of the sweep -- with a long tail len_b = len_a + band and
    narrow bands on long reads (e.g. 5 kbp at band 32) can never pass it."""
    if tail is None:
        tail = band // 2
    blen = alen + tail
    a = np.empty(alen + 8, dtype=np.uint8)
    b = np.empty(blen, dtype=np.uint8)
    R = (band - 0.5) / alen
    if nedits is None:
        nedits = max(0, int(alen * R / 2) - 2)
    n = C.c_int(0)
    _L().pbs_sweep_pair(seed, idx, alen, blen, R, nedits, a.ctypes.data, b.ctypes.data, C.byref(n))
    return a[: n.value].tobytes(), b.tobytes(), R
