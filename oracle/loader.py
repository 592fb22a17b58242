"""ctypes loader for oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm.
The product (goleft_b200/, cli/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

if not os.path.exists(LIB_PATH):
    subprocess.check_call(["make", "-s", "oracle"], cwd=_ROOT)
lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_i64p = C.POINTER(C.c_int64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _proto(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_proto("orc_free", None, _vp)
_proto("orc_pileup_brute", C.c_int64, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, _vp)
_proto("orc_pileup_diff", C.c_int64, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, _vp)
_proto("orc_chrom_start_end", C.c_int, C.c_char_p, C.c_char_p, _i64p, _i64p)
_proto("orc_gen_chunks", C.c_int64, C.c_int64, C.c_int64, _vp, _vp, C.c_int64)
_proto("orc_walk_chunk", C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp,
       C.POINTER(_vp), _i64p, C.POINTER(_vp), _i64p)
_proto("orc_samtools_text", C.c_int, C.c_char_p, C.c_int64, C.c_int64, _vp, C.POINTER(_vp), _i64p)
_proto("orc_walk_text", C.c_int, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_vp), _i64p,
       C.POINTER(_vp), _i64p)
_proto("orc_window_sums", C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, C.c_int64)
_proto("orc_class_runs", C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, C.c_int64)


def pileup_brute(start, end, rs: int, re: int) -> np.ndarray:
    start = np.ascontiguousarray(start, np.int32)
    end = np.ascontiguousarray(end, np.int32)
    out = np.empty(re - rs, np.int32)
    lib.orc_pileup_brute(_ptr(start), _ptr(end), start.size, rs, re, _ptr(out))
    return out


def pileup_diff(start, end, rs: int, re: int) -> np.ndarray:
    start = np.ascontiguousarray(start, np.int32)
    end = np.ascontiguousarray(end, np.int32)
    out = np.empty(re - rs, np.int32)
    lib.orc_pileup_diff(_ptr(start), _ptr(end), start.size, rs, re, _ptr(out))
    return out


def chrom_start_end(line: str) -> Tuple[str, int, int]:
    b = line.encode()
    chrom = C.create_string_buffer(len(b) + 1)
    s, e = C.c_int64(0), C.c_int64(0)
    if lib.orc_chrom_start_end(b, chrom, C.byref(s), C.byref(e)) != 0:
        raise ValueError(f"couldn't get region from line {line!r}")
    return chrom.value.decode(), s.value, e.value


def gen_chunks(length: int, W: int):
    n = lib.orc_gen_chunks(length, W, None, None, 0)
    s = np.empty(n, np.int64)
    e = np.empty(n, np.int64)
    lib.orc_gen_chunks(length, W, _ptr(s), _ptr(e), n)
    return list(zip(s.tolist(), e.tolist()))


def _take(p: _vp, n: C.c_int64) -> bytes:
    try:
        return C.string_at(p, n.value)
    finally:
        lib.orc_free(p)


def walk_chunk(chrom: str, rs: int, re: int, W: int, mincov: int, maxmean: int, depth: np.ndarray) -> Tuple[bytes, bytes]:
    """(depth.bed rows, callable.bed rows) of one chunk from its per-base depth array."""
    depth = np.ascontiguousarray(depth, np.int32)
    assert depth.size == re - rs
    d, c = _vp(), _vp()
    dl, cl = C.c_int64(0), C.c_int64(0)
    rc = lib.orc_walk_chunk(chrom.encode(), rs, re, W, mincov, maxmean, _ptr(depth), C.byref(d), C.byref(dl),
                            C.byref(c), C.byref(cl))
    if rc != 0:
        raise ValueError("orc_walk_chunk failed")
    return _take(d, dl), _take(c, cl)


def samtools_text(chrom: str, rs: int, re: int, depth: np.ndarray) -> bytes:
    depth = np.ascontiguousarray(depth, np.int32)
    p, n = _vp(), C.c_int64(0)
    lib.orc_samtools_text(chrom.encode(), rs, re, _ptr(depth), C.byref(p), C.byref(n))
    return _take(p, n)


def walk_text(txt: bytes, W: int, mincov: int, maxmean: int) -> Tuple[bytes, bytes]:
    d, c = _vp(), _vp()
    dl, cl = C.c_int64(0), C.c_int64(0)
    buf = C.create_string_buffer(txt, len(txt))
    rc = lib.orc_walk_text(C.cast(buf, _vp), len(txt), W, mincov, maxmean, C.byref(d), C.byref(dl), C.byref(c),
                           C.byref(cl))
    if rc != 0:
        raise ValueError("orc_walk_text failed")
    return _take(d, dl), _take(c, cl)


def window_sums(depth: np.ndarray, rs: int, re: int, W: int):
    depth = np.ascontiguousarray(depth, np.int32)
    n = (re - 1) // W - rs // W + 1
    s = np.empty(n, np.int64)
    m = np.empty(n, np.int32)
    k = lib.orc_window_sums(_ptr(depth), rs, re, W, _ptr(s), _ptr(m), n)
    assert k == n
    return s, m


def class_runs(depth: np.ndarray, rs: int, re: int, mincov: int, maxmean: int, run_break: int = 0):
    depth = np.ascontiguousarray(depth, np.int32)
    n = lib.orc_class_runs(_ptr(depth), rs, re, mincov, maxmean, run_break, None, None, 0)
    a = np.empty(n, np.int32)
    c = np.empty(n, np.uint8)
    lib.orc_class_runs(_ptr(depth), rs, re, mincov, maxmean, run_break, _ptr(a), _ptr(c), n)
    return a, c
