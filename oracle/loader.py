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
_proto("orc_faidx_stats", C.c_int, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp)
_proto("orc_stats_text", C.c_int, _vp, C.c_char_p, C.c_int64)
_proto("orc_depth_jobs_mt", C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, _i64p)


def faidx_stats(fasta: bytes, rec, start: int, end: int):
    """Faidx.Stats restatement: rec = (offset, length, line_bases, line_width) from the .fai; -> (GC, CpG, Masked) or None
    where faidx panics."""
    buf = np.frombuffer(fasta, np.uint8)
    out = np.zeros(3, np.float64)
    rc = lib.orc_faidx_stats(_ptr(buf), buf.size, rec[0], rec[1], rec[2], rec[3], start, end, _ptr(out))
    return None if rc < 0 else out


def stats_text(st3) -> bytes:
    b = C.create_string_buffer(128)
    n = lib.orc_stats_text(_ptr(np.ascontiguousarray(st3, np.float64)), b, 128)
    return b.raw[:n]


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


class _ChunkJob(C.Structure):
    _fields_ = [("chrom", C.c_char_p), ("start", _vp), ("end", _vp), ("n", C.c_int64), ("rs", C.c_int64), ("re", C.c_int64)]


def depth_jobs_mt(contigs, W: int, mincov: int, maxmean: int, threads: int, slack: int = 20000):
    """bench CPU arm: contigs = [(name, length, start_sorted int32, end int32)]; every 10 Mb chunk (gen_chunks) of every
    contig is one job (pileup + callback walk + BED text) on `threads` C workers.  -> (jobs done, text bytes)"""
    jobs, keep = [], []
    for name, L, s, e in contigs:
        s, e = np.ascontiguousarray(s, np.int32), np.ascontiguousarray(e, np.int32)
        nm = name.encode()
        keep.append((s, e, nm))
        for cs, ce in gen_chunks(L, W):
            jobs.append((ce - cs, nm, s, e, cs, ce))
    jobs.sort(key=lambda j: -j[0])                        # longest chunks first
    arr = (_ChunkJob * len(jobs))()
    for k, (_, nm, s, e, cs, ce) in enumerate(jobs):
        arr[k].chrom = nm; arr[k].start = s.ctypes.data; arr[k].end = e.ctypes.data; arr[k].n = s.size; arr[k].rs = cs; arr[k].re = ce
    tb = C.c_int64(0)
    done = lib.orc_depth_jobs_mt(arr, len(jobs), W, mincov, maxmean, slack, threads, C.byref(tb))
    assert done == len(jobs)
    return done, tb.value


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


# ------------------------------------------------------------------ indexcov / covstats / depthwed
_proto("orc_ic_sizes", C.c_int64, _vp, _vp, C.c_int32, _vp, _vp)
_proto("orc_ic_median", C.c_int64, _vp, C.c_int64)
_proto("orc_ic_normalize", None, _vp, C.c_int64, C.c_double, _vp)
_proto("orc_ic_counts", None, _vp, C.c_int64, _vp)
_proto("orc_ic_roc", None, _vp, _vp)
_proto("orc_ic_bins", None, _vp, C.c_int64, C.c_int64, _vp)
_proto("orc_ic_xnorm", None, _vp, _vp, C.c_int32, C.c_int32)
_proto("orc_ic_getcn", C.c_double, _vp, C.c_int64)
_proto("orc_cs_mean_std", None, _vp, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double))
_proto("orc_cs_mad_filter", C.c_int64, _vp, C.c_int64, C.c_int)
_proto("orc_cs_tail", C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_int32, _vp, _vp, C.c_int64)
_proto("orc_bincount", None, _vp, C.c_int64, C.c_int32, C.c_int32, _vp)
_proto("orc_depthwed", C.c_int64, _vp, C.c_int32, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64)


def ic_sizes(voff, ref_ptr):
    voff = np.ascontiguousarray(voff, np.uint64)
    ref_ptr = np.ascontiguousarray(ref_ptr, np.int64)
    n_refs = ref_ptr.size - 1
    sizes = np.empty(max(voff.size, 1), np.int64)
    size_ptr = np.zeros(n_refs + 1, np.int64)
    k = lib.orc_ic_sizes(_ptr(voff), _ptr(ref_ptr), n_refs, _ptr(sizes), _ptr(size_ptr))
    if k < 0:
        raise ValueError("expected positive change in vOffset")
    return sizes[:k], size_ptr


def ic_median(sizes) -> int:
    sizes = np.ascontiguousarray(sizes, np.int64)
    return int(lib.orc_ic_median(_ptr(sizes), sizes.size))


def ic_normalize(sizes, median: float) -> np.ndarray:
    sizes = np.ascontiguousarray(sizes, np.int64)
    out = np.empty(sizes.size, np.float32)
    lib.orc_ic_normalize(_ptr(sizes), sizes.size, float(median), _ptr(out))
    return out


def ic_counts(depth, counts=None) -> np.ndarray:
    depth = np.ascontiguousarray(depth, np.float32)
    if counts is None:
        counts = np.zeros(70, np.int32)
    lib.orc_ic_counts(_ptr(depth), depth.size, _ptr(counts))
    return counts


def ic_roc(counts) -> np.ndarray:
    counts = np.ascontiguousarray(counts, np.int32)
    roc = np.empty(70, np.float32)
    lib.orc_ic_roc(_ptr(counts), _ptr(roc))
    return roc


def ic_bins(depth, longest: int, out4=None) -> np.ndarray:
    depth = np.ascontiguousarray(depth, np.float32)
    if out4 is None:
        out4 = np.zeros(4, np.int64)
    lib.orc_ic_bins(_ptr(depth), depth.size, longest, _ptr(out4))
    return out4


def ic_xnorm(depths, lens) -> np.ndarray:
    d = np.ascontiguousarray(depths, np.float32).copy()
    lens = np.ascontiguousarray(lens, np.int32)
    lib.orc_ic_xnorm(_ptr(d), _ptr(lens), d.shape[0], d.shape[1])
    return d


def ic_getcn(depth) -> float:
    depth = np.ascontiguousarray(depth, np.float32)
    return float(lib.orc_ic_getcn(_ptr(depth), depth.size))


def bincount(v, lo: int, hi: int) -> np.ndarray:
    v = np.ascontiguousarray(v, np.int32)
    h = np.empty(hi - lo, np.uint64)
    lib.orc_bincount(_ptr(v), v.size, lo, hi, _ptr(h))
    return h


def cs_tail(insert, tmpl, max_read_len: int):
    insert = np.ascontiguousarray(insert, np.int32).copy()
    tmpl = np.ascontiguousarray(tmpl, np.int32).copy()
    out6 = np.zeros(6, np.float64)
    H = np.zeros(1 << 16, np.float64)
    hn = lib.orc_cs_tail(_ptr(insert), insert.size, _ptr(tmpl), tmpl.size, max_read_len, _ptr(out6), _ptr(H), H.size)
    return out6, H[:hn]


def depthwed(means, starts, ends, chrom_id, size: int):
    means = np.ascontiguousarray(means, np.float64)
    S, R = means.shape
    starts = np.ascontiguousarray(starts, np.int32)
    ends = np.ascontiguousarray(ends, np.int32)
    chrom_id = np.ascontiguousarray(chrom_id, np.int32)
    cap = max(R, 1)
    o_s, o_e, o_c = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32)
    out = np.empty((cap, S), np.int64)
    k = lib.orc_depthwed(_ptr(means), S, R, _ptr(starts), _ptr(ends), _ptr(chrom_id), size, _ptr(o_s), _ptr(o_e), _ptr(o_c),
                         _ptr(out), cap)
    return o_s[:k], o_e[:k], o_c[:k], out[:k]


_proto("orc_crai_sizes", C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, C.c_int64)


_proto("orc_indexsplit", C.c_int64, _vp, _vp, C.c_int, C.c_int, C.POINTER(C.c_char_p), _vp, C.c_int, _vp, _vp, _vp, C.c_int64, C.c_char_p, C.c_int64)


def indexsplit(sample_sizes, names, ref_lens, N: int, problems=()) -> bytes:
    """indexsplit.Split restatement.  sample_sizes: per sample a list (per reference) of int64 tile-size arrays."""
    S, R = len(sample_sizes), len(names)
    flat, ptr = [], np.zeros(S * (R + 1), np.int64)
    off = 0
    for s, per_ref in enumerate(sample_sizes):
        for r in range(R):
            ptr[s * (R + 1) + r] = off
            if r < len(per_ref):
                flat.append(np.asarray(per_ref[r], np.int64)); off += flat[-1].size
        ptr[s * (R + 1) + R] = off
    sizes = np.concatenate(flat) if flat else np.zeros(0, np.int64)
    arr = (C.c_char_p * R)(*[n.encode() for n in names])
    lens = np.asarray(ref_lens, np.int64)
    pr = np.asarray([p[0] for p in problems], np.int32)
    ps = np.asarray([p[1] for p in problems], np.int64)
    pe = np.asarray([p[2] for p in problems], np.int64)
    cap = 1 << 24
    out = C.create_string_buffer(cap)
    n = lib.orc_indexsplit(_ptr(sizes), _ptr(ptr), S, R, arr, _ptr(lens), N, _ptr(pr), _ptr(ps), _ptr(pe), len(problems), out, cap)
    if n < 0:
        raise RuntimeError("orc_indexsplit: output buffer too small")
    return out.raw[:n]


def crai_sizes(start, span, nbytes) -> np.ndarray:
    start = np.ascontiguousarray(start, np.int64)
    span = np.ascontiguousarray(span, np.int64)
    nbytes = np.ascontiguousarray(nbytes, np.int32)
    n = lib.orc_crai_sizes(_ptr(start), _ptr(span), _ptr(nbytes), start.size, None, 0)
    if n < 0:
        raise ValueError("tilewidth logic error")
    out = np.empty(n, np.int64)
    lib.orc_crai_sizes(_ptr(start), _ptr(span), _ptr(nbytes), start.size, _ptr(out), n)
    return out
