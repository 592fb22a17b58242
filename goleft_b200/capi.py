"""ctypes binding of libgoleft_b200.so (the C ABI in include/goleft_b200.h).

This is what tests/ and bench.py call; it adds nothing but argument marshalling.  There is no
CPU fallback: if the shared library is missing the import fails, and every call on a box
without a CUDA device raises GlError (GL_ECUDA).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GOLEFT_B200_LIB: another build of the same library (A/B runs); tests and bench use the default
LIB_PATH = os.environ.get("GOLEFT_B200_LIB") or os.path.join(_HERE, "libgoleft_b200.so")

GL_OK, GL_EINVAL, GL_ECUDA, GL_ENOMEM, GL_ESTATE, GL_ERANGE, GL_ENCCL = 0, -1, -2, -3, -4, -5, -6
CLASS_NAMES = ("NO_COVERAGE", "LOW_COVERAGE", "CALLABLE", "EXCESSIVE_COVERAGE")
INDEXCOV_SLOTS = 70


class GlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"goleft_b200 error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make lib` (or __graft_entry__.build()); "
            "goleft_b200 has no CPU fallback")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p


def _proto(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_proto("gl_version", C.c_char_p)
_proto("gl_device_count", C.c_int, C.POINTER(C.c_int))
_proto("gl_ctx_create", C.c_int, C.c_int, C.POINTER(_vp))
_proto("gl_ctx_destroy", C.c_int, _vp)
_proto("gl_last_error", C.c_char_p, _vp)
_proto("gl_sync", C.c_int, _vp)
_proto("gl_launch_count", C.c_int, _vp, _i64p)
_proto("gl_stream_handle", C.c_int, _vp, _u64p)
_proto("gl_dev_alloc", C.c_int, _vp, C.c_int64, C.POINTER(_vp))
_proto("gl_dev_free", C.c_int, _vp, _vp)
_proto("gl_host_alloc_pinned", C.c_int, _vp, C.c_int64, C.POINTER(_vp))
_proto("gl_host_free_pinned", C.c_int, _vp, _vp)
_proto("gl_memcpy_h2d", C.c_int, _vp, _vp, _vp, C.c_int64)
_proto("gl_memcpy_d2h", C.c_int, _vp, _vp, _vp, C.c_int64)
_proto("gl_timer_start", C.c_int, _vp)
_proto("gl_timer_stop_ms", C.c_int, _vp, C.POINTER(C.c_float))
_proto("gl_flush_l2", C.c_int, _vp)
_proto("gl_profile_enable", C.c_int, _vp, C.c_int)
_proto("gl_profile_read", C.c_int, _vp, _vp, C.c_int64, _i64p)

_proto("gl_bind_numa_for_device", C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int))
_proto("gl_device_numa_node", C.c_int, C.c_int, C.POINTER(C.c_int))
_proto("gl_lpt_assign", C.c_int, _vp, C.c_int32, C.c_int32, _vp, _vp)
_proto("gl_depth_begin", C.c_int, _vp, C.c_int64, C.c_int64)
_proto("gl_depth_add_segments", C.c_int, _vp, _vp, _vp, C.c_int64)
_proto("gl_depth_add_segments_device", C.c_int, _vp, _vp, _vp, C.c_int64)
_proto("gl_depth_reduce", C.c_int, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64)
_proto("gl_depth_result_sizes", C.c_int, _vp, _i64p, _i64p, _i32p)
_proto("gl_depth_get_windows", C.c_int, _vp, _vp, C.c_int64)
_proto("gl_depth_last_path", C.c_int, _vp, _i32p)
_proto("gl_depth_set_path", C.c_int, _vp, C.c_int32)
_proto("gl_depth_get_runs", C.c_int, _vp, _vp, _vp, _vp, C.c_int64)
_proto("gl_depth_windows", C.c_int, _vp, C.c_int32, _vp, _vp, C.c_int64)
_proto("gl_depth_classes", C.c_int, _vp, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depth_perbase", C.c_int, _vp, _vp)
_proto("gl_depth_interval_sums", C.c_int, _vp, _vp, _vp, C.c_int64, _vp)
_proto("gl_fasta_load", C.c_int, _vp, _vp, C.c_int64)
_proto("gl_fasta_stats", C.c_int, _vp, _vp, _vp, C.c_int64, _vp, _vp)
_proto("gl_depth_chunk_rows", C.c_int, C.c_int64, C.c_int64, C.c_int32, _vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depth_region", C.c_int, _vp, C.c_int64, C.c_int64, _vp, _vp, C.c_int64, C.c_int32, C.c_int32,
       C.c_int32, C.c_int64, _vp, C.c_int64, _i64p, _vp, _vp, C.c_int64, _i64p)
_proto("gl_indexcov_sizes", C.c_int, _vp, _vp, _vp, C.c_int32, _vp, _vp)
_proto("gl_indexcov_scale", C.c_int, _vp, _vp, C.c_int64, _i64p)
_proto("gl_indexcov_normalize", C.c_int, _vp, _vp, C.c_int64, C.c_double, _vp)
_proto("gl_indexcov_cohort", C.c_int, _vp, _vp, _vp, C.c_int32, _vp, _vp)
_proto("gl_indexcov_cohort_device", C.c_int, _vp, _vp, _vp, C.c_int32, _vp, _vp)
_proto("gl_indexcov_cohort_fallbacks", C.c_int, _vp, _i32p)
_proto("gl_indexcov_sizes_batch_device", C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int64, _vp)
_proto("gl_indexcov_counts", C.c_int, _vp, _vp, C.c_int64, _vp)
_proto("gl_indexcov_bins", C.c_int, _vp, _vp, C.c_int64, C.c_int64, _vp)
_proto("gl_indexcov_counts_batch", C.c_int, _vp, _vp, _vp, _vp, C.c_int32, _vp, _vp)
_proto("gl_indexcov_counts_batch_device", C.c_int, _vp, _vp, _vp, _vp, C.c_int32, _vp, _vp)
_proto("gl_indexcov_counts_segs_device", C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int32, _vp, _vp)
_proto("gl_indexcov_xnorm", C.c_int, _vp, _vp, _vp, C.c_int32, C.c_int32)
_proto("gl_format_g3", C.c_int, _vp, _vp, C.c_int64, _vp)
_proto("gl_format_g3_device", C.c_int, _vp, _vp, C.c_int64, _vp)
_proto("gl_bincount_i32", C.c_int, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp)
_proto("gl_depthwed_aggregate", C.c_int, _vp, _vp, C.c_int32, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp,
       C.c_int64, _i64p)
_proto("gl_depthwed_aggregate_device", C.c_int, _vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, _vp)
_proto("gl_depthwed_aggregate_i32", C.c_int, _vp, _vp, C.c_int32, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depthwed_aggregate_i32_device", C.c_int, _vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, C.c_int64, _vp, _vp)
_proto("gl_depthwed_aggregate_i32_p2p", C.c_int, _vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, C.c_int64, _vp, C.c_int32, C.c_int64, C.c_int32, _vp)
_proto("gl_ipc_export", C.c_int, _vp, _vp, _vp)
_proto("gl_ipc_open", C.c_int, _vp, _vp, C.POINTER(_vp))
_proto("gl_ipc_close", C.c_int, _vp, _vp)
_proto("gl_allgather_device_async", C.c_int, _vp, _vp, _vp, C.c_int64)
_proto("gl_comm_wait", C.c_int, _vp)
_proto("gl_comm_unique_id", C.c_int, _vp)
_proto("gl_comm_init", C.c_int, _vp, _vp, C.c_int, C.c_int)
_proto("gl_comm_destroy", C.c_int, _vp)
_proto("gl_allgather_device", C.c_int, _vp, _vp, _vp, C.c_int64)
_proto("gl_bam_decode_segments", C.c_int, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp), _vp, C.c_int64)
_proto("gl_segset_n_refs", C.c_int, _vp, _i32p, _i64p, _i64p)
_proto("gl_segset_ref", C.c_int, _vp, C.c_int32, C.POINTER(C.c_char_p), _i64p, C.POINTER(_vp), C.POINTER(_vp), _i64p)
_proto("gl_segset_free", None, _vp)
_proto("gl_bam_open", C.c_int, C.c_char_p, C.POINTER(_vp), _vp, C.c_int64)
_proto("gl_bam_close", None, _vp)
_proto("gl_bam_info", C.c_int, _vp, _i32p, _i32p)
_proto("gl_bam_ref", C.c_int, _vp, C.c_int32, C.POINTER(C.c_char_p), _i64p, _i64p)
_proto("gl_bam_decode", C.c_int, _vp, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_int64)
_proto("gl_bam_decode_device", C.c_int, _vp, _vp, C.c_int32, C.c_int32, C.POINTER(_vp), C.POINTER(_vp), _i64p, _vp)
_proto("gl_bgzf_inflate_device", C.c_int, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp)
_proto("gl_bai_read", C.c_int, C.c_char_p, C.POINTER(_vp), _vp, C.c_int64)
_proto("gl_bai_n_refs", C.c_int, _vp, _i32p, _u64p)
_proto("gl_bai_ref", C.c_int, _vp, C.c_int32, C.POINTER(_vp), _i64p, _u64p, _u64p, _i32p)
_proto("gl_bai_free", None, _vp)
_proto("gl_pack_segments16_bound", C.c_int64, C.c_int64)
_proto("gl_pack_segments16", C.c_int, _vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_pack_segments16_mt", C.c_int, _vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depth_add_segments_packed16", C.c_int, _vp, _vp, _vp, _vp, C.c_int64)
_proto("gl_depth_region_packed16", C.c_int, _vp, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32,
       C.c_int32, C.c_int64, _vp, C.c_int64, _i64p, _vp, _vp, C.c_int64, _i64p)
_proto("gl_pack_segments8_bound", C.c_int64, C.c_int64)
_proto("gl_pack_segments8", C.c_int, _vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depth_transport_phases", C.c_int, _vp, C.POINTER(C.c_double))
_proto("gl_depth_transport_stats", C.c_int, _vp, _i32p, C.POINTER(C.c_double), _i64p, _i64p)
_proto("gl_pack_segments16_fixed_mt", C.c_int, _vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_pack_segments8_mt", C.c_int, _vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depth_add_segments_packed8", C.c_int, _vp, _vp, _vp, _vp, C.c_int64)
_proto("gl_depth_add_segments_packed8_device", C.c_int, _vp, _vp, _vp, _vp, C.c_int64)
_proto("gl_depth_region_packed8", C.c_int, _vp, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32,
       C.c_int32, C.c_int64, _vp, C.c_int64, _i64p, _vp, _vp, C.c_int64, _i64p)
_proto("gl_depth_text", C.c_int, _vp, C.c_char_p, _vp, C.c_int64, _i64p, _vp, C.c_int64, _i64p)
_proto("gl_depth_text_bound", C.c_int64, C.c_char_p, C.c_int64)
_proto("gl_depth_format_rows", C.c_int, _vp, C.c_char_p, _vp, _vp, _vp, C.c_int64, _vp, C.c_int64, _i64p)
_proto("gl_depth_bed_contig", C.c_int, _vp, C.c_char_p, C.c_int64, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
       C.c_int32, _vp, C.c_int64, _i64p, _vp, C.c_int64, _i64p)
_proto("gl_depth_bed_contig_packed8", C.c_int, _vp, C.c_char_p, C.c_int64, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
       C.c_int64, _vp, C.c_int64, _i64p, _vp, C.c_int64, _i64p)
_proto("gl_indexsplit_accumulate", C.c_int, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp)
_proto("gl_indexsplit_chunks", C.c_int, _vp, _vp, C.c_int32, C.POINTER(C.c_char_p), _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_int64,
       C.POINTER(_vp), _i64p)
_proto("gl_crai_make_sizes", C.c_int, _vp, _vp, _vp, C.c_int64, _vp, C.c_int64, _i64p)
_proto("gl_depth_format_chunk", C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.c_int32, _vp, C.c_int64, _vp, _vp,
       C.c_int64, C.POINTER(_vp), _i64p, C.POINTER(_vp), _i64p)
_proto("gl_free_text", None, _vp)


def version() -> str:
    return lib.gl_version().decode()


def device_count() -> int:
    n = C.c_int(0)
    rc = lib.gl_device_count(C.byref(n))
    if rc != GL_OK:
        return 0
    return n.value


def bind_numa_for_device(device: int, share_index: int = 0, share_count: int = 1) -> Tuple[int, int]:
    """(numa node or -1, cpus bound) — call before the first pinned allocation / pool use"""
    node, n = C.c_int(-1), C.c_int(0)
    lib.gl_bind_numa_for_device(device, share_index, share_count, C.byref(node), C.byref(n))
    return node.value, n.value


def device_numa_node(device: int) -> int:
    node = C.c_int(-1)
    lib.gl_device_numa_node(device, C.byref(node))
    return node.value


def lpt_assign(weights, bins: int):
    """(bin_of int32[n], bin_load int64[bins]) — longest first onto the least loaded bin"""
    w = np.ascontiguousarray(weights, np.int64)
    bin_of, load = np.empty(w.size, np.int32), np.empty(bins, np.int64)
    rc = lib.gl_lpt_assign(w.ctypes.data_as(_vp), w.size, bins, bin_of.ctypes.data_as(_vp), load.ctypes.data_as(_vp))
    if rc != GL_OK:
        raise GlError(rc, "gl_lpt_assign: bad arguments")
    return bin_of, load


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def _as(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def format_chunk(chrom: str, rs: int, re: int, W: int, win_sum: np.ndarray, run_start: np.ndarray,
                 run_class: np.ndarray) -> Tuple[bytes, bytes]:
    """Host-only: the reference's rows for one chunk (no GPU needed)."""
    win_sum = _as(win_sum, np.int64)
    run_start = _as(run_start, np.int32)
    run_class = _as(run_class, np.uint8)
    d, c = _vp(), _vp()
    dl, cl = C.c_int64(0), C.c_int64(0)
    rc = lib.gl_depth_format_chunk(chrom.encode(), rs, re, W, _ptr(win_sum), len(win_sum), _ptr(run_start),
                                   _ptr(run_class), len(run_start), C.byref(d), C.byref(dl), C.byref(c), C.byref(cl))
    if rc != GL_OK:
        raise GlError(rc, "gl_depth_format_chunk: bad arguments")
    try:
        return C.string_at(d, dl.value), C.string_at(c, cl.value)
    finally:
        lib.gl_free_text(d)
        lib.gl_free_text(c)


def chunk_rows(rs: int, re: int, W: int, run_start: np.ndarray, run_class: np.ndarray):
    """Host-only: (s, e) of the window rows gl_depth_format_chunk writes for a chunk, in order."""
    run_start = _as(run_start, np.int32)
    run_class = _as(run_class, np.uint8)
    n = C.c_int64(0)
    lib.gl_depth_chunk_rows(rs, re, W, _ptr(run_start), _ptr(run_class), run_start.size, None, None, 0, C.byref(n))
    s, e = np.empty(n.value, np.int64), np.empty(n.value, np.int64)
    rc = lib.gl_depth_chunk_rows(rs, re, W, _ptr(run_start), _ptr(run_class), run_start.size, _ptr(s), _ptr(e), n.value, C.byref(n))
    if rc != GL_OK:
        raise GlError(rc, "gl_depth_chunk_rows: bad arguments")
    return s, e


def pack_segments16(start: np.ndarray, end: np.ndarray, threads: Optional[int] = None):
    """Host-only: (anchors int32[nb], off uint16[nb*256], len uint16[nb*256]) — the feeder's compact format.
    threads=None: single-threaded; an int: gl_pack_segments16_mt (0 = the whole pool)."""
    start, end = _as(start, np.int32), _as(end, np.int32)
    nb = C.c_int64(0)
    cap = max(1, int(lib.gl_pack_segments16_bound(start.size)) // 8 + start.size // 256 + 2)
    while True:
        a = np.empty(cap, np.int32)
        o = np.empty(cap * 256, np.uint16)
        ln = np.empty(cap * 256, np.uint16)
        if threads is None:
            rc = lib.gl_pack_segments16(_ptr(start), _ptr(end), start.size, _ptr(a), _ptr(o), _ptr(ln), cap, C.byref(nb))
        else:
            rc = lib.gl_pack_segments16_mt(_ptr(start), _ptr(end), start.size, threads, _ptr(a), _ptr(o), _ptr(ln), cap, C.byref(nb))
        if rc == GL_OK:
            k = nb.value
            return a[:k], o[: k * 256], ln[: k * 256]
        if rc != GL_ERANGE:
            raise GlError(rc, "gl_pack_segments16: bad arguments")
        cap = nb.value + 1


def indexsplit_layout(sample_sizes, R: int):
    """CSR layout gl_indexsplit_accumulate wants: (sizes int64, ptr int64[S*(R+1)], out_ptr int64[R+1])"""
    S = len(sample_sizes)
    flat, ptr, maxlen = [], np.zeros(S * (R + 1), np.int64), np.zeros(R, np.int64)
    off = 0
    for s, per_ref in enumerate(sample_sizes):
        for r in range(R):
            ptr[s * (R + 1) + r] = off
            if r < len(per_ref):
                a = np.asarray(per_ref[r], np.int64)
                flat.append(a); off += a.size
                maxlen[r] = max(maxlen[r], a.size)
        ptr[s * (R + 1) + R] = off
    sizes = np.concatenate(flat) if flat else np.zeros(0, np.int64)
    return sizes, ptr, np.concatenate([[0], np.cumsum(maxlen)]).astype(np.int64)


def indexsplit_chunks(tile_sum, out_ptr, names, ref_lens, ref_ids, N: int, problems=()) -> bytes:
    """Host-only: gl_indexsplit_chunks -> the output lines."""
    tile_sum = _as(tile_sum, np.float64)
    out_ptr = _as(out_ptr, np.int64)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    lens, ids = _as(np.asarray(ref_lens), np.int64), _as(np.asarray(ref_ids), np.int32)
    pr = _as(np.asarray([p[0] for p in problems]), np.int32)
    ps = _as(np.asarray([p[1] for p in problems]), np.int64)
    pe = _as(np.asarray([p[2] for p in problems]), np.int64)
    t, n = _vp(), C.c_int64(0)
    rc = lib.gl_indexsplit_chunks(_ptr(tile_sum), _ptr(out_ptr), out_ptr.size - 1, arr, _ptr(lens), _ptr(ids), len(names), N,
                                  _ptr(pr), _ptr(ps), _ptr(pe), len(problems), C.byref(t), C.byref(n))
    if rc != GL_OK:
        raise GlError(rc, "gl_indexsplit_chunks: bad arguments")
    try:
        return C.string_at(t, n.value)
    finally:
        lib.gl_free_text(t)


def pack_segments16_fixed(start: np.ndarray, end: np.ndarray, threads: int = 0, esc_cap: Optional[int] = None):
    """fixed-block packed16 (gl_pack_segments16_fixed_mt): (anchors, off, len, esc_start, esc_end)"""
    start, end = _as(start, np.int32), _as(end, np.int32)
    n = start.size
    nb = (n + 255) // 256
    anchors, off, ln = np.empty(nb, np.int32), np.empty(nb * 256, np.uint16), np.empty(nb * 256, np.uint16)
    cap = n if esc_cap is None else esc_cap
    es, ee = np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.int32)
    m = C.c_int64(0)
    rc = lib.gl_pack_segments16_fixed_mt(_ptr(start), _ptr(end), n, threads, _ptr(anchors), _ptr(off), _ptr(ln), _ptr(es), _ptr(ee), cap, C.byref(m))
    if rc != 0:
        raise GlError(rc, "gl_pack_segments16_fixed_mt (needs %d escapes)" % m.value)
    return anchors, off, ln, es[: m.value].copy(), ee[: m.value].copy()


def pack_segments8(start: np.ndarray, end: np.ndarray, threads: Optional[int] = None):
    """Host-only: (anchors int32[nb], dstart uint8[nb*64], len uint8[nb*64]) — the feeder's densest format (short reads).
    threads=None: the single-threaded packer; an int: gl_pack_segments8_mt (0 = all pool threads)."""
    start, end = _as(start, np.int32), _as(end, np.int32)
    nb = C.c_int64(0)
    cap = max(1, start.size // 48 + 64)
    while True:
        a = np.empty(cap, np.int32)
        d = np.empty(cap * 64, np.uint8)
        ln = np.empty(cap * 64, np.uint8)
        if threads is None:
            rc = lib.gl_pack_segments8(_ptr(start), _ptr(end), start.size, _ptr(a), _ptr(d), _ptr(ln), cap, C.byref(nb))
        else:
            rc = lib.gl_pack_segments8_mt(_ptr(start), _ptr(end), start.size, threads, _ptr(a), _ptr(d), _ptr(ln), cap, C.byref(nb))
        if rc == GL_OK:
            k = nb.value
            return a[:k], d[: k * 64], ln[: k * 64]
        if rc != GL_ERANGE:
            raise GlError(rc, "gl_pack_segments8: bad arguments")
        cap = nb.value + 1


def unpack_segments8(anchors, dstart, ln):
    """numpy decode of packed8 (what depth_unpack8_kernel computes): -> (start, end) int64, empty slots dropped"""
    nb = anchors.size
    d = dstart.reshape(nb, 64).astype(np.int64)
    s = anchors.astype(np.int64)[:, None] + np.cumsum(d, axis=1)
    l = ln.reshape(nb, 64).astype(np.int64)
    keep = l > 0
    return s[keep], (s + l)[keep]


def bam_segments(path: str, min_mapq: int = 1, threads: int = 4, only_tid: int = -1):
    """Host-only feeder: {"refs": [(name, length)], "segments": {tid: (start, end)}, "n_records", "n_pass"}."""
    h = _vp()
    err = C.create_string_buffer(512)
    rc = lib.gl_bam_decode_segments(path.encode(), min_mapq, threads, only_tid, C.byref(h), C.cast(err, _vp), 512)
    if rc != GL_OK:
        raise GlError(rc, err.value.decode())
    try:
        n, nrec, npass = C.c_int32(0), C.c_int64(0), C.c_int64(0)
        lib.gl_segset_n_refs(h, C.byref(n), C.byref(nrec), C.byref(npass))
        refs, segs = [], {}
        for tid in range(n.value):
            name, ln, ps, pe, k = C.c_char_p(), C.c_int64(0), _vp(), _vp(), C.c_int64(0)
            lib.gl_segset_ref(h, tid, C.byref(name), C.byref(ln), C.byref(ps), C.byref(pe), C.byref(k))
            refs.append((name.value.decode(), ln.value))
            if k.value:
                s_ = np.ctypeslib.as_array(C.cast(ps, _i32p), (k.value,)).copy()
                e_ = np.ctypeslib.as_array(C.cast(pe, _i32p), (k.value,)).copy()
                segs[tid] = (s_, e_)
        return {"refs": refs, "segments": segs, "n_records": nrec.value, "n_pass": npass.value}
    finally:
        lib.gl_segset_free(h)


class _BamSegments(C.Structure):
    _fields_ = [("format", C.c_int32), ("units", C.c_int32), ("max_len", C.c_int32), ("_pad", C.c_int32), ("n", C.c_int64),
                ("a0", _vp), ("a1", _vp), ("a2", _vp), ("n_records", C.c_int64), ("n_pass", C.c_int64), ("bytes_in", C.c_int64),
                ("bytes_out", C.c_int64), ("inflate_s", C.c_double), ("parse_s", C.c_double), ("wall_s", C.c_double)]


class Bam:
    """Index-guided parallel feeder (gl_bam_*): host-only."""

    def __init__(self, path: str):
        h = _vp()
        err = C.create_string_buffer(512)
        rc = lib.gl_bam_open(path.encode(), C.byref(h), C.cast(err, _vp), 512)
        if rc != GL_OK:
            raise GlError(rc, err.value.decode())
        self.h = h
        n, hi = C.c_int32(0), C.c_int32(0)
        lib.gl_bam_info(h, C.byref(n), C.byref(hi))
        self.has_index = bool(hi.value)
        self.refs = []
        for tid in range(n.value):
            name, ln, nm = C.c_char_p(), C.c_int64(0), C.c_int64(0)
            lib.gl_bam_ref(h, tid, C.byref(name), C.byref(ln), C.byref(nm))
            self.refs.append((name.value.decode(), ln.value, nm.value))

    def close(self):
        if self.h:
            lib.gl_bam_close(self.h)
            self.h = None

    def decode(self, tid: int, beg: int = 0, end: int = 1 << 40, min_mapq: int = 1, threads: int = 0, want: int = 0):
        """-> dict(format, arrays (copies), stats)"""
        o = _BamSegments()
        err = C.create_string_buffer(512)
        rc = lib.gl_bam_decode(self.h, tid, beg, end, min_mapq, threads, want, C.byref(o), C.cast(err, _vp), 512)
        if rc != GL_OK:
            raise GlError(rc, err.value.decode())
        out = {"format": o.format, "n": o.n, "units": o.units, "max_len": o.max_len, "n_records": o.n_records, "n_pass": o.n_pass,
               "bytes_in": o.bytes_in, "bytes_out": o.bytes_out, "inflate_s": o.inflate_s, "parse_s": o.parse_s, "wall_s": o.wall_s}
        if o.format == 8 and o.n:
            out["anchors"] = np.ctypeslib.as_array(C.cast(o.a0, _i32p), (o.n,)).copy()
            out["dstart"] = np.ctypeslib.as_array(C.cast(o.a1, _u8p), (o.n * 64,)).copy()
            out["len"] = np.ctypeslib.as_array(C.cast(o.a2, _u8p), (o.n * 64,)).copy()
        elif o.format == 32 and o.n:
            out["start"] = np.ctypeslib.as_array(C.cast(o.a0, _i32p), (o.n,)).copy()
            out["end"] = np.ctypeslib.as_array(C.cast(o.a1, _i32p), (o.n,)).copy()
        return out


def bam_decode_device(ctx: "Ctx", bam: Bam, tid: int, min_mapq: int = 1):
    """The GPU feeder (BGZF inflate + record parse on the device) -> dict(start, end (copies), stats); GlError(GL_ESTATE) when
    the reference cannot be done on the device."""
    o = _BamSegments()
    ps, pe, n = _vp(), _vp(), C.c_int64(0)
    rc = lib.gl_bam_decode_device(ctx.h, bam.h, tid, min_mapq, C.byref(ps), C.byref(pe), C.byref(n), C.byref(o))
    if rc != GL_OK:
        raise GlError(rc, lib.gl_last_error(ctx.h).decode())
    out = {"n": n.value, "units": o.units, "max_len": o.max_len, "n_records": o.n_records, "n_pass": o.n_pass, "bytes_in": o.bytes_in,
           "bytes_out": o.bytes_out, "inflate_s": o.inflate_s, "parse_s": o.parse_s, "wall_s": o.wall_s, "d_start": ps.value, "d_end": pe.value}
    if n.value:
        s_, e_ = np.empty(n.value, np.int32), np.empty(n.value, np.int32)
        ctx._ck(lib.gl_memcpy_d2h(ctx.h, _ptr(s_), ps.value, n.value * 4))
        ctx._ck(lib.gl_memcpy_d2h(ctx.h, _ptr(e_), pe.value, n.value * 4))
        out["start"], out["end"] = s_, e_
    else:
        out["start"], out["end"] = np.zeros(0, np.int32), np.zeros(0, np.int32)
    return out


def crai_make_sizes(start, span, nbytes) -> np.ndarray:
    """Host-only: 16 KB pseudo-tile sizes of one reference from its CRAM slices (crai.go:56-127)."""
    start, span, nbytes = _as(start, np.int64), _as(span, np.int64), _as(nbytes, np.int32)
    cap = int((start[-1] + max(int(span[-1]), 0)) // 16384 + 64) if start.size else 1
    out = np.empty(cap, np.int64)
    n = C.c_int64(0)
    rc = lib.gl_crai_make_sizes(_ptr(start), _ptr(span), _ptr(nbytes), start.size, _ptr(out), cap, C.byref(n))
    if rc != GL_OK:
        raise GlError(rc, "gl_crai_make_sizes: tile-width logic error (the reference panics) or capacity")
    return out[: n.value].copy()


def bai_read(path: str):
    h = _vp()
    err = C.create_string_buffer(512)
    rc = lib.gl_bai_read(path.encode(), C.byref(h), C.cast(err, _vp), 512)
    if rc != GL_OK:
        raise GlError(rc, err.value.decode())
    try:
        n, nc = C.c_int32(0), C.c_uint64(0)
        lib.gl_bai_n_refs(h, C.byref(n), C.byref(nc))
        out = {"ioffsets": [], "mapped": [], "unmapped": [], "n_no_coor": nc.value}
        for tid in range(n.value):
            p, k, m, u, hs = _vp(), C.c_int64(0), C.c_uint64(0), C.c_uint64(0), C.c_int32(0)
            lib.gl_bai_ref(h, tid, C.byref(p), C.byref(k), C.byref(m), C.byref(u), C.byref(hs))
            out["ioffsets"].append(np.ctypeslib.as_array(C.cast(p, _u64p), (k.value,)).copy() if k.value else np.zeros(0, np.uint64))
            out["mapped"].append(m.value); out["unmapped"].append(u.value)
        return out
    finally:
        lib.gl_bai_free(h)


class DevBuf:
    """A device allocation owned by a Ctx."""

    def __init__(self, ctx: "Ctx", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = _vp()
        ctx._ck(lib.gl_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, a: np.ndarray) -> "DevBuf":
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        self.ctx._ck(lib.gl_memcpy_h2d(self.ctx.h, self.ptr, _ptr(a), a.nbytes))
        return self

    def download(self, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert out.nbytes <= self.nbytes
        self.ctx._ck(lib.gl_memcpy_d2h(self.ctx.h, _ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib.gl_dev_free(self.ctx.h, self.ptr)
            self.ptr = None


class Ctx:
    """One GPU context (gl_ctx)."""

    def __init__(self, device: int = 0):
        h = _vp()
        rc = lib.gl_ctx_create(device, C.byref(h))
        if rc != GL_OK:
            raise GlError(rc, lib.gl_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib.gl_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int):
        if rc != GL_OK:
            raise GlError(rc, lib.gl_last_error(self.h).decode())

    # ---- plumbing
    def sync(self):
        self._ck(lib.gl_sync(self.h))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        self._ck(lib.gl_launch_count(self.h, C.byref(n)))
        return n.value

    def dev_array(self, a: np.ndarray) -> DevBuf:
        a = np.ascontiguousarray(a)
        return DevBuf(self, max(a.nbytes, 16)).upload(a)

    def dev_empty(self, nbytes: int) -> DevBuf:
        return DevBuf(self, nbytes)

    def pinned_empty(self, count: int, dtype) -> np.ndarray:
        dt = np.dtype(dtype)
        p = _vp()
        self._ck(lib.gl_host_alloc_pinned(self.h, max(count * dt.itemsize, 16), C.byref(p)))
        buf = (C.c_char * (count * dt.itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=count)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p.value)
        return arr

    def timer_start(self):
        self._ck(lib.gl_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float(0)
        self._ck(lib.gl_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool):
        self._ck(lib.gl_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """[(kernel name, ms), ...] for every kernel launched since the last read"""
        buf = C.create_string_buffer(1 << 20)
        need = C.c_int64(0)
        self._ck(lib.gl_profile_read(self.h, C.cast(buf, _vp), len(buf), C.byref(need)))
        out = []
        for ln in buf.value.decode().splitlines():
            nm, ms = ln.rsplit(" ", 1)
            out.append((nm, float(ms)))
        return out

    def flush_l2(self):
        self._ck(lib.gl_flush_l2(self.h))

    # ---- depth
    def depth_begin(self, rs: int, re: int):
        self._ck(lib.gl_depth_begin(self.h, rs, re))

    def depth_add_segments(self, start: np.ndarray, end: np.ndarray):
        start, end = _as(start, np.int32), _as(end, np.int32)
        assert start.shape == end.shape
        self._ck(lib.gl_depth_add_segments(self.h, _ptr(start), _ptr(end), start.size))

    def depth_add_segments_device(self, d_start: DevBuf, d_end: DevBuf, n: int, offset: int = 0):
        self._ck(lib.gl_depth_add_segments_device(self.h, d_start.ptr + 4 * offset, d_end.ptr + 4 * offset, n))

    def depth_add_segments_packed16(self, anchors: np.ndarray, off: np.ndarray, ln: np.ndarray):
        self._ck(lib.gl_depth_add_segments_packed16(self.h, _ptr(anchors), _ptr(off), _ptr(ln), anchors.size))

    def depth_add_segments_packed8(self, anchors: np.ndarray, dstart: np.ndarray, ln: np.ndarray):
        self._ck(lib.gl_depth_add_segments_packed8(self.h, _ptr(anchors), _ptr(dstart), _ptr(ln), anchors.size))

    def depth_add_segments_packed8_device(self, d_anchors: "DevBuf", d_dstart: "DevBuf", d_len: "DevBuf", n_blocks: int):
        self._ck(lib.gl_depth_add_segments_packed8_device(self.h, d_anchors.ptr, d_dstart.ptr, d_len.ptr, n_blocks))

    def depth_region_packed8(self, rs: int, re: int, anchors, dstart, ln, W: int, mincov: int = 4, maxmean: int = 0,
                             run_break: int = 0, out=None):
        n_win = (re - 1) // W - rs // W + 1
        if out is None:
            out = (np.empty(n_win, np.int64), np.empty(max(1024, (re - rs) // 8), np.int32),
                   np.empty(max(1024, (re - rs) // 8), np.uint8))
        s, r0, rc_ = out
        nw, nr = C.c_int64(0), C.c_int64(0)
        self._ck(lib.gl_depth_region_packed8(self.h, rs, re, _ptr(anchors), _ptr(dstart), _ptr(ln), anchors.size, W, mincov,
                                             maxmean, run_break, _ptr(s), s.size, C.byref(nw), _ptr(r0), _ptr(rc_),
                                             min(r0.size, rc_.size), C.byref(nr)))
        return s[: nw.value], r0[: nr.value], rc_[: nr.value]

    def depth_region_packed16(self, rs: int, re: int, anchors, off, ln, W: int, mincov: int = 4, maxmean: int = 0,
                              run_break: int = 0, out=None):
        n_win = (re - 1) // W - rs // W + 1
        if out is None:
            out = (np.empty(n_win, np.int64), np.empty(max(1024, (re - rs) // 8), np.int32),
                   np.empty(max(1024, (re - rs) // 8), np.uint8))
        s, r0, rc_ = out
        nw, nr = C.c_int64(0), C.c_int64(0)
        self._ck(lib.gl_depth_region_packed16(self.h, rs, re, _ptr(anchors), _ptr(off), _ptr(ln), anchors.size, W, mincov,
                                              maxmean, run_break, _ptr(s), s.size, C.byref(nw), _ptr(r0), _ptr(rc_),
                                              min(r0.size, rc_.size), C.byref(nr)))
        return s[: nw.value], r0[: nr.value], rc_[: nr.value]

    def depth_reduce(self, W: int, mincov: int = 4, maxmean: int = 0, run_break: int = 0):
        self._ck(lib.gl_depth_reduce(self.h, W, mincov, maxmean, run_break))

    def depth_result_sizes(self) -> Tuple[int, int, int]:
        nw, nr, md = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        self._ck(lib.gl_depth_result_sizes(self.h, C.byref(nw), C.byref(nr), C.byref(md)))
        return nw.value, nr.value, md.value

    def depth_get_windows(self) -> np.ndarray:
        nw, _, _ = self.depth_result_sizes()
        s = np.empty(nw, np.int64)
        self._ck(lib.gl_depth_get_windows(self.h, _ptr(s), nw))
        return s

    def depth_last_path(self) -> int:
        p = C.c_int32(0)
        self._ck(lib.gl_depth_last_path(self.h, C.byref(p)))
        return p.value

    def depth_transport_stats(self):
        """(transport, host pack seconds, bytes sent host->device, escaped segments) of the last depth_bed_contig call"""
        t, ps, hb, ne = C.c_int32(0), C.c_double(0), C.c_int64(0), C.c_int64(0)
        self._ck(lib.gl_depth_transport_stats(self.h, C.byref(t), C.byref(ps), C.byref(hb), C.byref(ne)))
        return t.value, ps.value, hb.value, ne.value

    def depth_transport_phases(self):
        ph = (C.c_double * 3)()
        self._ck(lib.gl_depth_transport_phases(self.h, ph))
        return [ph[0], ph[1], ph[2]]

    def depth_set_path(self, path: int):
        self._ck(lib.gl_depth_set_path(self.h, path))

    def depth_get_runs(self, want_end: bool = False):
        _, nr, _ = self.depth_result_sizes()
        rs_ = np.empty(nr, np.int32)
        rc_ = np.empty(nr, np.uint8)
        re_ = np.empty(nr, np.int32) if want_end else None
        self._ck(lib.gl_depth_get_runs(self.h, _ptr(rs_), _ptr(re_), _ptr(rc_), nr))
        return (rs_, re_, rc_) if want_end else (rs_, rc_)

    def depth_windows(self, W: int, n_windows: int, want_min: bool = True):
        s = np.empty(n_windows, np.int64)
        m = np.empty(n_windows, np.int32) if want_min else None
        self._ck(lib.gl_depth_windows(self.h, W, _ptr(s), _ptr(m), n_windows))
        return s, m

    def depth_classes(self, mincov: int, maxmean: int, cap: int):
        rs_ = np.empty(cap, np.int32)
        re_ = np.empty(cap, np.int32)
        rc_ = np.empty(cap, np.uint8)
        n = C.c_int64(0)
        self._ck(lib.gl_depth_classes(self.h, mincov, maxmean, _ptr(rs_), _ptr(re_), _ptr(rc_), cap, C.byref(n)))
        return rs_[: n.value], re_[: n.value], rc_[: n.value]

    def indexsplit_accumulate(self, sizes, ptr, S: int, R: int, out_ptr) -> np.ndarray:
        sizes, ptr, out_ptr = _as(sizes, np.int64), _as(ptr, np.int64), _as(out_ptr, np.int64)
        out = np.zeros(int(out_ptr[-1]), np.float64)
        self._ck(lib.gl_indexsplit_accumulate(self.h, _ptr(sizes), _ptr(ptr), S, R, _ptr(out_ptr), _ptr(out)))
        return out

    def depth_interval_sums(self, a, b) -> np.ndarray:
        a, b = _as(a, np.int32), _as(b, np.int32)
        out = np.empty(a.size, np.int64)
        self._ck(lib.gl_depth_interval_sums(self.h, _ptr(a), _ptr(b), a.size, _ptr(out)))
        return out

    def fasta_load(self, record_bytes):
        a = np.frombuffer(record_bytes, np.uint8) if not isinstance(record_bytes, np.ndarray) else _as(record_bytes, np.uint8)
        self._ck(lib.gl_fasta_load(self.h, _ptr(a), a.size))

    def fasta_stats(self, byte_start, byte_end):
        """-> (counts[n,4] = G+C, lower, ACGT, CpG ; stats[n,3] = GC, CpG, Masked)"""
        a, b = _as(byte_start, np.int64), _as(byte_end, np.int64)
        counts = np.zeros((a.size, 4), np.int64)
        stats = np.zeros((a.size, 3), np.float64)
        self._ck(lib.gl_fasta_stats(self.h, _ptr(a), _ptr(b), a.size, _ptr(counts), _ptr(stats)))
        return counts, stats

    def depth_perbase(self, length: int) -> np.ndarray:
        out = np.empty(length, np.int32)
        self._ck(lib.gl_depth_perbase(self.h, _ptr(out)))
        return out

    def depth_region(self, rs: int, re: int, start: np.ndarray, end: np.ndarray, W: int, mincov: int = 4,
                     maxmean: int = 0, run_break: int = 0, out=None):
        """One call, host in / host out.  `out` = (sum, run_start, run_class) preallocated arrays."""
        n_win = (re - 1) // W - rs // W + 1
        if out is None:
            out = (np.empty(n_win, np.int64), np.empty(max(1024, (re - rs) // 8), np.int32),
                   np.empty(max(1024, (re - rs) // 8), np.uint8))
        s, r0, rc_ = out
        nw, nr = C.c_int64(0), C.c_int64(0)
        self._ck(lib.gl_depth_region(self.h, rs, re, _ptr(start), _ptr(end), start.size, W, mincov, maxmean, run_break,
                                     _ptr(s), s.size, C.byref(nw), _ptr(r0), _ptr(rc_), min(r0.size, rc_.size),
                                     C.byref(nr)))
        return s[: nw.value], r0[: nr.value], rc_[: nr.value]

    # ---- depth text (device formatter)
    def _text_bufs(self, chrom: str, n_win: int, out):
        if out is None:
            out = (np.empty(int(lib.gl_depth_text_bound(chrom.encode(), n_win)), np.uint8), np.empty(1 << 16, np.uint8))
        return out

    def depth_text(self, chrom: str, out=None) -> Tuple[bytes, bytes]:
        """(depth.bed bytes, callable.bed bytes) of the last reduce, formatted on the device."""
        nw, _, _ = self.depth_result_sizes()
        hd, ca = self._text_bufs(chrom, nw, out)
        hl, cl = C.c_int64(0), C.c_int64(0)
        rc = lib.gl_depth_text(self.h, chrom.encode(), _ptr(hd), hd.size, C.byref(hl), _ptr(ca), ca.size, C.byref(cl))
        if rc == GL_ERANGE and out is None:
            hd, ca = np.empty(hl.value + 16, np.uint8), np.empty(cl.value + 16, np.uint8)
            rc = lib.gl_depth_text(self.h, chrom.encode(), _ptr(hd), hd.size, C.byref(hl), _ptr(ca), ca.size, C.byref(cl))
        self._ck(rc)
        return hd[: hl.value].tobytes(), ca[: cl.value].tobytes()

    def depth_format_rows(self, chrom: str, row_s, row_e, row_sum) -> bytes:
        row_s, row_e, row_sum = _as(row_s, np.int32), _as(row_e, np.int32), _as(row_sum, np.int64)
        out = np.empty(row_s.size * (len(chrom) + 33) + 16, np.uint8)
        ln = C.c_int64(0)
        self._ck(lib.gl_depth_format_rows(self.h, chrom.encode(), _ptr(row_s), _ptr(row_e), _ptr(row_sum), row_s.size, _ptr(out),
                                          out.size, C.byref(ln)))
        return out[: ln.value].tobytes()

    def depth_bed_contig(self, chrom: str, length: int, start, end, W: int, mincov: int = 4, maxmean: int = 0,
                         step: int = 10_000_000, threads: int = 0, out=None, raw: bool = False):
        """One call: host int32 segments in -> (depth.bed, callable.bed) bytes out.  raw=True returns the two lengths only
        (the bytes are in `out`)."""
        n_win = (length - 1) // W + 1
        keep = out is not None
        hd, ca = self._text_bufs(chrom, n_win, out)
        hl, cl = C.c_int64(0), C.c_int64(0)
        cb = chrom.encode()
        rc = lib.gl_depth_bed_contig(self.h, cb, length, _ptr(start), _ptr(end), start.size, W, mincov, maxmean, step, threads,
                                     _ptr(hd), hd.size, C.byref(hl), _ptr(ca), ca.size, C.byref(cl))
        if rc == GL_ERANGE and not keep:
            hd, ca = np.empty(hl.value + 16, np.uint8), np.empty(cl.value + 16, np.uint8)
            rc = lib.gl_depth_bed_contig(self.h, cb, length, _ptr(start), _ptr(end), start.size, W, mincov, maxmean, step, threads,
                                         _ptr(hd), hd.size, C.byref(hl), _ptr(ca), ca.size, C.byref(cl))
        self._ck(rc)
        if raw:
            return hl.value, cl.value
        return hd[: hl.value].tobytes(), ca[: cl.value].tobytes()

    def depth_bed_contig_packed8(self, chrom: str, length: int, anchors, dstart, ln, W: int, mincov: int = 4, maxmean: int = 0,
                                 step: int = 10_000_000, out=None, raw: bool = False):
        n_win = (length - 1) // W + 1
        keep = out is not None
        hd, ca = self._text_bufs(chrom, n_win, out)
        hl, cl = C.c_int64(0), C.c_int64(0)
        cb = chrom.encode()
        args = (self.h, cb, length, _ptr(anchors), _ptr(dstart), _ptr(ln), anchors.size, W, mincov, maxmean, step)
        rc = lib.gl_depth_bed_contig_packed8(*args, _ptr(hd), hd.size, C.byref(hl), _ptr(ca), ca.size, C.byref(cl))
        if rc == GL_ERANGE and not keep:
            hd, ca = np.empty(hl.value + 16, np.uint8), np.empty(cl.value + 16, np.uint8)
            rc = lib.gl_depth_bed_contig_packed8(*args, _ptr(hd), hd.size, C.byref(hl), _ptr(ca), ca.size, C.byref(cl))
        self._ck(rc)
        if raw:
            return hl.value, cl.value
        return hd[: hl.value].tobytes(), ca[: cl.value].tobytes()

    def bgzf_inflate(self, data: bytes):
        """GPU inflate of a whole BGZF byte string -> (inflated bytes, per-block status int32[]); test / tool helper"""
        b = np.frombuffer(data, np.uint8)
        offs, outs, p = [0], [0], 0
        while p + 18 <= b.size:
            bs = int(b[p + 16]) + (int(b[p + 17]) << 8) + 1
            if p + bs > b.size:
                break
            isize = int.from_bytes(bytes(b[p + bs - 4:p + bs]), "little")
            p += bs
            offs.append(p); outs.append(outs[-1] + isize)
        nb = len(offs) - 1
        d_c, d_co, d_oo = self.dev_array(b), self.dev_array(np.array(offs, np.int64)), self.dev_array(np.array(outs, np.int64))
        d_out, d_st = self.dev_empty(max(outs[-1], 16)), self.dev_array(np.full(max(nb, 1), -1, np.int32))
        try:
            self._ck(lib.gl_bgzf_inflate_device(self.h, d_c.ptr, d_co.ptr, d_oo.ptr, nb, d_out.ptr, d_st.ptr))
            self.sync()
            return d_out.download(np.uint8, outs[-1]).tobytes(), d_st.download(np.int32, nb)
        finally:
            for x in (d_c, d_co, d_oo, d_out, d_st):
                x.free()

    # ---- indexcov
    def indexcov_sizes(self, voff: np.ndarray, ref_ptr: np.ndarray):
        voff, ref_ptr = _as(voff, np.uint64), _as(ref_ptr, np.int64)
        n_refs = ref_ptr.size - 1
        size_ptr = np.zeros(n_refs + 1, np.int64)
        sizes = np.empty(max(int(voff.size), 1), np.int64)
        self._ck(lib.gl_indexcov_sizes(self.h, _ptr(voff), _ptr(ref_ptr), n_refs, _ptr(sizes), _ptr(size_ptr)))
        return sizes[: size_ptr[-1]], size_ptr

    def indexcov_scale(self, sizes: np.ndarray) -> int:
        sizes = _as(sizes, np.int64)
        m = C.c_int64(0)
        self._ck(lib.gl_indexcov_scale(self.h, _ptr(sizes), sizes.size, C.byref(m)))
        return m.value

    def indexcov_normalize(self, sizes: np.ndarray, median: float) -> np.ndarray:
        sizes = _as(sizes, np.int64)
        out = np.empty(sizes.size, np.float32)
        self._ck(lib.gl_indexcov_normalize(self.h, _ptr(sizes), sizes.size, float(median), _ptr(out)))
        return out

    def indexcov_cohort(self, sizes: np.ndarray, sample_ptr: np.ndarray, want_depth: bool = True):
        sizes, sample_ptr = _as(sizes, np.int64), _as(sample_ptr, np.int64)
        S = sample_ptr.size - 1
        med = np.empty(S, np.float64)
        dep = np.empty(sizes.size, np.float32) if want_depth else None
        self._ck(lib.gl_indexcov_cohort(self.h, _ptr(sizes), _ptr(sample_ptr), S, _ptr(med), _ptr(dep)))
        return med, dep

    def indexcov_cohort_fallbacks(self) -> int:
        n = C.c_int32(0)
        self._ck(lib.gl_indexcov_cohort_fallbacks(self.h, C.byref(n)))
        return n.value

    def indexcov_sizes_batch(self, voff: np.ndarray, voff_off, n_intv, size_off, total_sizes: int) -> np.ndarray:
        """I1 for many (sample, reference) descriptors in one launch (arrays uploaded here; the device form keeps them resident)"""
        d_v, d_o = self.dev_array(_as(voff, np.uint64)), self.dev_array(_as(voff_off, np.int64))
        d_n, d_s = self.dev_array(_as(n_intv, np.int32)), self.dev_array(_as(size_off, np.int64))
        d_out = self.dev_empty(max(total_sizes, 1) * 8)
        try:
            self._ck(lib.gl_indexcov_sizes_batch_device(self.h, d_v.ptr, d_o.ptr, d_n.ptr, d_s.ptr, len(n_intv), d_out.ptr))
            return d_out.download(np.int64, total_sizes)
        finally:
            for b in (d_v, d_o, d_n, d_s, d_out):
                b.free()

    def indexcov_counts(self, depth: np.ndarray, counts: Optional[np.ndarray] = None) -> np.ndarray:
        depth = _as(depth, np.float32)
        if counts is None:
            counts = np.zeros(INDEXCOV_SLOTS, np.int32)
        self._ck(lib.gl_indexcov_counts(self.h, _ptr(depth), depth.size, _ptr(counts)))
        return counts

    def indexcov_bins(self, depth: np.ndarray, longest: int, out4: Optional[np.ndarray] = None) -> np.ndarray:
        depth = _as(depth, np.float32)
        if out4 is None:
            out4 = np.zeros(4, np.int64)
        self._ck(lib.gl_indexcov_bins(self.h, _ptr(depth), depth.size, longest, _ptr(out4)))
        return out4

    def indexcov_counts_batch(self, depth: np.ndarray, seg_ptr: np.ndarray, longest: Optional[np.ndarray] = None):
        depth, seg_ptr = _as(depth, np.float32), _as(seg_ptr, np.int64)
        n = seg_ptr.size - 1
        lg = None if longest is None else _as(longest, np.int64)
        counts = np.empty((n, INDEXCOV_SLOTS), np.int32)
        bins = np.empty((n, 4), np.int64)
        self._ck(lib.gl_indexcov_counts_batch(self.h, _ptr(depth), _ptr(seg_ptr), _ptr(lg), n, _ptr(counts), _ptr(bins)))
        return counts, bins

    def indexcov_xnorm(self, depths: np.ndarray, lens: np.ndarray) -> np.ndarray:
        d = np.ascontiguousarray(depths, np.float32).copy()
        lens = _as(lens, np.int32)
        S, T = d.shape
        self._ck(lib.gl_indexcov_xnorm(self.h, _ptr(d), _ptr(lens), S, T))
        return d

    def indexcov_cohort_device(self, d_sizes: DevBuf, d_sample_ptr: DevBuf, S: int, d_medians: DevBuf, d_depth: Optional[DevBuf]):
        self._ck(lib.gl_indexcov_cohort_device(self.h, d_sizes.ptr, d_sample_ptr.ptr, S, d_medians.ptr,
                                               d_depth.ptr if d_depth is not None else None))

    def depthwed_aggregate_device(self, d_means: DevBuf, S: int, R: int, d_grp: Optional[DevBuf], n_out: int, d_out: DevBuf):
        self._ck(lib.gl_depthwed_aggregate_device(self.h, d_means.ptr, S, R, d_grp.ptr if d_grp is not None else None, n_out,
                                                  d_out.ptr))

    def format_g3(self, vals: np.ndarray) -> np.ndarray:
        """(n,10) uint8 tokens: text in [0:len), len in byte 9"""
        vals = _as(vals, np.float32)
        tok = np.empty((vals.size, 10), np.uint8)
        self._ck(lib.gl_format_g3(self.h, _ptr(vals), vals.size, _ptr(tok)))
        return tok

    # ---- covstats / depthwed
    def bincount(self, v: np.ndarray, lo: int, hi: int) -> np.ndarray:
        v = _as(v, np.int32)
        h = np.empty(hi - lo, np.uint64)
        self._ck(lib.gl_bincount_i32(self.h, _ptr(v), v.size, lo, hi, _ptr(h)))
        return h

    def depthwed_aggregate(self, means: np.ndarray, starts, ends, chrom_id, size: int):
        means = np.ascontiguousarray(means, np.float64)
        S, R = means.shape
        starts, ends, chrom_id = _as(starts, np.int32), _as(ends, np.int32), _as(chrom_id, np.int32)
        cap = max(R, 1)
        o_s, o_e, o_c = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32)
        out = np.empty((cap, S), np.int64)
        n = C.c_int64(0)
        self._ck(lib.gl_depthwed_aggregate(self.h, _ptr(means), S, R, _ptr(starts), _ptr(ends), _ptr(chrom_id), size,
                                           _ptr(o_s), _ptr(o_e), _ptr(o_c), _ptr(out), cap, C.byref(n)))
        k = n.value
        return o_s[:k], o_e[:k], o_c[:k], out[:k]

    def depthwed_aggregate_i32(self, depth: np.ndarray, starts, ends, chrom_id, size: int):
        """int32 form: depth = int(0.5 + mean) per (sample, row), as the reference rounds at parse time (depthwed.go:103)"""
        depth = np.ascontiguousarray(depth, np.int32)
        S, R = depth.shape
        starts, ends, chrom_id = _as(starts, np.int32), _as(ends, np.int32), _as(chrom_id, np.int32)
        cap = max(R, 1)
        o_s, o_e, o_c = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32)
        out = np.empty((cap, S), np.int32)
        n = C.c_int64(0)
        self._ck(lib.gl_depthwed_aggregate_i32(self.h, _ptr(depth), S, R, _ptr(starts), _ptr(ends), _ptr(chrom_id), size,
                                               _ptr(o_s), _ptr(o_e), _ptr(o_c), _ptr(out), cap, C.byref(n)))
        k = n.value
        return o_s[:k], o_e[:k], o_c[:k], out[:k]

    def depthwed_aggregate_i32_device(self, d_depth: DevBuf, S: int, R: int, d_grp: Optional[DevBuf], g_begin: int, g_end: int,
                                      d_out_ptr: int, d_overflow: DevBuf):
        self._ck(lib.gl_depthwed_aggregate_i32_device(self.h, d_depth.ptr, S, R, d_grp.ptr if d_grp is not None else None, g_begin, g_end,
                                                      d_out_ptr, d_overflow.ptr))

    def ipc_export(self, buf: DevBuf) -> bytes:
        h = (C.c_uint8 * 64)()
        self._ck(lib.gl_ipc_export(self.h, buf.ptr, C.cast(h, _vp)))
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = _vp()
        self._ck(lib.gl_ipc_open(self.h, C.cast(h, _vp), C.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        self._ck(lib.gl_ipc_close(self.h, ptr))

    def depthwed_aggregate_i32_p2p(self, d_depth: DevBuf, S: int, R: int, d_grp: Optional[DevBuf], g_begin: int, g_end: int, dst_ptrs,
                                   row_stride: int, col_off: int, d_overflow: DevBuf):
        arr = (_vp * len(dst_ptrs))(*dst_ptrs)
        self._ck(lib.gl_depthwed_aggregate_i32_p2p(self.h, d_depth.ptr, S, R, d_grp.ptr if d_grp is not None else None, g_begin, g_end,
                                                   C.cast(arr, _vp), len(dst_ptrs), row_stride, col_off, d_overflow.ptr))

    def allgather_device_async(self, send_ptr: int, recv_ptr: int, nbytes: int):
        self._ck(lib.gl_allgather_device_async(self.h, send_ptr, recv_ptr, nbytes))

    def comm_wait(self):
        self._ck(lib.gl_comm_wait(self.h))

    # ---- multi-GPU
    def comm_init(self, id128: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        self._ck(lib.gl_comm_init(self.h, C.cast(buf, _vp), rank, world))

    def allgather_device(self, d_send: DevBuf, d_recv: DevBuf, nbytes: int):
        self._ck(lib.gl_allgather_device(self.h, d_send.ptr, d_recv.ptr, nbytes))


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = lib.gl_comm_unique_id(C.cast(buf, _vp))
    if rc != GL_OK:
        raise GlError(rc, lib.gl_last_error(None).decode())
    return bytes(buf)
