#!/usr/bin/env python
"""How the one-call int32 entry (gl_depth_bed_contig) spends its time on chr20 30x for both transports (GL_BED_PACK=0 plain
int32, =16 fixed-block packed16), plus the host packer alone at several thread counts.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))
from goleft_b200 import capi  # noqa: E402
import glsynth  # noqa: E402


def main():
    node, n_cpus = capi.bind_numa_for_device(0, 0, 1)
    ctx = capi.Ctx(0)
    L = int(os.environ.get("PROBE_LEN", glsynth.CHR20_LEN))
    s, e = glsynth.segments(L, 19, threads=n_cpus, alloc=ctx.pinned_empty)
    n = s.size
    out = {"numa_node": node, "cpus": n_cpus, "pool": int(capi.lib.glhost_pool_size()), "segments": int(n)}
    n_win = (L - 1) // 500 + 1
    o_hd = ctx.pinned_empty(int(capi.lib.gl_depth_text_bound(b"chr20", n_win)), np.uint8)
    o_ca = ctx.pinned_empty(1 << 20, np.uint8)
    for mode in (() if os.environ.get("PROBE_CALLS_ONLY") == "0" else ("0", "16", "0", "16")):
        os.environ["GL_BED_PACK"] = mode
        for _ in range(3):
            ctx.depth_bed_contig("chr20", L, s, e, 500, 4, 0, 10_000_000, out=(o_hd, o_ca), raw=True)
        ts, ph, pk = [], [], []
        for _ in range(int(os.environ.get("PROBE_REPS", "10"))):
            ctx.flush_l2(); ctx.sync()
            t0 = time.perf_counter()
            ctx.depth_bed_contig("chr20", L, s, e, 500, 4, 0, 10_000_000, out=(o_hd, o_ca), raw=True)
            ts.append((time.perf_counter() - t0) * 1e3)
            ph.append([x * 1e3 for x in ctx.depth_transport_phases()])
            pk.append(ctx.depth_transport_stats()[1] * 1e3)
        st = ctx.depth_transport_stats()
        ctx.profile_enable(True); ctx.profile_read()
        ctx.depth_bed_contig("chr20", L, s, e, 500, 4, 0, 10_000_000, out=(o_hd, o_ca), raw=True)
        kms = {}
        for nm, t in ctx.profile_read():
            kms[nm] = kms.get(nm, 0.0) + t
        ctx.profile_enable(False)
        out.setdefault("kernel_ms_GL_BED_PACK_" + mode, []).append(kms)
        out.setdefault("call_ms_GL_BED_PACK_" + mode, []).append(
            {"median": float(np.median(ts)), "min": float(np.min(ts)), "max": float(np.max(ts)), "all": [round(x, 2) for x in ts],
             "phases_all": [[round(y, 2) for y in x] for x in ph], "path": ctx.depth_last_path(),
             "transport": st[0], "h2d_bytes": st[2], "escaped": st[3], "host_pack_ms_median": float(np.median(pk)),
             "phases_ms_median": [float(x) for x in np.median(np.array(ph), axis=0)]})
    os.environ.pop("GL_BED_PACK", None)
    if os.environ.get("PROBE_CALLS_ONLY") == "1":
        print(json.dumps(out)); return
    # the packer alone, pinned output
    nb = (n + 255) // 256
    a = ctx.pinned_empty(nb, np.int32); o = ctx.pinned_empty(nb * 256, np.uint16); ln = ctx.pinned_empty(nb * 256, np.uint16)
    es = np.empty(n // 16 + 4096, np.int32); ee = np.empty_like(es)
    m = C.c_int64(0)
    out["packer_alone_ms"] = {}
    for thr in (8, 16, 32, 64, 0):
        ts = []
        for _ in range(8):
            t0 = time.perf_counter()
            capi.lib.gl_pack_segments16_fixed_mt(s.ctypes.data, e.ctypes.data, n, thr, a.ctypes.data, o.ctypes.data, ln.ctypes.data,
                                                 es.ctypes.data, ee.ctypes.data, es.size, C.byref(m))
            ts.append((time.perf_counter() - t0) * 1e3)
        out["packer_alone_ms"][str(thr)] = {"min": float(np.min(ts)), "median": float(np.median(ts)), "GBps_in": n * 8 / (np.min(ts) * 1e-3) / 1e9}
    # the packer in 8 block ranges back to back (what the call does), host only ...
    L = capi.lib
    L.gl_pack_segments16_fixed_range_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32] + [C.c_void_p] * 5 + [C.c_int64, C.POINTER(C.c_int64)]

    def ranges(K, thr=0):
        m.value = 0
        t0 = time.perf_counter()
        per = []
        for c in range(K):
            t1 = time.perf_counter()
            L.gl_pack_segments16_fixed_range_mt(s.ctypes.data, e.ctypes.data, n, nb * c // K, nb * (c + 1) // K, thr, a.ctypes.data, o.ctypes.data,
                                                ln.ctypes.data, es.ctypes.data, ee.ctypes.data, es.size, C.byref(m))
            per.append((time.perf_counter() - t1) * 1e3)
        return (time.perf_counter() - t0) * 1e3, per
    for thr in (0, 32, 16):
        out["packer_ranges_host_only_ms_threads_%d" % thr] = {str(K): min(ranges(K, thr)[0] for _ in range(6)) for K in (1, 2, 4, 8, 16)}
    out["packer_8_ranges_host_only_ms"] = min(ranges(8)[0] for _ in range(8))
    out["packer_8_ranges_host_only_per_range_ms"] = ranges(8)[1]
    # ... and the same while an unrelated H2D copy out of pinned memory is running (does the DMA slow the CPU's streaming pass?)
    try:
        import torch
        big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        st = torch.cuda.Stream()
        res = []
        for _ in range(int(os.environ.get("PROBE_DMA_REPS", "4"))):
            with torch.cuda.stream(st):
                dev.copy_(big, non_blocking=True)
            t_all, per = ranges(8)
            busy = not st.query()
            st.synchronize()
            res.append({"ms": t_all, "copy_still_running_after": bool(busy), "per_range_ms": per})
        out["packer_8_ranges_during_h2d_dma"] = res
        # and the other way round: H2D time of 1 GiB alone and while the packer loops
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            ev0.record(st); dev.copy_(big, non_blocking=True); ev1.record(st)
        st.synchronize()
        out["h2d_1GiB_alone_ms"] = ev0.elapsed_time(ev1)
        with torch.cuda.stream(st):
            ev0.record(st); dev.copy_(big, non_blocking=True); ev1.record(st)
        while not st.query():
            ranges(8)
        st.synchronize()
        out["h2d_1GiB_while_packing_ms"] = ev0.elapsed_time(ev1)
    except Exception as ex:
        out["dma_interference_error"] = str(ex)[:200]
    # plain H2D of the same bytes for scale
    d = ctx.dev_empty(n * 8)
    ts = []
    for _ in range(5):
        ctx.sync(); t0 = time.perf_counter()
        capi.lib.gl_memcpy_h2d(ctx.h, d.ptr, s.ctypes.data, n * 4); capi.lib.gl_memcpy_h2d(ctx.h, d.ptr + n * 4, e.ctypes.data, n * 4)
        ts.append((time.perf_counter() - t0) * 1e3)
    out["plain_h2d_ms"] = {"min": float(np.min(ts)), "GBps": n * 8 / (np.min(ts) * 1e-3) / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
