"""Times gl_depth_bed_contig (host int32 segments -> BED bytes) on synthetic 30x chr20, pinned host buffers.
GL_BED_PACK=0 uploads int32, 16 (default) repacks to packed16 pipelined with the upload, 1 packs to packed8;
GL_THREADS sets the pool size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from goleft_b200 import capi, synth
import ctypes as C

L = synth.CHR20_LEN
s, e = synth.chr20_like()
W, STEP = 500, 10_000_000
ctx = capi.Ctx(0)
hs, he = ctx.pinned_empty(s.size, np.int32), ctx.pinned_empty(e.size, np.int32)
hs[:] = s; he[:] = e
n_win = (L - 1) // W + 1
hd = ctx.pinned_empty(int(capi.lib.gl_depth_text_bound(b"chr20", n_win)), np.uint8)
ca = ctx.pinned_empty(1 << 20, np.uint8)
for thr in [int(x) for x in os.environ.get("PROBE_THREADS", "0").split(",")]:
    ts = []
    for it in range(12):
        ctx.flush_l2(); ctx.sync()
        time.sleep(float(os.environ.get("PROBE_SLEEP", "0")))
        t0 = time.perf_counter()
        hl, cl = ctx.depth_bed_contig("chr20", L, hs, he, W, 4, 0, STEP, threads=thr, out=(hd, ca), raw=True)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("threads", thr, "GL_BED_PACK", os.environ.get("GL_BED_PACK"), "path", ctx.depth_last_path(), "text bytes", hl, cl,
          "ms: min %.3f med %.3f max %.3f" % (min(ts[2:]), float(np.median(ts[2:])), max(ts[2:])), flush=True)
cap = s.size // 200 + 1000
a = np.zeros(cap, np.int32); o = np.zeros(cap * 256, np.uint16); ln = np.zeros(cap * 256, np.uint16)
nb = C.c_int64(0)
for thr in (0, 32, 16, 8, 1):
    ts = []
    for it in range(8):
        time.sleep(float(os.environ.get("PROBE_SLEEP", "0")))
        t0 = time.perf_counter()
        capi.lib.gl_pack_segments16_mt(capi._ptr(hs), capi._ptr(he), s.size, thr, capi._ptr(a), capi._ptr(o), capi._ptr(ln), cap, C.byref(nb))
        ts.append((time.perf_counter() - t0) * 1e3)
    print("pack16_mt threads", thr, "ms min %.3f med %.3f" % (min(ts[2:]), float(np.median(ts[2:]))), "cpus", os.cpu_count(), flush=True)
