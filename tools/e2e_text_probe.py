"""Times gl_depth_bed_contig (host int32 segments -> BED bytes) on synthetic 30x chr20, pinned host buffers.
GL_BED_PACK=0 uploads int32, =1 packs on the host threads first; GL_THREADS sets the pool size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from goleft_b200 import capi, synth

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
        t0 = time.perf_counter()
        hl, cl = ctx.depth_bed_contig("chr20", L, hs, he, W, 4, 0, STEP, threads=thr, out=(hd, ca), raw=True)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("threads", thr, "GL_BED_PACK", os.environ.get("GL_BED_PACK"), "path", ctx.depth_last_path(), "text bytes", hl, cl,
          "ms: min %.3f med %.3f" % (min(ts[2:]), float(np.median(ts[2:]))), flush=True)
# pack alone
import ctypes as C
cap = s.size // 40 + 1024
a = np.empty(cap, np.int32); d = np.empty(cap * 64, np.uint8); ln = np.empty(cap * 64, np.uint8)
a[:] = 0; d[:] = 0; ln[:] = 0
nb = C.c_int64(0)
for thr in (0, 64, 32, 16, 8):
    ts = []
    for it in range(8):
        t0 = time.perf_counter()
        capi.lib.gl_pack_segments8_mt(capi._ptr(hs), capi._ptr(he), s.size, thr, capi._ptr(a), capi._ptr(d), capi._ptr(ln), cap, C.byref(nb))
        ts.append((time.perf_counter() - t0) * 1e3)
    print("pack_mt threads", thr, "ms min %.3f med %.3f" % (min(ts[2:]), float(np.median(ts[2:]))), "cpus", os.cpu_count(), flush=True)
ctx.profile_enable(True); ctx.profile_read()
ctx.depth_bed_contig("chr20", L, hs, he, W, 4, 0, STEP, out=(hd, ca), raw=True)
print(ctx.profile_read())
