#!/usr/bin/env python
"""e2e timing probe for gl_depth_region_packed8: device-event time and host wall time per call (GL_P8_CHUNKS=k)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from goleft_b200 import capi, synth

L = synth.CHR20_LEN
s, e = synth.chr20_like()
a, d, ln = capi.pack_segments8(s, e)
ctx = capi.Ctx(0)
hs = []
for x in (a, d, ln):
    h = ctx.pinned_empty(x.size, x.dtype); h[:] = x; hs.append(h)
nw = (L - 1) // 500 + 1
cap = L // 16 + 4096
out = (ctx.pinned_empty(nw, np.int64), ctx.pinned_empty(cap, np.int32), ctx.pinned_empty(cap, np.uint8))
for _ in range(3):
    ctx.depth_region_packed8(0, L, hs[0], hs[1], hs[2], 500, 4, 0, 10_000_000, out=out)
dev, wall = [], []
for _ in range(20):
    ctx.flush_l2(); ctx.sync()
    t0 = time.perf_counter(); ctx.timer_start()
    ctx.depth_region_packed8(0, L, hs[0], hs[1], hs[2], 500, 4, 0, 10_000_000, out=out)
    dev.append(ctx.timer_stop_ms()); wall.append((time.perf_counter() - t0) * 1e3)
print("chunks", os.environ.get("GL_P8_CHUNKS", "8"), "dev ms %.3f  wall ms %.3f  (min dev %.3f)" % (np.mean(dev), np.mean(wall), np.min(dev)))
# raw H2D of the same bytes for reference
db = [ctx.dev_empty(x.nbytes) for x in hs]
import ctypes as C
for rep in range(3):
    ctx.sync(); t0 = time.perf_counter()
    for h, b in zip(hs, db):
        b.upload(h)
    ctx.sync(); t = (time.perf_counter() - t0) * 1e3
print("plain upload of %d bytes: %.3f ms" % (sum(x.nbytes for x in hs), t))
