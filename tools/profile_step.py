"""A few resident chr20 passes (packed8, int32, general) for ncu to capture:
    ncu --set full --clock-control none --import-source on -k regex:depth_fused8 -s 2 -c 1 -o gpurun_out/prof python tools/profile_step.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth"))
from goleft_b200 import capi
import glsynth

L = glsynth.CHR20_LEN
s, e = glsynth.segments(L, 19)
ctx = capi.Ctx(0)
qa, qd, ql = capi.pack_segments8(s, e, threads=0)
d_a, d_d, d_l = ctx.dev_array(qa), ctx.dev_array(qd), ctx.dev_array(ql)
d_s, d_e = ctx.dev_array(s), ctx.dev_array(e)
which = sys.argv[1] if len(sys.argv) > 1 else "p8"
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    ctx.flush_l2()
    ctx.depth_begin(0, L)
    if which == "p8":
        ctx.depth_add_segments_packed8_device(d_a, d_d, d_l, qa.size)
    else:
        ctx.depth_add_segments_device(d_s, d_e, s.size)
    if which == "general":
        ctx.depth_set_path(4)
    if which == "hbm":
        ctx.depth_set_path(2)
    ctx.depth_reduce(500, 4, 0, 10_000_000)
print("path", ctx.depth_last_path(), ctx.depth_result_sizes())
