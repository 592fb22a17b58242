#!/usr/bin/env python
"""depthwed matrix assembly across GPUs (BASELINE config 4 at reduced row count): every rank aggregates its sample shard on
its own GPU (depthwed_kernel), then ONE NCCL all-gather over NVLink (gl_allgather_device) gives every rank all row-major
blocks.  Launch with torchrun; torch.distributed (gloo) is used only to hand the NCCL unique id to the other ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_depthwed_nccl.py
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=500)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist.init_process_group("gloo")
    from goleft_b200 import capi, multigpu
    ctx = capi.Ctx(local)
    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)

    S, R = args.samples, args.rows
    lo, hi = multigpu.shard_range(S, rank, world)
    width = multigpu.padded_width(S, world)
    rng = np.random.default_rng(1234)
    base = rng.gamma(9, 3.3, R)
    d_means = ctx.dev_empty(width * R * 8)
    for k in range(width):                                   # rows of samples lo..hi (padding samples repeat the last one)
        col = base + float(min(lo + k, S - 1))
        capi.lib.gl_memcpy_h2d(ctx.h, d_means.ptr + k * R * 8, col.ctypes.data, R * 8)
    d_local = ctx.dev_empty(R * width * 8)
    d_all = ctx.dev_empty(R * width * 8 * world)
    t_agg, t_gather = [], []
    for it in range(args.reps + 1):
        dist.barrier()
        ctx.timer_start()
        ctx.depthwed_aggregate_device(d_means, width, R, None, R, d_local)
        a = ctx.timer_stop_ms()
        dist.barrier()
        ctx.timer_start()
        ctx.allgather_device(d_local, d_all, R * width * 8)
        g = ctx.timer_stop_ms()
        if it:
            t_agg.append(a); t_gather.append(g)
    # check a few rows on every rank: block r holds samples shard_range(S, r, world)
    flat = d_all.download(np.int64, 64 * width * world * 0 + R * width * world)
    ok = True
    for r in range(world):
        rlo, rhi = multigpu.shard_range(S, r, world)
        blk = flat[r * R * width:(r + 1) * R * width].reshape(R, width)
        for k in (0, rhi - rlo - 1):
            exp = (0.5 + base[:1000] + float(rlo + k)).astype(np.int64)
            ok &= bool(np.array_equal(blk[:1000, k], exp))
    tm = np.array([np.mean(t_agg), np.mean(t_gather)])
    import torch
    tt = torch.tensor(tm)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if rank == 0:
        bytes_per_rank = R * width * 8
        total = bytes_per_rank * world
        out = {"world": world, "samples": S, "rows": R, "block_bytes_per_rank": bytes_per_rank,
               "aggregate_ms": float(tt[0]), "allgather_ms": float(tt[1]),
               "allgather_algbw_gbs": total / (float(tt[1]) * 1e-3) / 1e9,
               "allgather_busbw_gbs": total * (world - 1) / world / (float(tt[1]) * 1e-3) / 1e9,
               "cells_per_s_end_to_end": S * R / ((float(tt[0]) + float(tt[1])) * 1e-3), "verified": all(oks)}
        print(json.dumps(out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
