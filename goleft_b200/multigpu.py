"""Multi-GPU host logic (one process per GPU).

depth    : contigs shard across ranks, no data-path collective (the reference's unit of independence is the
           10 Mb chunk / contig, depth/depth.go:132,150-154).
depthwed : samples (columns) shard across ranks; each rank aggregates its columns on its GPU; ONE all-gather of
           equal-sized row-major blocks assembles the n-sites x n-samples matrix on every rank
           (depthwed/depthwed.go:64-71 writes one TSV row per site with all samples).
The collective is injected (`allgather(block) -> list of blocks`) so the same code runs over NCCL
(Ctx.allgather_device) on GPUs and over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced: the first n % world shards get one extra item"""
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def padded_width(n: int, world: int) -> int:
    return (n + world - 1) // world


def assign_contigs(lengths: Sequence[int], world: int) -> List[List[int]]:
    """longest-processing-time-first assignment of contigs to ranks (deterministic)"""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += lengths[i]
    return [sorted(x) for x in out]


def depthwed_pad_block(local: np.ndarray, width: int) -> np.ndarray:
    """local: n_out x S_local int64 -> n_out x width (zero padded) so that every rank sends the same size"""
    n_out, s = local.shape
    blk = np.zeros((n_out, width), np.int64)
    blk[:, :s] = local
    return blk


def depthwed_assemble(blocks: Sequence[np.ndarray], S: int, world: int) -> np.ndarray:
    """blocks[r]: n_out x width from rank r -> n_out x S, samples in their original order"""
    cols = []
    for r, b in enumerate(blocks):
        lo, hi = shard_range(S, r, world)
        cols.append(np.asarray(b)[:, : hi - lo])
    return np.concatenate(cols, axis=1)


def depthwed_sharded(aggregate_local: Callable[[np.ndarray], np.ndarray], means: np.ndarray, rank: int, world: int,
                     allgather: Callable[[np.ndarray], List[np.ndarray]]) -> np.ndarray:
    """means: S x R (all samples; each rank only touches its own rows). aggregate_local(S_local x R) -> n_out x S_local."""
    S = means.shape[0]
    lo, hi = shard_range(S, rank, world)
    local = aggregate_local(means[lo:hi])
    blk = depthwed_pad_block(local, padded_width(S, world))
    return depthwed_assemble(allgather(blk), S, world)


# ---- int32 matrix, row-chunked, gather overlapped with aggregation (BASELINE config 4: 500 samples x 6.18 M windows)
def chunk_bounds(n_rows: int, n_chunks: int) -> List[Tuple[int, int]]:
    n_chunks = max(1, min(n_chunks, n_rows)) if n_rows else 1
    return [(n_rows * c // n_chunks, n_rows * (c + 1) // n_chunks) for c in range(n_chunks)]


def depthwed_gather_overlapped(ctx, d_depth, width: int, R: int, world: int, d_local, d_all, d_overflow, n_chunks: int = 8):
    """Every rank: aggregate its `width` sample columns chunk by chunk (depthwed_i32_kernel on the compute stream) and
    all-gather each finished chunk on the communication stream while the next one is being aggregated.
    d_all layout: [chunk][rank][rows of the chunk][width] int32.  Returns after both streams have drained."""
    for g0, g1 in chunk_bounds(R, n_chunks):
        ctx.depthwed_aggregate_i32_device(d_depth, width, R, None, g0, g1, d_local.ptr + g0 * width * 4, d_overflow)
        ctx.allgather_device_async(d_local.ptr + g0 * width * 4, d_all.ptr + g0 * width * 4 * world, (g1 - g0) * width * 4)
    ctx.comm_wait()


def depthwed_assemble_chunked(flat: np.ndarray, R: int, width: int, world: int, S: int, n_chunks: int) -> np.ndarray:
    """flat: the int32 contents of d_all -> R x S with the samples in their original order"""
    out = np.empty((R, S), np.int32)
    for g0, g1 in chunk_bounds(R, n_chunks):
        base = g0 * width * world
        for r in range(world):
            lo, hi = shard_range(S, r, world)
            blk = flat[base + r * (g1 - g0) * width: base + (r + 1) * (g1 - g0) * width].reshape(g1 - g0, width)
            out[g0:g1, lo:hi] = blk[:, : hi - lo]
    return out
