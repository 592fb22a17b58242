"""Multi-process paths.  CPU: world_size-2 gloo runs of the sharding / assembly logic (the per-rank compute is the
oracle there, standing in for the GPU).  GPU (needs 2 devices): the same with the CUDA kernels and the NCCL all-gather
behind the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from goleft_b200 import multigpu  # noqa: E402


def _problem(seed=0, S=7, n_rows=300):
    rng = np.random.default_rng(seed)
    starts = (np.arange(n_rows) * 250).astype(np.int32)
    ends = starts + 250
    chrom = (np.arange(n_rows) >= 200).astype(np.int32)
    starts[200:] -= starts[200]
    ends[200:] = starts[200:] + 250
    means = np.array([[float("%.4g" % x) for x in rng.gamma(9, 3.3, n_rows)] for _ in range(S)])
    return means, starts, ends, chrom


def test_shard_helpers():
    for n in range(0, 20):
        for w in (1, 2, 3, 8):
            spans = [multigpu.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= multigpu.padded_width(n, w)
    lens = [248956422, 242193529, 198295559, 64444167, 16569, 57227415, 156040895]
    a = multigpu.assign_contigs(lens, 3)
    assert sorted(i for r in a for i in r) == list(range(len(lens)))
    loads = [sum(lens[i] for i in r) for r in a]
    assert max(loads) - min(loads) < max(lens)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    from oracle import loader as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    means, starts, ends, chrom = _problem()

    def agg(m):
        return orc.depthwed(m, starts, ends, chrom, 1000)[3]

    def allgather(blk):
        t = torch.from_numpy(np.ascontiguousarray(blk))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.numpy() for o in outs]

    full = multigpu.depthwed_sharded(agg, means, rank, world, allgather)
    # depth: contigs shard with no collective; gather only the per-contig checksums to compare
    lens = [50_000, 20_001, 16_571, 4_096, 9_000]
    mine = multigpu.assign_contigs(lens, world)[rank]
    sums = {}
    for ci in mine:
        rng = np.random.default_rng(100 + ci)
        s = np.sort(rng.integers(0, lens[ci], 3000)).astype(np.int32)
        e = (s + 150).astype(np.int32)
        ws, _ = orc.window_sums(orc.pileup_diff(s, e, 0, lens[ci]), 0, lens[ci], 500)
        sums[ci] = int(ws.sum())
    gathered = [None] * world
    dist.all_gather_object(gathered, sums)
    if rank == 0:
        q.put((full, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_depthwed_and_contig_sharding():
    import torch.multiprocessing as mp
    from oracle import loader as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    means, starts, ends, chrom = _problem()
    assert np.array_equal(full, orc.depthwed(means, starts, ends, chrom, 1000)[3])
    merged = {}
    for g in gathered:
        assert not set(g) & set(merged)
        merged.update(g)
    assert sorted(merged) == [0, 1, 2, 3, 4]
    for ci, tot in merged.items():
        assert tot == 3000 * 150 - 0 or tot <= 3000 * 150        # clipped at the contig end


def _nccl_worker(rank, world, uid, q, shared=None, bar=None):
    from goleft_b200 import capi
    from oracle import loader as orc  # noqa: F401
    c = capi.Ctx(rank)
    c.comm_init(uid, rank, world)
    means, starts, ends, chrom = _problem(seed=3, S=9)

    def agg(m):
        return c.depthwed_aggregate(m, starts, ends, chrom, 1000)[3]

    def allgather(blk):
        send = c.dev_array(blk)
        recv = c.dev_empty(blk.nbytes * world)
        c.allgather_device(send, recv, blk.nbytes)
        flat = recv.download(np.int64, blk.size * world)
        send.free(); recv.free()
        return [flat[r * blk.size:(r + 1) * blk.size].reshape(blk.shape) for r in range(world)]

    full = multigpu.depthwed_sharded(agg, means, rank, world, allgather)
    # the int32, row-chunked, overlapped form (what bench.py runs at 500 x 6.18 M): one row per output line
    S2, R2, n_chunks = 11, 5000, 4
    lo, hi = multigpu.shard_range(S2, rank, world)
    width = multigpu.padded_width(S2, world)
    depth = ((np.arange(R2, dtype=np.int64)[None, :] * 7919 + np.arange(S2)[:, None] * 104729) % 1000).astype(np.int32)
    loc = np.zeros((width, R2), np.int32)
    loc[: hi - lo] = depth[lo:hi]
    d_depth = c.dev_array(loc)
    d_local, d_all, d_ovf = c.dev_empty(R2 * width * 4), c.dev_empty(R2 * width * 4 * world), c.dev_array(np.zeros(4, np.int32))
    multigpu.depthwed_gather_overlapped(c, d_depth, width, R2, world, d_local, d_all, d_ovf, n_chunks)
    got = multigpu.depthwed_assemble_chunked(d_all.download(np.int32, R2 * width * world), R2, width, world, S2, n_chunks)
    ok32 = bool(np.array_equal(got, depth.T)) and int(d_ovf.download(np.int32, 1)[0]) == 0
    # the fused form: every rank's kernel stores its columns into both ranks' row-major matrices over peer memory
    if shared is not None:
        wpad = (width + 3) // 4 * 4
        row_stride = wpad * world
        d_full = c.dev_array(np.full(R2 * row_stride, -1, np.int32))
        shared[rank * 64:(rank + 1) * 64] = list(c.ipc_export(d_full))
        bar.wait()
        ptrs = [d_full.ptr if r == rank else c.ipc_open(bytes(shared[r * 64:(r + 1) * 64])) for r in range(world)]
        c.depthwed_aggregate_i32_p2p(d_depth, width, R2, None, 0, R2, ptrs, row_stride, rank * wpad, d_ovf)
        c.sync()
        bar.wait()
        m = d_full.download(np.int32, R2 * row_stride).reshape(R2, row_stride)
        for r in range(world):
            rlo, rhi = multigpu.shard_range(S2, r, world)
            ok32 = ok32 and bool(np.array_equal(m[:, r * wpad:r * wpad + (rhi - rlo)], depth[rlo:rhi].T))
        bar.wait()
        for r in range(world):
            if r != rank:
                c.ipc_close(ptrs[r])
    q.put((rank, full, ok32))
    c.close()


@pytest.mark.gpu
def test_nccl_world2_depthwed_allgather():
    from goleft_b200 import capi
    from oracle import loader as orc
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    uid = capi.comm_unique_id()
    q = ctx.Queue()
    shared, bar = ctx.Array("B", 128), ctx.Barrier(2)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, uid, q, shared, bar)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    res = {g[0]: g[1] for g in got}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    means, starts, ends, chrom = _problem(seed=3, S=9)
    exp = orc.depthwed(means, starts, ends, chrom, 1000)[3]
    assert np.array_equal(res[0], exp) and np.array_equal(res[1], exp)
    assert all(g[2] for g in got)                               # int32 chunked + overlapped gather, verified on every rank


def test_chunked_assembly_layout():
    """CPU: the [chunk][rank][rows][width] layout of the overlapped gather reassembles to R x S"""
    S, R, world, n_chunks = 11, 103, 4, 5
    width = multigpu.padded_width(S, world)
    depth = (np.arange(S)[:, None] * 1000 + np.arange(R)[None, :]).astype(np.int32)
    flat = np.zeros(R * width * world, np.int32)
    for g0, g1 in multigpu.chunk_bounds(R, n_chunks):
        for r in range(world):
            lo, hi = multigpu.shard_range(S, r, world)
            blk = np.zeros((g1 - g0, width), np.int32)
            blk[:, : hi - lo] = depth[lo:hi, g0:g1].T
            o = g0 * width * world + r * (g1 - g0) * width
            flat[o:o + blk.size] = blk.ravel()
    assert np.array_equal(multigpu.depthwed_assemble_chunked(flat, R, width, world, S, n_chunks), depth.T)
