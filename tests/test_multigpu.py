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


def _nccl_worker(rank, world, uid, q):
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
    q.put((rank, full))
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
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    means, starts, ends, chrom = _problem(seed=3, S=9)
    exp = orc.depthwed(means, starts, ends, chrom, 1000)[3]
    assert np.array_equal(res[0], exp) and np.array_equal(res[1], exp)
