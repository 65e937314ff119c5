"""N>1 host logic on CPU: world_size-2 gloo processes shard a clip range with no data-path collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yad_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_bounds(n_items, rank, world)
    # stand-in for the per-rank pipeline: every clip i yields (i % 3) segments
    local_bidx = torch.cat([torch.full((i % 3,), i - lo, dtype=torch.int64) for i in range(lo, hi)] or [torch.zeros(0, dtype=torch.int64)])
    counts = parallel.gather_counts(local_bidx.numel())
    gl = parallel.globalize_batch_idxs(local_bidx, n_items, rank, world)
    t = parallel.max_over_ranks(10.0 + rank)
    q.put((rank, lo, hi, counts, gl.tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 64])
def test_two_rank_sharding(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, counts, gl, t in res:
        covered += list(range(lo, hi))
        assert t == 11.0                                   # max over ranks
        assert counts == [r[4].__len__() for r in res]     # every rank sees every count
    assert covered == list(range(n_items))                 # disjoint, complete, contiguous
    merged = [i for r in res for i in r[4]]
    assert merged == [i for i in range(n_items) for _ in range(i % 3)]


def test_shard_bounds_properties():
    for n in (0, 1, 5, 512, 65536):
        for w in (1, 2, 4, 8):
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _ar_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)          # rank r holds (r+1) * arange
    parallel.allreduce_mean_(g, bucket_bytes=1024)                    # 4 buckets
    q.put((rank, g.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce():
    """The train step's only collective: mean of the flat gradient arena over the data-parallel ranks."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ar_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(1000, dtype=torch.float32) * 1.5).tolist()
    for _, got in res:
        assert got == want
    assert parallel.allreduce_mean_(torch.ones(4)).tolist() == [1.0] * 4     # not initialised: no-op


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    # buckets in the order a backward completes them: tail of the arena first (neck + layer4), then layer3, then the head
    bar = parallel.BucketedAllReduce(g, [[(600, 800), (900, 1000)], (250, 600), None, [(0, 250), (800, 900)]])
    mid = []
    for i in range(4):
        bar.ready(i)
        mid.append(len(bar._works))
    bar.wait()
    q.put((rank, g.tolist(), mid))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_overlap_api():
    """The overlapped form of the train step's collective: each bucket is reduced as it becomes ready, wait() joins them;
    the result equals the one-shot mean over the whole arena."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(1000, dtype=torch.float32) * 1.5).tolist()
    for _, got, mid in res:
        assert got == want
        assert mid == [2, 3, 3, 5]           # one collective per contiguous run; an empty bucket launches nothing
    bar = parallel.BucketedAllReduce(torch.ones(8), [(0, 8)])
    bar.ready(0); bar.wait()                 # not initialised: no-op
    assert bar.flat.tolist() == [1.0] * 8
    with pytest.raises(ValueError):
        parallel.BucketedAllReduce(torch.ones(8), [(4, 12)])
