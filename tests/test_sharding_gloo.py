"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous shards, padded all-gather, no collective
anywhere else."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import iris_b200
from iris_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 10000, 10001):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    assert sharding.shard_range(10000, 7, 8) == (8750, 10000)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, D, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = sharding.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)

    def fn(lo, hi):  # stands in for the per-shard GPU feature extraction
        idx = torch.arange(lo, hi, dtype=torch.float32)
        return idx[:, None] * 10 + torch.arange(D, dtype=torch.float32)[None]

    full = sharding.sharded_map(n, fn)
    expect = torch.arange(n, dtype=torch.float32)[:, None] * 10 + torch.arange(D, dtype=torch.float32)[None]
    q.put((rank, bool(torch.equal(full, expect)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 1])
def test_sharded_map_all_gather_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok and shape == (n, 5) for _, ok, shape in res)


def _worker_rows(rank, world, port, n, D, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sharding.init_from_env(backend="gloo")
    rows = sharding.RowGatherer(n, D, "cpu", chunk_rows=2)
    assert (rows.lo, rows.hi) == sharding.shard_range(n, rank, world) and rows.local.shape == (rows.hi - rows.lo, D)
    done = 0
    for i in range(rows.lo, rows.hi, 3):     # "batches" of 3 rows written in place, chunks of 2 gathered as they complete
        j = min(rows.hi, i + 3)
        idx = torch.arange(i, j, dtype=torch.float32)
        rows.local[i - rows.lo: j - rows.lo] = idx[:, None] * 10 + torch.arange(D, dtype=torch.float32)[None]
        done = j - rows.lo
        rows.flush(done)
    full = rows.finish()
    expect = torch.arange(n, dtype=torch.float32)[:, None] * 10 + torch.arange(D, dtype=torch.float32)[None]
    q.put((rank, bool(torch.equal(full, expect)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 1, 13])
def test_row_gatherer_chunked_world2(n):
    """The in-place, chunk-wise all-gather of the feature rows (ragged last shard, a rank with no rows, chunk boundaries
    that do not divide the shard) returns the full matrix on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rows, args=(r, 2, port, n, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and shape == (n, 5) for _, ok, shape in res)


def test_row_gatherer_single_process():
    rows = sharding.RowGatherer(5, 3, "cpu")
    rows.local[:] = 1.0
    rows.flush(5)
    assert torch.equal(rows.finish(), torch.ones(5, 3))
