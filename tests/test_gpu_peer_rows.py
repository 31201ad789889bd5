"""The peer row exchange of BASELINE config 3 (sharding.PeerRows: CUDA IPC + pushed device-to-device copies) with TWO
processes on the one GPU of the test box (gloo carries the handle exchange and the barriers: NCCL refuses two ranks on one
device; the 2- and 8-GPU runs use NCCL, scratch/peer_rows_check.py)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, D, q):
    import torch.distributed as dist

    from iris_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    sharding.init_from_env(backend="gloo")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ok = True
    for rep in range(2):  # second round: cached buffers and mappings are reused
        rows = sharding.make_row_exchange(n, D, dev, chunk_rows=3)
        kind = type(rows).__name__
        lo, hi = rows.lo, rows.hi
        done = 0
        for b0 in range(lo, hi, 2):   # "batches" of two rows written in place, pushed as they complete
            b1 = min(hi, b0 + 2)
            idx = torch.arange(b0, b1, device=dev, dtype=torch.float32)
            rows.local[done:done + (b1 - b0)] = idx[:, None] * 10 + torch.arange(D, device=dev)[None] + 1000 * rep
            done += b1 - b0
            rows.flush(done)
        full = rows.finish()
        expect = torch.arange(n, device=dev, dtype=torch.float32)[:, None] * 10 + torch.arange(D, device=dev)[None] + 1000 * rep
        ok = ok and bool(torch.equal(full, expect))
    q.put((rank, ok, kind))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 4, 1])
def test_peer_rows_two_processes_one_gpu(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(kind == "PeerRows" for _, _, kind in res), res   # the IPC path really ran (no silent fall-back)
