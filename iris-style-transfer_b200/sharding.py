"""Multi-GPU plumbing: one process per GPU (torch.distributed over NCCL/NVLink on the box, gloo in the CPU
tests).  Every eye image is an independent problem, so the image list is cut into
contiguous shards -- rank r owns [r*ceil(n/G), min(n, (r+1)*ceil(n/G))) -- with NO collective on the NST
inner loop; the only exchange is ONE all-gather of the per-eye feature vectors for the iris classifier
(BASELINE config 3), padded to uniform counts."""
from __future__ import annotations

import os
from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def init_from_env(backend: str = None):
    """Join the process group the launcher (torchrun) describes; single-process when WORLD_SIZE is absent."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if backend == "nccl" else None)
    return rank, local, world


def all_gather_rows(local_rows: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather row shards produced with shard_range (uniform padded counts, last shard trimmed).
    local_rows: [n_local, D] on the rank's device -> [n_total, D] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local_rows.shape[0] == n_total
        return local_rows
    world = dist.get_world_size()
    per = (n_total + world - 1) // world
    D = local_rows.shape[1]
    padded = local_rows.new_zeros(per, D)
    padded[: local_rows.shape[0]] = local_rows
    out = local_rows.new_empty(world * per, D)
    dist.all_gather_into_tensor(out, padded)   # NCCL all-gather over NVLink / NVSwitch on the B200 box
    return out[:n_total]


def sharded_map(n_total: int, fn: Callable[[int, int], torch.Tensor]) -> torch.Tensor:
    """Run fn(lo, hi) -> [hi-lo, D] on this rank's shard and all-gather the rows of every rank."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    rows = fn(lo, hi)
    assert rows.shape[0] == hi - lo
    return all_gather_rows(rows, n_total)
