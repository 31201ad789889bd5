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


class RowGatherer:
    """Row-sharded [n_total, D] matrix that every rank ends up holding in full (BASELINE config 3: the feature rows of
    the iris classifier).  Rank r produces rows [lo, hi) = shard_range(n_total, r, world) by writing them IN PLACE into
    `local` (a view of the final buffer: no padded copy); `flush(rows_done)` all-gathers every complete chunk of
    `chunk_rows` rows on a side stream, so the exchange of chunk k rides under the computation of chunk k+1 (NCCL over
    NVLink / NVSwitch; on CPU tensors / gloo the same calls run synchronously); `finish()` gathers the rest and
    returns the [n_total, D] matrix."""

    def __init__(self, n_total: int, D: int, device, dtype=torch.float32, chunk_rows: int = 128):
        self.dist = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.dist else 1
        self.rank = dist.get_rank() if self.dist else 0
        self.n_total, self.D = n_total, D
        self.per = (n_total + self.world - 1) // self.world if n_total else 0
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        self.device = torch.device(device)
        self.full = torch.empty(self.world * self.per, D, device=self.device, dtype=dtype)
        self.local = self.full[self.rank * self.per: self.rank * self.per + (self.hi - self.lo)]
        self.chunk = max(1, int(chunk_rows))
        self.sent = 0                      # rows of every rank's shard already exchanged (same schedule on all ranks)
        self.cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(self.device) if (self.cuda and self.dist) else None

    def _gather(self, r0: int, r1: int):
        """Exchange rows [r0, r1) of every rank's (padded) shard."""
        n = r1 - r0
        mine = self.full[self.rank * self.per + r0: self.rank * self.per + r1]
        view3 = self.full.view(self.world, self.per, self.D)
        if self.comm is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                tmp = torch.empty(self.world, n, self.D, device=self.device, dtype=self.full.dtype)
                dist.all_gather_into_tensor(tmp.view(self.world * n, self.D), mine)
                view3[:, r0:r1].copy_(tmp)
        else:
            tmp = torch.empty(self.world * n, self.D, device=self.device, dtype=self.full.dtype)
            dist.all_gather_into_tensor(tmp, mine.contiguous())
            view3[:, r0:r1].copy_(tmp.view(self.world, n, self.D))

    def flush(self, rows_done: int):
        """Call after `rows_done` local rows have been enqueued: complete chunks go out.  Every rank must reach the same
        chunk boundaries, so only rows that EVERY rank owns (the padded tail of the last shard excluded) are sent here."""
        if not self.dist:
            return
        # collectives must match across ranks: chunks are cut only from the rows EVERY shard has (the last shard is the
        # shortest), so each rank issues the same floor(common / chunk) gathers here, whatever its own progress
        common = max(0, self.n_total - (self.world - 1) * self.per)
        limit = min(rows_done, common)
        while self.sent + self.chunk <= limit:
            self._gather(self.sent, self.sent + self.chunk)
            self.sent += self.chunk

    def finish(self) -> torch.Tensor:
        if not self.dist:
            return self.local
        if self.sent < self.per:
            self._gather(self.sent, self.per)
            self.sent = self.per
        if self.comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)
        return self.full[: self.n_total]


def sharded_map(n_total: int, fn: Callable[[int, int], torch.Tensor]) -> torch.Tensor:
    """Run fn(lo, hi) -> [hi-lo, D] on this rank's shard and all-gather the rows of every rank."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    rows = fn(lo, hi)
    assert rows.shape[0] == hi - lo
    return all_gather_rows(rows, n_total)
