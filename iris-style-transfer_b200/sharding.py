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


class PeerRows:
    """The same contract as RowGatherer (`local`, `flush(rows_done)`, `finish()`), without a collective kernel: every rank
    maps the row matrices of the other ranks of the box into its own address space (CUDA IPC, peer access over NVLink /
    NVSwitch: isx_ipc_export / isx_ipc_open) and PUSHES each finished batch of rows straight into them with device-to-device
    copies on a side stream -- copy engines only.  An NCCL all-gather kernel needs SMs, and the persistent conv CTAs own every
    SM with >= 204 KB of shared memory each: at 8 GPUs the chunked NCCL exchange cost 9 % of the feature throughput
    (profiles/r02_scale_n8.txt), the pushed rows cost nothing measurable.  NCCL still carries the handle exchange and the two
    barriers.  Buffers and mappings are cached per (n_total, D): the matrix returned by finish() stays valid until the next
    gather of the same shape on this process group."""

    _cache = {}

    @classmethod
    def get(cls, n_total: int, D: int, device, dtype=torch.float32):
        key = (n_total, D, str(torch.device(device)), dtype, dist.get_world_size())
        obj = cls._cache.get(key)
        if obj is None:
            obj = cls._cache[key] = cls(n_total, D, device, dtype)
        return obj

    def __init__(self, n_total: int, D: int, device, dtype=torch.float32):
        import ctypes

        from . import _lib

        assert dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.n_total, self.D = n_total, D
        self.per = (n_total + self.world - 1) // self.world
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        self.device = torch.device(device)
        self.esz = torch.empty(0, dtype=dtype).element_size()
        self.full = torch.empty(self.world * self.per, D, device=self.device, dtype=dtype)
        self.local = self.full[self.rank * self.per: self.rank * self.per + (self.hi - self.lo)]
        self.comm = torch.cuda.Stream(self.device)
        self._lib = _lib
        handle = (ctypes.c_ubyte * 64)()
        off = ctypes.c_int64()
        mine = None
        try:
            with torch.cuda.device(self.device):
                _lib.call("isx_ipc_export", ctypes.c_void_p(self.full.data_ptr()), handle, ctypes.byref(off))
            mine = (bytes(handle), int(off.value))
        except Exception as e:  # keep the collective below matched on every rank
            self.error = e
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self.peer_ptr = [None] * self.world
        self._bases = []
        self.ok = all(x is not None for x in everyone)
        if self.ok:
            try:
                with torch.cuda.device(self.device):
                    for r, (h, o) in enumerate(everyone):
                        if r == self.rank:
                            continue
                        base = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                        _lib.call("isx_ipc_open", buf, ctypes.byref(base))
                        self._bases.append(base.value)
                        self.peer_ptr[r] = base.value + o
            except Exception as e:
                self.ok = False
                self.error = e
        self.sent = 0

    def _begin(self):
        """Every rank must be done with the previous result before anybody overwrites it."""
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier()
        self.sent = 0

    def flush(self, rows_done: int):
        """Rows [sent, rows_done) of this rank's shard have been enqueued on the current stream: push them to every peer."""
        if rows_done <= self.sent:
            return
        import ctypes

        r0, r1 = self.sent, rows_done
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        off = (self.rank * self.per + r0) * self.D * self.esz
        nbytes = (r1 - r0) * self.D * self.esz
        src = self.full.data_ptr() + off
        with torch.cuda.device(self.device):
            self.comm.wait_event(ev)
            for k in range(1, self.world):        # start with a different peer on every rank: no hot receiver
                r = (self.rank + k) % self.world
                self._lib.call("isx_copy_d2d_async", ctypes.c_void_p(self.peer_ptr[r] + off), ctypes.c_void_p(src),
                               self._lib.i64(nbytes), ctypes.c_void_p(self.comm.cuda_stream))
        self.sent = rows_done

    def finish(self) -> torch.Tensor:
        self.flush(self.hi - self.lo)
        self.comm.synchronize()          # my pushes have landed ...
        dist.barrier()                   # ... and so have everybody else's
        torch.cuda.current_stream(self.device).wait_stream(self.comm)
        return self.full[: self.n_total]

    def close(self):
        for b in self._bases:
            try:
                import ctypes

                self._lib.call("isx_ipc_close", ctypes.c_void_p(b))
            except Exception:
                pass
        self._bases = []


def make_row_exchange(n_total: int, D: int, device, dtype=torch.float32, chunk_rows: int = 128, peer: bool = True):
    """The row exchange of extract_features_sharded: pushed peer rows on a CUDA box (see PeerRows), the chunked NCCL / gloo
    all-gather otherwise (single process, CPU tensors, or a platform without CUDA IPC -- decided collectively, so every rank
    takes the same path)."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    dev = torch.device(device)
    if multi and peer and dev.type == "cuda":
        obj = PeerRows.get(n_total, D, dev, dtype)   # collective-safe: never raises on a subset of the ranks
        flag = torch.tensor([1 if obj.ok else 0], device=dev if dist.get_backend() == "nccl" else "cpu", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # no IPC somewhere (expandable segments, no peer access): fall back together
        if int(flag.item()) == 1:
            obj._begin()
            return obj
        if not obj.ok and dist.get_rank() == 0:
            import warnings

            warnings.warn("iris_b200: peer row exchange unavailable (%r); using the NCCL all-gather" % (getattr(obj, "error", None),))
    return RowGatherer(n_total, D, device, dtype, chunk_rows)


def sharded_map(n_total: int, fn: Callable[[int, int], torch.Tensor]) -> torch.Tensor:
    """Run fn(lo, hi) -> [hi-lo, D] on this rank's shard and all-gather the rows of every rank."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    rows = fn(lo, hi)
    assert rows.shape[0] == hi - lo
    return all_gather_rows(rows, n_total)
