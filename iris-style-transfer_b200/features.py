"""Style-feature extraction for iris classification (iris_classification.py:66-71,94-98;
iris_style_transfer_openeds2019.py:82-84): VGG-19 forward to the deepest style tap, then per layer the
mean/std statistics Classifier2 consumes (models/classifiers/classifiers.py:71 -> 1920 floats per eye) and,
optionally, the Gram matrices (utils.py:242-257) as upper triangles (BASELINE config 3)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from .engine import tap_channels
from .vgg import VGG19


_COPY_STREAMS = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    """One side stream per device for the life of the process (a new stream per call would get its own allocator pool)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(dev)
    return _COPY_STREAMS[key]


@torch.no_grad()
def style_features_batch(vgg: VGG19, x: torch.Tensor, gram: bool = True, stats: bool = True,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [B,1|3,H,W] fp32 on a CUDA device -> [B, D]; D = 2*sum(C_l) (stats) + sum(C_l(C_l+1)/2) (gram).
    One forward + ONE libisx call (isx_nst_style_features): the statistics and the Gram upper triangles are computed
    on the activations where they lie in the workspace and written straight into the rows of `out` (given or new)."""
    if not (gram or stats):
        raise ValueError("nothing to extract")
    eng = vgg.run_forward(x, full=False, lean=True)
    chans = [tap_channels(c) for c in vgg.style_convs]
    D = feature_dim(chans, gram, stats)
    if out is None:
        out = torch.empty(eng.cfg.B, D, device=eng.device, dtype=torch.float32)
    assert out.shape[0] == eng.cfg.B and out.shape[1] >= D
    with torch.cuda.device(eng.device):
        eng.style_features(out, stats=stats, gram=gram)
    return out


def feature_dim(channels: Sequence[int], gram: bool = True, stats: bool = True) -> int:
    d = 0
    if stats:
        d += 2 * sum(channels)
    if gram:
        d += sum(c * (c + 1) // 2 for c in channels)
    return d


@torch.no_grad()
def extract_features_sharded(vgg: VGG19, images, batch: int = 64, gram: bool = True, stats: bool = True,
                             device=None, gather_chunk: int = 128, peer: bool = True) -> torch.Tensor:
    """`images`: indexable [n,1|3,H,W] (host or device).  Each rank extracts its contiguous shard in batches of `batch`,
    every batch writing its rows in place into the final matrix; complete chunks of `gather_chunk` rows are all-gathered
    on a side stream while later batches compute (sharding.RowGatherer) -- or, on a CUDA box (peer=True), every batch of rows
    is pushed straight into the other ranks' matrices over NVLink by the copy engines (sharding.PeerRows; the returned matrix
    is then a cached buffer, valid until the next extraction of the same shape); every rank returns the full [n, D] matrix."""
    from .sharding import make_row_exchange

    n = len(images)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    chans = [tap_channels(c) for c in vgg.style_convs]
    rows = make_row_exchange(n, feature_dim(chans, gram, stats), dev, chunk_rows=gather_chunk, peer=peer)
    lo, hi = rows.lo, rows.hi
    # Host -> device copies run on a side stream, one batch ahead of the kernels, into two device buffers that are
    # allocated once (a fresh allocation per batch on a second stream makes the caching allocator fall back to
    # cudaMalloc, which serialises the device: measured 83 ms vs 140-220 ms per 512 eyes).  Pageable sources go through
    # two pinned staging buffers.
    starts = list(range(lo, hi, batch))
    main = torch.cuda.current_stream(dev)
    copy = _copy_stream(dev)
    copy.wait_stream(main)
    staging = [None, None]
    dbuf = [None, None]
    consumed = [None, None]  # event on `main`: the kernels reading dbuf[j] have been enqueued and finished

    def fetch(k: int):
        i = starts[k]
        xb = torch.as_tensor(images[i:min(hi, i + batch)])
        if xb.device.type == "cuda":
            return xb.to(dev, torch.float32), None
        xb = xb.to(torch.float32)
        j = k & 1
        if not xb.is_pinned():
            buf = staging[j]
            if buf is None or buf.shape != xb.shape:
                buf = staging[j] = torch.empty(xb.shape, dtype=torch.float32).pin_memory()
            buf.copy_(xb)
            xb = buf
        if dbuf[j] is None or dbuf[j].shape[1:] != xb.shape[1:] or dbuf[j].shape[0] < xb.shape[0]:
            dbuf[j] = torch.empty(xb.shape, dtype=torch.float32, device=dev)
            dbuf[j].record_stream(copy)
        xd = dbuf[j][:xb.shape[0]]
        with torch.cuda.stream(copy):
            if consumed[j] is not None:
                copy.wait_event(consumed[j])
            xd.copy_(xb, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return xd, ev

    row = 0
    nxt = fetch(0) if starts else None
    for k in range(len(starts)):
        xd, ev = nxt
        if ev is not None:
            main.wait_event(ev)
        if k + 1 < len(starts):
            if ev is not None and staging[(k + 1) & 1] is not None:
                ev.synchronize()  # pageable source: staging buffer (k+1) & 1 was last read by the copy of batch k-1 (<= ev)
            nxt = fetch(k + 1)
        style_features_batch(vgg, xd, gram=gram, stats=stats, out=rows.local[row:row + xd.shape[0]])
        if ev is not None:
            consumed[k & 1] = torch.cuda.Event()
            consumed[k & 1].record(main)
        row += xd.shape[0]
        rows.flush(row)
    return rows.finish()


@torch.no_grad()
def cache_classifier_inputs(vgg: VGG19, images, batch: int = 32, device=None):
    """What the frozen-VGG classifier training (iris_classification.py:59-108) needs from the VGG, computed ONCE instead of
    once per epoch: for every image the Classifier2 input (mean | unbiased std of the style taps, fp32 [n, 2*sum C_l]) and the
    Classifier1 input (AdaptiveAvgPool2d(7,7) + Flatten of pool5, bf16 [n, 25088]).  `images`: indexable [n,1|3,H,W]."""
    n = len(images)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    chans = [tap_channels(c) for c in vgg.style_convs]
    D = feature_dim(chans, gram=False, stats=True)
    stats = torch.empty(n, D, device=dev, dtype=torch.float32)
    pool = torch.empty(n, 25088, device=dev, dtype=torch.bfloat16)
    for lo in range(0, n, batch):
        xb = torch.as_tensor(images[lo:lo + batch]).to(dev, torch.float32)
        eng = vgg.run_forward(xb, full=True, lean=True)
        b = xb.shape[0]
        with torch.cuda.device(dev):
            eng.style_features(stats[lo:lo + b], stats=True, gram=False)
            p5 = eng.feature_view(1, 4)                       # bf16 NHWC [b,h,w,512]
            mpad = (b + 63) // 64 * 64
            xp = torch.empty(mpad, 25088 + 64, device=dev, dtype=torch.bfloat16)
            _lib.call("isx_pool7_flatten_pack", p5, b, p5.shape[1], p5.shape[2], 512, mpad, xp, _lib.stream_ptr())
            pool[lo:lo + b] = xp[:b, :25088]
    return stats, pool
