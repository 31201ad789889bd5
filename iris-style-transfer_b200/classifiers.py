"""Drop-ins for models/classifiers/classifiers.py (inference): Classifier1 (projection head on VGG's pool5 output) and
Classifier2 (projection head on the style statistics), as the drivers use them in eval mode
(iris_style_transfer_openeds2019.py:82-84,144-146; iris_classification.py:94-98), plus the feature cache the frozen-VGG
training loop lacks (iris_classification.py:66-71 recomputes the VGG forward every epoch although vgg is frozen,
:52-55,133).  The Linear layers run as weight-streaming tcgen05 GEMMs (csrc/heads.cu); training the heads stays with the
caller's optimiser (out of the accelerated path)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib


def _ceil64(n: int) -> int:
    return (n + 63) // 64 * 64


class _Head(torch.nn.Module):
    """Three Linear layers (K -> 4096 -> 4096 -> num_class), ReLU between, packed once per device."""

    def __init__(self, in_features: int, num_class: int, state_dict: Optional[Dict[str, torch.Tensor]], linear_idx: Sequence[int]):
        super().__init__()
        dims = [(in_features, 4096), (4096, 4096), (4096, num_class)]
        self.num_class = num_class
        self.in_features = in_features
        self.host: List[tuple] = []
        for (k, n), li in zip(dims, linear_idx):
            if state_dict is not None:
                w = state_dict["model.%d.weight" % li] if "model.%d.weight" % li in state_dict else state_dict["%d.weight" % li]
                b = state_dict["model.%d.bias" % li] if "model.%d.bias" % li in state_dict else state_dict["%d.bias" % li]
            else:  # torch.nn.Linear's default initialisation
                lin = torch.nn.Linear(k, n)
                w, b = lin.weight, lin.bias
            assert tuple(w.shape) == (n, k)
            self.host.append((w.detach().float().cpu().contiguous(), b.detach().float().cpu().contiguous()))
        self._packed: Dict[str, list] = {}
        self._device = torch.device("cuda:0")

    def to(self, device=None, *args, **kwargs):
        if device is not None:
            self._device = torch.device(device)
        return self

    def eval(self):
        return self

    def _weights(self, dev):
        key = str(dev)
        if key not in self._packed:
            out = []
            with torch.cuda.device(dev):
                for w, b in self.host:
                    n, k = w.shape
                    wp = torch.empty(n, k + 64, device=dev, dtype=torch.bfloat16)
                    _lib.call("isx_linear_pack", w.to(dev), b.to(dev), n, k, n, wp, _lib.stream_ptr())
                    out.append(wp)
                torch.cuda.current_stream().synchronize()
            self._packed[key] = out
        return self._packed[key]

    @torch.no_grad()
    def _mlp(self, xp: torch.Tensor, M: int) -> torch.Tensor:
        """xp: packed input X' bf16 [Mpad, K+64] -> logits fp32 [M, num_class]."""
        dev = xp.device
        Mpad = xp.shape[0]
        ws = self._weights(dev)
        with torch.cuda.device(dev):
            for li, wp in enumerate(ws):
                n, kp = wp.shape
                outT = torch.empty(n, Mpad, device=dev, dtype=torch.bfloat16)
                _lib.call("isx_linear_fwd", xp, wp, outT, Mpad, n, kp, int(li < 2), _lib.stream_ptr())
                if li < 2:
                    xp = torch.empty(Mpad, n + 64, device=dev, dtype=torch.bfloat16)
                    _lib.call("isx_transpose_pack", outT, n, Mpad, M, xp, _lib.stream_ptr())
            logits = torch.empty(M, self.num_class, device=dev, dtype=torch.float32)
            _lib.call("isx_transpose_out", outT, self.num_class, Mpad, M, logits, _lib.stream_ptr())
        return logits

    def _rows(self, rows: torch.Tensor) -> torch.Tensor:
        """fp32 [M, in_features] (CUDA) -> logits, in chunks of <= 256 rows."""
        if not rows.is_cuda:
            rows = rows.to(self._device)
        rows = rows.to(torch.float32)
        outs = []
        for lo in range(0, rows.shape[0], 256):
            r = rows[lo:lo + 256]
            M, K = r.shape
            Mpad = _ceil64(M)
            with torch.cuda.device(r.device):
                xp = torch.empty(Mpad, K + 64, device=r.device, dtype=torch.bfloat16)
                _lib.call("isx_rows_pack", r, _lib.i64(r.stride(0)), M, K, Mpad, xp, _lib.stream_ptr())
            outs.append(self._mlp(xp, M))
        return torch.cat(outs) if len(outs) > 1 else outs[0]


class Classifier1(_Head):
    """classifiers.py:3-36: AdaptiveAvgPool2d((7,7)) -> Flatten -> 25088 -> 4096 -> 4096 -> num_class on the VGG output."""

    def __init__(self, num_class: int = 152, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__(25088, num_class, state_dict, (2, 5, 8))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: VGG19's first return value (pool5) as fp32 NCHW [B,512,h,w] like the reference, or the bf16 NHWC map
        [B,h,w,512] straight from the device buffers (engine.feature_view(1, 4))."""
        if not x.is_cuda:
            x = x.to(self._device)
        if x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[-1] == 512:
            p5 = x.contiguous()
        else:
            p5 = x.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        B, h, w, C = p5.shape
        outs = []
        for lo in range(0, B, 256):
            part = p5[lo:lo + 256]
            M = part.shape[0]
            Mpad = _ceil64(M)
            with torch.cuda.device(p5.device):
                xp = torch.empty(Mpad, C * 49 + 64, device=p5.device, dtype=torch.bfloat16)
                _lib.call("isx_pool7_flatten_pack", part, M, h, w, C, Mpad, xp, _lib.stream_ptr())
            outs.append(self._mlp(xp, M))
        return torch.cat(outs) if len(outs) > 1 else outs[0]


class Classifier2(_Head):
    """classifiers.py:38-72: cat(mean, std) per style tap -> in_features -> 4096 -> 4096 -> num_class."""

    def __init__(self, in_features: int = (64 + 128 + 256 + 512) * 2, num_class: int = 152,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__(in_features, num_class, state_dict, (0, 3, 6))

    @torch.no_grad()
    def forward(self, style_features) -> torch.Tensor:
        """style_features: the list of style feature maps VGG19 returns (classifiers.py:71 reduces them to mean | unbiased
        std per channel), or the already reduced [B, in_features] matrix (features.style_features_batch(..., gram=False))."""
        if isinstance(style_features, torch.Tensor):
            rows = style_features
        else:
            from .utils import style_features as reduce

            rows = reduce(style_features)
        return self._rows(rows)
