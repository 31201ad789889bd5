"""Drop-in for models/vgg/vgg.py: `VGG19(content_layers, style_layers, bn).forward(x, mask)` ->
(pool5 output, [content features], [style features]) computed by libisx (bf16 tensor-core convs)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .engine import (CONV_OF_FEATURE_INDEX, N_CONVS, POOL_OF_FEATURE_INDEX, VGG19_LAYERS, NstEngine, PackedVGG)

vgg19_layers = dict(VGG19_LAYERS)  # same public name as models/vgg/vgg.py:6


def random_vgg19_weights(seed: int = 0) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """BASELINE's "random-init VGG-19": torchvision vgg19(weights=None) under torch.manual_seed(seed)."""
    import torchvision.models as tvm

    torch.manual_seed(seed)
    net = tvm.vgg19(weights=None).features
    return [(m.weight.detach().clone(), m.bias.detach().clone()) for m in net if isinstance(m, torch.nn.Conv2d)]


class VGG19(torch.nn.Module):
    """models/vgg/vgg.py:19-92.  `weights`: 'imagenet' (the reference's IMAGENET1K_V1 via torchvision; needs the
    checkpoint to be available), 'random' (torchvision init under `seed`), or the 16 (weight, bias) pairs."""

    def __init__(self, content_layers: Sequence[str] = ("relu4_2",),
                 style_layers: Sequence[str] = ("relu1_1", "relu2_1", "relu3_1", "relu4_1"), bn: bool = False,
                 weights="imagenet", seed: int = 0) -> None:
        super().__init__()
        if bn:
            raise NotImplementedError("vgg19_bn (models/vgg/vgg.py:41-42) is outside the accelerated path")
        self.content_layers = list(content_layers)
        self.style_layers = list(style_layers)
        self.content_layers_idx = [VGG19_LAYERS[i] for i in self.content_layers]
        self.style_layers_idx = [VGG19_LAYERS[i] for i in self.style_layers]
        for idx in self.content_layers_idx + self.style_layers_idx:
            if idx not in CONV_OF_FEATURE_INDEX:
                raise NotImplementedError("taps on pooling layers are not supported by the accelerated path")
        self.content_convs = [CONV_OF_FEATURE_INDEX[i] for i in self.content_layers_idx]
        self.style_convs = [CONV_OF_FEATURE_INDEX[i] for i in self.style_layers_idx]
        if isinstance(weights, str):
            if weights == "imagenet":
                import torchvision.models as tvm

                net = tvm.vgg19(weights=tvm.VGG19_Weights.IMAGENET1K_V1).features
                weights = [(m.weight.detach(), m.bias.detach()) for m in net if isinstance(m, torch.nn.Conv2d)]
            elif weights == "random":
                weights = random_vgg19_weights(seed)
            else:
                raise ValueError("weights must be 'imagenet', 'random' or 16 (weight, bias) pairs")
        assert len(weights) == N_CONVS
        self.host_weights = [(w.detach().float().cpu(), b.detach().float().cpu()) for w, b in weights]
        self._packed: Dict[str, PackedVGG] = {}
        self._engines: Dict[tuple, NstEngine] = {}
        self._device = torch.device("cuda:0")

    # the reference calls vgg.to(device) (pipelines.py:46)
    def to(self, device=None, *args, **kwargs):  # noqa: D401
        if device is not None:
            self._device = torch.device(device)
        return self

    def packed(self, device=None) -> PackedVGG:
        dev = torch.device(device if device is not None else self._device)
        if dev.type != "cuda":
            raise _lib.IsxError("iris_b200.VGG19 runs on CUDA (B200) only; got device %s" % dev)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        key = str(dev)
        if key not in self._packed:
            self._packed[key] = PackedVGG(self.host_weights, dev)
        return self._packed[key]

    def _engine(self, B, H, W, xc, n_conv, device) -> NstEngine:
        key = (B, H, W, xc, n_conv, str(device))
        if key not in self._engines:
            self._engines.clear()  # one cached workspace at a time
            self._engines[key] = NstEngine(self.packed(device), B, H, W, xc, self.content_convs, self.style_convs,
                                           n_conv=n_conv)
        return self._engines[key]

    @torch.no_grad()
    def features_nhwc(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, full: bool = False):
        """bf16 NHWC features straight from the device buffers: (pool5 or None, content list, style list)."""
        unbatched = x.dim() == 3
        if unbatched:
            x = x[None]
        dev = x.device if x.is_cuda else self._device
        x = x.detach().to(dev, torch.float32).contiguous()
        B, xc, H, W = x.shape
        taps = self.content_convs + self.style_convs
        n_conv = N_CONVS if full else max(taps) + 1
        eng = self._engine(B, H, W, xc, n_conv, dev)
        eng.set_input_mask(mask)
        with torch.cuda.device(dev):
            eng.forward(x, with_last_pool=full)
            last = eng.feature(1, 4) if full else None
            c = [eng.feature(0, i) for i in self.content_convs]
            s = [eng.feature(0, i) for i in self.style_convs]
        return last, c, s, unbatched

    @torch.no_grad()
    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """models/vgg/vgg.py:69-92.  Returns fp32 NCHW tensors like the reference (not differentiable: the
        backward of this stack lives in the fused nst() driver)."""
        last, c, s, unbatched = self.features_nhwc(x, mask, full=True)

        def nchw(t):
            t = t.permute(0, 3, 1, 2).float()
            return t[0] if unbatched else t

        return nchw(last), [nchw(t) for t in c], [nchw(t) for t in s]
