"""Drop-in for models/vgg/vgg.py: `VGG19(content_layers, style_layers, bn).forward(x, mask)` ->
(pool5 output, [content features], [style features]) computed by libisx (bf16 tensor-core convs)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .engine import (CONV_OF_FEATURE_INDEX, N_CONVS, POOL_OF_FEATURE_INDEX, TAP_POOL0, VGG19_LAYERS, NstEngine, PackedVGG,
                     tap_conv)

vgg19_layers = dict(VGG19_LAYERS)  # same public name as models/vgg/vgg.py:6

# models/vgg/vgg.py:12-17 (torchvision vgg19_bn.features indices): conv -> bn -> relu triples
vgg19_bn_layers: Dict[str, int] = {}
_BN_CONV_OF_INDEX: Dict[int, int] = {}    # bn / relu index -> conv ordinal (the in-place ReLU overwrites the BN output)
_BN_RAW_CONV_INDEX = set()                # conv indices: their PRE-BatchNorm output is a separate tensor
_i, _blk, _sub, _c = 0, 1, 1, 0
for _v in [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]:
    if _v == "M":
        vgg19_bn_layers["pool%d" % _blk] = _i
        _i, _blk, _sub = _i + 1, _blk + 1, 1
    else:
        vgg19_bn_layers["conv%d_%d" % (_blk, _sub)] = _i
        vgg19_bn_layers["bn%d_%d" % (_blk, _sub)] = _i + 1
        vgg19_bn_layers["relu%d_%d" % (_blk, _sub)] = _i + 2
        _BN_RAW_CONV_INDEX.add(_i)
        _BN_CONV_OF_INDEX[_i + 1] = _c
        _BN_CONV_OF_INDEX[_i + 2] = _c
        _i, _sub, _c = _i + 3, _sub + 1, _c + 1
assert vgg19_bn_layers["relu4_2"] == 32 and vgg19_bn_layers["pool5"] == 52


def fold_batchnorm(w, b, gamma, beta, mean, var, eps: float = 1e-5):
    """Conv2d followed by an eval-mode BatchNorm2d == one Conv2d: w' = w * g/sqrt(var+eps), b' = (b - mean) * g/sqrt(var+eps) + beta."""
    scale = (gamma.double() / torch.sqrt(var.double() + eps))
    return (w.double() * scale[:, None, None, None]).float(), ((b.double() - mean.double()) * scale + beta.double()).float()


def random_vgg19_bn_weights(seed: int = 0):
    """torchvision vgg19_bn(weights=None) under torch.manual_seed(seed), BatchNorm folded into the 16 (weight, bias) pairs."""
    import torchvision.models as tvm

    torch.manual_seed(seed)
    net = tvm.vgg19_bn(weights=None).features.eval()
    return vgg19_bn_pairs(net)


def vgg19_bn_pairs(features: torch.nn.Module):
    """The 16 folded (weight, bias) pairs of a torchvision vgg19_bn `.features` stack in eval mode."""
    mods = list(features)
    out = []
    for i, m in enumerate(mods):
        if isinstance(m, torch.nn.Conv2d):
            bn = mods[i + 1]
            assert isinstance(bn, torch.nn.BatchNorm2d)
            out.append(fold_batchnorm(m.weight.detach(), m.bias.detach(), bn.weight.detach(), bn.bias.detach(),
                                      bn.running_mean.detach(), bn.running_var.detach(), bn.eps))
    assert len(out) == N_CONVS
    return out


def random_vgg19_weights(seed: int = 0) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """BASELINE's "random-init VGG-19": torchvision vgg19(weights=None) under torch.manual_seed(seed)."""
    import torchvision.models as tvm

    torch.manual_seed(seed)
    net = tvm.vgg19(weights=None).features
    return [(m.weight.detach().clone(), m.bias.detach().clone()) for m in net if isinstance(m, torch.nn.Conv2d)]


class _VGGForward(torch.autograd.Function):
    """Differentiable VGG19.forward: forward = isx_nst_forward, backward = isx_nst_backward (tcgen05 dgrad chain)."""

    @staticmethod
    def forward(ctx, x, vgg, mask):
        last, c, s, unbatched = vgg.features_nhwc(x, mask, full=True)
        ctx.vgg, ctx.eng, ctx.version, ctx.unbatched = vgg, vgg._last_engine, vgg._fwd_version, unbatched
        ctx.x_shape = tuple(x.shape)

        def nchw(t):
            t = t.permute(0, 3, 1, 2).float()
            return t[0] if unbatched else t

        return (nchw(last),) + tuple(nchw(t) for t in c) + tuple(nchw(t) for t in s)

    @staticmethod
    def backward(ctx, g_last, *g_feats):
        vgg, eng = ctx.vgg, ctx.eng
        if vgg._fwd_version != ctx.version:
            raise RuntimeError("iris_b200.VGG19: backward() after another forward() of the same module -- the stored "
                               "activations were overwritten (the module is not re-entrant, like the reference's "
                               "FeatureExtractor, models/vgg/vgg.py:107)")

        def nhwc(g):
            if ctx.unbatched:
                g = g[None]
            return g.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()

        grads = {}
        convs = list(vgg.content_convs) + list(vgg.style_convs)
        for conv, g in zip(convs, g_feats):
            if g is None:
                continue
            gn = nhwc(g)
            grads[conv] = grads[conv] + gn if conv in grads else gn
        lp = nhwc(g_last) if g_last is not None and bool((g_last != 0).any()) else None
        if not grads and lp is None:
            return torch.zeros(ctx.x_shape, device=g_last.device), None, None
        B, xc, H, W = eng.cfg.B, eng.cfg.xc, eng.cfg.H, eng.cfg.W
        dx = torch.empty(B, xc, H, W, device=eng.device, dtype=torch.float32)
        with torch.cuda.device(eng.device):
            eng.backward(grads, lp, dx)
        return (dx[0] if ctx.unbatched else dx), None, None


class VGG19(torch.nn.Module):
    """models/vgg/vgg.py:19-92.  `weights`: 'imagenet' (the reference's IMAGENET1K_V1 via torchvision; needs the
    checkpoint to be available), 'random' (torchvision init under `seed`), or the 16 (weight, bias) pairs.
    `bn=True` (vgg19_bn, vgg.py:41-44): the model is frozen in eval mode (vgg.py:49-53), so every BatchNorm2d is an affine
    map of its conv and is folded into the 16 conv kernels once; `bn*` / `relu*` taps are the usual post-ReLU tensors
    (the in-place ReLU overwrites the BatchNorm output), `conv*` taps of the bn model -- the PRE-BatchNorm tensor -- are
    not representable after folding and raise."""

    def __init__(self, content_layers: Sequence[str] = ("relu4_2",),
                 style_layers: Sequence[str] = ("relu1_1", "relu2_1", "relu3_1", "relu4_1"), bn: bool = False,
                 weights="imagenet", seed: int = 0) -> None:
        super().__init__()
        self.bn = bool(bn)
        self.content_layers = list(content_layers)
        self.style_layers = list(style_layers)
        table = vgg19_bn_layers if bn else VGG19_LAYERS
        conv_of = dict(_BN_CONV_OF_INDEX if bn else CONV_OF_FEATURE_INDEX)
        for k in range(5):   # pooling layers are taps too (vgg.py:6-17): tap id 16 + k
            conv_of[table["pool%d" % (k + 1)]] = TAP_POOL0 + k
        self.content_layers_idx = [table[i] for i in self.content_layers]
        self.style_layers_idx = [table[i] for i in self.style_layers]
        for idx in self.content_layers_idx + self.style_layers_idx:
            if bn and idx in _BN_RAW_CONV_INDEX:
                raise NotImplementedError("conv* taps of vgg19_bn are the pre-BatchNorm tensors; BatchNorm is folded into the "
                                          "convolutions here -- tap bn* / relu* instead")
            if idx not in conv_of:
                raise ValueError("unknown feature layer index %d" % idx)
        self.content_convs = [conv_of[i] for i in self.content_layers_idx]
        self.style_convs = [conv_of[i] for i in self.style_layers_idx]
        if isinstance(weights, str):
            if weights == "imagenet":
                import torchvision.models as tvm

                if bn:
                    weights = vgg19_bn_pairs(tvm.vgg19_bn(weights=tvm.VGG19_BN_Weights.IMAGENET1K_V1).features.eval())
                else:
                    net = tvm.vgg19(weights=tvm.VGG19_Weights.IMAGENET1K_V1).features
                    weights = [(m.weight.detach(), m.bias.detach()) for m in net if isinstance(m, torch.nn.Conv2d)]
            elif weights == "random":
                weights = random_vgg19_bn_weights(seed) if bn else random_vgg19_weights(seed)
            else:
                raise ValueError("weights must be 'imagenet', 'random' or 16 (weight, bias) pairs")
        assert len(weights) == N_CONVS
        self.host_weights = [(w.detach().float().cpu(), b.detach().float().cpu()) for w, b in weights]
        self._packed: Dict[str, PackedVGG] = {}
        self._engines: Dict[tuple, NstEngine] = {}
        self._device = torch.device("cuda:0")
        self._last_engine: Optional[NstEngine] = None
        self._fwd_version = 0

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
    def run_forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, full: bool = False,
                    lean: bool = False) -> NstEngine:
        """Forward pass only; the activations stay in the returned engine's workspace (engine.feature_view).  lean: pre-pool
        activations that are not taps are not materialised (feature extraction: nobody reads them)."""
        if x.dim() == 3:
            x = x[None]
        dev = x.device if x.is_cuda else self._device
        x = x.detach().to(dev, torch.float32).contiguous()
        B, xc, H, W = x.shape
        taps = self.content_convs + self.style_convs
        n_conv = N_CONVS if full else max(tap_conv(t) for t in taps) + 1
        eng = self._engine(B, H, W, xc, n_conv, dev)
        eng.set_input_mask(mask)
        self._last_engine = eng
        self._fwd_version += 1
        with torch.cuda.device(dev):
            eng.forward(x, with_last_pool=full, lean=lean)
        return eng

    @torch.no_grad()
    def features_nhwc(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, full: bool = False):
        """bf16 NHWC features (copies of the device buffers): (pool5 or None, content list, style list)."""
        unbatched = x.dim() == 3
        eng = self.run_forward(x, mask, full)
        with torch.cuda.device(eng.device):
            last = eng.feature(1, 4) if full else None
            c = [eng.tap(i) for i in self.content_convs]
            s = [eng.tap(i) for i in self.style_convs]
        return last, c, s, unbatched

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """models/vgg/vgg.py:69-92.  Returns fp32 NCHW tensors like the reference.  Differentiable w.r.t. x (the weights are
        frozen, vgg.py:52-53): autograd runs the tcgen05 dgrad chain (isx_nst_backward).  nst() does not go through
        here -- its closure is the fused isx_nst_eval."""
        if torch.is_grad_enabled() and x.requires_grad:
            outs = _VGGForward.apply(x, self, mask)
            nc = len(self.content_convs)
            return outs[0], list(outs[1:1 + nc]), list(outs[1 + nc:])
        with torch.no_grad():
            return self._forward_nograd(x, mask)

    def _forward_nograd(self, x, mask):
        last, c, s, unbatched = self.features_nhwc(x, mask, full=True)

        def nchw(t):
            t = t.permute(0, 3, 1, 2).float()
            return t[0] if unbatched else t

        return nchw(last), [nchw(t) for t in c], [nchw(t) for t in s]
