"""CPU oracle: an fp32 restatement of the reference hot path (iris-masked Gatys NST + style
feature extraction).  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module, and only as the checker or as the timed CPU baseline.  The product package
(iris-style-transfer_b200/) never imports it and has no CPU fallback.

Parity pin: tests/golden/*.npz were produced by the UNMODIFIED reference, imported live from
/root/reference by tests/golden/make_golden.py (oracle/ref_loader.py), and
tests/test_oracle_golden.py checks every function below against them.  The reference itself
ships no tests or golden vectors for this path.

All citations are relative to /root/reference/.  Arithmetic that lives in third-party code
(torch 2.6.0+cu126 / torchvision 0.21.0 pinned by environment.yml:97-99; 2.11.0 / 0.26.0
installed) is restated from the installed sources: torch/optim/lbfgs.py:333-537 (L-BFGS, no
line search), torchvision/models/vgg.py:73-94 (cfg "E"), ATen UpSampleKernel.cpp
(_upsample_bilinear2d_aa), torchvision rgb_to_grayscale.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# VGG-19 feature stack (models/vgg/vgg.py:6-10, 19-92; torchvision/models/vgg.py cfg "E")
# --------------------------------------------------------------------------------------
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]

# models/vgg/vgg.py:6-10 -- name -> index into torchvision vgg19().features
VGG19_LAYERS: Dict[str, int] = {}
_i = 0
_blk, _sub = 1, 1
for _v in VGG19_CFG:
    if _v == "M":
        VGG19_LAYERS["pool%d" % _blk] = _i
        _i += 1
        _blk += 1
        _sub = 1
    else:
        VGG19_LAYERS["conv%d_%d" % (_blk, _sub)] = _i
        VGG19_LAYERS["relu%d_%d" % (_blk, _sub)] = _i + 1
        _i += 2
        _sub += 1
assert VGG19_LAYERS["relu4_2"] == 22 and VGG19_LAYERS["pool5"] == 36 and VGG19_LAYERS["conv5_1"] == 28

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # models/vgg/vgg.py:66
IMAGENET_STD = (0.229, 0.224, 0.225)

DEFAULT_CONTENT = ("relu4_2",)  # models/vgg/vgg.py:25
DEFAULT_STYLE = ("relu1_1", "relu2_1", "relu3_1", "relu4_1")  # models/vgg/vgg.py:26


def random_vgg19_weights(seed: int = 0) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """BASELINE.json's "random-init VGG-19": torchvision `vgg19(weights=None)` built under
    `torch.manual_seed(seed)` (kaiming_normal_(fan_out, relu), bias 0 -- torchvision
    models/vgg.py:55-57; SURVEY.md note N4).  Returns the 16 (weight OIHW, bias) pairs."""
    import torchvision.models as tvm

    torch.manual_seed(seed)
    fn = getattr(tvm.vgg19, "_isx_orig", tvm.vgg19)  # un-patched even if oracle/ref_loader.py is active
    net = fn(weights=None).features
    out = []
    for m in net:
        if isinstance(m, torch.nn.Conv2d):
            out.append((m.weight.detach().clone(), m.bias.detach().clone()))
    assert len(out) == 16
    return out


class _RoundSTE(torch.autograd.Function):
    """Round to a narrower storage type in the forward pass, identity in the backward pass."""

    @staticmethod
    def forward(ctx, x, dtype):
        return x.to(dtype).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g, None


def vgg19_forward(
    x: torch.Tensor,
    weights: Sequence[Tuple[torch.Tensor, torch.Tensor]],
    content_layers: Sequence[str] = DEFAULT_CONTENT,
    style_layers: Sequence[str] = DEFAULT_STYLE,
    mask: Optional[torch.Tensor] = None,
    full: bool = True,
    operand_dtype: Optional[torch.dtype] = None,
):
    """VGG19.forward (models/vgg/vgg.py:69-92): Normalize(mean,std) -> optional `* mask` ->
    vgg19.features; returns (last activation, [content feats], [style feats]).

    torchvision's ReLUs are inplace, so a `conv*` tap aliases the following ReLU's output
    (SURVEY.md note N2): every tap is the post-ReLU tensor.  `full=False` stops after the
    deepest tap (what the product computes inside nst(); the reference always runs to pool5,
    vgg.py:87, but nothing on the NST path reads that output).

    `operand_dtype` (None = the reference's fp32): EMULATION of the precision BASELINE.json's north star prescribes
    for the accelerated path -- conv operands stored in that type (torch.bfloat16), fp32 accumulation: the weights
    of conv1_2.. and every post-ReLU activation are rounded to it (straight-through in the backward pass).  Used by
    the tests to measure how far the REFERENCE ALGORITHM itself moves under that rounding (the calibration of the
    trajectory bounds); never by the product."""
    unbatched = x.dim() == 3
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype).view(-1, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype).view(-1, 1, 1)
    h = (x - mean) / std  # broadcasts 1->3 channels for (1,H,W) input (SURVEY.md note N3)
    if mask is not None:
        h = h * mask  # models/vgg/vgg.py:84-85
    want = {VGG19_LAYERS[n] for n in list(content_layers) + list(style_layers)}
    # conv tap -> the relu right after it (aliasing)
    want_eff = {(i + 1) if _is_conv_index(i) else i for i in want}
    deepest = max(want_eff) if want_eff else 36
    feats: Dict[int, torch.Tensor] = {}
    idx = 0
    ci = 0
    for v in VGG19_CFG:
        if v == "M":
            h = F.max_pool2d(h, 2, 2)
            feats[idx] = h
            idx += 1
        else:
            w, b = weights[ci]
            ci += 1
            if operand_dtype is not None:
                w = w.to(operand_dtype).to(torch.float32)
            h = F.relu(F.conv2d(h, w, b, padding=1))
            if operand_dtype is not None:
                h = _RoundSTE.apply(h, operand_dtype)
            feats[idx] = h  # conv index aliases post-ReLU
            feats[idx + 1] = h
            idx += 2
        if not full and idx > deepest:
            break
    c = [feats[VGG19_LAYERS[n]] for n in content_layers]
    s = [feats[VGG19_LAYERS[n]] for n in style_layers]
    return h, c, s


def _is_conv_index(i: int) -> bool:
    return any(k.startswith("conv") and v == i for k, v in VGG19_LAYERS.items())


# --------------------------------------------------------------------------------------
# Losses (utils.py:242-355) and style features (models/classifiers/classifiers.py:71)
# --------------------------------------------------------------------------------------
def gram_matrix(x: torch.Tensor) -> torch.Tensor:
    """utils.py:242-257.  n = x[0].numel() AFTER flatten: C*H*W for (B,C,H,W), H*W for (C,H,W)."""
    x = x.flatten(start_dim=-2)
    n = x[0].numel()
    return (x @ x.transpose(-2, -1)) / n


def masked_gram_matrix(x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """Row G' of SURVEY.md §8a (extension): GramMatrix(F * m_l); equals gram_matrix for m == 1."""
    return gram_matrix(x * m)


def layer_masks(mask: torch.Tensor, shapes: Sequence[Tuple[int, int]]) -> List[torch.Tensor]:
    """SURVEY.md note N5: m_1 = bool iris mask at frame resolution, m_{l+1} = 2x2 stride-2
    average pool of m_l (values k/4^l, exact in fp32/bf16)."""
    m = mask.to(torch.float32)
    out = []
    for hw in shapes:
        while tuple(m.shape[-2:]) != tuple(hw):
            m = F.avg_pool2d(m, 2, 2)
        out.append(m)
    return out


def content_loss_l2(preds, targets, weights=None) -> torch.Tensor:
    """ContentLoss_L2.forward (utils.py:274-290): 0.5 * sum_l w_l * mse(p, t) (mean over ALL elements)."""
    weights = [1.0] * len(targets) if weights is None else weights
    loss = 0
    for p, t, w in zip(preds, targets, weights):
        loss = loss + F.mse_loss(p, t) * w
    return loss * 0.5


def style_loss_gram(preds, target_grams, weights=None) -> torch.Tensor:
    """StyleLoss_Gram.forward (utils.py:308-322): 0.25 * sum_l w_l * sum((G(p)-T)^2)."""
    weights = [1.0] * len(target_grams) if weights is None else weights
    loss = 0
    for p, t, w in zip(preds, target_grams, weights):
        loss = loss + ((gram_matrix(p) - t) ** 2).sum() * w
    return loss * 0.25


def bn_stats(x: torch.Tensor):
    """mean and UNBIASED std over (H,W) (utils.py:337-338, classifiers.py:71)."""
    return x.mean(dim=(-2, -1)), x.std(dim=(-2, -1))


def style_loss_bn(preds, target_means, target_stds, weights=None) -> torch.Tensor:
    """StyleLoss_BN.forward (utils.py:341-355)."""
    weights = [1.0] * len(target_means) if weights is None else weights
    loss = 0
    for p, tm, ts, w in zip(preds, target_means, target_stds, weights):
        pm, ps = bn_stats(p)
        loss = loss + ((pm - tm) ** 2 + (ps - ts) ** 2).sum() * w / pm.shape[-1]
    return loss


def style_features(style_feats: Sequence[torch.Tensor]) -> torch.Tensor:
    """Classifier2's feature reduction (models/classifiers/classifiers.py:71): per layer
    cat(mean, std) over (H,W), concatenated over layers -> (B, 2*sum(C_l))."""
    return torch.cat([torch.cat([x.mean(dim=(-2, -1)), x.std(dim=(-2, -1))], dim=1) for x in style_feats], dim=1)


# --------------------------------------------------------------------------------------
# L-BFGS (torch/optim/lbfgs.py:333-537, line_search_fn=None) -- restated, not imported
# --------------------------------------------------------------------------------------
class LBFGS:
    """Restatement of torch.optim.LBFGS([x], lr) with every other argument default
    (pipelines.py:59): max_iter=20, max_eval=25, tolerance_grad=1e-7, tolerance_change=1e-9,
    history_size=100, no line search.  `closure()` must return (loss: float, flat_grad)."""

    def __init__(self, x: torch.Tensor, lr: float = 1.0, max_iter: int = 20, history_size: int = 100,
                 tolerance_grad: float = 1e-7, tolerance_change: float = 1e-9):
        self.x = x
        self.lr = lr
        self.max_iter = max_iter
        self.max_eval = max_iter * 5 // 4  # lbfgs.py:261-262
        self.history_size = history_size
        self.tolerance_grad = tolerance_grad
        self.tolerance_change = tolerance_change
        self.n_iter = 0
        self.func_evals = 0
        self.d = None
        self.t = None
        self.old_dirs: List[torch.Tensor] = []
        self.old_stps: List[torch.Tensor] = []
        self.ro: List[torch.Tensor] = []
        self.H_diag = 1
        self.prev_flat_grad = None
        self.prev_loss = None
        self.al = [None] * history_size

    def direction(self, flat_grad: torch.Tensor) -> torch.Tensor:
        """lbfgs.py:396-442 (memory update + two-loop recursion)."""
        if self.n_iter == 1:
            self.d = flat_grad.neg()
            self.old_dirs, self.old_stps, self.ro = [], [], []
            self.H_diag = 1
            return self.d
        y = flat_grad.sub(self.prev_flat_grad)
        s = self.d.mul(self.t)
        ys = y.dot(s)
        if ys > 1e-10:
            if len(self.old_dirs) == self.history_size:
                self.old_dirs.pop(0)
                self.old_stps.pop(0)
                self.ro.pop(0)
            self.old_dirs.append(y)
            self.old_stps.append(s)
            self.ro.append(1.0 / ys)
            self.H_diag = ys / y.dot(y)
        num_old = len(self.old_dirs)
        al = self.al
        q = flat_grad.neg()
        for i in range(num_old - 1, -1, -1):
            al[i] = self.old_stps[i].dot(q) * self.ro[i]
            q.add_(self.old_dirs[i], alpha=-al[i])
        self.d = r = torch.mul(q, self.H_diag)
        for i in range(num_old):
            be_i = self.old_dirs[i].dot(r) * self.ro[i]
            r.add_(self.old_stps[i], alpha=al[i] - be_i)
        return self.d

    @torch.no_grad()
    def step(self, closure: Callable[[], Tuple[float, torch.Tensor]]) -> float:
        """lbfgs.py:333-537."""
        loss, flat_grad = closure()
        orig_loss = loss
        current_evals = 1
        self.func_evals += 1
        if flat_grad.abs().max() <= self.tolerance_grad:  # lbfgs.py:370-374
            return orig_loss
        n_iter = 0
        while n_iter < self.max_iter:
            n_iter += 1
            self.n_iter += 1
            d = self.direction(flat_grad)
            if self.prev_flat_grad is None:
                self.prev_flat_grad = flat_grad.clone()
            else:
                self.prev_flat_grad.copy_(flat_grad)
            self.prev_loss = loss
            if self.n_iter == 1:  # lbfgs.py:454-457
                self.t = min(1.0, 1.0 / float(flat_grad.abs().sum())) * self.lr
            else:
                self.t = self.lr
            gtd = flat_grad.dot(d)
            if gtd > -self.tolerance_change:  # lbfgs.py:463
                break
            self.x.view(-1).add_(d, alpha=self.t)  # lbfgs.py:492 / 313
            ls_func_evals = 0
            if n_iter != self.max_iter:  # lbfgs.py:493-502
                loss, flat_grad = closure()
                opt_cond = flat_grad.abs().max() <= self.tolerance_grad
                ls_func_evals = 1
            current_evals += ls_func_evals
            self.func_evals += ls_func_evals
            if n_iter == self.max_iter:
                break
            if current_evals >= self.max_eval:
                break
            if opt_cond:
                break
            if d.mul(self.t).abs().max() <= self.tolerance_change:  # lbfgs.py:522
                break
            if abs(loss - self.prev_loss) < self.tolerance_change:  # lbfgs.py:525
                break
        return orig_loss


# --------------------------------------------------------------------------------------
# nst() (pipelines.py:8-110)
# --------------------------------------------------------------------------------------
def nst(
    c_img: torch.Tensor,
    s_img: torch.Tensor,
    weights,
    clone_content: bool = True,
    BN_loss: bool = True,
    c_loss_weight: float = 1,
    s_loss_weight: float = 1,
    lr: float = 1,
    epochs: int = 200,
    content_layers: Sequence[str] = DEFAULT_CONTENT,
    style_layers: Sequence[str] = DEFAULT_STYLE,
    x0: Optional[torch.Tensor] = None,
    keep_hist: bool = True,
    full_forward: bool = False,
    operand_dtype: Optional[torch.dtype] = None,
    grad_noise: float = 0.0,
    noise_seed: int = 0,
):
    """pipelines.py:8-110 restated on CPU fp32.  A batch is ONE L-BFGS problem exactly as in the
    reference; call with B=1 for the per-image semantics the product shards on.
    `x0` replaces torch.rand (pipelines.py:54) when clone_content is False so the test controls it.
    `operand_dtype` (see vgg19_forward) and `grad_noise` (multiply every gradient element by 1 + grad_noise * N(0,1))
    are SENSITIVITY PROBES of the reference algorithm for the tests' calibrated trajectory bounds; both default off."""
    c_img = c_img.to(torch.float32)
    s_img = s_img.to(torch.float32)
    if clone_content:
        x = c_img.clone()
    else:
        x = x0.clone() if x0 is not None else torch.rand(c_img.shape)
    x = x.contiguous()
    with torch.no_grad():
        _, c_feats, _ = vgg19_forward(c_img, weights, content_layers, style_layers, full=full_forward,
                                      operand_dtype=operand_dtype)
        _, _, s_feats = vgg19_forward(s_img, weights, content_layers, style_layers, full=full_forward,
                                      operand_dtype=operand_dtype)
        if BN_loss:
            t_mean = [t.mean(dim=(-2, -1)) for t in s_feats]
            t_std = [t.std(dim=(-2, -1)) for t in s_feats]
        else:
            t_gram = [gram_matrix(t) for t in s_feats]
    opt = LBFGS(x, lr=lr)
    x_hist: List[torch.Tensor] = []
    c_hist: List[float] = []
    s_hist: List[float] = []
    n_evals = [0]
    gen = torch.Generator().manual_seed(noise_seed)

    def closure():
        with torch.no_grad():
            x.clamp_(0, 1)  # pipelines.py:81-82
        xv = x.detach().requires_grad_(True)
        with torch.enable_grad():
            _, x_c, x_s = vgg19_forward(xv, weights, content_layers, style_layers, full=full_forward,
                                        operand_dtype=operand_dtype)
            c_loss = content_loss_l2(x_c, c_feats)
            s_loss = style_loss_bn(x_s, t_mean, t_std) if BN_loss else style_loss_gram(x_s, t_gram)
            loss = c_loss * c_loss_weight + s_loss * s_loss_weight
            (g,) = torch.autograd.grad(loss, xv)
        if grad_noise > 0.0:
            g = g * (1.0 + grad_noise * torch.randn(g.shape, generator=gen))
        if keep_hist:
            x_hist.append(x.detach().clone())
        c_hist.append(float(c_loss))
        s_hist.append(float(s_loss))
        n_evals[0] += 1
        return float(loss), g.reshape(-1)

    while n_evals[0] < epochs:  # pipelines.py:79
        opt.step(closure)
    x = x.detach()
    x.clamp_(0, 1)
    return x, x_hist, c_hist, s_hist


def nst_eval(x, c_feats, targets, weights, BN_loss, c_loss_weight, s_loss_weight,
             content_layers=DEFAULT_CONTENT, style_layers=DEFAULT_STYLE, layer_mask=None):
    """One closure evaluation (pipelines.py:80-91) at a given x: returns (c_loss, s_loss, grad).
    `layer_mask`: optional frame-resolution mask for the G' extension (masked Gram)."""
    xv = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        _, x_c, x_s = vgg19_forward(xv, weights, content_layers, style_layers, full=False)
        c_loss = content_loss_l2(x_c, c_feats)
        if layer_mask is not None:
            ms = layer_masks(layer_mask, [tuple(f.shape[-2:]) for f in x_s])
            x_s = [f * m for f, m in zip(x_s, ms)]
        if BN_loss:
            s_loss = style_loss_bn(x_s, targets[0], targets[1])
        else:
            s_loss = style_loss_gram(x_s, targets)
        loss = c_loss * c_loss_weight + s_loss * s_loss_weight
        (g,) = torch.autograd.grad(loss, xv)
    return float(torch.as_tensor(c_loss).detach()), float(torch.as_tensor(s_loss).detach()), g


# --------------------------------------------------------------------------------------
# Mask / bbox / crop (pipelines.py:112-166, utils.py:44-72) -- integer work, bit-exact
# --------------------------------------------------------------------------------------
def crop_bbox(image: np.ndarray) -> Tuple[int, int, int, int]:
    """utils.py:57-64: bbox of NONZERO PIXELS of `image` ((h,w) or (1,h,w)):
    (x_min, y_min, x_max, y_max) = (row_min, col_min, row_max, col_max), inclusive."""
    a = np.asarray(image)
    if a.ndim == 3 and a.shape[0] == 1:
        a = a[0]
    elif a.ndim != 2:
        raise Exception("image shape wrong:", a.shape)  # utils.py:66
    rows, cols = np.nonzero(a)
    if rows.size == 0:
        raise RuntimeError("crop_bbox: image has no nonzero pixel")  # torch: min() of empty tensor
    return int(rows.min()), int(cols.min()), int(rows.max()), int(cols.max())


def mask_and_crop(x: np.ndarray, seg: np.ndarray, glint_threshold: float = 0.8):
    """pipelines.py:139-165 with the RITnet label map `seg` given (the mask PRODUCER is out of
    scope): m = (seg == 2) * (x <= thr); x*m; bbox of nonzero PIXELS; slice; repeat(3,1,1)."""
    x = np.asarray(x, dtype=np.float32)
    m = (np.asarray(seg) == 2) & (x <= np.float32(glint_threshold))
    xm = x * m
    x_min, y_min, x_max, y_max = crop_bbox(xm)
    xc = xm[:, x_min:x_max + 1, y_min:y_max + 1]
    mc = m[:, x_min:x_max + 1, y_min:y_max + 1]
    return np.repeat(xc, 3, axis=0), mc, x_min, y_min, x_max, y_max


# --------------------------------------------------------------------------------------
# Composite (iris_style_transfer_openeds2019.py:111-130; …2020.py:121-139)
# --------------------------------------------------------------------------------------
def rgb_to_grayscale(x: np.ndarray) -> np.ndarray:
    """torchvision.transforms.v2.functional.rgb_to_grayscale: 0.2989 R + 0.587 G + 0.114 B."""
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(0.2989) * x[..., 0:1, :, :] + np.float32(0.587) * x[..., 1:2, :, :]
            + np.float32(0.114) * x[..., 2:3, :, :]).astype(np.float32)


def aa_weights(in_size: int, out_size: int):
    """ATen UpSampleKernel.cpp HelperInterpLinear / _compute_indices_min_size_weights_aa
    (bilinear, antialias=True, align_corners=False): per output index (xmin, xsize, weights)."""
    scale = in_size / out_size
    support = scale if scale >= 1.0 else 1.0  # interp_size/2 * scale, interp_size = 2
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    out = []
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(xsize, dtype=np.float64)
        for j in range(xsize):
            a = abs((j + xmin - center + 0.5) * invscale)
            w[j] = 1.0 - a if a < 1.0 else 0.0
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        out.append((xmin, xsize, w.astype(np.float32)))
    return out


def resize_bilinear_aa(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """transforms.v2.Resize((out_h,out_w)) on a float tensor == F.interpolate(bilinear,
    antialias=True): separable, W pass then H pass, fp32."""
    x = np.asarray(x, dtype=np.float32)
    in_h, in_w = x.shape[-2:]
    ww = aa_weights(in_w, out_w)
    tmp = np.zeros(x.shape[:-1] + (out_w,), dtype=np.float32)
    for i, (x0, n, w) in enumerate(ww):
        tmp[..., i] = (x[..., x0:x0 + n] * w).sum(axis=-1, dtype=np.float32)
    wh = aa_weights(in_h, out_h)
    out = np.zeros(x.shape[:-2] + (out_h, out_w), dtype=np.float32)
    for i, (y0, n, w) in enumerate(wh):
        out[..., i, :] = (tmp[..., y0:y0 + n, :] * w[:, None]).sum(axis=-2, dtype=np.float32)
    return out


def composite(frame: np.ndarray, new_iris_rgb: np.ndarray, mask: np.ndarray, bbox) -> np.ndarray:
    """…2019.py:111-130 for one image: gray -> Resize(bbox shape, bilinear AA) -> * mask[bbox]
    -> frame[bbox] = frame[bbox] * ~mask + new.  frame (1,H,W) fp32, new_iris_rgb (3,h',w'),
    mask (1,H,W) bool, bbox = (x_min,y_min,x_max,y_max) rows/cols inclusive."""
    x_min, y_min, x_max, y_max = bbox
    out = np.array(frame, dtype=np.float32, copy=True)
    g = rgb_to_grayscale(new_iris_rgb)
    g = resize_bilinear_aa(g, x_max - x_min + 1, y_max - y_min + 1)
    m = np.asarray(mask)[:, x_min:x_max + 1, y_min:y_max + 1]
    g = g * m
    out[:, x_min:x_max + 1, y_min:y_max + 1] *= ~m
    out[:, x_min:x_max + 1, y_min:y_max + 1] += g
    return out
