"""CPU oracle of the mask PRODUCER: the reference's RITnet segmenter
(models/ritnet/ritnet.py:8-223) restated on the CPU.  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as
oracle/nst_oracle.py: only tests/ and __graft_entry__.smoke() import it).

Two parts:
  * RITnet_transform (ritnet.py:64-98): x*255 -> uint8 -> gamma table (cv2.LUT, then np.uint8 truncation) ->
    cv2.createCLAHE(clipLimit=1.5, tileGridSize=(8,8)).apply -> ToImage/ToDtype(scale)/Normalize(0.5,0.5).
    `clahe_numpy` restates OpenCV's algorithm (modules/imgproc/src/clahe.cpp: CLAHE_CalcLut_Body /
    CLAHE_Interpolation_Body; opencv-python 4.13 installed, unpinned by environment.yml) in numpy so that the CUDA kernel
    has a line-by-line model; tests pin it bit-exactly against cv2 itself, which is present wherever the tests run.
  * DenseNet2D (ritnet.py:100-223) in eval mode: functional torch fp32 restatement taking the state dict
    (dropout is the identity in eval mode; BatchNorm uses its running statistics).
Pinned by tests/test_ritnet_oracle.py against label maps the UNMODIFIED reference produced (tests/golden/ritnet.npz,
tests/golden/make_golden_ritnet.py) and, through the existing fixtures, against the two shipped eye PNGs."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# RITnet_transform (ritnet.py:64-98)
# ------------------------------------------------------------------------------------------------
def gamma_table_u8() -> np.ndarray:
    """ritnet.py:72 `self.table = 255.0 * (np.linspace(0, 1, 256)**0.8)` looked up by cv2.LUT (ritnet.py:93) and
    truncated by `np.uint8` (ritnet.py:94)."""
    table = 255.0 * (np.linspace(0, 1, 256) ** 0.8)
    return np.uint8(table)


def normalize_table_f32() -> np.ndarray:
    """ToImage -> ToDtype(float32, scale=True) -> Normalize([0.5],[0.5]) (ritnet.py:73-77) of every uint8 value, computed
    with torchvision itself so that the division / multiplication order is whatever the installed version does."""
    import torchvision.transforms.v2 as transforms

    t = transforms.Compose([transforms.ToImage(), transforms.ToDtype(torch.float32, scale=True),
                            transforms.Normalize([0.5], [0.5])])
    return t(np.arange(256, dtype=np.uint8).reshape(16, 16)).reshape(-1).numpy().copy()


def clahe_numpy(src: np.ndarray, clip_limit: float = 1.5, tiles=(8, 8)) -> np.ndarray:
    """cv2.createCLAHE(clipLimit, tileGridSize=tiles).apply(src) for a uint8 (H,W) image, restated.
    tiles = (tilesX, tilesY) like cv2's tileGridSize."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape
    tiles_x, tiles_y = tiles
    hist_size = 256
    if W % tiles_x == 0 and H % tiles_y == 0:
        ext = src
    else:  # copyMakeBorder(src, 0, tilesY - H % tilesY, 0, tilesX - W % tilesX, BORDER_REFLECT_101)
        pb, pr = tiles_y - (H % tiles_y), tiles_x - (W % tiles_x)
        ext = np.pad(src, ((0, pb), (0, pr)), mode="reflect")
    th, tw = ext.shape[0] // tiles_y, ext.shape[1] // tiles_x
    area = th * tw
    lut_scale = np.float32(hist_size - 1) / np.float32(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / hist_size), 1)
    lut = np.zeros((tiles_y * tiles_x, hist_size), dtype=np.uint8)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            tile = ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            hist = np.bincount(tile.reshape(-1), minlength=hist_size).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // hist_size
                residual = clipped - batch * hist_size
                hist = hist + batch
                if residual != 0:
                    step = max(hist_size // residual, 1)
                    i = 0
                    while i < hist_size and residual > 0:
                        hist[i] += 1
                        i += step
                        residual -= 1
            cum = np.cumsum(hist).astype(np.float32) * lut_scale           # int * float -> float
            lut[ty * tiles_x + tx] = np.clip(np.rint(cum), 0, 255).astype(np.uint8)   # saturate_cast<uchar> = cvRound
    # interpolation (CLAHE_Interpolation_Body), fp32 throughout, no FMA contraction
    inv_tw, inv_th = np.float32(1.0) / np.float32(tw), np.float32(1.0) / np.float32(th)
    xs = np.arange(W, dtype=np.float32)
    txf = xs * inv_tw - np.float32(0.5)
    tx1 = np.floor(txf).astype(np.int64)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = (np.float32(1.0) - xa).astype(np.float32)
    tx2 = np.minimum(tx1 + 1, tiles_x - 1)
    tx1 = np.maximum(tx1, 0)
    ys = np.arange(H, dtype=np.float32)
    tyf = ys * inv_th - np.float32(0.5)
    ty1 = np.floor(tyf).astype(np.int64)
    ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    ya1 = (np.float32(1.0) - ya).astype(np.float32)
    ty2 = np.minimum(ty1 + 1, tiles_y - 1)
    ty1 = np.maximum(ty1, 0)
    v = src.astype(np.int64)
    lutf = lut.astype(np.float32)
    p11 = lutf[(ty1[:, None] * tiles_x + tx1[None, :]), v]
    p12 = lutf[(ty1[:, None] * tiles_x + tx2[None, :]), v]
    p21 = lutf[(ty2[:, None] * tiles_x + tx1[None, :]), v]
    p22 = lutf[(ty2[:, None] * tiles_x + tx2[None, :]), v]
    top = (p11 * xa1[None, :]).astype(np.float32) + (p12 * xa[None, :]).astype(np.float32)
    bot = (p21 * xa1[None, :]).astype(np.float32) + (p22 * xa[None, :]).astype(np.float32)
    res = (top.astype(np.float32) * ya1[:, None]).astype(np.float32) + (bot.astype(np.float32) * ya[:, None]).astype(np.float32)
    return np.clip(np.rint(res.astype(np.float32)), 0, 255).astype(np.uint8)


def ritnet_transform(x: torch.Tensor, use_cv2: bool = True) -> torch.Tensor:
    """RITnet_transform.forward (ritnet.py:79-98) for ONE image (1,h,w) or (h,w) in [0,1] -> (1,1,h,w) fp32."""
    if x.dim() == 3:
        x = x[0]
    u8 = (x.to("cpu") * 255).to(torch.uint8).numpy()
    g = gamma_table_u8()[u8]
    if use_cv2:
        import cv2

        e = cv2.createCLAHE(clipLimit=1.5, tileGridSize=(8, 8)).apply(np.ascontiguousarray(g))
    else:
        e = clahe_numpy(g)
    return torch.from_numpy(normalize_table_f32()[e.astype(np.int64)])[None, None]


# ------------------------------------------------------------------------------------------------
# DenseNet2D (ritnet.py:100-223), eval mode
# ------------------------------------------------------------------------------------------------
def _conv(sd, name, x, pad):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad)


def _down(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, pool: bool) -> torch.Tensor:
    """DenseNet2D_down_block.forward (ritnet.py:118-135); AvgPool2d although the attribute is called max_pool."""
    if pool:
        x = F.avg_pool2d(x, 2)
    x1 = F.leaky_relu(_conv(sd, p + ".conv1", x, 1))
    x21 = torch.cat((x, x1), dim=1)
    x22 = F.leaky_relu(_conv(sd, p + ".conv22", _conv(sd, p + ".conv21", x21, 0), 1))
    x31 = torch.cat((x21, x22), dim=1)
    out = F.leaky_relu(_conv(sd, p + ".conv32", _conv(sd, p + ".conv31", x31, 0), 1))
    return F.batch_norm(out, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"], sd[p + ".bn.bias"],
                        training=False, eps=1e-5)


def _up(sd: Dict[str, torch.Tensor], p: str, skip: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """DenseNet2D_up_block_concat.forward (ritnet.py:151-162)."""
    x = F.interpolate(x, scale_factor=(2, 2), mode="nearest")
    x = torch.cat((x, skip), dim=1)
    x1 = F.leaky_relu(_conv(sd, p + ".conv12", _conv(sd, p + ".conv11", x, 0), 1))
    x21 = torch.cat((x, x1), dim=1)
    return F.leaky_relu(_conv(sd, p + ".conv22", _conv(sd, p + ".conv21", x21, 0), 1))


@torch.no_grad()
def densenet2d_logits(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """DenseNet2D.forward (ritnet.py:209-222): x (B,1,H,W) normalised -> logits (B,4,H,W)."""
    x1 = _down(sd, "down_block1", x, False)
    x2 = _down(sd, "down_block2", x1, True)
    x3 = _down(sd, "down_block3", x2, True)
    x4 = _down(sd, "down_block4", x3, True)
    x5 = _down(sd, "down_block5", x4, True)
    x6 = _up(sd, "up_block1", x4, x5)
    x7 = _up(sd, "up_block2", x3, x6)
    x8 = _up(sd, "up_block3", x2, x7)
    x9 = _up(sd, "up_block4", x1, x8)
    return _conv(sd, "out_conv1", x9, 0)


@torch.no_grad()
def ritnet_labels(sd: Dict[str, torch.Tensor], x: torch.Tensor, use_cv2: bool = True) -> torch.Tensor:
    """RITnet.forward (ritnet.py:40-58) for one image (1,h,w): int64 label map (1,h,w), classes 0..3 (2 = iris)."""
    logits = densenet2d_logits(sd, ritnet_transform(x, use_cv2))
    return logits.max(1)[1]
