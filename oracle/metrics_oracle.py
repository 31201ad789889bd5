"""CPU oracle of the drivers' evaluation metrics around the NST path: utils.cal_IoUs (utils.py:163-194) and
utils.angular_distance (utils.py:216-240) restated in numpy.  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as
oracle/nst_oracle.py).  Pinned by tests/test_oracle_golden.py-style checks in tests/test_metrics_oracle.py against outputs
of the UNMODIFIED reference functions (tests/golden/metrics.npz, tests/golden/make_golden_metrics.py)."""
from __future__ import annotations

import numpy as np


def cal_ious(preds: np.ndarray, targets: np.ndarray, num_class: int = 4, eps: float = 1e-6):
    """utils.py:163-194 for label maps (b, h, w): the reference sums float32 0/1 maps -- exact integers below 2^24 -- and
    divides in float32: iou[b, c] = float32(inter) / (float32(union) + float32(eps)); miou = mean over the classes."""
    p = np.asarray(preds)
    t = np.asarray(targets)
    b = p.shape[0]
    iou = np.zeros((b, num_class), dtype=np.float32)
    for c in range(num_class):
        pc, tc = (p == c), (t == c)
        inter = (pc & tc).reshape(b, -1).sum(axis=1).astype(np.float32)
        union = (pc | tc).reshape(b, -1).sum(axis=1).astype(np.float32)
        iou[:, c] = inter / (union + np.float32(eps))
    acc = np.zeros(b, dtype=np.float32)
    for c in range(num_class):
        acc = acc + iou[:, c]
    return iou, acc / np.float32(num_class)


def angular_distance(v1: np.ndarray, v2: np.ndarray):
    """utils.py:216-240: acos(clamp(sum(v1 * v2, dim=1), -1, 1)) in float32, and degrees."""
    a = np.asarray(v1, dtype=np.float32)
    b = np.asarray(v2, dtype=np.float32)
    dot = np.zeros(a.shape[0], dtype=np.float32)
    for k in range(a.shape[1]):
        dot = dot + a[:, k] * b[:, k]
    rad = np.arccos(np.clip(dot, np.float32(-1), np.float32(1))).astype(np.float32)
    return rad, (rad * np.float32(180.0 / np.pi)).astype(np.float32)
