"""CPU oracle of the downstream evaluator's feature path: `extract_eye_landmarks` and the
gaze-estimator heads of models/gaze_estimators/gaze_estimators.py restated on the CPU.  THIS IS TEST INFRASTRUCTURE, NOT
PRODUCT CODE (same rules as oracle/nst_oracle.py: only tests/ and __graft_entry__.smoke() import it).

The arithmetic of gaze_estimators.py:55-106 lives in a third-party dependency, OpenCV (`opencv-python`, unpinned by
environment.yml; 4.13 installed here and on the GPU box): `cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)`,
`cv2.contourArea`, `cv2.fitEllipse`.  Its published algorithms are restated here in plain Python / numpy so that the
CUDA kernel has a line-by-line model:
  * external_contours: Suzuki-Abe border following as OpenCV implements it (modules/imgproc/src/contours.cpp,
    cvFindNextContour / icvFetchContour): raster scan for 0 -> 1 transitions on unmarked pixels, a start is rejected in
    RETR_EXTERNAL mode when the last marked pixel met on the row carries a positive mark; the trace marks every border
    pixel (negative when its right-hand neighbour was examined as background) and emits a point wherever the chain
    direction changes (CHAIN_APPROX_SIMPLE).  cv2 returns the contours in REVERSE order of discovery.
  * contour_area: Green's formula over the emitted points (shapedescr.cpp contourArea, oriented=false).
  * fit_ellipse: the least-squares conic fit of shapedescr.cpp fitEllipseNoDirect (mean-centred, scaled points; five-
    parameter fit with b = 10000, centre from the gradient equations, three-parameter refit with b = 1, angle and axes).
Pinned by tests/test_landmarks_oracle.py against cv2 itself (random blobs, rings, specks, thin walls: contour lists equal
point for point, ellipses within float32 rounding) and against landmarks the UNMODIFIED reference function produced
(tests/golden/landmarks.npz, tests/golden/make_golden_landmarks.py)."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

# chain code s -> (dx, dy): 0 = east, then counter-clockwise on the screen (y grows downwards): NE, N, NW, W, SW, S, SE
CODE_DELTAS = ((1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1))
NBD = 2          # mark of a visited border pixel (external mode never needs distinct border numbers)
NBD_NEG = -126   # (schar)(2 | -128): border pixel whose right-hand neighbour was examined as background


def _trace(img: np.ndarray, x0: int, y0: int) -> List[Tuple[int, int]]:
    """icvFetchContour for an OUTER border starting at (x0, y0) of the zero-framed int8 image `img` (modified in place),
    CHAIN_APPROX_SIMPLE: returns the emitted points (frame coordinates)."""
    pts: List[Tuple[int, int]] = []
    s_end = s = 4
    while True:
        s = (s - 1) & 7
        dx, dy = CODE_DELTAS[s]
        x1, y1 = x0 + dx, y0 + dy
        if img[y1, x1] != 0 or s == s_end:
            break
    if s == s_end:                      # single-pixel domain
        img[y0, x0] = NBD_NEG
        return [(x0, y0)]
    x3, y3 = x0, y0
    px, py = x0, y0
    prev_s = s ^ 4
    while True:
        s_end = s
        x4 = y4 = 0
        while s < 15:
            s += 1
            dx, dy = CODE_DELTAS[s & 7]
            x4, y4 = x3 + dx, y3 + dy
            if img[y4, x4] != 0:
                break
        s &= 7
        if ((s - 1) & 0xFFFFFFFF) < s_end:      # the east neighbour was among the examined background pixels
            img[y3, x3] = NBD_NEG
        elif img[y3, x3] == 1:
            img[y3, x3] = NBD
        if s != prev_s:
            pts.append((px, py))
            prev_s = s
        px += CODE_DELTAS[s][0]
        py += CODE_DELTAS[s][1]
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return pts


def external_contours(mask: np.ndarray) -> List[np.ndarray]:
    """cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)[0] restated: list of (n,2) int32 arrays (x, y) in cv2's
    order (last discovered first)."""
    m = np.asarray(mask)
    H, W = m.shape
    img = np.zeros((H + 2, W + 2), dtype=np.int8)      # one-pixel zero frame like OpenCV's copyMakeBorder
    img[1:-1, 1:-1] = (m != 0)
    found: List[np.ndarray] = []
    for y in range(1, H + 1):
        row = img[y]
        # candidate columns of this row: any pixel that is nonzero (marks change while we trace, so re-read per pixel)
        xs = np.flatnonzero(row)
        lnbd_x = 0
        k = 0
        prev = 0
        x = 0
        # walk only over the nonzero pixels; a zero pixel between two of them resets `prev` to 0
        while k < len(xs):
            x = int(xs[k])
            prev = int(row[x - 1])
            p = int(row[x])
            if prev == 0 and p == 1:                    # outer border start?
                if not (img[y, lnbd_x] > 0):            # RETR_EXTERNAL: rejected inside a positively marked border
                    pts = _trace(img, x, y)
                    found.append(np.asarray(pts, dtype=np.int32) - 1)
                    p = int(row[x])
            if p != 0 and p != 1:                       # `if (prev & -2) lnbd.x = x` on the next iteration
                lnbd_x = x
            k += 1
    return found[::-1]


def contour_area(pts: np.ndarray) -> float:
    """cv2.contourArea(pts) (oriented = false)."""
    p = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    if len(p) == 0:
        return 0.0
    q = np.roll(p, 1, axis=0)
    return abs(0.5 * float(np.sum(q[:, 0] * p[:, 1] - q[:, 1] * p[:, 0])))


def is_degenerate(pts: np.ndarray) -> bool:
    """True where the CUDA kernel does not claim cv2.fitEllipse's result (info flag bit 0): (i) for EXACTLY five points cv2
    switches to fitEllipseDirect (the ellipse-constrained eigenvector fit); (ii) when the five-parameter system is rank
    deficient (sigma_max * FLT_EPSILON > sigma_min: collinear specks) fitEllipseNoDirect perturbs the points with an
    mt19937 stream before it solves; (iii) see below.  All three only happen when the LARGEST component of a class is a
    speck of a handful of pixels -- never for a pupil or an iris.  Tests skip such point sets."""
    p = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    if len(p) == 5:
        return True
    c = p.sum(axis=0, dtype=np.float32) / np.float32(len(p))
    d = (p - c).astype(np.float64)
    s = float(np.sum(np.abs(d)))
    d *= 100.0 / max(s, float(np.finfo(np.float32).eps))
    A = np.stack([-d[:, 0] ** 2, -d[:, 1] ** 2, -d[:, 0] * d[:, 1], d[:, 0], d[:, 1]], axis=1)
    w = np.linalg.svd(A, compute_uv=False)
    if w[0] * float(np.finfo(np.float32).eps) > w[-1]:
        return True
    # (iii) the refit with the centre fixed is EXACTLY rank deficient (a speck symmetric about its fitted centre: two columns
    # coincide).  cv2's answer is then the minimum-norm solution its SVD back-substitution picks -- `fit_ellipse` below
    # reproduces that through lstsq, the CUDA kernel (normal equations) does not and reports the contour instead.
    gfp = np.linalg.lstsq(A, np.full(len(p), 10000.0), rcond=None)[0]
    A2 = np.array([[2 * gfp[0], gfp[2]], [gfp[2], 2 * gfp[1]]])
    r = np.linalg.lstsq(A2, np.array([gfp[3], gfp[4]]), rcond=None)[0]
    A3 = np.stack([(d[:, 0] - r[0]) ** 2, (d[:, 1] - r[1]) ** 2, (d[:, 0] - r[0]) * (d[:, 1] - r[1])], axis=1)
    w3 = np.linalg.svd(A3, compute_uv=False)
    return bool(w3[-1] < 1e-10 * w3[0])


def fit_ellipse(pts: np.ndarray):
    """cv2.fitEllipse(pts) for n >= 5 integer points: ((cx, cy), (width, height), angle) as float32 values."""
    p = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    n = len(p)
    assert n >= 5
    c = p.sum(axis=0, dtype=np.float32) / np.float32(n)            # Point2f accumulation
    # OpenCV accumulates c += p in float; for integer coordinates below 2^24/n the sum is exact either way
    d = p - c                                                      # float32
    s = float(np.sum(np.abs(d[:, 0].astype(np.float64)) + np.abs(d[:, 1].astype(np.float64))))
    scale = 100.0 / max(s, float(np.finfo(np.float32).eps))
    px = d[:, 0].astype(np.float64) * scale
    py = d[:, 1].astype(np.float64) * scale
    A = np.stack([-px * px, -py * py, -px * py, px, py], axis=1)
    b = np.full(n, 10000.0)
    gfp = np.linalg.lstsq(A, b, rcond=None)[0]
    A2 = np.array([[2 * gfp[0], gfp[2]], [gfp[2], 2 * gfp[1]]])
    rp = np.zeros(5)
    rp[:2] = np.linalg.lstsq(A2, np.array([gfp[3], gfp[4]]), rcond=None)[0]
    A3 = np.stack([(px - rp[0]) ** 2, (py - rp[1]) ** 2, (px - rp[0]) * (py - rp[1])], axis=1)
    g = np.linalg.lstsq(A3, np.ones(n), rcond=None)[0]
    rp[4] = -0.5 * math.atan2(g[2], g[1] - g[0])
    if abs(g[2]) > 1e-8:
        t = g[2] / math.sin(-2.0 * rp[4])
    else:
        t = g[1] - g[0]
    rp[2] = abs(g[0] + g[1] - t)
    if rp[2] > 1e-8:
        rp[2] = math.sqrt(2.0 / rp[2])
    rp[3] = abs(g[0] + g[1] + t)
    if rp[3] > 1e-8:
        rp[3] = math.sqrt(2.0 / rp[3])
    cx = np.float32(rp[0] / scale) + c[0]
    cy = np.float32(rp[1] / scale) + c[1]
    w = np.float32(rp[2] * 2 / scale)
    h = np.float32(rp[3] * 2 / scale)
    ang = np.float32(0)                  # RotatedRect's default: OpenCV assigns the angle only in the swapped branch
    if w > h:                            # (always taken for a real ellipse: t >= 0 makes rp[2] the longer semi-axis)
        w, h = h, w
        ang = np.float32(90 + rp[4] * 180 / math.pi)
    if ang < -180:
        ang += np.float32(360)
    if ang > 360:
        ang -= np.float32(360)
    return (float(cx), float(cy)), (float(w), float(h)), float(ang)


def find_ellipse_features(mask: np.ndarray):
    """gaze_estimators.py:55-83."""
    cs = external_contours(mask)
    if len(cs) == 0:
        return None, None, None, None, None
    areas = [contour_area(c) for c in cs]
    best = cs[int(np.argmax(areas))]          # max(contours, key=cv2.contourArea): first maximum in cv2's order
    if len(best) < 5:
        return None, None, None, None, None
    (cx, cy), (w, h), ang = fit_ellipse(best)
    return cx, cy, w, h, ang


def find_eye_corners(mask: np.ndarray):
    """gaze_estimators.py:85-106 (left / right = min / max column, 'bottom' / 'top' = min / max row)."""
    ys, xs = np.where(mask > 0)
    if len(xs) == 0:
        return None, None, None, None
    return int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())


def extract_eye_landmarks(segmentation: np.ndarray, epsilon: float = 1e-6) -> np.ndarray:
    """gaze_estimators.py:108-178 for one (H, W) label map -> float32 [19]."""
    seg = np.asarray(segmentation).astype(np.uint8)
    pcx, pcy, pmaj, pmin, pang = find_ellipse_features((seg == 3).astype(np.uint8))
    icx, icy, imaj, imin, iang = find_ellipse_features((seg == 2).astype(np.uint8))
    left, right, bottom, top = find_eye_corners((seg == 1).astype(np.uint8))
    if left is not None:
        ew = right - left
        eh = top - bottom
        ear = eh / (ew + epsilon)
    else:
        ew = eh = ear = None
    if pcx is not None and left is not None:
        npx = (pcx - (left + right) / 2) / (ew + epsilon)
        npy = (pcy - (bottom + top) / 2) / (eh + epsilon)
    else:
        npx = npy = None
    lm = [pcx, pcy, pmaj, pmin, pang, icx, icy, imaj, imin, iang, left, right, bottom, top, ew, eh, ear, npx, npy]
    return np.asarray([0 if v is None else v for v in lm], dtype=np.float32)


def gaze_head(x: np.ndarray, params) -> np.ndarray:
    """GazeEstimator1.model / GazeEstimator2.model in eval mode (gaze_estimators.py:24-32,51-53 / :196-204,221-223):
    Linear -> ReLU -> [Dropout = identity] -> Linear -> ReLU -> Linear, then x / ||x||_2 per row.  params = (W1, b1, W2,
    b2, W3, b3) with torch's (out, in) weight layout; fp32 like the reference."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(a, dtype=np.float32) for a in params]
    h = np.maximum(np.asarray(x, dtype=np.float32) @ W1.T + b1, 0)
    h = np.maximum(h @ W2.T + b2, 0)
    o = h @ W3.T + b3
    return o / np.sqrt(np.sum(o * o, axis=1, keepdims=True, dtype=np.float32))
