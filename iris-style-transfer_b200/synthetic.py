"""Deterministic synthetic eye frames / iris label maps (numpy PCG64, stable across versions).

Stands in for the licensed OpenEDS2019 (1,640,400) / OpenEDS2020 (1,400,640) frames the
reference loads in data_preprocessing.py (out of scope): a smooth low-frequency
background, a textured annulus "iris" (label 2), a dark pupil (label 3), a bright sclera
(label 1) and a few saturated glint blobs (> 0.8) so that the glint mask of
pipelines.py:143-144 is exercised.  The iris covers ~6-10 % of the frame like the two eye
PNGs shipped with the reference.
"""
from __future__ import annotations

import numpy as np


def _smooth_noise(rng: np.random.Generator, h: int, w: int, cells: int) -> np.ndarray:
    """Bilinear up-sampling of a coarse random grid -> smooth field in [0,1]."""
    gh, gw = max(2, h // cells + 2), max(2, w // cells + 2)
    g = rng.random((gh, gw), dtype=np.float64)
    ys = np.linspace(0, gh - 1.001, h)
    xs = np.linspace(0, gw - 1.001, w)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)


def synthetic_eye(seed: int, h: int = 640, w: int = 400):
    """Return (frame float32 [1,h,w] in [0,1], labels int64 [1,h,w] in {0,1,2,3})."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cy = h * (0.45 + 0.1 * rng.random())
    cx = w * (0.45 + 0.1 * rng.random())
    r_iris = min(h, w) * (0.19 + 0.03 * rng.random())
    r_pupil = r_iris * (0.30 + 0.1 * rng.random())
    ecc = 0.85 + 0.15 * rng.random()
    rr = np.sqrt(((yy - cy) / ecc) ** 2 + (xx - cx) ** 2)
    theta = np.arctan2(yy - cy, xx - cx)

    bg = 0.25 + 0.35 * _smooth_noise(rng, h, w, 64)
    sclera = rr < r_iris * 2.2
    img = bg.copy()
    img[sclera] = 0.55 + 0.2 * _smooth_noise(rng, h, w, 32)[sclera]

    # iris texture: radial streaks + fine noise
    k = int(rng.integers(24, 48))
    streak = 0.5 + 0.5 * np.sin(k * theta + 6.0 * _smooth_noise(rng, h, w, 16))
    fine = rng.random((h, w))
    iris_tex = 0.30 + 0.25 * streak + 0.15 * fine
    iris = (rr < r_iris) & (rr >= r_pupil)
    img[iris] = iris_tex[iris]
    pupil = rr < r_pupil
    img[pupil] = 0.04 + 0.04 * fine[pupil]

    # eyelid: cut the top of the iris like a real eye
    lid = yy < cy - r_iris * (0.55 + 0.25 * rng.random())
    img[lid] = bg[lid]

    labels = np.zeros((h, w), dtype=np.int64)
    labels[sclera & ~lid] = 1
    labels[iris & ~lid] = 2
    labels[pupil & ~lid] = 3

    # glints: saturated blobs, some inside the iris
    for _ in range(int(rng.integers(3, 7))):
        a = rng.random() * 2 * np.pi
        rad = r_iris * rng.random()
        gy, gx = cy + rad * np.sin(a), cx + rad * np.cos(a)
        gr = 2.0 + 0.02 * min(h, w) * rng.random()
        blob = ((yy - gy) ** 2 + (xx - gx) ** 2) < gr * gr
        img[blob] = 0.85 + 0.15 * rng.random()

    img = np.clip(img, 0.0, 1.0).astype(np.float32)
    return img[None], labels[None]


def synthetic_batch(seeds, h: int = 640, w: int = 400):
    """Stack several synthetic eyes: (frames [B,1,h,w] float32, labels [B,1,h,w] int64)."""
    frames, labels = zip(*(synthetic_eye(s, h, w) for s in seeds))
    return np.stack(frames), np.stack(labels)


def synthetic_iris_crops(seeds, size: int = 224):
    """3-channel square iris crops as the drivers feed nst() (…2019.py:66-79): a textured
    disc on a zero background, replicated to 3 channels.  float32 [B,3,size,size]."""
    out = []
    for s in seeds:
        rng = np.random.default_rng(1000003 + s)
        yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
        c = (size - 1) / 2.0
        rr = np.sqrt((yy - c) ** 2 + (xx - c) ** 2)
        theta = np.arctan2(yy - c, xx - c)
        k = int(rng.integers(24, 48))
        tex = 0.30 + 0.25 * (0.5 + 0.5 * np.sin(k * theta + 6.0 * _smooth_noise(rng, size, size, 16)))
        tex = tex + 0.15 * rng.random((size, size))
        m = (rr < c) & (rr > c * 0.33)
        img = np.where(m, tex, 0.0).astype(np.float32)
        out.append(np.repeat(img[None], 3, axis=0))
    return np.stack(out)


def synthetic_label_map(seed: int, h: int = 400, w: int = 640, speck: float = 0.0, drop=()):
    """Label map int64 [h,w] in {0,1,2,3} as a segmenter would emit it for an OpenEDS2020-shaped frame (the input of
    `extract_eye_landmarks`, gaze_estimators.py:108): the synthetic eye's labels, optionally with a fraction `speck` of
    the pixels relabelled at random (stray components, holes, ragged borders) and with the classes in `drop` erased."""
    _, lab = synthetic_eye(seed, h, w)
    lab = lab[0].copy()
    rng = np.random.default_rng(7000003 + seed)
    if speck > 0:
        hit = rng.random((h, w)) < speck
        lab[hit] = rng.integers(0, 4, size=int(hit.sum()))
    for c in drop:
        lab[lab == c] = 0
    return lab


def landmark_cases():
    """(name, label map [400,640]) pairs: clean eyes, speckled eyes, missing classes, an empty map."""
    cases = []
    for seed in (11, 12, 13):
        cases.append(("clean%d" % seed, synthetic_label_map(seed)))
    for seed, speck in ((21, 0.002), (22, 0.02), (23, 0.2)):
        cases.append(("speck%d" % seed, synthetic_label_map(seed, speck=speck)))
    cases.append(("nopupil", synthetic_label_map(31, drop=(3,))))
    cases.append(("nosclera", synthetic_label_map(32, speck=0.001, drop=(1,))))
    cases.append(("irisonly", synthetic_label_map(33, drop=(1, 3))))
    cases.append(("empty", np.zeros((400, 640), dtype=np.int64)))
    return cases


def gaze_head_case(in_dim, hidden=64, out_dim=3, batch=37, seed=0):
    """Seeded weights (torch's (out, in) layout) and inputs of a gaze head."""
    rng = np.random.default_rng(9000 + seed + in_dim)

    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    params = [u((hidden, in_dim), in_dim), u((hidden,), in_dim), u((hidden, hidden), hidden), u((hidden,), hidden),
              u((out_dim, hidden), hidden), u((out_dim,), hidden)]
    x = rng.standard_normal((batch, in_dim)).astype(np.float32)
    if in_dim == 19:   # landmark-sized magnitudes (pixel coordinates, angles)
        x = (np.abs(x) * 100).astype(np.float32)
    return params, x


def iou_case(shape=(640, 400), seeds=(61, 62, 63, 64, 65)):
    """(preds, targets) int64 [b,h,w] for cal_IoUs: targets = synthetic label maps, preds = the same maps with 0 / 0.5 / 5 %
    of the pixels relabelled, one with a class missing, one all background (an empty union: iou = 0 / eps = 0)."""
    h, w = shape
    t = np.stack([synthetic_label_map(s, h, w) for s in seeds])
    p = np.stack([synthetic_label_map(seeds[0], h, w), synthetic_label_map(seeds[1], h, w, speck=0.005),
                  synthetic_label_map(seeds[2], h, w, speck=0.05), synthetic_label_map(seeds[3], h, w, drop=(3,)),
                  np.zeros((h, w), dtype=np.int64)])
    t[4][t[4] == 3] = 0          # class 3 absent from both maps of the last pair
    return p, t


def gaze_vector_case(n=257, seed=0):
    """Two sets of unit vectors (n, 3) for angular_distance, incl. identical and opposite rows (dot = +-1 up to rounding)."""
    rng = np.random.default_rng(31337 + seed)
    a = rng.standard_normal((n, 3)).astype(np.float32)
    b = rng.standard_normal((n, 3)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    b[0] = a[0]
    b[1] = -a[1]
    return a, b
