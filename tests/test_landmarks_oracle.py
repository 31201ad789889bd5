"""Pin oracle/landmarks_oracle.py (CPU restatement of the reference's `extract_eye_landmarks` + gaze heads,
models/gaze_estimators/gaze_estimators.py:8-223) against OpenCV itself -- the third-party library the reference calls
(cv2.findContours / contourArea / fitEllipse, gaze_estimators.py:70-81) -- and against landmarks the UNMODIFIED reference
produced (tests/golden/landmarks.npz, make_golden_landmarks.py)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import landmarks_oracle as L

cv2 = pytest.importorskip("cv2")


def _synthetic():
    spec = importlib.util.spec_from_file_location("isx_synthetic", os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iris-style-transfer_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    return syn


def random_mask(rng, kind, H, W):
    """Speckle, smooth blobs, rings with something inside (a component in another's hole), clipped ellipses."""
    if kind == 0:
        return (rng.random((H, W)) < rng.choice([0.05, 0.3, 0.5, 0.7])).astype(np.uint8)
    if kind == 1:
        g = cv2.resize(rng.random((H // 4 + 2, W // 4 + 2)), (W, H), interpolation=cv2.INTER_CUBIC)
        return (g > rng.choice([0.4, 0.5, 0.6])).astype(np.uint8)
    m = np.zeros((H, W), np.uint8)
    if kind == 2:
        for _ in range(int(rng.integers(1, 4))):
            c = (int(rng.integers(0, W)), int(rng.integers(0, H)))
            r = int(rng.integers(2, min(H, W) // 2 + 1))
            cv2.circle(m, c, r, 1, int(rng.integers(1, 4)))
            if rng.random() < 0.7:
                cv2.circle(m, c, max(1, r // 3), 1, -1)
        return m
    cv2.ellipse(m, (W // 2, H // 2), (int(rng.integers(2, W // 2 + 1)), int(rng.integers(2, H // 2 + 1))),
                float(rng.uniform(0, 180)), 0, 360, 1, -1)
    m[: int(rng.integers(0, H // 2))] = 0
    return m


def test_external_contours_equal_cv2():
    rng = np.random.default_rng(0)
    for it in range(300):
        H, W = int(rng.integers(8, 60)), int(rng.integers(8, 80))
        m = random_mask(rng, it % 4, H, W)
        ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        mine = L.external_contours(m)
        assert len(ref) == len(mine), (it, H, W)
        for a, b in zip(ref, mine):                     # same contours, same order, same points
            assert np.array_equal(a.reshape(-1, 2), b), (it, H, W)
            assert L.contour_area(b) == cv2.contourArea(a)


def _as_vec(e):
    return np.array([e[0][0], e[0][1], e[1][0], e[1][1], e[2]], dtype=np.float64)


def test_fit_ellipse_matches_cv2_on_eye_sized_contours():
    rng = np.random.default_rng(1)
    n = 0
    for it in range(150):
        m = np.zeros((400, 640), np.uint8)
        cv2.ellipse(m, (int(rng.integers(100, 540)), int(rng.integers(100, 300))),
                    (int(rng.integers(4, 150)), int(rng.integers(4, 150))), float(rng.uniform(0, 180)), 0, 360, 1, -1)
        if it % 3 == 1:
            m[: int(rng.integers(60, 250))] = 0         # eyelid
        if it % 3 == 2:
            m[rng.random(m.shape) < 0.2] = 0            # ragged border, holes
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if not cs:
            continue
        c = max(cs, key=cv2.contourArea)
        if len(c) < 5 or L.is_degenerate(c):
            continue
        np.testing.assert_allclose(_as_vec(L.fit_ellipse(c)), _as_vec(cv2.fitEllipse(c)), rtol=2e-6, atol=2e-5)
        n += 1
    assert n > 120


def test_fit_ellipse_small_blobs_and_the_documented_exception():
    """Tiny blobs: equal to cv2 (angles compared modulo 180: for axis-aligned fits the sign of a rounding-level xy term picks
    0 or 180) unless the system is rank deficient, where OpenCV perturbs the points pseudo-randomly (`is_degenerate`)."""
    rng = np.random.default_rng(2)
    n = deg = 0
    for it in range(1200):
        H, W = int(rng.integers(6, 24)), int(rng.integers(6, 24))
        g = cv2.resize(rng.random((H // 3 + 2, W // 3 + 2)), (W, H), interpolation=cv2.INTER_LINEAR)
        cs, _ = cv2.findContours((g > 0.5).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if not cs:
            continue
        c = max(cs, key=cv2.contourArea)
        if len(c) < 5:
            continue
        if L.is_degenerate(c):
            deg += 1
            continue
        r, q = _as_vec(cv2.fitEllipse(c)), _as_vec(L.fit_ellipse(c))
        if not (np.all(np.isfinite(r)) and np.all(np.isfinite(q))):
            continue
        np.testing.assert_allclose(q[:4], r[:4], rtol=1e-4, atol=1e-3)
        da = abs(q[4] - r[4]) % 180.0
        assert min(da, 180.0 - da) < 1e-2, (c.reshape(-1, 2).tolist(), r, q)
        n += 1
    assert n > 1000 and deg < 20


def test_landmarks_equal_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "landmarks.npz"))
    for name, lab in _synthetic().landmark_cases():
        got = L.extract_eye_landmarks(lab)
        ref = gold["lm_" + name]
        assert got.dtype == np.float32 and got.shape == (19,)
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-5, err_msg=name)
        # the integer landmarks (eye corners, width, height) and the absent-class zeros are exact
        assert np.array_equal(got[10:16], ref[10:16]), name
        assert np.array_equal(got == 0, ref == 0), name


@pytest.mark.parametrize("in_dim", [19, 2048])
def test_gaze_head_equals_reference_golden(golden_dir, in_dim):
    gold = np.load(os.path.join(golden_dir, "landmarks.npz"))
    params, x = _synthetic().gaze_head_case(in_dim)
    out = L.gaze_head(x, params)
    np.testing.assert_allclose(out, gold["head%d_out" % in_dim], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, rtol=1e-5)
