"""The host+device core of the landmark kernels (iris-style-transfer_b200/csrc/landmarks_core.cuh: bit-plane raster scan,
border trace, normal-equation ellipse fit, landmark assembly) compiled with g++ and pinned against cv2 and the oracle on
the CPU.  The CUDA kernels of landmarks.cu call these same functions and add only block-level glue; the `-m gpu` tests
check the kernels themselves."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import landmarks_oracle as L
from test_landmarks_oracle import _synthetic, random_mask

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lmhost") / "liblmhost.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(ROOT, "iris-style-transfer_b200", "csrc"),
                    os.path.join(ROOT, "tests", "landmarks_core_host.cpp"), "-o", out], check=True)
    return ctypes.CDLL(out)


def features(host, mask, cap=16384):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.zeros(5, np.float32)
    info = np.zeros(4, np.int32)
    pts = np.zeros(cap, np.uint32)
    rc = host.lm_host_ellipse_features(m.ctypes.data_as(ctypes.c_void_p), m.shape[0], m.shape[1], cap, out.ctypes.data_as(ctypes.c_void_p),
                                       info.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, "the determinant bound contradicted the eigenvalue test"
    n = int(info[0])
    p = np.stack([pts[:min(n, cap)] & 0xFFFF, pts[:min(n, cap)] >> 16], axis=1).astype(np.int32)
    return out, info, p


def test_chosen_contour_equals_cv2(host):
    """Scan + trace + area selection: the contour the core picks == max(cv2 contours, key=contourArea), point for point."""
    rng = np.random.default_rng(5)
    for it in range(400):
        H, W = int(rng.integers(8, 70)), int(rng.integers(8, 100))
        m = random_mask(rng, it % 4, H, W)
        ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        out, info, p = features(host, m)
        assert info[1] == len(ref), (it, H, W)
        if not ref:
            assert info[0] == 0 and info[3] == 0
            continue
        best = max(ref, key=cv2.contourArea).reshape(-1, 2)
        assert np.array_equal(p, best), (it, H, W)


def test_word_boundaries_and_wide_frames(host):
    """Frames wider than one 32-bit word with components straddling word boundaries, inside holes, and on the frame edge."""
    rng = np.random.default_rng(6)
    for it in range(60):
        H, W = int(rng.integers(30, 120)), int(rng.choice([31, 32, 33, 62, 63, 64, 65, 95, 96, 97, 160, 640]))
        m = random_mask(rng, it % 4, H, W)
        m[0, :] |= rng.random(W) < 0.3
        m[:, 0] |= rng.random(H) < 0.3
        m[:, -1] |= rng.random(H) < 0.3
        m[-1, :] |= rng.random(W) < 0.3
        ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        out, info, p = features(host, m)
        assert info[1] == len(ref), (it, H, W)
        assert np.array_equal(p, max(ref, key=cv2.contourArea).reshape(-1, 2)), (it, H, W)


def test_ellipse_equals_cv2(host):
    rng = np.random.default_rng(7)
    n = exact = quick = 0
    for it in range(200):
        m = np.zeros((400, 640), np.uint8)
        cv2.ellipse(m, (int(rng.integers(100, 540)), int(rng.integers(100, 300))),
                    (int(rng.integers(4, 150)), int(rng.integers(4, 150))), float(rng.uniform(0, 180)), 0, 360, 1, -1)
        if it % 3 == 1:
            m[: int(rng.integers(60, 250))] = 0
        if it % 3 == 2:
            m[rng.random(m.shape) < 0.2] = 0
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if not cs:
            continue
        c = max(cs, key=cv2.contourArea)
        out, info, p = features(host, m)
        if len(c) < 5:
            assert info[3] == 0
            continue
        assert bool(info[2] & 1) == L.is_degenerate(c)
        if info[2] & 1:
            continue
        e = cv2.fitEllipse(c)
        ref = np.array([e[0][0], e[0][1], e[1][0], e[1][1], e[2]], np.float32)
        np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-5)
        n += 1
        exact += int(np.array_equal(out, ref))
        quick += int(not (info[2] & 16))
    assert n > 150 and exact > 0.9 * n      # float32-identical almost everywhere
    assert quick == n                       # eye-sized contours: the determinant bound proves full rank, no Jacobi sweeps


def test_point_cap_is_reported(host):
    m = np.zeros((64, 64), np.uint8)
    m[::2, ::2] = 1
    m[10:50, 10:50] = np.indices((40, 40)).sum(0) % 2      # one diagonal-connected component with hundreds of corners
    out, info, p = features(host, m, cap=16)
    assert info[2] & 2 and info[3] == 0 and info[0] > 16


def test_assemble_equals_reference_golden(host, golden_dir):
    """Per-class core results + lm_assemble == the 19 landmarks of the unmodified reference."""
    gold = np.load(os.path.join(golden_dir, "landmarks.npz"))
    for name, lab in _synthetic().landmark_cases():
        lab8 = lab.astype(np.uint8)
        pup, pi, _ = features(host, lab8 == 3)
        iri, ii, _ = features(host, lab8 == 2)
        ys, xs = np.nonzero(lab8 == 1)
        has_s = int(len(xs) > 0)
        bbox = np.array([xs.min(), xs.max(), ys.min(), ys.max()] if has_s else [0, 0, 0, 0], np.int32)
        out = np.zeros(19, np.float32)
        host.lm_host_assemble(pup.ctypes.data_as(ctypes.c_void_p), int(pi[3]), iri.ctypes.data_as(ctypes.c_void_p), int(ii[3]),
                              bbox.ctypes.data_as(ctypes.c_void_p), has_s, ctypes.c_double(1e-6), out.ctypes.data_as(ctypes.c_void_p))
        np.testing.assert_allclose(out, gold["lm_" + name], rtol=2e-6, atol=2e-5, err_msg=name)
        assert np.array_equal(out[10:16], gold["lm_" + name][10:16])


@pytest.mark.parametrize("shape", [(24, 31), (57, 97), (400, 640)])
def test_specks_and_ill_conditioned_contours_equal_cv2(host, shape):
    """Speckled masks: the largest component is often a handful of pixels and the conic system ill conditioned (six points
    on almost a line pair: cond 3e6).  The double-double normal equations must still give cv2's SVD solution wherever
    OpenCV itself keeps the general algorithm (`is_degenerate` false)."""
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    n = 0
    for b in range(24 if H * W < 100000 else 8):
        m3 = random_mask(rng, b % 4, H, W)
        m2 = random_mask(rng, (b + 1) % 4, H, W) & (1 - m3)
        for m in (m3, m2):
            cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            out, info, p = features(host, m)
            assert info[1] == len(cs)
            if not cs:
                continue
            c = max(cs, key=cv2.contourArea)
            assert np.array_equal(p, c.reshape(-1, 2))
            if len(c) < 5:
                continue
            if len(c) == 5:
                assert info[2] & 1
            if (info[2] & 1) or L.is_degenerate(c):      # specks where cv2.fitEllipse leaves the general algorithm
                continue
            e = cv2.fitEllipse(c)
            ref = np.array([e[0][0], e[0][1], e[1][0], e[1][1], e[2]], np.float32)
            if not np.all(np.isfinite(ref)):
                continue
            np.testing.assert_allclose(out[:4], ref[:4], rtol=1e-4, atol=1e-3)
            da = abs(float(out[4]) - float(ref[4])) % 180.0
            assert min(da, 180.0 - da) < 2e-2
            n += 1
    assert n > 4


def test_core_under_asan(tmp_path):
    """The plane / neighbour / trace index arithmetic under AddressSanitizer + UBSan (compute-sanitizer's stand-in: the GPU
    pool does not offer it): 600 random masks with widths around the word boundaries and foreground on the frame edge."""
    exe = str(tmp_path / "lm_asan")
    csrc = os.path.join(ROOT, "iris-style-transfer_b200", "csrc")
    r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-x", "c++", "-I", csrc,
                        os.path.join(ROOT, "tests", "landmarks_core_host.cpp"), os.path.join(ROOT, "tests", "landmarks_core_asan_main.cpp"),
                        "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ -fsanitize=address,undefined not usable here: %s" % r.stderr[-300:])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    env.pop("LD_PRELOAD", None)
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert run.returncode == 0 and "asan driver ok" in run.stdout, (run.stdout[-500:], run.stderr[-2000:])


def test_warp_row_scan_equals_serial_scan(host):
    """The word-parallel formulation of the row scan (what lm_row_first_start_warp does with one shuffle and two ballots, here
    with loops over 32 lanes) == the serial scan, on random planes with random marks -- far more mark patterns than real
    traces leave -- for a fresh row (x_after = 0) and for a rescan after every marked pixel."""
    rng = np.random.default_rng(11)
    total = 0
    for it in range(300):
        Ww = int(rng.integers(1, 33))
        rows = int(rng.integers(1, 12))
        dens = rng.choice([0.02, 0.2, 0.5, 0.9])
        F = (rng.random((rows, Ww * 32)) < dens)
        F[:, 0] = False                                             # the frame column
        Mk = F & (rng.random(F.shape) < rng.choice([0.0, 0.1, 0.5]))   # marks live on foreground pixels only
        Nk = Mk & (rng.random(F.shape) < 0.5)

        def pack(a):
            return np.ascontiguousarray(np.packbits(a.reshape(rows, Ww, 32), axis=2, bitorder="little").view(np.uint32).reshape(rows, Ww))

        f, m, n = pack(F), pack(Mk), pack(Nk)
        bad = host.lm_host_row_scan_mismatches(f.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p),
                                               n.ctypes.data_as(ctypes.c_void_p), rows, Ww)
        assert bad == 0, (it, Ww, rows)
        total += rows
    assert total > 1000
