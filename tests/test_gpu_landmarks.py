"""Row f4 on the B200: iris_b200.extract_eye_landmarks(_batch) and the gaze heads against (i) the 19 landmarks the UNMODIFIED
reference produced (tests/golden/landmarks.npz), (ii) OpenCV itself -- the library the reference calls
(gaze_estimators.py:70-81) -- on random masks: the chosen contour's point count and the number of external contours are
index work and asserted EQUAL, the ellipse within float32 rounding, (iii) the oracle (oracle/landmarks_oracle.py)."""
import os

import numpy as np
import pytest
import torch

from test_landmarks_oracle import random_mask

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def env(golden_dir):
    import iris_b200
    from oracle import landmarks_oracle as L

    return dict(ib=iris_b200, L=L, gold=np.load(os.path.join(golden_dir, "landmarks.npz")))


def test_landmarks_equal_reference_golden(env):
    ib, gold = env["ib"], env["gold"]
    cases = ib.synthetic.landmark_cases()
    segs = torch.from_numpy(np.stack([lab for _, lab in cases])).cuda()           # [10,400,640] int64: one call
    out, info = ib.extract_eye_landmarks_batch(segs, return_info=True)
    assert out.shape == (len(cases), 19) and out.dtype == torch.float32 and out.is_cuda
    out, info = out.cpu().numpy(), info.cpu().numpy()
    for k, (name, lab) in enumerate(cases):
        ref = gold["lm_" + name]
        print(name, "pupil pts/contours %d/%d iris %d/%d flags %d %d  max |diff| %.2e" % (
            info[k, 0], info[k, 1], info[k, 3], info[k, 4], info[k, 2], info[k, 5], float(np.abs(out[k] - ref).max())))
        np.testing.assert_allclose(out[k], ref, rtol=2e-6, atol=5e-5, err_msg=name)
        assert np.array_equal(out[k, 10:16], ref[10:16]), name                     # eye corners, width, height: exact
        assert np.array_equal(out[k] == 0, ref == 0), name                         # absent classes -> zeros
        assert info[k, 2] == 0 and info[k, 5] == 0
        # the reference-shaped single-frame call
        one = ib.extract_eye_landmarks(torch.from_numpy(lab).cuda())
        assert one.shape == (19,) and np.array_equal(one.cpu().numpy(), out[k])


def test_dtypes_and_4d_input(env):
    ib = env["ib"]
    lab = torch.from_numpy(np.stack([ib.synthetic.synthetic_label_map(s, speck=0.01) for s in (41, 42)])).cuda()
    a = ib.extract_eye_landmarks_batch(lab)
    assert torch.equal(a, ib.extract_eye_landmarks_batch(lab.to(torch.uint8)))
    assert torch.equal(a, ib.extract_eye_landmarks_batch(lab.to(torch.int32)))
    assert torch.equal(a, ib.extract_eye_landmarks_batch(lab[:, None]))
    assert torch.equal(a, ib.extract_eye_landmarks_batch((lab + 256).cpu()))       # .astype(np.uint8) wraps (gaze_estimators.py:127)
    with pytest.raises(ValueError):
        ib.extract_eye_landmarks_batch(lab[0, 0])
    with pytest.raises(AssertionError):
        ib.extract_eye_landmarks(lab[0, :100])


@pytest.mark.parametrize("shape", [(24, 31), (40, 64), (57, 97), (120, 200), (400, 640), (640, 400)])
def test_contours_and_ellipses_equal_cv2(env, shape):
    """Random pupil / iris masks (speckle, blobs, rings with a component inside the hole, clipped ellipses) in one batch."""
    ib, L = env["ib"], env["L"]
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    B = 24 if H * W < 100000 else 8
    labs = np.zeros((B, H, W), np.int64)
    for b in range(B):
        m3 = random_mask(rng, b % 4, H, W)
        m2 = random_mask(rng, (b + 1) % 4, H, W) & (1 - m3)
        labs[b][m2 > 0] = 2
        labs[b][m3 > 0] = 3
        labs[b][(rng.random((H, W)) < 0.05) & (labs[b] == 0)] = 1
    out, info = ib.extract_eye_landmarks_batch(torch.from_numpy(labs).cuda(), return_info=True)
    out, info = out.cpu().numpy(), info.cpu().numpy()
    checked = 0
    for b in range(B):
        for cls, o, base in ((3, 0, 0), (2, 5, 3)):
            m = (labs[b] == cls).astype(np.uint8)
            cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            assert info[b, base + 1] == len(cs), (b, cls)
            if not cs:
                assert info[b, base] == 0 and not out[b, o:o + 5].any()
                continue
            c = max(cs, key=cv2.contourArea)
            assert info[b, base] == len(c), (b, cls)
            if len(c) < 5:
                assert not out[b, o:o + 5].any()
                continue
            if len(c) == 5:
                assert info[b, base + 2] & 1, (b, cls)
            if (info[b, base + 2] & 1) or L.is_degenerate(c):      # specks where cv2.fitEllipse leaves the general algorithm (isx.h)
                continue
            e = cv2.fitEllipse(c)
            ref = np.array([e[0][0], e[0][1], e[1][0], e[1][1], e[2]], np.float32)
            if not np.all(np.isfinite(ref)):
                continue
            np.testing.assert_allclose(out[b, o:o + 4], ref[:4], rtol=1e-4, atol=1e-3, err_msg=str((b, cls)))
            da = abs(float(out[b, o + 4]) - float(ref[4])) % 180.0
            assert min(da, 180.0 - da) < 2e-2, (b, cls, out[b, o:o + 5], ref)
            checked += 1
        ys, xs = np.nonzero(labs[b] == 1)
        if len(xs):
            assert out[b, 10:14].tolist() == [xs.min(), xs.max(), ys.min(), ys.max()]
    assert checked > B // 2


def test_point_cap_flag(env):
    ib = env["ib"]
    lab = np.zeros((2, 64, 64), np.int64)
    lab[0, 10:50, 10:50] = 3 * (np.indices((40, 40)).sum(0) % 2)       # a diagonal-connected checkerboard: hundreds of corners
    lab[1, 20:40, 20:40] = 3
    out, info = ib.extract_eye_landmarks_batch(torch.from_numpy(lab).cuda(), max_points=16, return_info=True)
    info = info.cpu().numpy()
    assert info[0, 2] & 2 and info[0, 0] > 16 and not out[0, :5].any()
    assert info[1, 2] == 0 and info[1, 0] == 4                             # a square: four corners, fewer than five points
    full = ib.extract_eye_landmarks_batch(torch.from_numpy(lab).cuda(), max_points=4096, return_info=True)[1].cpu().numpy()
    assert full[0, 2] & 2 == 0 and full[0, 0] == info[0, 0]


@pytest.mark.parametrize("in_dim,cls", [(19, "GazeEstimator1"), (2048, "GazeEstimator2")])
def test_gaze_heads(env, in_dim, cls):
    ib, L, gold = env["ib"], env["L"], env["gold"]
    params, x = ib.synthetic.gaze_head_case(in_dim)
    names = ["model.0.weight", "model.0.bias", "model.3.weight", "model.3.bias", "model.6.weight", "model.6.bias"]
    sd = {k: torch.from_numpy(p) for k, p in zip(names, params)}
    net = getattr(ib, cls)(extract_feature=False, state_dict=sd).to("cuda:0")
    out = net(torch.from_numpy(x).cuda())
    assert out.shape == (x.shape[0], 3) and out.is_cuda
    np.testing.assert_allclose(out.cpu().numpy(), gold["head%d_out" % in_dim], rtol=1e-4, atol=2e-5)   # the unmodified reference module
    np.testing.assert_allclose(out.cpu().numpy(), L.gaze_head(x, params), rtol=1e-4, atol=2e-5)
    # a plain fp32 PyTorch reference of the same op, on the device, for odd batch sizes and strided rows
    for B in (1, 3, 16, 17, 130):
        xb = torch.randn(B, in_dim + 5, device="cuda", generator=torch.Generator("cuda").manual_seed(B))[:, 2:2 + in_dim]
        ref = net.model(xb)
        ref = ref / torch.norm(ref, dim=1, keepdim=True)
        torch.testing.assert_close(net(xb), ref, rtol=1e-4, atol=2e-5)
    net.train()
    with pytest.raises(RuntimeError):
        net(torch.from_numpy(x).cuda())


def test_gaze_estimator1_end_to_end(env):
    """Label maps in -> gaze vectors out (extract_feature=True, gaze_estimators.py:48-53) == oracle landmarks -> oracle head."""
    ib, L = env["ib"], env["L"]
    params, _ = ib.synthetic.gaze_head_case(19)
    names = ["model.0.weight", "model.0.bias", "model.3.weight", "model.3.bias", "model.6.weight", "model.6.bias"]
    net = ib.GazeEstimator1(extract_feature=True, state_dict={k: torch.from_numpy(p) for k, p in zip(names, params)}).to("cuda:0")
    labs = np.stack([ib.synthetic.synthetic_label_map(s, speck=sp) for s, sp in ((51, 0.0), (52, 0.005), (53, 0.05))])
    got = net(torch.from_numpy(labs).cuda()).cpu().numpy()
    ref = L.gaze_head(np.stack([L.extract_eye_landmarks(l) for l in labs]), params)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5)
    with pytest.raises(ValueError):
        ib.GazeEstimator2(extract_feature=True)
