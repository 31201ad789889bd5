"""The drivers' evaluation metrics on the B200: iris_b200.cal_IoUs (one fused pass, exact counts: asserted EQUAL to the
unmodified reference's golden values) and iris_b200.angular_distance."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(golden_dir):
    import iris_b200
    from oracle import metrics_oracle as M

    return dict(ib=iris_b200, M=M, gold=np.load(os.path.join(golden_dir, "metrics.npz")))


@pytest.mark.parametrize("name,shape", [("2019", (640, 400)), ("small", (37, 53))])
def test_ious_equal_reference_golden(env, name, shape):
    ib, gold = env["ib"], env["gold"]
    p, t = ib.synthetic.iou_case(shape)
    per_class, miou = ib.cal_IoUs(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda())
    assert len(per_class) == 4 and all(v.shape == (5,) and v.is_cuda for v in per_class) and miou.shape == (5,)
    got = torch.stack(per_class, dim=1).cpu().numpy()
    assert np.array_equal(got, gold["iou_" + name])
    np.testing.assert_allclose(miou.cpu().numpy(), gold["miou_" + name], rtol=2e-7, atol=0)


def test_ious_shapes_classes_and_dtypes(env):
    ib, M = env["ib"], env["M"]
    rng = np.random.default_rng(3)
    for (b, h, w, nc) in [(1, 1, 1, 4), (3, 5, 7, 2), (2, 33, 31, 8), (7, 64, 96, 4), (1, 400, 640, 4), (2, 17, 1025, 3)]:
        p = rng.integers(-1, nc + 2, size=(b, h, w))         # labels outside 0..nc-1 belong to no class
        t = rng.integers(-1, nc + 2, size=(b, h, w))
        per_class, miou = ib.cal_IoUs(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda(), num_class=nc)
        iou, m = M.cal_ious(p, t, nc)
        assert np.array_equal(torch.stack(per_class, dim=1).cpu().numpy(), iou), (b, h, w, nc)
        np.testing.assert_allclose(miou.cpu().numpy(), m, rtol=2e-7)
    # uint8 predictions (RITnet labels stored compactly), CPU targets, broadcasting a single ground truth like data_preprocessing.py:168
    p, t = ib.synthetic.iou_case((37, 53))
    a = ib.cal_IoUs(torch.from_numpy(p).to(torch.uint8).cuda(), torch.from_numpy(t))[1]
    assert torch.equal(a, ib.cal_IoUs(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda())[1])
    one = ib.cal_IoUs(torch.from_numpy(p[:1]).cuda(), torch.from_numpy(t[0]).cuda().unsqueeze(0))[1]
    assert one.shape == (1,)
    with pytest.raises(ValueError):
        ib.cal_IoUs(torch.zeros(4, 4).cuda(), torch.zeros(4, 4).cuda())
    with pytest.raises(ib._lib.IsxError):
        ib.cal_IoUs(torch.zeros(1, 4, 4).cuda(), torch.zeros(1, 4, 4).cuda(), num_class=9)


def test_angular_distance(env):
    ib, gold = env["ib"], env["gold"]
    a, b = ib.synthetic.gaze_vector_case()
    rad, deg = ib.angular_distance(torch.from_numpy(a).cuda(), torch.from_numpy(b))
    np.testing.assert_allclose(rad.cpu().numpy(), gold["rad"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(deg.cpu().numpy(), gold["deg"], rtol=2e-6, atol=2e-4)
    # a plain fp32 PyTorch reference of the same op on the device
    va, vb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dot = torch.sum(va * vb, dim=1)
    ref = torch.acos(torch.clamp(dot, -1.0, 1.0))
    ok = dot.abs() < 0.999          # acos is ill conditioned at +-1: one ulp of the dot product (summation order) moves it by 1e-4
    torch.testing.assert_close(rad[ok], ref[ok], rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(deg[ok], torch.rad2deg(ref)[ok], rtol=2e-6, atol=2e-4)
    torch.testing.assert_close(torch.cos(rad), torch.clamp(dot, -1.0, 1.0), rtol=0, atol=1e-6)
