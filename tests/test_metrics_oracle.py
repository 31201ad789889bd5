"""Pin oracle/metrics_oracle.py (utils.cal_IoUs, utils.angular_distance restated) against outputs of the UNMODIFIED
reference functions (tests/golden/metrics.npz, make_golden_metrics.py)."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as M
from test_landmarks_oracle import _synthetic


@pytest.mark.parametrize("name,shape", [("2019", (640, 400)), ("small", (37, 53))])
def test_ious_equal_reference_golden(golden_dir, name, shape):
    gold = np.load(os.path.join(golden_dir, "metrics.npz"))
    p, t = _synthetic().iou_case(shape)
    iou, miou = M.cal_ious(p, t)
    assert np.array_equal(iou, gold["iou_" + name])              # exact counts, one float32 division: bit-identical
    np.testing.assert_allclose(miou, gold["miou_" + name], rtol=2e-7, atol=0)
    assert np.all(iou[0] > 0.999999) and iou[4, 3] == 0.0   # identical maps (n / (n + 1e-6) rounds below 1 for a tiny class); empty union -> 0 / eps


def test_angular_distance_equals_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "metrics.npz"))
    a, b = _synthetic().gaze_vector_case()
    rad, deg = M.angular_distance(a, b)
    np.testing.assert_allclose(rad, gold["rad"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(deg, gold["deg"], rtol=1e-6, atol=1e-4)
    assert rad[0] < 1e-3 and abs(deg[1] - 180.0) < 0.1
