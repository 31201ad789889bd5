"""Pin oracle/ritnet_oracle.py (CPU restatement of the reference's RITnet mask producer) against the label maps the
UNMODIFIED reference produced (tests/golden/ritnet.npz, make_golden_ritnet.py) and its CLAHE restatement against cv2."""
import os

import numpy as np
import pytest
import torch

from oracle import ritnet_oracle as R

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ritnet.npz"))


@pytest.fixture(scope="module")
def sd(gold):
    return {k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w:")}


def _syn(seed, h, w):
    import importlib.util

    spec = importlib.util.spec_from_file_location("isx_synthetic", os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iris-style-transfer_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    return torch.from_numpy(syn.synthetic_eye(seed, h, w)[0])


@pytest.mark.parametrize("k,seed,h,w", [(2, 7, 64, 48), (3, 8, 160, 96), (0, 5, 640, 400)])
def test_labels_match_reference(gold, sd, k, seed, h, w):
    x = _syn(seed, h, w)
    xt = R.ritnet_transform(x)
    assert float(xt.double().sum()) == pytest.approx(float(gold["syn%d_transform_sum" % k]), rel=1e-12)
    logits = R.densenet2d_logits(sd, xt)
    assert float(logits.double().abs().sum()) == pytest.approx(float(gold["syn%d_logit_abs_sum" % k]), rel=1e-5)
    assert np.array_equal(logits.max(1)[1].numpy().astype(np.uint8), gold["syn%d_labels" % k])
    if k == 2:
        assert np.array_equal(xt.numpy(), gold["syn2_transform"])
        np.testing.assert_allclose(logits.numpy(), gold["syn2_logits"], rtol=1e-5, atol=1e-4)


def test_clahe_restatement_is_bit_exact_against_cv2():
    import cv2

    rng = np.random.default_rng(0)
    for (h, w) in [(640, 400), (400, 640), (64, 48), (75, 101), (33, 17), (8, 8), (641, 400), (640, 403)]:
        for img in (rng.integers(0, 256, (h, w), dtype=np.uint8),
                    (rng.integers(0, 40, (h, w)) + (np.arange(w)[None, :] * 200 // w)).astype(np.uint8)):
            ref = cv2.createCLAHE(clipLimit=1.5, tileGridSize=(8, 8)).apply(img)
            assert np.array_equal(R.clahe_numpy(img), ref), (h, w)


def test_transform_tables_and_package_packing(sd):
    from iris_b200 import ritnet

    assert np.array_equal(ritnet.gamma_table_u8(), R.gamma_table_u8())
    assert np.array_equal(ritnet.normalize_table_f32(), R.normalize_table_f32())
    g = R.gamma_table_u8()
    assert g[0] == 0 and g[255] == 255 and (np.diff(g.astype(int)) >= 0).all()
    blob = ritnet.pack_ritnet_params(sd)
    assert blob.dtype == torch.float32 and blob.numel() == 248900
    # first conv (1 -> 32, 3x3) is stored as [tap][cin][32]
    w = sd["down_block1.conv1.weight"]
    assert torch.equal(blob[:288].reshape(9, 1, 32), w.permute(2, 3, 1, 0).reshape(9, 1, 32))
