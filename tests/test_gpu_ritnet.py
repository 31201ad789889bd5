"""The mask producer on the B200: iris_b200.RITnet against the label maps of the UNMODIFIED reference
(tests/golden/ritnet.npz) -- label maps are index work: asserted EQUAL -- and RITnet_transform bit-exact against the
reference's own OpenCV round trip (cv2 is the third-party library the reference calls, ritnet.py:93-94)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(golden_dir):
    import iris_b200
    from oracle import nst_oracle as O, ritnet_oracle as R

    gold = np.load(os.path.join(golden_dir, "ritnet.npz"))
    sd = {k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w:")}
    return dict(ib=iris_b200, O=O, R=R, gold=gold, sd=sd, net=iris_b200.RITnet(state_dict=sd))


def test_transform_bit_exact_against_cv2(env):
    import cv2

    R, net = env["R"], env["net"]
    g = torch.Generator().manual_seed(1)
    for (h, w) in [(640, 400), (400, 640), (64, 48), (75, 101), (33, 17), (8, 8), (641, 400)]:
        x = torch.rand(2, 1, h, w, generator=g)
        x[1] = torch.from_numpy(env["ib"].synthetic.synthetic_eye(3 + h, h, w)[0])
        got = net.transform(x.cuda()).cpu()
        for i in range(2):
            ref = R.ritnet_transform(x[i], use_cv2=True)     # uint8 -> gamma LUT -> cv2 CLAHE -> ToDtype/Normalize
            assert torch.equal(got[i:i + 1], ref), (h, w, i, int((got[i:i + 1] != ref).sum()))


@pytest.mark.parametrize("k,seed,h,w", [(0, 5, 640, 400), (1, 6, 400, 640), (2, 7, 64, 48), (3, 8, 160, 96)])
def test_labels_equal_reference(env, k, seed, h, w):
    ib, gold, net = env["ib"], env["gold"], env["net"]
    x = torch.from_numpy(ib.synthetic.synthetic_eye(seed, h, w)[0]).cuda()      # (1,h,w) like the drivers pass it
    lab, logits = net(x, return_logits=True)
    assert lab.dtype == torch.int64 and tuple(lab.shape) == (1, h, w) and lab.is_cuda
    ref = gold["syn%d_labels" % k]
    diff = int((lab.cpu().numpy().astype(np.uint8) != ref).sum())
    print("syn%d %dx%d: %d labels differ, class counts %s" % (k, h, w, diff, np.bincount(ref.reshape(-1), minlength=4).tolist()))
    assert diff == 0
    assert float(logits.double().abs().sum()) == pytest.approx(float(gold["syn%d_logit_abs_sum" % k]), rel=1e-4)
    if k == 2:
        np.testing.assert_allclose(logits.cpu().numpy(), gold["syn2_logits"], rtol=1e-4, atol=2e-3)


def test_batched_equals_single_and_feeds_the_mask_stage(env):
    """A batch in one call == per-image calls; the label map drives mask_and_crop_iris / stylize_frames like the reference's
    ritnet argument (pipelines.py:133-141), bbox and mask equal to the oracle chain on the reference's labels."""
    ib, O, gold, net = env["ib"], env["O"], env["gold"], env["net"]
    frames, _ = ib.synthetic.synthetic_batch([5, 9, 10], 640, 400)
    xb = torch.from_numpy(frames).cuda()
    lab = net(xb)
    assert tuple(lab.shape) == (3, 640, 400)
    for i in range(3):
        assert torch.equal(lab[i:i + 1], net(xb[i]))
    assert np.array_equal(lab[0].cpu().numpy().astype(np.uint8), gold["syn0_labels"][0])
    xc, mc, *bb = ib.mask_and_crop_iris(xb[0], ritnet=net, device="cuda:0")
    rxc, rmc, *rbb = O.mask_and_crop(frames[0], gold["syn0_labels"].astype(np.int64))
    assert bb == rbb and np.array_equal(mc.cpu().numpy(), rmc) and np.array_equal(xc.cpu().numpy(), rxc)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 100, 64, device="cuda"))        # 100 is not a multiple of 16: the reference fails in torch.cat
