"""K10 / K9 on the B200 against the oracle and the golden vectors of the reference: mask + bbox
(bit-exact integer work), crop, antialiased resize and the composite back into the eye frame."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import iris_b200
    from oracle import nst_oracle as O

    return iris_b200, O


def rand_img(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)


def test_mask_bbox_bit_exact(mods, golden_dir):
    ib, O = mods
    m = np.load(os.path.join(golden_dir, "mask_bbox.npz"))
    for k, (seed, h, w) in enumerate([(5, 640, 400), (6, 400, 640), (7, 64, 48)]):
        frame, seg = ib.synthetic.synthetic_eye(seed, h, w)
        xt, st = torch.from_numpy(frame).cuda(), torch.from_numpy(seg).cuda()
        xc, mc, x0, y0, x1, y1 = ib.mask_and_crop_iris(xt, seg=st, device="cuda:0")
        rxc, rmc, rx0, ry0, rx1, ry1 = O.mask_and_crop(frame, seg)
        assert [x0, y0, x1, y1] == list(m["syn%d_bbox" % k]) == [rx0, ry0, rx1, ry1]
        assert mc.dtype == torch.bool and np.array_equal(mc.cpu().numpy(), rmc)
        assert np.array_equal(xc.cpu().numpy(), rxc)  # x * m is exact
        assert int(mc.sum()) == int(m["syn%d_mask_count" % k])
        # same thing through a callable segmenter, like the reference's ritnet argument
        xc2, _, *bb = ib.mask_and_crop_iris(xt, ritnet=lambda t: st, device="cuda:0")
        assert bb == [x0, y0, x1, y1] and torch.equal(xc2, xc)
    # real eye frames through the shipped RITnet (fixtures hold the two boolean masks)
    for k in range(2):
        shape = tuple(m["real%d_shape" % k])
        n = int(np.prod(shape))
        iris = np.unpackbits(m["real%d_iris_bits" % k])[:n].reshape(shape).astype(bool)
        nog = np.unpackbits(m["real%d_noglint_bits" % k])[:n].reshape(shape).astype(bool)
        img = torch.from_numpy((iris & nog).astype(np.float32)).cuda()
        assert list(ib.crop_image(img, return_idx=True)) == list(m["real%d_bbox" % k])
    assert list(m["real0_bbox"]) == [223, 92, 393, 297]  # notebook cell 2: crop 171 x 206


def test_crop_image_edge_cases(mods):
    ib, O = mods
    with pytest.raises(Exception):
        ib.crop_image(torch.zeros(2, 4, 4, device="cuda"))
    with pytest.raises(RuntimeError):
        ib.crop_image(torch.zeros(1, 4, 4, device="cuda"))
    a = torch.zeros(5, 7, device="cuda")
    a[2, 3] = 1e-30
    assert ib.crop_image(a, return_idx=True) == (2, 3, 2, 3)
    a[4, 0] = -1.0
    assert ib.crop_image(a, return_idx=True) == (2, 0, 4, 3)
    assert tuple(ib.crop_image(a[None]).shape) == (1, 3, 4)
    # batched recipe == per-image recipe
    frames, segs = ib.synthetic.synthetic_batch([1, 2, 3], 128, 96)
    mask, bbox = ib.iris_masks_and_bboxes(torch.from_numpy(frames).cuda(), torch.from_numpy(segs).cuda())
    for i in range(3):
        _, rmc, *rbb = O.mask_and_crop(frames[i], segs[i])
        assert bbox[i].tolist() == rbb
        assert int(mask[i].sum()) == int(((segs[i] == 2) & (frames[i] <= np.float32(0.8))).sum())


def test_composite_and_resize(mods, golden_dir):
    ib, O = mods
    k = np.load(os.path.join(golden_dir, "composite.npz"))
    frames, masks, bboxes, news = [], [], [], []
    for idx, (seed, h, w) in enumerate([(5, 640, 400), (6, 400, 640)]):
        frame, seg = ib.synthetic.synthetic_eye(seed, h, w)
        ft, st = torch.from_numpy(frame).cuda()[None], torch.from_numpy(seg).cuda()[None]
        mask, bbox = ib.iris_masks_and_bboxes(ft, st)
        assert bbox[0].tolist() == list(k["comp%d_bbox" % idx])
        new_rgb = rand_img(40 + idx, (1, 3, 224, 224)).cuda()
        out = ib.composite_irises(ft.clone(), new_rgb, mask, bbox)
        x_min, y_min = bbox[0, 0].item(), bbox[0, 1].item()
        got = out[0].cpu().numpy()
        np.testing.assert_allclose(got[:, x_min:x_min + 24, y_min + 40:y_min + 64], k["comp%d_patch" % idx], atol=1e-5)
        np.testing.assert_allclose(got, k["comp%d_out" % idx].astype(np.float32), atol=1e-3)
        assert float(got.astype(np.float64).sum()) == pytest.approx(float(k["comp%d_out_sum" % idx]), rel=1e-6)
        ref = O.composite(frame, new_rgb[0].cpu().numpy(), (seg == 2) & (frame <= np.float32(0.8)), bbox[0].tolist())
        np.testing.assert_allclose(got, ref, atol=1e-5)
        # outside the mask the frame is untouched, bit for bit
        keep = ~mask[0].bool().cpu().numpy()
        assert np.array_equal(got[keep], frame[keep])
        # forward resize of the drivers: crop -> 224 x 224 x 3
        crops = ib.crop_resize_irises(ft, mask, bbox)
        assert tuple(crops.shape) == (1, 3, 224, 224) and torch.equal(crops[:, 0], crops[:, 2])
        np.testing.assert_allclose(crops[0, :1, 100:116, 100:116].cpu().numpy(), k["resize%d_224_patch" % idx], atol=1e-5)
        assert float(crops[0, 0].double().sum()) == pytest.approx(float(k["resize%d_224_sum" % idx]), rel=1e-6)


def test_composite_batched_ragged_bboxes(mods):
    """One launch for a batch whose bounding boxes all differ (the reference loops in Python, …2019.py:116-130)."""
    ib, O = mods
    seeds = [11, 12, 13, 14]
    frames, segs = ib.synthetic.synthetic_batch(seeds, 200, 160)
    ft, st = torch.from_numpy(frames).cuda(), torch.from_numpy(segs).cuda()
    mask, bbox = ib.iris_masks_and_bboxes(ft, st)
    new = rand_img(77, (4, 3, 64, 64)).cuda()
    out = ib.composite_irises(ft.clone(), new, mask, bbox).cpu().numpy()
    for i in range(4):
        m = (segs[i] == 2) & (frames[i] <= np.float32(0.8))
        ref = O.composite(frames[i], new[i].cpu().numpy(), m, bbox[i].tolist())
        np.testing.assert_allclose(out[i], ref, atol=1e-5)


def test_stylize_frames_end_to_end(mods):
    """The drivers' per-batch body (…2019.py:64-79,93-100,111-137) as ONE call: masks / bboxes bit-exact, the 224x224
    crops and the composite equal to the oracle's per-image chain fed with the same new irises, frames without an iris
    returned untouched, losses decreasing."""
    ib, O = mods
    H, W = 120, 96
    frames, segs = ib.synthetic.synthetic_batch([21, 22, 23], H, W)
    segs[1] = 0                                            # frame 1: the segmenter found no iris
    sfr, sseg = ib.synthetic.synthetic_eye(31, H, W)
    vgg = ib.VGG19(weights="random", seed=0)
    ft = torch.from_numpy(frames)
    # style iris prepared like …2020.py:238-249: masked, cropped, resized, UNBATCHED (1,h,w)
    sm, sb = ib.iris_masks_and_bboxes(torch.from_numpy(sfr)[None].cuda(), torch.from_numpy(sseg)[None].cuda())
    s_iris = ib.crop_resize_irises(torch.from_numpy(sfr)[None].cuda(), sm, sb, size=(64, 64))[0, :1]
    out, info = ib.stylize_frames(ft, s_iris, segs=torch.from_numpy(segs), vgg=vgg, s_loss_weight=1e4, epochs=20, size=(64, 64))
    assert info["valid"].tolist() == [True, False, True] and info["evals"] == 20
    assert tuple(out.shape) == (3, 1, H, W) and out.is_cuda
    got = out.cpu().numpy()
    assert np.array_equal(got[1], frames[1])               # untouched
    new = info["irises"].cpu().numpy()                     # the NST results of the two valid frames
    for k, i in enumerate([0, 2]):
        m = (segs[i] == 2) & (frames[i] <= np.float32(0.8))
        xc, mc, *bb = O.mask_and_crop(frames[i], segs[i])
        assert info["bboxes"][i].tolist() == bb and np.array_equal(info["masks"][i].cpu().numpy().astype(bool), m)
        ref = O.composite(frames[i], new[k], m, bb)
        np.testing.assert_allclose(got[i], ref, atol=1e-5)
        assert np.abs(got[i] - frames[i])[m].mean() > 1e-4   # the iris texture did change
    sh = info["s_loss_hist"]
    assert np.isfinite(sh).all() and sh[-1] < sh[0]
    # segmenter callable instead of label maps
    out2, info2 = ib.stylize_frames(ft, s_iris, segmenter=lambda x: torch.from_numpy(segs).cuda(), vgg=vgg, s_loss_weight=1e4,
                                    epochs=20, size=(64, 64))
    assert float((out2 - out).abs().max()) < 1e-4


def test_crop_resize_masked_matches_oracle_resize(mods):
    """isx_crop_resize_masked (mask multiply on the source taps inside the kernel) == oracle: (frame*mask)[bbox] -> AA resize,
    for down- and up-scaling windows and a size that needs more taps than the shared-memory table holds."""
    ib, O = mods
    frames, segs = ib.synthetic.synthetic_batch([5, 6], 640, 400)
    ft, st = torch.from_numpy(frames).cuda(), torch.from_numpy(segs).cuda()
    mask, bbox = ib.iris_masks_and_bboxes(ft, st)
    for size in ((224, 224), (300, 64), (7, 5)):
        crops = ib.crop_resize_irises(ft, mask, bbox, size=size).cpu().numpy()
        for i in range(2):
            m = (segs[i] == 2) & (frames[i] <= np.float32(0.8))
            x0, y0, x1, y1 = bbox[i].tolist()
            ref = O.resize_bilinear_aa((frames[i] * m)[:, x0:x1 + 1, y0:y1 + 1], size[0], size[1])
            np.testing.assert_allclose(crops[i, :1], ref, atol=1e-5)
            assert np.array_equal(crops[i, 0], crops[i, 2])
