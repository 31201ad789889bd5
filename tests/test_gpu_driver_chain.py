"""The drivers' per-batch body end to end on the B200, every stage from this library and nothing leaving the device in
between (iris_style_transfer_openeds2019.py:64-160 / …2020.py:78-150 with the shipped RITnet as the segmenter):
frames -> RITnet labels -> mask / bbox / crop / resize -> nst -> composite -> RITnet again -> cal_IoUs(pre, post) ->
extract_eye_landmarks -> GazeEstimator1 -> angular_distance(pre, post).  Every intermediate is checked against the oracle
on the same inputs (index work exact), the chain as a whole for the invariants the reference's design implies."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_2020_driver_flow(golden_dir, vgg_weights):
    import iris_b200 as ib
    from oracle import landmarks_oracle as L, metrics_oracle as M

    gold = np.load(os.path.join(golden_dir, "ritnet.npz"))
    rit = ib.RITnet(state_dict={k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w:")})
    vgg = ib.VGG19(weights=vgg_weights)
    B, H, W = 4, 400, 640
    frames_np, _ = ib.synthetic.synthetic_batch([71, 72, 73, 74], H, W)
    frames = torch.from_numpy(frames_np).cuda()
    style_np, style_seg = ib.synthetic.synthetic_eye(75, H, W)
    sx = torch.from_numpy(style_np).cuda()[None]
    smask, sbb = ib.iris_masks_and_bboxes(sx, torch.from_numpy(style_seg).cuda()[None])
    s_iris = ib.crop_resize_irises(sx, smask, sbb, size=(64, 64))[0, :1].contiguous()

    pre = rit(frames)                                                   # [B,H,W] int64 on the device
    assert pre.shape == (B, H, W) and pre.dtype == torch.int64 and pre.is_cuda
    out, info = ib.stylize_frames(frames, s_iris, segs=pre, vgg=vgg, s_loss_weight=1e4, epochs=20, size=(64, 64))
    assert info["evals"] == 20 and out.shape == frames.shape and bool(torch.isfinite(out).all())
    masks = info["masks"].bool()
    valid = info["valid"]
    # the composite touches iris pixels only (…2020.py:137: frame[bbox] * ~m + new): everything outside the mask is bit-equal
    assert torch.equal(out[~masks], frames[~masks])
    if bool(valid.any()):
        assert float((out - frames).abs().sum()) > 0

    post = rit(out)
    per_class, miou = ib.cal_IoUs(post, pre)
    iou_ref, miou_ref = M.cal_ious(post.cpu().numpy(), pre.cpu().numpy())
    assert np.array_equal(torch.stack(per_class, dim=1).cpu().numpy(), iou_ref)
    np.testing.assert_allclose(miou.cpu().numpy(), miou_ref, rtol=2e-7)
    print("IoU pre/post per class (mean over frames):", iou_ref.mean(axis=0).round(3).tolist(), "mIoU", miou_ref.round(3).tolist())

    lm_pre, lm_post = ib.extract_eye_landmarks_batch(pre), ib.extract_eye_landmarks_batch(post)
    for k in range(B):
        for lm, lab in ((lm_pre, pre), (lm_post, post)):
            want = L.extract_eye_landmarks(lab[k].cpu().numpy())
            np.testing.assert_allclose(lm[k].cpu().numpy(), want, rtol=2e-6, atol=5e-5)
    params, _ = ib.synthetic.gaze_head_case(19)
    names = ["model.0.weight", "model.0.bias", "model.3.weight", "model.3.bias", "model.6.weight", "model.6.bias"]
    net = ib.GazeEstimator1(extract_feature=True, state_dict={k: torch.from_numpy(p) for k, p in zip(names, params)}).to("cuda:0")
    g_pre, g_post = net(pre), net(post)
    assert g_pre.shape == (B, 3)
    torch.testing.assert_close(g_pre.norm(dim=1), torch.ones(B, device="cuda"), rtol=1e-5, atol=1e-5)
    rad, deg = ib.angular_distance(g_pre, g_post)
    r_ref, d_ref = M.angular_distance(g_pre.cpu().numpy(), g_post.cpu().numpy())
    ok = np.abs(np.sum(g_pre.cpu().numpy() * g_post.cpu().numpy(), axis=1)) < 0.999
    np.testing.assert_allclose(rad.cpu().numpy()[ok], r_ref[ok], rtol=1e-5, atol=1e-5)
    print("gaze change after stylisation (degrees):", deg.cpu().numpy().round(3).tolist())
