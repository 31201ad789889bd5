"""Generate tests/golden/*.npz from the UNMODIFIED reference imported live from /root/reference.

Run once in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Everything stored is an OUTPUT of reference code (pipelines.nst, utils.GramMatrix / StyleLoss_* /
ContentLoss_L2 / crop_image, pipelines.mask_and_crop_iris, models.VGG19, models.RITnet) or of the
third-party ops the drivers call for the composite (torchvision v2 rgb_to_grayscale / Resize),
on seeded synthetic inputs.  VGG weights = torchvision vgg19(weights=None) under
torch.manual_seed(0); only a checksum of them is stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "iris-style-transfer_b200"))
import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
pipelines, utils, models = ref_loader.load(seed=0)


def rand_img(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)


def weights_checksum(vgg):
    tot = 0.0
    for p in vgg.model.parameters():
        tot += float(p.double().abs().sum())
    return tot


def main():
    vgg = models.VGG19()
    vgg5 = models.VGG19(style_layers=["relu1_1", "relu2_1", "relu3_1", "relu4_1", "relu5_1"])
    g = {"weights_abs_sum": np.float64(weights_checksum(vgg))}

    # ---- (1) features / Gram / losses / gradient at a given x -------------------------------
    H, W = 48, 64
    c = rand_img(11, (2, 3, H, W))
    s = rand_img(12, (2, 3, H, W))
    xq = rand_img(13, (2, 3, H, W))
    with torch.no_grad():
        p5, c_f, s_f = vgg(c)
        _, _, s_t = vgg(s)
    g["eval_pool5"] = p5.numpy()
    g["eval_content_feat_sum"] = np.array([float(f.double().sum()) for f in c_f])
    for i, f in enumerate(s_f):
        G = utils.GramMatrix(f)
        if G.shape[-1] <= 128:
            g["eval_gram_c_%d" % i] = G.numpy()
        g["eval_gram_c_%d_corner" % i] = G[:, :32, -32:].numpy()
        g["eval_gram_c_%d_sums" % i] = np.array([float(G.double().sum()), float(G.double().abs().sum()),
                                                  float((G.double() ** 2).sum())])
    # Classifier2 feature reduction (classifiers.py:71)
    g["eval_style_features"] = torch.cat(
        [torch.cat([x.mean(dim=(-2, -1)), x.std(dim=(-2, -1))], dim=1) for x in s_f], dim=1).numpy()
    for name, BN in (("gram", False), ("bn", True)):
        xr = xq.clone().requires_grad_(True)
        closs = utils.ContentLoss_L2(targets=c_f)
        sloss = (utils.StyleLoss_BN if BN else utils.StyleLoss_Gram)(targets=s_t)
        _, x_c, x_s = vgg(xr)
        cl = closs(x_c)
        sl = sloss(x_s)
        (cl * 1.0 + sl * 1e6).backward()
        g["eval_%s_c_loss" % name] = np.float64(cl.item())
        g["eval_%s_s_loss" % name] = np.float64(sl.item())
        g["eval_%s_grad" % name] = xr.grad.numpy()
    # 5-layer (relu5_1) style tap
    with torch.no_grad():
        _, _, s5 = vgg5(c)
    G5 = utils.GramMatrix(s5[4])
    g["eval_gram5_c_4_corner"] = G5[:, :32, -32:].numpy()
    g["eval_gram5_c_4_sums"] = np.array([float(G5.double().sum()), float(G5.double().abs().sum()),
                                         float((G5.double() ** 2).sum())])
    # unbatched GramMatrix (SURVEY note N3): (C,H,W) input, n = H*W
    g["eval_gram_unbatched"] = utils.GramMatrix(s_f[1][0]).numpy()
    np.savez_compressed(os.path.join(OUT, "eval_48x64.npz"), **g)

    # ---- (2) nst trajectories -----------------------------------------------------------------
    t = {}

    def run(tag, c_img, s_img, seed_before=None, **kw):
        if seed_before is not None:
            torch.manual_seed(seed_before)
        x, x_hist, c_hist, s_hist = pipelines.nst(c_img, s_img, vgg=vgg, use_tqdm=False, device="cpu", **kw)
        t[tag + "_x"] = x.numpy()
        t[tag + "_c_hist"] = np.array(c_hist)
        t[tag + "_s_hist"] = np.array(s_hist)
        # NB: on CPU `x.detach().cpu()` (pipelines.py:93) aliases x, so every x_hist entry equals the
        # final image there; only on a CUDA device are they per-eval copies.  Not stored.
        print(tag, len(c_hist), "evals; moved MAE", float((x - c_img).abs().mean()), "s_loss", s_hist[0], "->", s_hist[-1])

    c1, s1 = rand_img(21, (1, 3, H, W)), rand_img(22, (1, 3, H, W))
    run("gram_b1", c1, s1, BN_loss=False, s_loss_weight=1e6, epochs=50)           # 60 evals
    run("bn_b1", c1, s1, BN_loss=True, s_loss_weight=1e4, epochs=40)
    run("gram_b2_coupled", c[:2], s[:2], BN_loss=False, s_loss_weight=1e6, epochs=20)
    run("gram_b2_style1", c[:2], s[:1], BN_loss=False, s_loss_weight=1e6, epochs=20)
    run("gram_rand_init", c1, s1, seed_before=123, clone_content=False, BN_loss=False, s_loss_weight=1e6, epochs=20)
    run("degenerate", c1, s1, BN_loss=False, s_loss_weight=1.0, epochs=5)          # |g| < 1e-7: x never moves
    run("gram_unbatched_style", c1, s1[0, :1], BN_loss=False, s_loss_weight=1e6, epochs=20)  # …2020.py:103 (N3)
    run("bn_unbatched_style", c1, s1[0, :1], BN_loss=True, s_loss_weight=1e4, epochs=20)
    # synthetic iris crops, 96x96 (drivers feed square crops)
    ic = torch.from_numpy(synthetic.synthetic_iris_crops([1, 2], 96))
    run("gram_iris96", ic[:1], ic[1:2], BN_loss=False, s_loss_weight=1e6, epochs=40)
    # long run that fills more of the history (history_size=100 ring): 128 evals at 32x32
    c3, s3 = rand_img(31, (1, 3, 32, 32)), rand_img(32, (1, 3, 32, 32))
    run("gram_long", c3, s3, BN_loss=False, s_loss_weight=1e6, epochs=130)
    np.savez_compressed(os.path.join(OUT, "nst_traj.npz"), **t)

    # ---- (3) mask / bbox / crop ---------------------------------------------------------------
    m = {}

    class FakeRITnet(torch.nn.Module):
        def __init__(self, seg):
            super().__init__()
            self.seg = seg

        def forward(self, x):
            return self.seg

    for k, (seed, h, w) in enumerate([(5, 640, 400), (6, 400, 640), (7, 64, 48)]):
        frame, seg = synthetic.synthetic_eye(seed, h, w)
        xt, st = torch.from_numpy(frame), torch.from_numpy(seg)
        xc, mc, x0, y0, x1, y1 = pipelines.mask_and_crop_iris(xt, ritnet=FakeRITnet(st), device="cpu")
        m["syn%d_bbox" % k] = np.array([int(x0), int(y0), int(x1), int(y1)])
        m["syn%d_crop_sum" % k] = np.float64(xc.double().sum())
        m["syn%d_mask_count" % k] = np.int64(mc.sum())
        m["syn%d_crop_shape" % k] = np.array(xc.shape)
        bb = utils.crop_image(xt * ((st == 2) * (xt <= 0.8)), return_idx=True)
        assert [int(v) for v in bb] == [int(x0), int(y0), int(x1), int(y1)]
    # real eye PNGs through the shipped RITnet (known answer: notebook cell 2 prints [171, 206])
    try:
        from PIL import Image
        import torchvision.transforms.v2 as T

        cwd = os.getcwd()
        os.chdir(ref_loader.REF)
        rit = models.RITnet()
        os.chdir(cwd)
        tt = T.Compose([T.ToImage(), T.ToDtype(torch.float32, scale=True)])
        for k, name in enumerate(["000000339816.png", "000000240703.png"]):
            img = tt(Image.open(os.path.join(ref_loader.REF, "images", name)))
            with torch.no_grad():
                seg = rit(img)
            xc, mc, x0, y0, x1, y1 = pipelines.mask_and_crop_iris(img, ritnet=rit, device="cpu")
            m["real%d_bbox" % k] = np.array([int(x0), int(y0), int(x1), int(y1)])
            m["real%d_iris_bits" % k] = np.packbits((seg == 2).numpy())
            m["real%d_noglint_bits" % k] = np.packbits((img <= 0.8).numpy())
            m["real%d_shape" % k] = np.array(img.shape)
            m["real%d_mask_count" % k] = np.int64(mc.sum())
            print("real", name, [int(v) for v in (x0, y0, x1, y1)], tuple(xc.shape), int(mc.sum()))
    except Exception as e:  # pragma: no cover
        print("real-image fixtures skipped:", e)
    np.savez_compressed(os.path.join(OUT, "mask_bbox.npz"), **m)

    # ---- (4) composite (…2019.py:111-130) -------------------------------------------------------
    import torchvision.transforms.v2 as transforms

    k = {}
    for idx, (seed, h, w) in enumerate([(5, 640, 400), (6, 400, 640)]):
        frame, seg = synthetic.synthetic_eye(seed, h, w)
        c_img = torch.from_numpy(frame)
        c_m = (torch.from_numpy(seg) == 2) * (c_img <= 0.8)
        x_min, y_min, x_max, y_max = [int(v) for v in utils.crop_image(c_img * c_m, return_idx=True)]
        raw_shape = (x_max - x_min + 1, y_max - y_min + 1)
        new_rgb = rand_img(40 + idx, (1, 3, 224, 224))
        gray = transforms.functional.rgb_to_grayscale(new_rgb)[0]
        new = transforms.Resize(raw_shape)(gray)
        cm = c_m[:, x_min:x_max + 1, y_min:y_max + 1]
        new = new * cm
        out = c_img.clone()
        out[:, x_min:x_max + 1, y_min:y_max + 1] *= ~cm
        out[:, x_min:x_max + 1, y_min:y_max + 1] += new
        k["comp%d_out" % idx] = out.numpy().astype(np.float16)  # compact; fp32 checksum below
        k["comp%d_out_sum" % idx] = np.float64(out.double().sum())
        k["comp%d_patch" % idx] = out[:, x_min:x_min + 24, y_min + 40:y_min + 64].numpy()
        k["comp%d_bbox" % idx] = np.array([x_min, y_min, x_max, y_max])
        # forward resize used by the drivers (crop -> 224x224), …2019.py:49,75
        crop = (c_img * c_m)[:, x_min:x_max + 1, y_min:y_max + 1]
        k["resize%d_224_sum" % idx] = np.float64(transforms.Resize((224, 224))(crop).double().sum())
        k["resize%d_224_patch" % idx] = transforms.Resize((224, 224))(crop)[:, 100:116, 100:116].numpy()
    np.savez_compressed(os.path.join(OUT, "composite.npz"), **k)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
