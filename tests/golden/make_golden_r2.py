"""Round-2 additions to the golden vectors, again produced by the UNMODIFIED reference imported live from
/root/reference (oracle/ref_loader.py).  Run once in the build container:
    python tests/golden/make_golden_r2.py        -> tests/golden/nst_traj_r2.npz
Cases:
  *_c3d_*      pipelines.nst called with an UNBATCHED content image (3,H,W) like the notebook's
               `nst(c_masked_cropped, s_masked_cropped, ...)`: returns (3,H,W); with BN_loss=False utils.GramMatrix
               divides the prediction's Gram by H*W (utils.py:253-254) while a batched style keeps C*H*W.
  bench640_*   BASELINE configs[0]/[1]: the iris-MASKED 640x400 frames bench.py feeds (bench.make_inputs, seeds 1.. /
               100001..), one image per call (B = 1 is the per-image semantics the product shards on), Gram loss,
               beta = 1e6; final images stored as uint16 fixed point (x * 65535, error 7.6e-6).
  bn_iris224_b4  what iris_style_transfer_openeds2019.py:93-100 calls: a batch of 224x224 iris crops as ONE L-BFGS
               problem with the default StyleLoss_BN (batch 4 instead of 64 to keep the CPU run short).
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


def main():
    vgg = models.VGG19()
    t = {}
    last_x = [None]

    def run(tag, c_img, s_img, store_x=True, **kw):
        x, x_hist, c_hist, s_hist = pipelines.nst(c_img, s_img, vgg=vgg, use_tqdm=False, device="cpu", **kw)
        last_x[0] = x
        if store_x:
            t[tag + "_x"] = x.numpy()
        t[tag + "_x_shape"] = np.array(x.shape)
        t[tag + "_c_hist"] = np.array(c_hist)
        t[tag + "_s_hist"] = np.array(s_hist)
        print(tag, tuple(x.shape), len(c_hist), "evals; moved MAE", float((x - c_img).abs().mean()), "s_loss", s_hist[0],
              "->", s_hist[-1])

    H, W = 48, 64
    c1, s1 = rand_img(21, (1, 3, H, W)), rand_img(22, (1, 3, H, W))
    run("gram_c3d_s3d", c1[0], s1[0], BN_loss=False, s_loss_weight=1e4, epochs=20)   # both unbatched: n = H*W on both sides
    run("bn_c3d_s3d", c1[0], s1[0], BN_loss=True, s_loss_weight=1e4, epochs=20)
    run("gram_c3d_s4d", c1[0], s1, BN_loss=False, s_loss_weight=1e4, epochs=20)      # prediction /HW, target /CHW (the quirk)
    ic = torch.from_numpy(synthetic.synthetic_iris_crops([1, 2, 3, 4, 11, 12, 13, 14], 224))
    run("bn_iris224_b4", ic[:4], ic[4:], BN_loss=True, s_loss_weight=1e4, epochs=40)
    import bench

    torch.set_num_threads(os.cpu_count() or 8)
    cb, sb = bench.make_inputs(6, 1)
    for tag, i, ep in (("bench640_img0_e20", 0, 20), ("bench640_img0_e50", 0, 50), ("bench640_img5_e50", 5, 50)):
        run(tag, cb[i:i + 1], sb[i:i + 1], store_x=False, BN_loss=False, s_loss_weight=1e6, epochs=ep)
        t[tag + "_moved"] = np.float64((last_x[0] - cb[i:i + 1]).abs().mean())
        t[tag + "_x_u16"] = np.round(last_x[0].numpy().astype(np.float64) * 65535.0).astype(np.uint16)
    np.savez_compressed(os.path.join(OUT, "nst_traj_r2.npz"), **t)
    print(os.path.getsize(os.path.join(OUT, "nst_traj_r2.npz")), "bytes")


if __name__ == "__main__":
    main()
