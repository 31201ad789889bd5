"""Golden vectors of the mask producer, from the UNMODIFIED reference RITnet (models/ritnet/ritnet.py) imported live from
/root/reference with its shipped weights (models/weights/ritnet_pretrained.pkl):
    python tests/golden/make_golden_ritnet.py      -> tests/golden/ritnet.npz
Stored: the 249 225 pretrained parameters (the state dict as float32 arrays -- DATA the tests need on the GPU box, where
/root/reference does not exist; no reference source is copied), label maps RITnet() returns for synthetic eye frames
(uint8, 0..3), checksums of the logits, and the RITnet_transform output for one frame.  The script also checks the oracle
(oracle/ritnet_oracle.py) against the reference on the two shipped eye PNGs, which are not stored."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, ritnet_oracle as R  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "iris-style-transfer_b200"))
import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
pipelines, utils, models = ref_loader.load(seed=0)


def main():
    cwd = os.getcwd()
    os.chdir(ref_loader.REF)
    rit = models.RITnet()
    os.chdir(cwd)
    sd = {k: v.clone() for k, v in rit.model.state_dict().items()}
    g = {"w:" + k: v.numpy() for k, v in sd.items() if v.dtype == torch.float32}
    for k, (seed, h, w) in enumerate([(5, 640, 400), (6, 400, 640), (7, 64, 48), (8, 160, 96)]):
        frame, _ = synthetic.synthetic_eye(seed, h, w)
        x = torch.from_numpy(frame)
        with torch.no_grad():
            lab = rit(x)                               # (1,h,w) int64
            xt = rit.t(x)
            logits = rit.model(xt)
        g["syn%d_labels" % k] = lab.numpy().astype(np.uint8)
        g["syn%d_logit_abs_sum" % k] = np.float64(logits.double().abs().sum())
        g["syn%d_transform_sum" % k] = np.float64(xt.double().sum())
        if k == 2:
            g["syn2_transform"] = xt.numpy()
            g["syn2_logits"] = logits.numpy()
        olab = R.ritnet_labels(sd, x)
        ol = R.densenet2d_logits(sd, R.ritnet_transform(x))
        print("syn%d" % k, (h, w), "classes", np.bincount(lab.numpy().reshape(-1), minlength=4).tolist(), "oracle label mismatches",
              int((olab != lab).sum()), "max logit diff %.2e" % float((ol - logits).abs().max()),
              "transform numpy-clahe == cv2:", bool(torch.equal(R.ritnet_transform(x, use_cv2=False), xt)))
    # the two shipped eye frames (not stored): oracle == reference, and == the iris bits of mask_bbox.npz
    from PIL import Image
    import torchvision.transforms.v2 as T

    tt = T.Compose([T.ToImage(), T.ToDtype(torch.float32, scale=True)])
    m = np.load(os.path.join(OUT, "mask_bbox.npz"))
    for k, name in enumerate(["000000339816.png", "000000240703.png"]):
        img = tt(Image.open(os.path.join(ref_loader.REF, "images", name)))
        with torch.no_grad():
            lab = rit(img)
        olab = R.ritnet_labels(sd, img)
        iris = np.unpackbits(m["real%d_iris_bits" % k])[:lab.numel()].reshape(tuple(lab.shape)).astype(bool)
        print("real", name, "oracle mismatches", int((olab != lab).sum()), "iris bits equal", bool(np.array_equal((lab == 2).numpy(), iris)))
    np.savez_compressed(os.path.join(OUT, "ritnet.npz"), **g)
    print(os.path.getsize(os.path.join(OUT, "ritnet.npz")), "bytes")


if __name__ == "__main__":
    main()
