"""Golden vectors of the drivers' evaluation metrics, from the UNMODIFIED reference (utils.cal_IoUs, utils.angular_distance)
imported live from /root/reference:   python tests/golden/make_golden_metrics.py   -> tests/golden/metrics.npz
Inputs are regenerated from seeds by the tests (synthetic.iou_case / gaze_vector_case); only the outputs are stored."""
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


def main():
    pipelines, utils, models = ref_loader.load(seed=0)
    g = {}
    for name, shape in (("2019", (640, 400)), ("small", (37, 53))):
        p, t = synthetic.iou_case(shape)
        per_class, miou = utils.cal_IoUs(torch.from_numpy(p), torch.from_numpy(t))
        g["iou_" + name] = torch.stack(per_class, dim=1).numpy()
        g["miou_" + name] = miou.numpy()
        print(name, g["iou_" + name].round(4).tolist(), g["miou_" + name].round(4).tolist())
    a, b = synthetic.gaze_vector_case()
    rad, deg = utils.angular_distance(torch.from_numpy(a), torch.from_numpy(b))
    g["rad"], g["deg"] = rad.numpy(), deg.numpy()
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **g)
    print("wrote metrics.npz", sorted(g))


if __name__ == "__main__":
    main()
