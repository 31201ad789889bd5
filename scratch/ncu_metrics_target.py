"""ncu target: the final landmark kernels (128 label maps, 0.2 % stray pixels) and cal_IoUs (128 pairs of 640x400 maps)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iris_b200  # noqa: E402

labs = np.stack([iris_b200.synthetic.synthetic_label_map(100 + i, speck=0.002) for i in range(128)])
seg = torch.from_numpy(labs).cuda()
p = torch.from_numpy(np.stack([iris_b200.synthetic.synthetic_label_map(i % 8, 640, 400, speck=0.01) for i in range(128)])).cuda()
t = torch.from_numpy(np.stack([iris_b200.synthetic.synthetic_label_map(i % 8, 640, 400) for i in range(128)])).cuda()
for _ in range(3):
    out = iris_b200.extract_eye_landmarks_batch(seg)
    iou = iris_b200.cal_IoUs(p, t)
torch.cuda.synchronize()
print("ok", float(out.sum()), float(iou[1].sum()))
