"""ncu target: extract_eye_landmarks_batch on 128 label maps (speck 0.002: ~120 stray contours per class), three calls."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iris_b200  # noqa: E402

speck = float(sys.argv[1]) if len(sys.argv) > 1 else 0.002
labs = np.stack([iris_b200.synthetic.synthetic_label_map(100 + i, speck=speck) for i in range(128)])
seg = torch.from_numpy(labs).cuda()
for _ in range(3):
    out = iris_b200.extract_eye_landmarks_batch(seg)
torch.cuda.synchronize()
print("ok", float(out.sum()))
