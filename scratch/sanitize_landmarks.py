"""compute-sanitizer target: the landmark kernels and the gaze head on small and reference-sized label maps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import iris_b200  # noqa: E402
from test_landmarks_oracle import random_mask  # noqa: E402

rng = np.random.default_rng(0)
for (H, W) in [(24, 31), (57, 97), (33, 64), (400, 640)]:
    B = 6 if H < 100 else 3
    labs = np.zeros((B, H, W), np.int64)
    for b in range(B):
        m3 = random_mask(rng, b % 4, H, W)
        m2 = random_mask(rng, (b + 1) % 4, H, W) & (1 - m3)
        labs[b][m2 > 0] = 2
        labs[b][m3 > 0] = 3
        labs[b][(rng.random((H, W)) < 0.05) & (labs[b] == 0)] = 1
    if H == 400:
        labs[0] = iris_b200.synthetic.synthetic_label_map(3, speck=0.01)
    for dt in (torch.int64, torch.uint8):
        out, info = iris_b200.extract_eye_landmarks_batch(torch.from_numpy(labs).to(dt).cuda(), return_info=True)
    torch.cuda.synchronize()
    print(H, W, "ok", float(out.abs().sum()), info[:, [1, 4]].max().item())
net = iris_b200.GazeEstimator1().to("cuda:0")
print(net(torch.randn(37, 19, device="cuda")).shape)
net2 = iris_b200.GazeEstimator2().to("cuda:0")
print(net2(torch.randn(5, 2048, device="cuda")).shape)
torch.cuda.synchronize()
