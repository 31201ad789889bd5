"""Where the time of extract_eye_landmarks_batch goes: 128 frames of (a) empty maps, (b) clean eyes, (c) eyes with 0.2 % /
0.6 % / 2 % of the pixels relabelled (hundreds / thousands of stray one-pixel contours per class)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iris_b200  # noqa: E402

B = 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(seg, reps=10):
    for _ in range(2):
        iris_b200.extract_eye_landmarks_batch(seg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0.record()
        out, info = iris_b200.extract_eye_landmarks_batch(seg, return_info=True)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), info


for name, speck in (("empty", None), ("clean", 0.0), ("speck 0.002", 0.002), ("speck 0.006", 0.006), ("speck 0.02", 0.02)):
    if speck is None:
        labs = np.zeros((B, 400, 640), np.int64)
    else:
        labs = np.stack([iris_b200.synthetic.synthetic_label_map(100 + i, speck=speck) for i in range(B)])
    ms, info = t(torch.from_numpy(labs).cuda())
    info = info.cpu().numpy()
    print("%-12s %.3f ms per %d frames; contours per class: mean %.0f max %d; points of the chosen contour: max %d" % (
        name, ms, B, info[:, [1, 4]].mean(), info[:, [1, 4]].max(), info[:, [0, 3]].max()))
