"""Row f4 timing on the GPU box: iris_b200.extract_eye_landmarks_batch on B 400x640 label maps (CUDA events, after warm-up)
next to the reference's own per-frame CPU path (cv2.findContours / contourArea / fitEllipse + np.where, what
gaze_estimators.py:108-178 executes) on the host cores.  python scratch/landmarks_timing.py [B] -> one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iris_b200  # noqa: E402
from oracle import landmarks_oracle as L  # noqa: E402  (checker only: compares the timed results)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
labs = np.stack([iris_b200.synthetic.synthetic_label_map(100 + i, speck=0.002 * (i % 4)) for i in range(B)])
seg = torch.from_numpy(labs).cuda()
for _ in range(3):
    out, info = iris_b200.extract_eye_landmarks_batch(seg, return_info=True)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for _ in range(reps):
    flush.zero_()                       # 256 MB > L2: the label maps come from HBM
    ev[0].record()
    out, info = iris_b200.extract_eye_landmarks_batch(seg, return_info=True)
    ev[1].record()
    torch.cuda.synchronize()
    ms.append(ev[0].elapsed_time(ev[1]))
ms = float(np.median(ms))
import cv2  # noqa: E402


def ref_frame(seg2d):
    s8 = seg2d.astype(np.uint8)
    res = []
    for cls in (3, 2):
        cs, _ = cv2.findContours((s8 == cls).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if cs:
            c = max(cs, key=cv2.contourArea)
            if len(c) >= 5:
                res.append(cv2.fitEllipse(c))
    ys, xs = np.where((s8 == 1) > 0)
    return res, (xs.min(), xs.max(), ys.min(), ys.max()) if len(xs) else None


t0 = time.perf_counter()
n_cpu = min(B, 64)
for i in range(n_cpu):
    ref_frame(labs[i])
cpu_ms = (time.perf_counter() - t0) * 1e3 / n_cpu
want = np.stack([L.extract_eye_landmarks(labs[i]) for i in range(min(B, 8))])
err = float(np.abs(out[:len(want)].cpu().numpy() - want).max())
print(json.dumps({"what": "extract_eye_landmarks, %d label maps 400x640 int64" % B, "gpu_ms_per_batch": ms,
                  "gpu_frames_per_s": B / ms * 1e3, "label_bytes_GBps": B * 400 * 640 * 8 / ms / 1e6,
                  "cpu_cv2_ms_per_frame": cpu_ms, "cpu_frames_per_s": 1e3 / cpu_ms,
                  "max_abs_diff_vs_oracle_first8": err, "flags": info[:, [2, 5]].max().item()}))
