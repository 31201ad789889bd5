"""lm_planes_kernel variants (option "lm_planes" = requests in flight per warp + 100 x halves of a resident wave): CUDA-event
time of the kernel alone (isx_prof family 3) over 20 calls of 128 label maps 400x640 int64, 256 MB written between calls."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iris_b200  # noqa: E402
from iris_b200 import _lib  # noqa: E402

lib = _lib.load()
B = 128
labs = np.stack([iris_b200.synthetic.synthetic_label_map(100 + i, speck=0.002 * (i % 4)) for i in range(B)])
seg = torch.from_numpy(labs).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for knob in (0, 5, 20, 110, 310, 410, 120, 320, 420, 105):
    assert lib.isx_set_option(b"lm_planes", knob) == 0
    for _ in range(3):
        iris_b200.extract_eye_landmarks_batch(seg)
    torch.cuda.synchronize()
    lib.isx_prof_enable(1)
    for _ in range(20):
        flush.zero_()
        iris_b200.extract_eye_landmarks_batch(seg)
    torch.cuda.synchronize()
    prof = (ctypes.c_double * 12)()
    _lib.call("isx_prof_collect", prof, 12)
    lib.isx_prof_enable(0)
    n, ms, by = prof[9], prof[10], prof[11]
    print("lm_planes=%3d: %2d launches, %.1f us each, %.0f GB/s" % (knob, n, 1e3 * ms / n, by / (ms / 1e3) / 1e9))
lib.isx_set_option(b"lm_planes", 0)
