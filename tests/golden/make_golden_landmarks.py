"""Golden vectors of the downstream evaluator's feature path, from the UNMODIFIED reference
(models/gaze_estimators/gaze_estimators.py) imported live from /root/reference:
    python tests/golden/make_golden_landmarks.py      -> tests/golden/landmarks.npz
Stored: the 19 landmarks `extract_eye_landmarks` returns for the label maps of `landmark_cases()` (regenerated from seeds
by the tests, not stored), and the outputs of GazeEstimator1 / GazeEstimator2 (eval mode, extract_feature=False) for
weights and inputs drawn from seeded numpy streams (`head_case`).  No reference source is copied."""
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


landmark_cases, head_case = synthetic.landmark_cases, synthetic.gaze_head_case


def main():
    pipelines, utils, models = ref_loader.load(seed=0)
    g = {}
    for name, lab in landmark_cases():
        lm = models.extract_eye_landmarks(torch.from_numpy(lab))
        g["lm_" + name] = lm.numpy()
        print(name, np.round(lm.numpy(), 3).tolist())
    for cls, in_dim in ((models.GazeEstimator1, 19), (models.GazeEstimator2, 2048)):
        params, x = head_case(in_dim)
        m = cls(extract_feature=False).eval()
        sd = m.state_dict()
        for k, p in zip(["model.0.weight", "model.0.bias", "model.3.weight", "model.3.bias", "model.6.weight", "model.6.bias"], params):
            assert sd[k].shape == p.shape
            sd[k] = torch.from_numpy(p)
        m.load_state_dict(sd)
        with torch.no_grad():
            g["head%d_out" % in_dim] = m(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(OUT, "landmarks.npz"), **g)
    print("wrote landmarks.npz:", sorted(g))


if __name__ == "__main__":
    main()
