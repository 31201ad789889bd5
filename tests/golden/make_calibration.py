"""Calibration of the trajectory bounds: how far does the REFERENCE ALGORITHM (oracle/nst_oracle.py, pinned to the
reference by test_oracle_golden.py) move its final image under perturbations of the size the north star's
precision implies?  The reference optimiser -- L-BFGS with lr = 1, no line search, clamp inside the closure
(pipelines.py:59,81-82) -- amplifies rounding-level differences, so "final image within 1e-2 MAE" cannot hold for
ANY bf16-operand implementation on inputs where the fp32 algorithm itself is this sensitive.  For every trajectory
case the GPU tests compare against, this script records

    sens_bf16    MAE(final image of the oracle run with bf16-rounded conv operands, fp32 run)
    sens_noise   MAE(oracle run whose gradient is multiplied by 1 + 2^-9 N(0,1), fp32 run), 3 seeds
    moved        MAE(final image, start image) of the fp32 run

into tests/golden/calibration_r2.json.  tests/test_gpu_parity_r2.py asserts
    MAE(GPU, reference) <= max(1e-2, 1.5 * max(sens_bf16, sens_noise...)).
Run in the build container (several minutes):  python tests/golden/make_calibration.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "iris-style-transfer_b200"))
from oracle import nst_oracle as O  # noqa: E402
import synthetic  # noqa: E402
import bench  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(os.cpu_count() or 8)


def rand_img(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)


def main():
    W = O.random_vgg19_weights(0)
    c1, s1 = rand_img(21, (1, 3, 48, 64)), rand_img(22, (1, 3, 48, 64))
    ic = torch.from_numpy(synthetic.synthetic_iris_crops([1, 2], 96))
    i224 = torch.from_numpy(synthetic.synthetic_iris_crops([1, 2, 3, 4, 11, 12, 13, 14], 224))
    c3, s3 = rand_img(31, (1, 3, 32, 32)), rand_img(32, (1, 3, 32, 32))
    cb, sb = bench.make_inputs(6, 1)
    cases = {
        "gram_b1": (c1, s1, dict(BN_loss=False, s_loss_weight=1e6, epochs=50)),
        "bn_b1": (c1, s1, dict(BN_loss=True, s_loss_weight=1e4, epochs=40)),
        "gram_iris96": (ic[:1], ic[1:2], dict(BN_loss=False, s_loss_weight=1e6, epochs=40)),
        "gram_long": (c3, s3, dict(BN_loss=False, s_loss_weight=1e6, epochs=130)),
        "bn_iris224_b4": (i224[:4], i224[4:], dict(BN_loss=True, s_loss_weight=1e4, epochs=40)),
        "bench640_img0_e20": (cb[0:1], sb[0:1], dict(BN_loss=False, s_loss_weight=1e6, epochs=20)),
        "bench640_img0_e50": (cb[0:1], sb[0:1], dict(BN_loss=False, s_loss_weight=1e6, epochs=50)),
        "bench640_img5_e50": (cb[5:6], sb[5:6], dict(BN_loss=False, s_loss_weight=1e6, epochs=50)),
    }
    out = {}
    path = os.path.join(OUT, "calibration_r2.json")
    for tag, (c, s, kw) in cases.items():
        x0, _, _, s0 = O.nst(c, s, W, keep_hist=False, **kw)
        xb, _, _, sb_ = O.nst(c, s, W, keep_hist=False, operand_dtype=torch.bfloat16, **kw)
        noise = []
        for seed in (1, 2, 3):
            xn, _, _, _ = O.nst(c, s, W, keep_hist=False, grad_noise=2.0 ** -9, noise_seed=seed, **kw)
            noise.append(float((xn - x0).abs().mean()))
        out[tag] = {"moved": float((x0 - c).abs().mean()), "sens_bf16": float((xb - x0).abs().mean()),
                    "sens_noise": noise, "evals": len(s0), "s_loss_first": s0[0], "s_loss_last": s0[-1],
                    "s_loss_last_bf16": sb_[-1]}
        print(tag, json.dumps(out[tag]), flush=True)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
