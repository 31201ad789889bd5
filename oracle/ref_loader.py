"""Loader for the UNMODIFIED reference (AnonymWriter/Iris-Style-Transfer) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used only by tests/golden/make_golden.py inside the build
container to pin oracle/nst_oracle.py against outputs of the reference itself.  Nothing on
the GPU box may call this (there is no /root/reference there).

Shims: three absent packages the hot path never calls are stubbed
(`skimage` utils.py:5, `matplotlib.pyplot` utils.py:9, `segmentation_models_pytorch`
models/efficientnet/efficientnet.py:4), and `torchvision.models.vgg19` is patched so that
`VGG19()` (models/vgg/vgg.py:43) builds `vgg19(weights=None)` under `torch.manual_seed(seed)`
instead of downloading ImageNet weights (no network) -- BASELINE.json's "random-init VGG-19".
No reference file is copied or edited.
"""
import os
import sys
import types

REF = os.environ.get("ISX_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "pipelines.py"))


def load(seed: int = 0):
    """Return (pipelines, utils, models) modules of the live reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    sys.dont_write_bytecode = True
    for name in ("skimage", "matplotlib", "matplotlib.pyplot", "segmentation_models_pytorch"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import torch
    import torchvision.models as tvm

    if not getattr(tvm.vgg19, "_isx_patched", False):
        _orig = tvm.vgg19

        def vgg19_random(weights=None, **kw):
            torch.manual_seed(seed)
            return _orig(weights=None, **kw)

        vgg19_random._isx_patched = True
        vgg19_random._isx_orig = _orig
        tvm.vgg19 = vgg19_random
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import pipelines
    import utils
    import models

    return pipelines, utils, models
