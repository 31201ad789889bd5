"""iris_b200 -- B200-native (sm_100a) drop-in for the NST hot path of AnonymWriter/Iris-Style-Transfer.

The arithmetic lives in libisx.so (hand-written CUDA, C ABI in include/isx.h); this package is the
host-side mirror of the reference's Python surface (pipelines.py, utils.py, models/vgg/vgg.py)."""
from . import _lib  # noqa: F401
from . import synthetic  # noqa: F401
