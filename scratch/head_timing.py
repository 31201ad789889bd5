"""conv1_1 head: us/image with 5 vs 8 resident CTAs per SM (option "head_ctas"), outputs compared bit for bit."""
import sys, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import _lib as L
lib = L.load()
dev = 'cuda'
H0, W0 = 640, 400
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sp = L.stream_ptr
for B in (32, 64):
    x = torch.rand(B, 3, H0, W0, device=dev); w0 = torch.randn(64, 3, 3, 3, device=dev) * 0.1; b0 = torch.randn(64, device=dev) * 0.1
    w0f = torch.empty(64, 64, device=dev, dtype=torch.bfloat16); L.call("isx_pack_conv1_1_fwd", w0, w0f, sp())
    outs = {}
    for occ in (5, 8):
        lib.isx_set_option(b"head_ctas", occ)
        a = torch.empty(B, H0, W0, 64, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: L.call("isx_conv1_1_fwd_tc", x, 3, None, 0, w0f, b0, a, B, H0, W0, sp()))
        outs[occ] = a
        print("B %d head_ctas %d: %.2f us/image  %.2f TB/s" % (B, occ, ms * 1e3 / B, B * H0 * W0 * 140 / ms / 1e9), flush=True)
    print("  equal:", bool(torch.equal(outs[5], outs[8])))
