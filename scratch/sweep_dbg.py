import sys, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import _lib as L
lib = L.load()
dev = 'cuda'
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sp = L.stream_ptr
for (B, H, W) in [(64, 640, 400), (64, 400, 640), (2, 640, 400), (2, 400, 640), (8, 640, 400)]:
    x = torch.randn(B, H, W, 64, device=dev).clamp_min(0).bfloat16()
    wt = torch.randn(64, 64, 3, 3, device=dev) * 0.03
    wf = torch.empty(9, 64, 64, device=dev, dtype=torch.bfloat16); wd = torch.empty(9, 64, 64, device=dev, dtype=torch.bfloat16)
    L.call("isx_pack_conv3x3_weights", wt, 64, 64, wf, wd, sp())
    bias = torch.zeros(64, device=dev); out = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
    lib.isx_set_option(b"sweep64", 2)
    res = []
    for dbg in (0, 1, 2, 3):
        lib.isx_set_option(b"sweep_dbg", dbg)
        t1 = timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp()), n=10 if B > 8 else 50)
        res.append("dbg%d %.2f" % (dbg, t1 * 1e3 / B))
    lib.isx_set_option(b"sweep_dbg", 0)
    lib.isx_set_option(b"sweep64", 0)
    t0 = timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp()), n=10 if B > 8 else 50)
    print("B %d %dx%d us/img: c64 %.2f | sweep %s" % (B, H, W, t0 * 1e3 / B, "  ".join(res)), flush=True)
    del x, out
