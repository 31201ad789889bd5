import sys, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import _lib as L
lib = L.load()
dev = 'cuda'
sp = L.stream_ptr
B, H, W = 64, 640, 400
x = torch.randn(B, H, W, 64, device=dev).clamp_min(0).bfloat16()
wt = torch.randn(64, 64, 3, 3, device=dev) * 0.03
wf = torch.empty(9, 64, 64, device=dev, dtype=torch.bfloat16); wd = torch.empty(9, 64, 64, device=dev, dtype=torch.bfloat16)
L.call("isx_pack_conv3x3_weights", wt, 64, 64, wf, wd, sp())
bias = torch.zeros(64, device=dev); out = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
lib.isx_set_option(b"sweep64", 2)
for dbg in (8,):
    print("---- dbg", dbg, flush=True)
    lib.isx_set_option(b"sweep_dbg", dbg)
    for _ in range(2):
        L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp())
        torch.cuda.synchronize()
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for dbg in (0, 16):
    lib.isx_set_option(b"sweep_dbg", dbg)
    t = timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp()))
    print("dbg %d: %.2f us/img" % (dbg, t * 1e3 / B))
