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
pool = torch.empty(B, H // 2, W // 2, 64, device=dev, dtype=torch.bfloat16); idx = torch.empty(B, H // 2, W // 2, 64, device=dev, dtype=torch.uint8)
lib.isx_set_option(b"sweep64", 2)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f_plain = lambda: L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp())
f_pool = lambda: L.call("isx_conv3x3_bias_relu_pool_idx_fwd", x, wf, bias, out, pool, idx, 1, B, H, W, 64, 64, 0, sp())
f_pool_store = lambda: L.call("isx_conv3x3_bias_relu_pool_idx_fwd", x, wf, bias, out, pool, idx, 0, B, H, W, 64, 64, 0, sp())
for rep in range(3):
    print("plain %.2f  pool+idx,skip_out %.2f  pool+idx+full store %.2f us/img" % tuple(timeit(f) * 1e3 / B for f in (f_plain, f_pool, f_pool_store)), flush=True)
lib.isx_set_option(b"sweep_dbg", 8)
print("---- pool+idx,skip_out", flush=True)
f_pool(); torch.cuda.synchronize()
