"""conv1_2 (64 -> 64, 640x400) forward / dgrad: conv_c64 (sweep64=0) vs the tap-stacked sweep kernel (sweep64=2)."""
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
def bf(*shape, relu=True):
    t = torch.randn(*shape, device=dev)
    if relu: t = t.clamp_min(0)
    return t.bfloat16().contiguous()
for (B, H, W) in [(32, 640, 400), (64, 640, 400), (64, 224, 224), (16, 1024, 1024), (64, 400, 640)]:
    x = bf(B, H, W, 64); wt = torch.randn(64, 64, 3, 3, device=dev) * 0.03
    wf = torch.empty(9, 64, 64, device=dev, dtype=torch.bfloat16); wd = torch.empty(9, 64, 64, device=dev, dtype=torch.bfloat16)
    L.call("isx_pack_conv3x3_weights", wt, 64, 64, wf, wd, sp())
    bias = torch.zeros(64, device=dev); out = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
    pool = torch.empty(B, H // 2, W // 2, 64, device=dev, dtype=torch.bfloat16); idx = torch.empty(B, H // 2, W // 2, 64, device=dev, dtype=torch.uint8)
    dy = bf(B, H, W, 64, relu=False); dxo = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
    D = (torch.randn(B, 64, 64, device=dev) * 0.01).bfloat16()
    fl = 2 * 9 * 64 * 64 * B * H * W
    for opt in (0, 2):
        lib.isx_set_option(b"sweep64", opt)
        t1 = timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp()))
        t2 = timeit(lambda: L.call("isx_conv3x3_bias_relu_pool_idx_fwd", x, wf, bias, out, pool, idx, 1, B, H, W, 64, 64, 0, sp()))
        t3 = timeit(lambda: L.call("isx_conv3x3_dgrad", dy, wd, dxo, B, H, W, 64, 64, x, None, None, None, 0, sp()))
        t4 = timeit(lambda: L.call("isx_conv3x3_dgrad_gram", dy, wd, dxo, B, H, W, 64, 64, x, D, sp()))
        print("B %d %dx%d sweep64=%d: fwd %.2f us/img (%.0f TF/s)  fwd+pool+idx,skip_out %.2f (%.0f)  dgrad+mask %.2f (%.0f)  dgrad+mask+gram %.2f (%.0f)" % (
            B, H, W, opt, t1 * 1e3 / B, fl / t1 / 1e9, t2 * 1e3 / B, fl / t2 / 1e9, t3 * 1e3 / B, fl / t3 / 1e9, t4 * 1e3 / B, (fl + 2 * 64 * 64 * B * H * W) / t4 / 1e9), flush=True)
    lib.isx_set_option(b"sweep64", 1)
    del x, out, pool, idx, dy, dxo
