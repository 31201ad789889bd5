"""Short launch sequence for `ncu --set full` (round 2): the symmetric Gram kernel at the bench shapes (batch 32 of 640x400
eyes): plain C = 64 / 128 / 256 / 512, C = 64 / 128 with fused channel sums, masked C = 64 / 512 with an iris-sized mask."""
import sys, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import _lib as L, engine as E
L.load(); sp = L.stream_ptr
B = 32; dev = 'cuda'
shapes = [(640, 400, 64), (320, 200, 128), (160, 100, 256), (80, 50, 512)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for (h, w, c) in shapes:
    f = torch.randn(B, h, w, c, device=dev).clamp_min(0).bfloat16().contiguous()
    m = torch.zeros(B, h, w, device=dev)
    m[:, h // 3: h // 3 + h // 4, w // 4: w // 4 + w // 3] = 1.0      # ~8 % of the frame
    for _ in range(reps):
        G = E.gram_of(f)
        Gm = E.masked_gram_of(f, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ws = torch.empty(max(256, L.call_i64("isx_gram_workspace_bytes", B, h * w, c)), device=dev, dtype=torch.uint8)
    Gout = torch.empty(B, c, c, device=dev)
    e0.record()
    for _ in range(5):
        L.call("isx_gram_fwd", f, B, h * w, c, L.f32(1.0 / (c * h * w)), ws, Gout, None, 1, L.f64(0.0), None, L.f32(0.0), None, sp())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("C=%3d  %dx%d  gram fwd+finalize %.3f ms per %d images = %.1f us/image, %.0f algorithmic TFLOP/s, %.2f TB/s of features" % (
        c, h, w, ms, B, 1e3 * ms / B, 2.0 * c * c * h * w * B / ms / 1e9, 2.0 * c * h * w * B / ms / 1e9))
print("ok")
