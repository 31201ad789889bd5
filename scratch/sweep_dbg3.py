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
dy = torch.randn(B, H, W, 64, device=dev).bfloat16(); dxo = torch.empty_like(out)
D = (torch.randn(B, 64, 64, device=dev) * 0.01).bfloat16()
lib.isx_set_option(b"sweep64", 2)
lib.isx_set_option(b"sweep_dbg", 8)
for name, fn in [("fwd", lambda: L.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, 64, 64, 1, 0, sp())),
                 ("dgrad+mask", lambda: L.call("isx_conv3x3_dgrad", dy, wd, dxo, B, H, W, 64, 64, x, None, None, None, 0, sp())),
                 ("dgrad+mask+gram", lambda: L.call("isx_conv3x3_dgrad_gram", dy, wd, dxo, B, H, W, 64, 64, x, D, sp()))]:
    print("----", name, flush=True)
    fn(); torch.cuda.synchronize()
    fn(); torch.cuda.synchronize()
