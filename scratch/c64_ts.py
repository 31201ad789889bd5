import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
lib=L.load(); sp=L.stream_ptr
B=4; h,w=640,400
x=torch.randn(B,h,w,64,device='cuda').clamp_min(0).bfloat16(); wf=(torch.randn(9,64,64,device='cuda')*0.03).bfloat16(); bias=torch.zeros(64,device='cuda'); out=torch.empty(B,h,w,64,device='cuda',dtype=torch.bfloat16)
w0d=(torch.randn(9,16,64,device='cuda')*0.03).bfloat16(); dx=torch.empty(B,3,h,w,device='cuda')
lib.isx_set_option(b"c64",2)
for _ in range(2):
    L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,64,64,1,0,sp())
    L.call("isx_conv1_1_dgrad_tc",x,w0d,None,0,dx,3,B,h,w,sp())
torch.cuda.synchronize()
for dbgv in (32+64, 32+128):
  lib.isx_set_option(b"c64_dbg",dbgv)
  print("== dbg", dbgv, file=sys.stderr, flush=True)
  L.call("isx_conv1_1_dgrad_tc",x,w0d,None,0,dx,3,B,h,w,sp())
lib.isx_set_option(b"c64_dbg",32)
print("== fwd", file=sys.stderr, flush=True)
L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,64,64,1,0,sp())
print("== tail", file=sys.stderr, flush=True)
L.call("isx_conv1_1_dgrad_tc",x,w0d,None,0,dx,3,B,h,w,sp())
