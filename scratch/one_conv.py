import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load(); sp=L.stream_ptr
B=8; h,w,cin,cout=640,400,64,64
x=torch.randn(B,h,w,cin,device='cuda').clamp_min(0).bfloat16(); wf=(torch.randn(9,cout,cin,device='cuda')*0.03).bfloat16(); bias=torch.zeros(cout,device='cuda'); out=torch.empty(B,h,w,cout,device='cuda',dtype=torch.bfloat16)
for _ in range(4): L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,cin,cout,1,0,sp())
torch.cuda.synchronize(); print('ok')
