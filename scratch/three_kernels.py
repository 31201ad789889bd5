import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load()
B=8; H0,W0=640,400; dev='cuda'; sp=L.stream_ptr
x=torch.rand(B,3,H0,W0,device=dev); w0=torch.randn(64,3,3,3,device=dev)*0.1; b0=torch.zeros(64,device=dev)
w0f=torch.empty(64,64,device=dev,dtype=torch.bfloat16); L.call("isx_pack_conv1_1_fwd",w0,w0f,sp())
w0d=torch.empty(9,16,64,device=dev,dtype=torch.bfloat16); L.call("isx_pack_conv1_1_dgrad",w0,w0d,sp())
a11=torch.empty(B,H0,W0,64,device=dev,dtype=torch.bfloat16)
g11=torch.randn(B,H0,W0,64,device=dev).bfloat16(); dx=torch.empty_like(x)
f=torch.randn(B,H0,W0,64,device=dev).clamp_min(0).bfloat16()
ws=torch.empty(L.call_i64("isx_gram_workspace_bytes",B,H0*W0,64),device=dev,dtype=torch.uint8)
G=torch.empty(B,64,64,device=dev)
for _ in range(3):
    L.call("isx_conv1_1_fwd_tc",x,3,None,0,w0f,b0,a11,B,H0,W0,sp())
    L.call("isx_conv1_1_dgrad_tc",g11,w0d,None,0,dx,3,B,H0,W0,sp())
    L.call("isx_gram_fwd",f,B,H0*W0,64,L.f32(1.0),ws,G,None,1,L.f64(0),None,L.f32(0),None,sp())
torch.cuda.synchronize(); print("ok")
