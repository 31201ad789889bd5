"""Short launch sequence for `ncu --set full`: the new conv kernels at the bench shapes (batch 16)."""
import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load(); sp=L.stream_ptr
B=16; dev='cuda'
def bf(*s, relu=True):
    t=torch.randn(*s,device=dev)
    return (t.clamp_min(0) if relu else t).bfloat16().contiguous()
def layer(h,w,cin,cout,reps):
    xin=bf(B,h,w,cin); wt=torch.randn(cout,cin,3,3,device=dev)*0.03
    wf=torch.empty(9,cout,cin,device=dev,dtype=torch.bfloat16); wd=torch.empty(9,cin,cout,device=dev,dtype=torch.bfloat16)
    L.call("isx_pack_conv3x3_weights",wt,cout,cin,wf,wd,sp())
    bias=torch.zeros(cout,device=dev); out=torch.empty(B,h,w,cout,device=dev,dtype=torch.bfloat16)
    dy=bf(B,h,w,cout); dxo=torch.empty(B,h,w,cin,device=dev,dtype=torch.bfloat16)
    D=(torch.randn(B,cin,cin,device=dev)*0.01).bfloat16()
    for _ in range(reps):
        L.call("isx_conv3x3_bias_relu_fwd",xin,wf,bias,out,B,h,w,cin,cout,1,0,sp())
        L.call("isx_conv3x3_dgrad_gram",dy,wd,dxo,B,h,w,cin,cout,xin,D,sp())
    torch.cuda.synchronize()
reps=int(sys.argv[1]) if len(sys.argv)>1 else 2
layer(640,400,64,64,reps)      # conv_c64
layer(320,200,128,128,reps)    # conv_halo<128>
layer(160,100,256,256,reps)    # conv_halo<128>, 4 input blocks
w0=torch.randn(64,3,3,3,device=dev)*0.1
w0d=torch.empty(9,16,64,device=dev,dtype=torch.bfloat16); L.call("isx_pack_conv1_1_dgrad",w0,w0d,sp())
g11=bf(B,640,400,64,relu=False); dx=torch.empty(B,3,640,400,device=dev)
for _ in range(reps): L.call("isx_conv1_1_dgrad_tc",g11,w0d,None,0,dx,3,B,640,400,sp())
torch.cuda.synchronize()
print("ok")
