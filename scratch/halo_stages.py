import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
lib=L.load(); sp=L.stream_ptr
B=32; H0,W0=640,400; dev='cuda'
lv=[(H0,W0),(H0//2,W0//2),(H0//4,W0//4),(H0//8,W0//8)]
def timeit(fn,n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for name,l,cin,cout in [("conv2_1",1,64,128),("conv2_2",1,128,128),("conv3_2",2,256,256)]:
    h,w=lv[l]
    xin=torch.randn(B,h,w,cin,device=dev).clamp_min(0).bfloat16(); wf=(torch.randn(9,cout,cin,device=dev)*0.03).bfloat16()
    bias=torch.zeros(cout,device=dev); out=torch.empty(B,h,w,cout,device=dev,dtype=torch.bfloat16)
    dy=torch.randn(B,h,w,cout,device=dev).bfloat16(); wd=(torch.randn(9,cin,cout,device=dev)*0.03).bfloat16(); dxo=torch.empty(B,h,w,cin,device=dev,dtype=torch.bfloat16)
    res=[]
    for st in (2,3,4,5,6):
        lib.isx_set_option(b"halo2_stages",st)
        ms=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",xin,wf,bias,out,B,h,w,cin,cout,1,0,sp()))
        ms2=timeit(lambda: L.call("isx_conv3x3_dgrad",dy,wd,dxo,B,h,w,cin,cout,xin,None,None,None,0,sp()))
        res.append("st%d fwd %.1f dgrad+mask %.1f"%(st, ms*1e3/B, ms2*1e3/B))
    lib.isx_set_option(b"halo2_stages",0)
    print(name, " | ".join(res), flush=True)
