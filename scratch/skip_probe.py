"""Is the generic tcgen05 conv bound by operand delivery (L2 -> shared memory)?  Time layers with the A and/or B TMA
loads skipped (results are garbage; only the time matters)."""
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
lib.isx_set_option(b"c64",0)
for name,l,cin,cout in [("conv1_2",0,64,64),("conv2_1",1,64,128),("conv2_2",1,128,128),("conv3_1",2,128,256),("conv3_2",2,256,256),("conv4_2",3,512,512)]:
    h,w=lv[l]
    xin=torch.randn(B,h,w,cin,device=dev).clamp_min(0).bfloat16(); wf=(torch.randn(9,cout,cin,device=dev)*0.03).bfloat16()
    bias=torch.zeros(cout,device=dev); out=torch.empty(B,h,w,cout,device=dev,dtype=torch.bfloat16)
    fl=2*9*cin*cout*B*h*w
    res=[]
    for skip in (0,1,2,3):
        lib.isx_set_option(b"conv_dbg_skip",skip)
        ms=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",xin,wf,bias,out,B,h,w,cin,cout,1,0,sp()))
        res.append("%s %.1f us/img %.0f TF"%(["full","noA","noB","noA+noB"][skip], ms*1e3/B, fl/ms/1e9))
    lib.isx_set_option(b"conv_dbg_skip",0)
    print(name, " | ".join(res), flush=True)
