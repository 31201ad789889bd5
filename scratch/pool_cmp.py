import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load(); sp=L.stream_ptr
B=32
def timeit(fn,n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for name,h,w,cin,cout in [("conv1_2",640,400,64,64),("conv2_2",320,200,128,128),("conv3_4",160,100,256,256)]:
    x=torch.randn(B,h,w,cin,device='cuda').clamp_min(0).bfloat16(); wf=(torch.randn(9,cout,cin,device='cuda')*0.03).bfloat16(); bias=torch.zeros(cout,device='cuda')
    out=torch.empty(B,h,w,cout,device='cuda',dtype=torch.bfloat16); pool=torch.empty(B,h//2,w//2,cout,device='cuda',dtype=torch.bfloat16)
    t0=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,cin,cout,1,0,sp()))
    t1=timeit(lambda: L.call("isx_conv3x3_bias_relu_pool_fwd",x,wf,bias,out,pool,B,h,w,cin,cout,0,sp()))
    def sep():
        L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,cin,cout,1,0,sp()); L.call("isx_maxpool2x2_fwd",out,pool,B,h,w,cout,sp())
    t2=timeit(sep)
    print("%s: conv %.1f us/img | conv+fused pool %.1f | conv + separate pool %.1f"%(name,t0*1e3/B,t1*1e3/B,t2*1e3/B),flush=True)
