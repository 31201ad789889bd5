import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
lib=L.load(); sp=L.stream_ptr
B=32; h,w=640,400
def timeit(fn,n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
x=torch.randn(B,h,w,64,device='cuda').clamp_min(0).bfloat16(); wf=(torch.randn(9,64,64,device='cuda')*0.03).bfloat16(); bias=torch.zeros(64,device='cuda'); out=torch.empty(B,h,w,64,device='cuda',dtype=torch.bfloat16)
w0d=(torch.randn(9,16,64,device='cuda')*0.03).bfloat16(); dx=torch.empty(B,3,h,w,device='cuda')
lib.isx_set_option(b"c64",1)
for slots in (4,):
  for dbg in (0,):
    lib.isx_set_option(b"c64_slots",slots); lib.isx_set_option(b"c64_dbg",dbg)
    t1=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,64,64,1,0,sp()))
    t2=timeit(lambda: L.call("isx_conv1_1_dgrad_tc",x,w0d,None,0,dx,3,B,h,w,sp()))
    print("slots %d dbg %d: conv1_2 fwd %.1f us/img | tail %.1f us/img"%(slots,dbg,t1*1e3/B,t2*1e3/B),flush=True)
