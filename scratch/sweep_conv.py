import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load()
B=int(sys.argv[1]) if len(sys.argv)>1 else 8
layers=[("conv1_2",400,640,64,64),("conv2_1",200,320,64,128),("conv2_2",200,320,128,128),("conv3_1",100,160,128,256),("conv3_2",100,160,256,256),("conv4_1",50,80,256,512),("conv4_2",50,80,512,512)]
def timeit(fn,n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for name,H,W,Cin,Cout in layers:
    x=torch.randn(B,H,W,Cin,device='cuda').clamp_min(0).bfloat16()
    w=torch.randn(9,Cout,Cin,device='cuda').bfloat16()*0.05
    bias=torch.zeros(Cout,device='cuda')
    out=torch.empty(B,H,W,Cout,device='cuda',dtype=torch.bfloat16)
    flops=2*9*Cin*Cout*H*W*B
    res=[]
    for bn in (64,128,256):
        if Cout%bn: continue
        for mt in (1,2):
            for st in (2,3,4,6):
                cfg=bn*100+mt*10+st
                try:
                    ms=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",x,w,bias,out,B,H,W,Cin,Cout,1,cfg,L.stream_ptr()))
                    res.append((flops/ms/1e9,cfg))
                except Exception as e:
                    pass
    res.sort(reverse=True)
    print(name,"B=%d"%B," ".join("%d:%.0f"%(c,t) for t,c in res[:8]),flush=True)
