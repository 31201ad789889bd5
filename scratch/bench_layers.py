import sys, torch, time
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load()
B=int(sys.argv[1]) if len(sys.argv)>1 else 8
layers=[("conv1_2",400,640,64,64),("conv2_1",200,320,64,128),("conv2_2",200,320,128,128),("conv3_1",100,160,128,256),("conv3_2",100,160,256,256),("conv4_1",50,80,256,512),("conv4_2",50,80,512,512)]
cfgs={64:[0,6410,6420,6413,6423,6416],128:[0,12810,12820,12813,12823],256:[0,25610,25620,12810,12820],512:[0,25610,25620,12820]}
def timeit(fn,n=10):
    for _ in range(3): fn()
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
    for cfg in cfgs[Cout]:
        try:
            ms=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",x,w,bias,out,B,H,W,Cin,Cout,1,cfg,L.stream_ptr()))
            res.append("%d: %.3f ms %.0f TF"%(cfg,ms,flops/ms/1e9))
        except Exception as e:
            res.append("%d: ERR %s"%(cfg,str(e)[:60]))
    print(name,B,"|"," | ".join(res),flush=True)
# gram
for C,H,W in [(64,400,640),(128,200,320),(256,100,160),(512,50,80)]:
    f=torch.randn(B,H,W,C,device='cuda').clamp_min(0).bfloat16()
    ws=torch.empty(L.call_i64("isx_gram_workspace_bytes",B,H*W,C),device='cuda',dtype=torch.uint8)
    G=torch.empty(B,C,C,device='cuda')
    ms=timeit(lambda: L.call("isx_gram_fwd",f,B,H*W,C,L.f32(1.0),ws,G,None,1,L.f64(0),None,L.f32(0),None,L.stream_ptr()))
    print("gram C=%d: %.3f ms %.0f TF %.0f GB/s"%(C,ms,2*C*C*H*W*B/ms/1e9, B*H*W*C*2/ms/1e6),flush=True)
