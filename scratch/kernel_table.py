"""Per-kernel CUDA-event timings of one NST evaluation at 640x400, batch B (isolated launches, warm)."""
import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load()
B=int(sys.argv[1]) if len(sys.argv)>1 else 32
H0,W0=640,400
dev='cuda'
def bf(*shape, relu=True):
    t=torch.randn(*shape,device=dev)
    if relu: t=t.clamp_min(0)
    return t.bfloat16().contiguous()
def timeit(fn,n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
rows=[]
def rec(name, ms, flops=0, bytes_=0):
    rows.append((name, ms*1e3/B, flops/ms/1e9 if flops else 0, bytes_/ms/1e6 if bytes_ else 0))
    print("%-34s %8.1f us/img  %7.0f TFLOP/s %7.0f GB/s"%rows[-1], flush=True)
sp=L.stream_ptr
lv=[(H0,W0),(H0//2,W0//2),(H0//4,W0//4),(H0//8,W0//8)]
# conv1_1 head / tail
x=torch.rand(B,3,H0,W0,device=dev); w0=torch.randn(64,3,3,3,device=dev)*0.1; b0=torch.zeros(64,device=dev)
w0f=torch.empty(64,64,device=dev,dtype=torch.bfloat16); L.call("isx_pack_conv1_1_fwd",w0,w0f,sp())
w0d=torch.empty(9,16,64,device=dev,dtype=torch.bfloat16); L.call("isx_pack_conv1_1_dgrad",w0,w0d,sp())
a11=torch.empty(B,H0,W0,64,device=dev,dtype=torch.bfloat16)
rec("conv1_1 head (tc)", timeit(lambda: L.call("isx_conv1_1_fwd_tc",x,3,None,0,w0f,b0,a11,B,H0,W0,sp())), 2*27*64*B*H0*W0, B*H0*W0*(12+128))
g11=bf(B,H0,W0,64,relu=False); dx=torch.empty_like(x)
rec("conv1_1 tail dgrad (tc, N=16)", timeit(lambda: L.call("isx_conv1_1_dgrad_tc",g11,w0d,None,0,dx,3,B,H0,W0,sp())), 2*27*64*B*H0*W0, B*H0*W0*(12+128))
specs=[("conv1_2",0,64,64),("conv2_1",1,64,128),("conv2_2",1,128,128),("conv3_1",2,128,256),("conv3_2",2,256,256),("conv4_1",3,256,512),("conv4_2",3,512,512)]
for name,l,cin,cout in specs:
    h,w=lv[l]
    xin=bf(B,h,w,cin); wt=torch.randn(cout,cin,3,3,device=dev)*0.03
    wf=torch.empty(9,cout,cin,device=dev,dtype=torch.bfloat16); wd=torch.empty(9,cin,cout,device=dev,dtype=torch.bfloat16)
    L.call("isx_pack_conv3x3_weights",wt,cout,cin,wf,wd,sp())
    bias=torch.zeros(cout,device=dev); out=torch.empty(B,h,w,cout,device=dev,dtype=torch.bfloat16)
    fl=2*9*cin*cout*B*h*w
    rec(name+" fwd", timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",xin,wf,bias,out,B,h,w,cin,cout,1,0,sp())), fl)
    dy=bf(B,h,w,cout,relu=False); dxo=torch.empty(B,h,w,cin,device=dev,dtype=torch.bfloat16)
    rec(name+" dgrad plain", timeit(lambda: L.call("isx_conv3x3_dgrad",dy,wd,dxo,B,h,w,cin,cout,None,None,None,None,0,sp())), fl)
    rec(name+" dgrad +mask", timeit(lambda: L.call("isx_conv3x3_dgrad",dy,wd,dxo,B,h,w,cin,cout,xin,None,None,None,0,sp())), fl)
    D=(torch.randn(B,cin,cin,device=dev)*0.01).bfloat16()
    rec(name+" dgrad +mask+gram", timeit(lambda: L.call("isx_conv3x3_dgrad_gram",dy,wd,dxo,B,h,w,cin,cout,xin,D,sp())), fl+2*cin*cin*B*h*w)
    del xin,out,dy,dxo
for C,l in [(64,0),(128,1),(256,2),(512,3)]:
    h,w=lv[l]; f=bf(B,h,w,C)
    ws=torch.empty(L.call_i64("isx_gram_workspace_bytes",B,h*w,C),device=dev,dtype=torch.uint8)
    tg=torch.zeros(B,C,C,device=dev); loss=torch.zeros(B,device=dev,dtype=torch.float64); D=torch.empty(B,C,C,device=dev,dtype=torch.bfloat16)
    rec("gram fwd+finalize C=%d"%C, timeit(lambda: L.call("isx_gram_fwd",f,B,h*w,C,L.f32(1.0),ws,None,tg,B,L.f64(0.25),loss,L.f32(1.0),D,sp())), 2*C*C*B*h*w, B*h*w*C*2)
    del f
for C,l in [(64,0),(128,1),(256,2)]:
    h,w=lv[l]; a=bf(B,h,w,C); o=torch.empty(B,h//2,w//2,C,device=dev,dtype=torch.bfloat16)
    rec("maxpool fwd C=%d"%C, timeit(lambda: L.call("isx_maxpool2x2_fwd",a,o,B,h,w,C,sp())), 0, B*h*w*C*2*1.25)
    dy=bf(B,h//2,w//2,C,relu=False); dxx=torch.empty_like(a)
    rec("maxpool bwd C=%d"%C, timeit(lambda: L.call("isx_maxpool2x2_bwd",dy,a,dxx,B,h,w,C,sp())), 0, B*h*w*C*2*2.25)
    del a,o,dy,dxx
h,w=lv[3]; p=bf(B,h,w,512); t=bf(B,h,w,512); g=torch.empty_like(p); loss=torch.zeros(B,device=dev,dtype=torch.float64)
rec("content mse", timeit(lambda: L.call("isx_content_mse_fwd_bwd",p,t,B,g,B,L.i64(h*w*512),L.f64(1.0),L.f32(1.0),loss,sp())),0,B*h*w*512*6)
tot=sum(r[1] for r in rows if ("dgrad plain" not in r[0] and "dgrad +mask" != r[0][-11:]))
print("note: an evaluation uses fwd + one dgrad variant per layer; see DESIGN.md")
