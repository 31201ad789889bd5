import sys, torch, torch.nn.functional as F
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import _lib as L
L.load(); sp=L.stream_ptr
torch.backends.cudnn.allow_tf32=False
def bf(*shape, relu=True):
    t=torch.randn(*shape,device='cuda')
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
for (B,H,W,Cin,Cout) in [(2,40,24,64,64),(1,50,80,128,128),(3,33,17,64,128)]:
    x=bf(B,H,W,Cin); w=torch.randn(Cout,Cin,3,3,device='cuda')*(2.0/(9*Cin))**0.5; bias=torch.randn(Cout,device='cuda')*0.1
    wf=torch.empty(9,Cout,Cin,device='cuda',dtype=torch.bfloat16); wd=torch.empty(9,Cin,Cout,device='cuda',dtype=torch.bfloat16)
    L.call("isx_pack_conv3x3_weights",w,Cout,Cin,wf,wd,sp())
    ref=F.relu(F.conv2d(x.float().permute(0,3,1,2),w.bfloat16().float(),bias,padding=1)).permute(0,2,3,1)
    bn=min(Cout,128)
    for mode in (0,1,2):
        out=torch.full((B,H,W,Cout),float('nan'),device='cuda',dtype=torch.bfloat16)
        cfg=mode*1000000+bn*100+10
        try:
            L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,H,W,Cin,Cout,1,cfg,sp()); torch.cuda.synchronize()
            err=(out.float()-ref).abs(); print((B,H,W,Cin,Cout),'mode',mode,'max err %.4f (ref max %.2f) nan %d'%(err.nan_to_num(99).max().item(),ref.abs().max().item(),int(out.float().isnan().sum())),flush=True)
        except Exception as e:
            print('mode',mode,'ERR',str(e)[:100]); 
# speed at real shapes
B=32
for name,h,w,cin,cout in [("conv1_2",640,400,64,64),("conv2_1",320,200,64,128),("conv2_2",320,200,128,128),("conv3_2",160,100,256,256)]:
    x=bf(B,h,w,cin); wf=(torch.randn(9,cout,cin,device='cuda')*0.03).bfloat16(); bias=torch.zeros(cout,device='cuda'); out=torch.empty(B,h,w,cout,device='cuda',dtype=torch.bfloat16)
    fl=2*9*cin*cout*B*h*w; res=[]
    for cfg in (0, 1000000+min(cout,256)*100+10, 2000000+min(cout,256)*100+10, 1000000+min(cout,128)*100+10):
        try:
            ms=timeit(lambda: L.call("isx_conv3x3_bias_relu_fwd",x,wf,bias,out,B,h,w,cin,cout,1,cfg,sp())); res.append("%d: %.0f TF"%(cfg,fl/ms/1e9))
        except Exception as e: res.append("%d ERR %s"%(cfg,str(e)[:60]))
    print(name,"|".join(res),flush=True)
