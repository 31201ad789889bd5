# Which bf16 roundings dominate the per-evaluation gradient error?  (CPU emulation, round 2)
import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/iris-style-transfer_b200')
from oracle import nst_oracle as O
import torch.nn.functional as F, synthetic
torch.set_num_threads(8)
class Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx,x): return x.bfloat16().float()
    @staticmethod
    def backward(ctx,g): return g
q=Q.apply
def fwdq(x, W, mode, wq=True):
    """mode: 'all' = every activation rounded, taps read the rounded map; 'tapfp32' = taps read the unrounded accumulator;
    'none' = no activation rounding"""
    mean=torch.tensor(O.IMAGENET_MEAN).view(-1,1,1); std=torch.tensor(O.IMAGENET_STD).view(-1,1,1)
    h=(x-mean)/std; feats={}; idx=0; ci=0
    for v in O.VGG19_CFG:
        if v=='M': h=F.max_pool2d(h,2,2); idx+=1
        else:
            w,b=W[ci]; ci+=1
            ww = w.bfloat16().float() if wq else w
            a=F.relu(F.conv2d(h,ww,b,padding=1))
            if mode=='none': h=a; feats[idx+1]=a
            elif mode=='tapfp32': h=q(a); feats[idx+1]=a
            else: h=q(a); feats[idx+1]=h
            idx+=2
        if idx>22: break
    return [feats[22]],[feats[i] for i in (1,6,11,20)]
W=O.random_vgg19_weights(0)
ic=torch.from_numpy(synthetic.synthetic_iris_crops([1,2],96)); c,s=ic[:1],ic[1:2]
fr,_=synthetic.synthetic_batch([1,2],160,100)
cases={'iris96':(c,s),'eye160':(torch.from_numpy(fr[0]).repeat(3,1,1)[None],torch.from_numpy(fr[1]).repeat(3,1,1)[None])}
for BN in (False, True):
  for name,(c,s) in cases.items():
    beta=1e6
    def targets(sf): return ([t.mean(dim=(-2,-1)) for t in sf],[t.std(dim=(-2,-1)) for t in sf]) if BN else [O.gram_matrix(t) for t in sf]
    def sloss(xs,tg): return O.style_loss_bn(xs,tg[0],tg[1]) if BN else O.style_loss_gram(xs,tg)
    with torch.no_grad():
        _,cf,_=O.vgg19_forward(c,W,full=False); _,_,sf=O.vgg19_forward(s,W,full=False); tg=targets(sf)
    xr,_,_,_=O.nst(c,s,W,BN_loss=BN,s_loss_weight=beta,epochs=5,keep_hist=False); xq=xr.clone()
    xv=xq.clone().requires_grad_(True)
    _,xc,xs=O.vgg19_forward(xv,W,full=False); loss=O.content_loss_l2(xc,cf)+beta*sloss(xs,tg); (g,)=torch.autograd.grad(loss,xv)
    for label,mode,wq in [('all acts+w bf16','all',1),('taps from fp32 acc, +w','tapfp32',1),('acts only','all',0),('taps fp32, fp32 w','tapfp32',0),('w only','none',1)]:
        with torch.no_grad():
            cfb,_=fwdq(c,W,mode,wq); _,sfb=fwdq(s,W,mode,wq); tgb=targets(sfb)
        xv=xq.clone().requires_grad_(True)
        xc,xs=fwdq(xv,W,mode,wq); loss=O.content_loss_l2(xc,cfb)+beta*sloss(xs,tgb); (gb,)=torch.autograd.grad(loss,xv)
        print('BN' if BN else 'Gram',name,'%-26s'%label,'grad rel L2 err %.4f cos %.5f'%(float((g-gb).norm()/g.norm()), float((g*gb).sum()/g.norm()/gb.norm())))
