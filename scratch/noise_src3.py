# fp16 vs bf16 forward operands (CPU emulation): gradient error per evaluation and final-image MAE
import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/iris-style-transfer_b200')
from oracle import nst_oracle as O
import torch.nn.functional as F, synthetic
torch.set_num_threads(8)
def mkq(dt, bwd_dt=None):
    class Q(torch.autograd.Function):
        @staticmethod
        def forward(ctx,x): return x.to(dt).float()
        @staticmethod
        def backward(ctx,g): return g.to(bwd_dt).float() if bwd_dt is not None else g
    return Q.apply
def fwdq(x, W, dt, bwd_dt=None):
    q=mkq(dt,bwd_dt)
    mean=torch.tensor(O.IMAGENET_MEAN).view(-1,1,1); std=torch.tensor(O.IMAGENET_STD).view(-1,1,1)
    h=(x-mean)/std; feats={}; idx=0; ci=0
    for v in O.VGG19_CFG:
        if v=='M': h=F.max_pool2d(h,2,2); idx+=1
        else:
            w,b=W[ci]; ci+=1
            ww = w.to(dt).float()
            h=q(F.relu(F.conv2d(h,ww,b,padding=1))); feats[idx+1]=h; idx+=2
        if idx>22: break
    return [feats[22]],[feats[i] for i in (1,6,11,20)]
W=O.random_vgg19_weights(0)
ic=torch.from_numpy(synthetic.synthetic_iris_crops([1,2],96)); c,s=ic[:1],ic[1:2]
fr,_=synthetic.synthetic_batch([1,2],160,100)
cases={'iris96':(c,s),'eye160':(torch.from_numpy(fr[0]).repeat(3,1,1)[None],torch.from_numpy(fr[1]).repeat(3,1,1)[None])}
def nst_q(c,s,epochs,dt,BN,beta=1e6):
    def targets(sf): return ([t.mean(dim=(-2,-1)) for t in sf],[t.std(dim=(-2,-1)) for t in sf]) if BN else [O.gram_matrix(t) for t in sf]
    def sloss(xs,tg): return O.style_loss_bn(xs,tg[0],tg[1]) if BN else O.style_loss_gram(xs,tg)
    with torch.no_grad():
        cf,_=fwdq(c,W,dt); _,sf=fwdq(s,W,dt); tg=targets(sf)
    x=c.clone(); opt=O.LBFGS(x); n=[0]
    def closure():
        with torch.no_grad(): x.clamp_(0,1)
        xv=x.detach().requires_grad_(True)
        with torch.enable_grad():
            xc,xs=fwdq(xv,W,dt,torch.bfloat16); cl=O.content_loss_l2(xc,cf); sl=sloss(xs,tg); loss=cl+sl*beta
            g,=torch.autograd.grad(loss,xv)
        n[0]+=1
        return float(loss), g.reshape(-1)
    while n[0]<epochs: opt.step(closure)
    return x.detach().clamp_(0,1)
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
    x40,_,_,_=O.nst(c,s,W,BN_loss=BN,s_loss_weight=beta,epochs=40,keep_hist=False)
    for label,dt in [('bf16',torch.bfloat16),('fp16',torch.float16)]:
        with torch.no_grad():
            cfb,_=fwdq(c,W,dt); _,sfb=fwdq(s,W,dt); tgb=targets(sfb)
        xv=xq.clone().requires_grad_(True)
        xc,xs=fwdq(xv,W,dt,torch.bfloat16); loss=O.content_loss_l2(xc,cfb)+beta*sloss(xs,tgb); (gb,)=torch.autograd.grad(loss,xv)
        xb=nst_q(c,s,40,dt,BN)
        print('BN' if BN else 'Gram',name,label,'grad rel L2 err %.4f cos %.5f'%(float((g-gb).norm()/g.norm()), float((g*gb).sum()/g.norm()/gb.norm())),
              '| 40-eval final MAE %.5f (moved %.5f)'%(float((xb-x40).abs().mean()), float((x40-c).abs().mean())))
