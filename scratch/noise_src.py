import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/iris-style-transfer_b200')
from oracle import nst_oracle as O
import torch.nn.functional as F, synthetic
torch.set_num_threads(8)
def mkQ(fw,bw):
    class Q(torch.autograd.Function):
        @staticmethod
        def forward(ctx,x): return x.bfloat16().float() if fw else x
        @staticmethod
        def backward(ctx,g): return g.bfloat16().float() if bw else g
    return Q.apply
def fwdq(x, W, q, wq=True):
    mean=torch.tensor(O.IMAGENET_MEAN).view(-1,1,1); std=torch.tensor(O.IMAGENET_STD).view(-1,1,1)
    h=(x-mean)/std; feats={}; idx=0; ci=0
    for v in O.VGG19_CFG:
        if v=='M': h=F.max_pool2d(h,2,2); idx+=1
        else:
            w,b=W[ci]; ci+=1
            ww = w.bfloat16().float() if (wq and ci>1) else w
            h=q(F.relu(F.conv2d(h,ww,b,padding=1))); feats[idx+1]=h; idx+=2
        if idx>22: break
    return [feats[22]],[feats[i] for i in (1,6,11,20)]
W=O.random_vgg19_weights(0)
ic=torch.from_numpy(synthetic.synthetic_iris_crops([1,2],96)); c,s=ic[:1],ic[1:2]
fr,_=synthetic.synthetic_batch([1,2],160,100)
cases={'iris96':(c,s),'eye160':(torch.from_numpy(fr[0]).repeat(3,1,1)[None],torch.from_numpy(fr[1]).repeat(3,1,1)[None])}
for name,(c,s) in cases.items():
    with torch.no_grad():
        _,cf,_=O.vgg19_forward(c,W,full=False); _,_,sf=O.vgg19_forward(s,W,full=False); tg=[O.gram_matrix(t) for t in sf]
    xq=c.clone()
    # take x after a few oracle steps to be at a typical point
    xr,_,_,_=O.nst(c,s,W,BN_loss=False,s_loss_weight=1e6,epochs=5,keep_hist=False); xq=xr.clone()
    cl,sl,g=O.nst_eval(xq,cf,tg,W,False,1.0,1e6)
    for label,fw,bw,wq,dq in [('fwd+bwd+w',1,1,1,0),('fwd only',1,0,0,0),('bwd only',0,1,0,0),('w only',0,0,1,0),('fwd+w',1,0,1,0)]:
        q=mkQ(fw,bw)
        with torch.no_grad():
            cfb,_=fwdq(c,W,q,wq); _,sfb=fwdq(s,W,q,wq); tgb=[O.gram_matrix(t) for t in sfb]
        xv=xq.clone().requires_grad_(True)
        xc,xs=fwdq(xv,W,q,wq); loss=O.content_loss_l2(xc,cfb)+1e6*O.style_loss_gram(xs,tgb); (gb,)=torch.autograd.grad(loss,xv)
        print(name,label,'grad rel L2 err %.4f cos %.5f'%(float((g-gb).norm()/g.norm()), float((g*gb).sum()/g.norm()/gb.norm())))
