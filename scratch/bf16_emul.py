import sys, torch, numpy as np, time
sys.path.insert(0,'/root/repo')
from oracle import nst_oracle as O
import torch.nn.functional as F
torch.set_num_threads(8)
class Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx,x): return x.bfloat16().float()
    @staticmethod
    def backward(ctx,g): return g.bfloat16().float()
q=Q.apply
def fwd_bf16(x, weights, content=O.DEFAULT_CONTENT, style=O.DEFAULT_STYLE):
    mean=torch.tensor(O.IMAGENET_MEAN).view(-1,1,1); std=torch.tensor(O.IMAGENET_STD).view(-1,1,1)
    h=(x-mean)/std
    feats={}; idx=0; ci=0
    want={O.VGG19_LAYERS[n] for n in list(content)+list(style)}; deepest=max(want)
    for v in O.VGG19_CFG:
        if v=='M':
            h=F.max_pool2d(h,2,2); feats[idx]=h; idx+=1
        else:
            w,b=weights[ci]; ci+=1
            wq = w.bfloat16().float()
            hin = h if ci==1 else h   # conv1_1 reads fp32 x
            h=q(F.relu(F.conv2d(hin, wq if ci>1 else w, b, padding=1)))
            feats[idx]=h; feats[idx+1]=h; idx+=2
        if idx>deepest: break
    return [feats[O.VGG19_LAYERS[n]] for n in content],[feats[O.VGG19_LAYERS[n]] for n in style]
def rand_img(seed, shape):
    g=torch.Generator().manual_seed(seed); return torch.rand(shape,generator=g)
W=O.random_vgg19_weights(0)
import importlib.util
sys.path.insert(0,'/root/repo/iris-style-transfer_b200'); import synthetic
H,Wd=int(sys.argv[1]),int(sys.argv[2]); epochs=int(sys.argv[3])
fr,_=synthetic.synthetic_batch([1,2],H,Wd)
c=torch.from_numpy(fr[0]).repeat(3,1,1)[None]; s=torch.from_numpy(fr[1]).repeat(3,1,1)[None]
def nst_bf16(c,s,epochs,beta=1e6):
    with torch.no_grad():
        cf,_=fwd_bf16(c,W); _,sf=fwd_bf16(s,W); tg=[O.gram_matrix(t) for t in sf]
    x=c.clone(); opt=O.LBFGS(x); n=[0]; ch=[];sh=[]
    def closure():
        with torch.no_grad(): x.clamp_(0,1)
        xv=x.detach().requires_grad_(True)
        with torch.enable_grad():
            xc,xs=fwd_bf16(xv,W); cl=O.content_loss_l2(xc,cf); sl=O.style_loss_gram(xs,tg); loss=cl+sl*beta
            g,=torch.autograd.grad(loss,xv)
        ch.append(float(cl));sh.append(float(sl)); n[0]+=1
        return float(loss), g.reshape(-1)
    while n[0]<epochs: opt.step(closure)
    return x.detach().clamp_(0,1), ch, sh
t=time.time()
xr,_,chr_,shr=O.nst(c,s,W,BN_loss=False,s_loss_weight=1e6,epochs=epochs,keep_hist=False); print('fp32',time.time()-t)
xb,chb,shb=nst_bf16(c,s,epochs)
print('moved MAE',float((xr-c).abs().mean()),'bf16-vs-fp32 MAE',float((xr-xb).abs().mean()), 'max', float((xr-xb).abs().max()))
sr=np.array(shr); sb=np.array(shb)
print('s_loss rel err per eval (first 10):', np.abs(sb-sr)[:10]/sr[:10])
print('s_loss final', sr[-1], sb[-1], 'c_loss final', chr_[-1], chb[-1])
# single-eval grad error
with torch.no_grad():
    _,cf,_=O.vgg19_forward(c,W,full=False); _,_,sf=O.vgg19_forward(s,W,full=False); tg=[O.gram_matrix(t) for t in sf]
xq=(c*0.7+0.3*s).clone()
cl,sl,g=O.nst_eval(xq,cf,tg,W,False,1.0,1e6)
xv=xq.clone().requires_grad_(True)
with torch.no_grad(): cfb,_=fwd_bf16(c,W); _,sfb=fwd_bf16(s,W); tgb=[O.gram_matrix(t) for t in sfb]
xc,xs=fwd_bf16(xv,W); clb=O.content_loss_l2(xc,cfb); slb=O.style_loss_gram(xs,tgb); (gb,)=torch.autograd.grad(clb+slb*1e6,xv)
print('eval: c', cl, float(clb), 's', sl, float(slb), 'grad rel L2 err', float((g-gb).norm()/g.norm()), 'cos', float((g*gb).sum()/g.norm()/gb.norm()))
for i,(a,b) in enumerate(zip(sf,sfb)):
    Ga,Gb=O.gram_matrix(a),O.gram_matrix(b); print('gram',i,'rel fro err',float((Ga-Gb).norm()/Ga.norm()), 'max rel', float(((Ga-Gb).abs()/Ga.abs().clamp_min(1e-12)).max()))
