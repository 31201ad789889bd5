import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/iris-style-transfer_b200')
from oracle import nst_oracle as O
import synthetic
torch.set_num_threads(8)
W=O.random_vgg19_weights(0)
H,Wd,epochs=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
fr,_=synthetic.synthetic_batch([1,2],H,Wd)
c=torch.from_numpy(fr[0]).repeat(3,1,1)[None]; s=torch.from_numpy(fr[1]).repeat(3,1,1)[None]
def run(noise, BN=False, beta=1e6):
    with torch.no_grad():
        _,cf,_=O.vgg19_forward(c,W,full=False); _,_,sf=O.vgg19_forward(s,W,full=False)
        tg=([t.mean(dim=(-2,-1)) for t in sf],[t.std(dim=(-2,-1)) for t in sf]) if BN else [O.gram_matrix(t) for t in sf]
    x=c.clone(); opt=O.LBFGS(x); n=[0]; sh=[]; xs=[]
    gen=torch.Generator().manual_seed(5)
    def closure():
        with torch.no_grad(): x.clamp_(0,1)
        cl,sl,g=O.nst_eval(x,cf,tg,W,BN,1.0,beta)
        g=g.reshape(-1)
        if noise>0: g=g*(1+noise*torch.randn(g.shape,generator=gen))
        sh.append(sl); n[0]+=1; xs.append(x.clone())
        return cl+sl*beta, g
    while n[0]<epochs: opt.step(closure)
    return x.clamp(0,1), sh, xs
x0,s0,xs0=run(0)
for nz in (1e-6,1e-4,1e-2):
    x1,s1,xs1=run(nz)
    print('noise',nz,'final MAE',float((x0-x1).abs().mean()),'moved',float((x0-c).abs().mean()),'s_loss',s0[-1],s1[-1], 'MAE@10,20,40:',[round(float((xs0[k]-xs1[k]).abs().mean()),5) for k in (10,20,min(40,len(xs0)-1))])
