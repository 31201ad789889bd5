import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/iris-style-transfer_b200')
from oracle import nst_oracle as O
import torch.nn.functional as F, synthetic
torch.set_num_threads(8)
def mkQ(mode):
    class Q(torch.autograd.Function):
        @staticmethod
        def forward(ctx,x): return x.half().float() if mode=='fp16' else (x.bfloat16().float() if mode=='bf16' else x)
        @staticmethod
        def backward(ctx,g): return g.bfloat16().float() if mode!='fp32' else g
    return Q.apply
def fwdq(x, W, q, mode):
    mean=torch.tensor(O.IMAGENET_MEAN).view(-1,1,1); std=torch.tensor(O.IMAGENET_STD).view(-1,1,1)
    h=(x-mean)/std; feats={}; idx=0; ci=0
    for v in O.VGG19_CFG:
        if v=='M': h=F.max_pool2d(h,2,2); idx+=1
        else:
            w,b=W[ci]; ci+=1
            ww = w if (mode=='fp32' or ci==1) else (w.half().float() if mode=='fp16' else w.bfloat16().float())
            h=q(F.relu(F.conv2d(h,ww,b,padding=1))); feats[idx+1]=h; idx+=2
        if idx>22: break
    return [feats[22]],[feats[i] for i in (1,6,11,20)]
W=O.random_vgg19_weights(0)
size=int(sys.argv[1]); ep=int(sys.argv[2])
ic=torch.from_numpy(synthetic.synthetic_iris_crops([1,2],size)); c,s=ic[:1],ic[1:2]
def run(mode,beta=1e6):
    q=mkQ(mode)
    with torch.no_grad():
        cf,_=fwdq(c,W,q,mode); _,sf=fwdq(s,W,q,mode); tg=[O.gram_matrix(t) for t in sf]
    x=c.clone(); opt=O.LBFGS(x); n=[0]; sh=[]
    def closure():
        with torch.no_grad(): x.clamp_(0,1)
        xv=x.detach().requires_grad_(True)
        with torch.enable_grad():
            xc,xs=fwdq(xv,W,q,mode); cl=O.content_loss_l2(xc,cf); sl=O.style_loss_gram(xs,tg); loss=cl+sl*beta
            g,=torch.autograd.grad(loss,xv)
        sh.append(float(sl)); n[0]+=1
        return float(loss), g.reshape(-1)
    while n[0]<ep: opt.step(closure)
    return x.detach().clamp(0,1), sh
x0,s0=run('fp32')
for mode in ('bf16','fp16'):
    x1,s1=run(mode)
    print(size,mode,'MAE %.5f moved %.5f'%(float((x0-x1).abs().mean()),float((x0-c).abs().mean())),'s_loss ref %.3g -> %.3g | %s %.3g -> %.3g ; max over run %.3g vs %.3g'%(s0[0],s0[-1],mode,s1[0],s1[-1],max(s1),max(s0)))
    print('   ref traj', np.array2string(np.array(s0[::4]),precision=2)); print('   '+mode, np.array2string(np.array(s1[::4]),precision=2))
