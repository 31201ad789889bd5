import sys, os, time, numpy as np, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import pipelines, vgg, synthetic
from oracle import nst_oracle as O
torch.set_num_threads(os.cpu_count())
W=O.random_vgg19_weights(0); net=vgg.VGG19(weights=W)
traj=np.load('tests/golden/nst_traj.npz')
def rand_img(seed, shape):
    g=torch.Generator().manual_seed(seed); return torch.rand(shape,generator=g)
def rep(tag,x,ch,sh,ref_x,rc,rs,c_img):
    mae=float((x-ref_x).abs().mean()); moved=float((ref_x-c_img).abs().mean())
    k=min(len(sh),len(rs))
    rel=np.abs(np.array(sh[:k])-np.array(rs[:k]))/np.maximum(np.array(rs[:k]),1e-30)
    print("%-22s evals %d/%d MAE %.5f moved %.5f ratio %.2f | s0 %.4g/%.4g sN %.4g/%.4g | rel s_loss first5 %s"%(tag,len(sh),len(rs),mae,moved,mae/max(moved,1e-9),sh[0],rs[0],sh[-1],rs[-1],np.round(rel[:5],4)),flush=True)
def run(c,s,**kw):
    x,xh,ch,sh=pipelines.nst(c,s,vgg=net,use_tqdm=False,device='cuda:0',x_hist_stride=0,**kw); return x.cpu(),ch,sh
c1,s1=rand_img(21,(1,3,48,64)),rand_img(22,(1,3,48,64))
for tag,kw,ci,si in [("gram_b1",dict(BN_loss=False,s_loss_weight=1e6,epochs=50),c1,s1),("bn_b1",dict(BN_loss=True,s_loss_weight=1e4,epochs=40),c1,s1),
    ("gram_b2_coupled",dict(BN_loss=False,s_loss_weight=1e6,epochs=20),rand_img(11,(2,3,48,64)),rand_img(12,(2,3,48,64))),
    ("gram_long",dict(BN_loss=False,s_loss_weight=1e6,epochs=130),rand_img(31,(1,3,32,32)),rand_img(32,(1,3,32,32)))]:
    x,ch,sh=run(ci,si,**kw); rep(tag,x,ch,sh,torch.from_numpy(traj[tag+"_x"]),traj[tag+"_c_hist"],traj[tag+"_s_hist"],ci)
ic=torch.from_numpy(synthetic.synthetic_iris_crops([1,2],96))
x,ch,sh=run(ic[:1],ic[1:2],BN_loss=False,s_loss_weight=1e6,epochs=40); rep("gram_iris96",x,ch,sh,torch.from_numpy(traj["gram_iris96_x"]),traj["gram_iris96_c_hist"],traj["gram_iris96_s_hist"],ic[:1])
# synthetic eyes vs live oracle on this box's CPU
for (H,Wd,ep,BN,beta) in [(160,100,40,False,1e6),(160,100,40,True,1e4),(320,200,60,False,1e6),(640,400,50,False,1e6)]:
    fr,_=synthetic.synthetic_batch([1,2],H,Wd)
    c=torch.from_numpy(fr[0]).repeat(3,1,1)[None]; s=torch.from_numpy(fr[1]).repeat(3,1,1)[None]
    t=time.time(); xr,_,cr,sr=O.nst(c,s,W,BN_loss=BN,s_loss_weight=beta,epochs=ep,keep_hist=False); tc=time.time()-t
    torch.cuda.synchronize(); t=time.time(); x,ch,sh=run(c,s,BN_loss=BN,s_loss_weight=beta,epochs=ep); torch.cuda.synchronize(); tg=time.time()-t
    rep("eye%dx%d%s"%(H,Wd,"bn" if BN else "gram"),x,ch,sh,xr,cr,sr,c); print("   cpu %.2fs (%.2f it/s, %d threads)  gpu %.3fs (%.1f it/s)"%(tc,len(sr)/tc,torch.get_num_threads(),tg,len(sh)/tg),flush=True)
