"""Throughput of nst() on the shapes the reference's drivers actually use (224x224 crops, batch 64 / 128, default BN
loss and Gram loss, 200 evaluations) and on BASELINE config 5 (1024x1024 RGB, Gram loss, one image)."""
import sys, time, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import synthetic
vgg=iris_b200.VGG19(weights="random")
def run(tag,c,s,ep,**kw):
    for rep in range(2):
        torch.cuda.synchronize(); t=time.perf_counter()
        x,_,ch,sh=iris_b200.nst(c,s,epochs=ep,vgg=vgg,use_tqdm=False,device='cuda:0',x_hist_stride=0,**kw)
        torch.cuda.synchronize(); dt=time.perf_counter()-t
    B=c.shape[0]
    print("%-40s %d evals %.3f s -> %.1f image-steps/s (s_loss %.3g -> %.3g)"%(tag,len(sh),dt,B*len(sh)/dt,sh[0],sh[-1]),flush=True)
for B in (64,128):
    c=torch.from_numpy(synthetic.synthetic_iris_crops(list(range(B)),224)).cuda()
    s=torch.from_numpy(synthetic.synthetic_iris_crops(list(range(500,500+B)),224)).cuda()
    run("224x224 B=%d BN loss (reference default)"%B,c,s,200,BN_loss=True,s_loss_weight=1e4)
    run("224x224 B=%d BN loss independent"%B,c,s,200,BN_loss=True,s_loss_weight=1e4,independent=True)
    run("224x224 B=%d Gram loss independent"%B,c,s,200,BN_loss=False,s_loss_weight=1e6,independent=True)
g=torch.Generator().manual_seed(1)
c=torch.rand(1,3,1024,1024,generator=g).cuda(); s=torch.rand(1,3,1024,1024,generator=g).cuda()
run("1024x1024 RGB B=1 Gram (config 5)",c,s,100,BN_loss=False,s_loss_weight=1e6)
