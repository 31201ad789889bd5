import sys, time, torch
sys.path.insert(0,'.')
import iris_b200, bench
from iris_b200 import synthetic
vgg=iris_b200.VGG19(weights="random")
for (B,H,W,ep) in [(1,640,400,100),(1,224,224,200),(4,224,224,200),(64,224,224,100)]:
    fr,_=synthetic.synthetic_batch(list(range(1,B+1)),H,W); c=torch.from_numpy(fr).repeat(1,3,1,1).cuda()
    fr2,_=synthetic.synthetic_batch(list(range(101,B+101)),H,W); s=torch.from_numpy(fr2).repeat(1,3,1,1).cuda()
    for g in (False,True):
        for rep in range(2):
            torch.cuda.synchronize(); t=time.perf_counter()
            x,_,ch,sh=iris_b200.nst(c,s,BN_loss=False,s_loss_weight=1e6,epochs=ep,vgg=vgg,use_tqdm=False,device='cuda:0',x_hist_stride=0,independent=True,cuda_graph=g)
            torch.cuda.synchronize(); dt=time.perf_counter()-t
        print("B=%d %dx%d graph=%s: %d evals %.3f s -> %.1f image-steps/s  s_loss %.3g -> %.3g"%(B,H,W,g,len(sh),dt,B*len(sh)/dt,sh[0],sh[-1]),flush=True)
