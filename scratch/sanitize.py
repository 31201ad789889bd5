"""Small end-to-end job for compute-sanitizer memcheck: every kernel family once (tcgen05 convs incl. fused pool /
Gram / masks, L-BFGS with a short history, image ops)."""
import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import synthetic
vgg=iris_b200.VGG19(weights="random")
fr,seg=synthetic.synthetic_batch([1,2],72,56)
c=torch.from_numpy(fr).repeat(1,3,1,1).cuda(); s=torch.from_numpy(fr[::-1].copy()).repeat(1,3,1,1).cuda()
for kw in (dict(BN_loss=False),dict(BN_loss=True),dict(BN_loss=False,independent=True,history_dtype=torch.bfloat16),
           dict(BN_loss=False,c_mask=torch.from_numpy(seg==2).cuda())):
    x,_,ch,sh=iris_b200.nst(c,s,s_loss_weight=1e6 if not kw.get('BN_loss') else 1e4,epochs=25,vgg=vgg,use_tqdm=False,device='cuda:0',x_hist_stride=0,history_size=5,**kw)
    torch.cuda.synchronize(); print(kw, len(sh), sh[0], sh[-1])
ft,st=torch.from_numpy(fr).cuda(),torch.from_numpy(seg).cuda()
mask,bbox=iris_b200.iris_masks_and_bboxes(ft,st)
crops=iris_b200.crop_resize_irises(ft,mask,bbox,(32,32))
out=iris_b200.composite_irises(ft.clone(),crops,mask,bbox)
xg=c.clone().requires_grad_(True); p5,xc,xs=vgg(xg); (p5.sum()+iris_b200.GramMatrix(xs[0]).sum()).backward()
torch.cuda.synchronize(); print("ok", float(out.sum()), float(xg.grad.abs().sum()))
