import sys, torch
sys.path.insert(0,'.')
import iris_b200
from iris_b200 import engine as E, synthetic
from oracle import nst_oracle as O
W_=O.random_vgg19_weights(0); net=iris_b200.VGG19(weights=W_); dev=torch.device('cuda:0')
for (H,W) in [(75,101),(80,104),(76,100),(75,104)]:
    fr,_=synthetic.synthetic_batch([3],H,W); c=torch.from_numpy(fr).repeat(1,3,1,1)
    fs,_=synthetic.synthetic_batch([9],H,W); s=torch.from_numpy(fs).repeat(1,3,1,1)
    g0=torch.Generator().manual_seed(5); xq=(c+0.02*torch.randn(c.shape,generator=g0)).clamp(0,1)
    eng=E.NstEngine(net.packed(dev),1,H,W,3,net.content_convs,net.style_convs,style_mode=0,c_weight=1.0,s_weight=1e6,coupled=True)
    eng.forward(c.to(dev)); eng.set_content_targets([eng.feature(0,i) for i in net.content_convs])
    eng.forward(s.to(dev)); eng.set_gram_targets([E.gram_of(eng.feature(0,i)) for i in net.style_convs])
    g=torch.empty(1,3,H,W,device=dev); eng.eval(xq.to(dev),g); torch.cuda.synchronize()
    with torch.no_grad():
        _,cf,_=O.vgg19_forward(c,W_,full=False); _,_,sf=O.vgg19_forward(s,W_,full=False); tg=[O.gram_matrix(t) for t in sf]
    rcl,rsl,rg=O.nst_eval(xq,cf,tg,W_,False,1.0,1e6)
    gc=g.cpu(); cos=float((gc*rg).sum()/(gc.norm()*rg.norm()))
    # error map: borders vs interior
    err=(gc-rg).abs()[0].sum(0); 
    print((H,W),'c %.4g/%.4g s %.4g/%.4g cos %.4f relL2 %.3f | err border rows %.3g interior %.3g last col %.3g'%(float(eng.loss_c.sum()),rcl,float(eng.loss_s.sum()),rsl,cos,float((gc-rg).norm()/rg.norm()), float(err[-2:].mean()), float(err[8:-8,8:-8].mean()), float(err[:,-2:].mean())))
