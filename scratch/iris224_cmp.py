import sys, torch, numpy as np
sys.path.insert(0,'.')
sys.path.insert(0,'iris-style-transfer_b200')
import synthetic
from oracle import nst_oracle as O
torch.set_num_threads(8)
W=O.random_vgg19_weights(0)
c=torch.from_numpy(synthetic.synthetic_iris_crops([0],224)); s=torch.from_numpy(synthetic.synthetic_iris_crops([500],224))
for BN,beta in ((True,1e4),(False,1e6)):
    x,_,ch,sh=O.nst(c,s,W,BN_loss=BN,s_loss_weight=beta,epochs=200,keep_hist=False)
    sh=np.array(sh); print('oracle BN=%s: evals %d s_loss %.4g -> %.4g  min %.4g max %.4g ; every 20:'%(BN,len(sh),sh[0],sh[-1],sh.min(),sh.max()), np.array2string(sh[::20],precision=2))
if torch.cuda.is_available():
    import iris_b200
    vgg=iris_b200.VGG19(weights=W)
    for BN,beta in ((True,1e4),(False,1e6)):
        x,_,ch,sh=iris_b200.nst(c,s,BN_loss=BN,s_loss_weight=beta,epochs=200,vgg=vgg,use_tqdm=False,device='cuda:0',x_hist_stride=0)
        sh=np.array(sh); print('isx    BN=%s: evals %d s_loss %.4g -> %.4g  min %.4g max %.4g ; every 20:'%(BN,len(sh),sh[0],sh[-1],sh.min(),sh.max()), np.array2string(sh[::20],precision=2))
