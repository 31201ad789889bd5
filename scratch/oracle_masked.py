import sys, torch, numpy as np, time
sys.path.insert(0,'.')
import bench
from oracle import nst_oracle as O
torch.set_num_threads(8)
W=O.random_vgg19_weights(0)
for seed in (1,2):
    c,s=bench.make_inputs(1,seed)
    t=time.time(); x,_,ch,sh=O.nst(c,s,W,BN_loss=False,s_loss_weight=1e6,epochs=60,keep_hist=False)
    print(seed,'t',round(time.time()-t,1),'moved',float((x-c).abs().mean()),'s_hist', np.array2string(np.array(sh[::3]),precision=2))
