"""Where the end-to-end call spends its wall clock (bench config: 64 eyes, 300 evaluations)."""
import sys, time, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import pipelines, synthetic
import bench
dev = torch.device('cuda:0')
vgg = iris_b200.VGG19(weights="random", seed=0)
c_host, s_host = bench.make_inputs(64, 1)
c_host, s_host = c_host.pin_memory(), s_host.pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    torch.cuda.empty_cache()
    t0 = T()
    c = c_host.to(dev, non_blocking=True); s = s_host.to(dev, non_blocking=True)
    t1 = T()
    job = pipelines.NstJob(c, s, vgg, dev, clone_content=True, BN_loss=False, c_loss_weight=1.0, s_loss_weight=1e6, lr=1.0,
                           epochs=300, independent=True)
    t2 = T()
    n = 0
    while n < 300:
        job.tick(); n += 1
    t3 = T()
    x = job.x.cpu(); hc = job.hist_c.cpu(); hs = job.hist_s.cpu()
    t4 = T()
    del job
    print("rep %d: h2d %.1f ms | job setup %.1f ms | 300 ticks %.1f ms | d2h %.1f ms | total %.1f ms" % (
        rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t4-t0)*1e3), flush=True)
# the public call
for rep in range(2):
    torch.cuda.empty_cache()
    t0 = T()
    x, _, ch, sh = iris_b200.nst(c_host, s_host, BN_loss=False, c_loss_weight=1.0, s_loss_weight=1e6, epochs=300, vgg=vgg,
                                 use_tqdm=False, device='cuda:0', independent=True, x_hist_stride=0)
    xh = x.cpu()
    t1 = T()
    print("nst(): %.1f ms, %d evals" % ((t1-t0)*1e3, len(sh)), flush=True)
