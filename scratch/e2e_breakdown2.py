import sys, time, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import pipelines
import bench
dev = torch.device('cuda:0')
vgg = iris_b200.VGG19(weights="random", seed=0)
c_host, s_host = bench.make_inputs(64, 1)
c_host, s_host = c_host.pin_memory(), s_host.pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
import cProfile, pstats
for rep in range(2):
    torch.cuda.empty_cache()
    t0 = T()
    job = pipelines.NstJob(c_host, s_host, vgg, dev, clone_content=True, BN_loss=False, c_loss_weight=1.0, s_loss_weight=1e6, lr=1.0,
                           epochs=300, independent=True)
    t1 = T()
    print("max_ticks", job.max_ticks)
    while job.ticks < job.max_ticks:
        job.tick()
        if job.ticks % 20 == 0 or job.ticks >= job.max_ticks:
            if bool((job.evals_done() > 0).all().item()):
                break
    t2 = T()
    out = job.finish()
    t3 = T()
    xh = out[0].cpu()
    t4 = T()
    print("rep %d: setup %.1f | ticks(%d) %.1f | finish %.1f | x.cpu %.1f | total %.1f ms" % (rep, (t1-t0)*1e3, job.ticks, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t4-t0)*1e3), flush=True)
    del job, out
pr = cProfile.Profile()
torch.cuda.empty_cache()
t0 = T()
pr.enable()
x, _, ch, sh = iris_b200.nst(c_host, s_host, BN_loss=False, c_loss_weight=1.0, s_loss_weight=1e6, epochs=300, vgg=vgg,
                             use_tqdm=False, device='cuda:0', independent=True, x_hist_stride=0)
pr.disable()
t1 = T()
print("nst(): %.1f ms" % ((t1-t0)*1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
