"""Where the feature-extraction leg (BASELINE config 3) spends its time, batch of 32 eyes at 640x400."""
import sys, time, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import features, synthetic
from iris_b200.engine import gram_of, stats_of
dev = torch.device('cuda:0')
vgg = iris_b200.VGG19(content_layers=[], weights="random", seed=0)
base, _ = synthetic.synthetic_batch(list(range(16)), 640, 400)
xb_host = torch.from_numpy(base).repeat(2, 1, 1, 1).pin_memory()
def T(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
xb = xb_host.to(dev)
print("h2d            %.2f ms" % T(lambda: xb_host.to(dev, non_blocking=True)))
print("forward        %.2f ms" % T(lambda: vgg.features_nhwc(xb, full=False)))
_, _, s, _ = vgg.features_nhwc(xb, full=False)
print("stats x4       %.2f ms" % T(lambda: [stats_of(f) for f in s]))
print("gram x4        %.2f ms" % T(lambda: [gram_of(f) for f in s]))
Gs = [gram_of(f) for f in s]
def triu():
    cols = []
    for G in Gs:
        iu = torch.triu_indices(G.shape[-1], G.shape[-1], device=dev)
        cols.append(G[:, iu[0], iu[1]])
    return torch.cat(cols, dim=1)
print("triu + cat     %.2f ms" % T(triu))
print("whole batch    %.2f ms" % T(lambda: features.style_features_batch(vgg, xb)))
