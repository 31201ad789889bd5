"""Feature-extraction leg, one batch of 32 eyes at 640x400, 4 style taps: for an ncu launch list (r02)."""
import sys, torch
sys.path.insert(0, '.')
import iris_b200
from iris_b200 import features, synthetic
dev = torch.device('cuda:0')
taps = ["relu1_1", "relu2_1", "relu3_1", "relu4_1"] + (["relu5_1"] if len(sys.argv) > 1 and sys.argv[1] == "5" else [])
vgg = iris_b200.VGG19(content_layers=[], style_layers=taps, weights="random", seed=0)
base, _ = synthetic.synthetic_batch(list(range(16)), 640, 400)
xb = torch.from_numpy(base).repeat(2, 1, 1, 1).to(dev)
for _ in range(3):
    out = features.style_features_batch(vgg, xb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = features.style_features_batch(vgg, xb)
e1.record(); torch.cuda.synchronize()
print("batch of 32: %.3f ms  -> %.1f images/s" % (e0.elapsed_time(e1) / 5, 32 * 5 / (e0.elapsed_time(e1) / 1e3)))
