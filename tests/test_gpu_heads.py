"""Classifier heads (models/classifiers/classifiers.py:3-72) as weight-streaming tcgen05 GEMMs against a plain PyTorch fp32
reference of the same modules, fed by the VGG outputs the drivers feed them (…2019.py:82-84); bf16 operands, fp32
accumulation: logits within 2e-2 of the largest logit, same arg-max wherever the fp32 top-2 margin exceeds that."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_classifier1(num_class):
    return torch.nn.Sequential(torch.nn.AdaptiveAvgPool2d((7, 7)), torch.nn.Flatten(), torch.nn.Linear(25088, 4096),
                               torch.nn.ReLU(True), torch.nn.Dropout(0.5), torch.nn.Linear(4096, 4096), torch.nn.ReLU(True),
                               torch.nn.Dropout(0.5), torch.nn.Linear(4096, num_class)).eval()


def _ref_classifier2(in_features, num_class):
    return torch.nn.Sequential(torch.nn.Linear(in_features, 4096), torch.nn.ReLU(True), torch.nn.Dropout(0.5),
                               torch.nn.Linear(4096, 4096), torch.nn.ReLU(True), torch.nn.Dropout(0.5),
                               torch.nn.Linear(4096, num_class)).eval()


def _check(got, ref):
    got, ref = got.cpu(), ref.cpu()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    top2 = ref.topk(2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 4e-2 * scale
    agree = (got.argmax(1) == ref.argmax(1))
    print("logits max |err| %.4g of scale %.4g; arg-max agreement %d/%d (%d with a clear margin)" % (
        err, scale, int(agree.sum()), len(agree), int(sure.sum())))
    assert err <= 2e-2 * scale
    assert bool(agree[sure].all())


@pytest.mark.parametrize("B,H,W,K", [(5, 224, 224, 152), (3, 640, 400, 80), (70, 64, 64, 152)])
def test_classifier_heads_match_torch(B, H, W, K):
    import iris_b200
    from iris_b200 import features

    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    vgg = iris_b200.VGG19(weights="random", seed=0)
    g = torch.Generator().manual_seed(B)
    x = torch.rand(B, 1, H, W, generator=g).to(dev)
    r1, r2 = _ref_classifier1(K), _ref_classifier2(1920, K)
    c1 = iris_b200.Classifier1(num_class=K, state_dict={"model." + k: v for k, v in r1.state_dict().items()})
    c2 = iris_b200.Classifier2(num_class=K, state_dict={"model." + k: v for k, v in r2.state_dict().items()})
    c1.to(dev); c2.to(dev)
    with torch.no_grad():
        p5, x_c, x_s = vgg(x.repeat(1, 3, 1, 1))            # fp32 NCHW like the reference's VGG19.forward
        got1, got2 = c1(p5), c2(x_s)
        ref1 = r1.to(dev)(p5)
        ref2 = r2.to(dev)(torch.cat([torch.cat([f.mean(dim=(-2, -1)), f.std(dim=(-2, -1))], dim=1) for f in x_s], dim=1))
    assert tuple(got1.shape) == tuple(got2.shape) == (B, K)
    _check(got1, ref1)
    _check(got2, ref2)
    # the cached inputs (computed once for a frozen VGG) give the same logits
    stats, pool = features.cache_classifier_inputs(vgg, x.cpu(), batch=4, device=dev)
    assert tuple(stats.shape) == (B, 1920) and tuple(pool.shape) == (B, 25088)
    got2c = c2(stats)
    assert float((got2c - got2).abs().max()) <= 1e-2 * float(got2.abs().max())
    flat = torch.nn.functional.adaptive_avg_pool2d(p5, (7, 7)).flatten(1)
    assert float((pool.float() - flat).abs().max()) <= 2e-2 * float(flat.abs().max()) + 1e-3
