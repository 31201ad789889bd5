"""torchrun -N: feature rows exchanged by pushed peer copies (sharding.PeerRows) == rows all-gathered over NCCL; timing of both."""
import os, sys, time, torch
sys.path.insert(0, '.')
import torch.distributed as dist
import iris_b200
from iris_b200 import features, sharding, synthetic
rank, local, world = sharding.init_from_env("nccl")
dev = torch.device("cuda", local)
vgg = iris_b200.VGG19(content_layers=[], style_layers=['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1'], weights="random", seed=0)
n = 96 * world + 5            # ragged last shard
base, _ = synthetic.synthetic_batch(list(range(8)), 160, 96)
imgs = torch.from_numpy(base)[torch.arange(n) % 8].contiguous().pin_memory()
out = {}
for peer in (False, True, True):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    rows = features.extract_features_sharded(vgg, imgs, batch=32, device=dev, peer=peer)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    out[peer] = rows.clone()
    if rank == 0:
        print("peer=%s: %d rows x %d in %.1f ms (%s)" % (peer, rows.shape[0], rows.shape[1], dt * 1e3, type(rows).__name__), flush=True)
same = bool(torch.equal(out[False], out[True]))
# every rank must hold the same matrix
ref = out[True].clone()
dist.broadcast(ref, 0)
same_ranks = bool(torch.equal(ref, out[True]))
print("rank %d: peer == nccl %s, same on all ranks %s, finite %s" % (rank, same, same_ranks, bool(torch.isfinite(out[True]).all())), flush=True)
assert same and same_ranks
dist.barrier()
dist.destroy_process_group()
