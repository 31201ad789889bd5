#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest nst bf16 + lbfgs"; timeout 600 python -m pytest tests/test_gpu_nst.py -m gpu -q -k "bf16 or lbfgs" -s 2>&1 | tail -8
echo "=== bench 2 gpus"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
echo "=== bench 1 gpu bf16 history"; timeout 900 python bench.py --steps 300 --no-cpu-baseline --no-e2e --no-features --history-bf16 2>&1 | tail -1
