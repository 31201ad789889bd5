#!/bin/bash
# feature legs only at N GPUs (peer row exchange), 512 eyes per GPU
N=$1
mkdir -p gpurun_out/r02
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for c in feat4 feat5; do
  timeout 600 $RUN bench.py --gpus $N --config $c 2> gpurun_out/r02/scale_${c}_n$N.err | grep -E '^\{' | tail -1 > gpurun_out/r02/scale_${c}_n$N.json
  python -c "
import json
d = json.loads(open('gpurun_out/r02/scale_${c}_n$N.json').read())
print('$c N', d['n_gpus'], 'value', round(d['value'], 1), 'per-gpu', round(d['value'] / d['n_gpus'], 1))"
done
