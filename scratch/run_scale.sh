#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{' | tail -1 > gpurun_out/scale_n$N.json
cat gpurun_out/scale_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value',round(d['value'],1),'per-gpu',round(d['value']/d['n_gpus'],1),'e2e',round(d['e2e']['value'],1),'feat',round(d['secondary']['value'],1),'ms/step',round(d['ms_per_step'],2), d['clocks'])"
