#!/bin/bash
# round 2: the driver's command line at N GPUs (default bench: NST leg, e2e, 4-tap / 5-tap feature legs) + the reference arm + frames2020
N=$1
mkdir -p gpurun_out/r02
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02/scale_n$N.err | grep -E '^\{' | tail -1 > gpurun_out/r02/scale_n$N.json
timeout 600 $RUN bench.py --gpus $N --config frames2020 2> gpurun_out/r02/scale_frames_n$N.err | grep -E '^\{' | tail -1 > gpurun_out/r02/scale_frames_n$N.json
timeout 600 $RUN bench.py --gpus $N --impl reference --steps 5 --warmup 1 2> gpurun_out/r02/scale_ref_n$N.err | grep -E '^\{' | tail -1 > gpurun_out/r02/scale_ref_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02/scale_n$N.json").read())
print('N', d['n_gpus'], 'value', round(d['value'], 1), 'per-gpu', round(d['value'] / d['n_gpus'], 1), 'e2e', round(d['e2e']['value'], 1),
      'feat4', round(d['secondary']['value'], 1), 'feat5', round(d['secondary_5tap']['value'], 1), 'ms/step', round(d['ms_per_step'], 2), d['clocks'])
f = json.loads(open("gpurun_out/r02/scale_frames_n$N.json").read())
print('frames2020', round(f['value'], 1), f.get('ms_per_step'))
r = json.loads(open("gpurun_out/r02/scale_ref_n$N.json").read())
print('reference arm', r.get('value'), r.get('unavailable'))
PY
tail -3 gpurun_out/r02/scale_n$N.err
